// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt_mem_multidim"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt_mem_multidim___024root final {
    CData/*0:0*/ clk;
    CData/*7:0*/ t__DOT__narrow;
    CData/*1:0*/ t__DOT__index0;
    CData/*1:0*/ t__DOT__index1;
    CData/*2:0*/ t__DOT__index2;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    IData/*31:0*/ t__DOT__cyc;
    VlWide<3>/*71:0*/ t__DOT__wide;
    VlWide<3>/*71:0*/ t__DOT__wread;
    IData/*31:0*/ __VactIterCount;
    QData/*63:0*/ t__DOT__crc;
    VlcuMdArray<VlcuLocalArray<VlWide<3>/*71:0*/, 36>, 3, 3, 4> t__DOT__memw;  // thread-local memory
    VlcuMdArray<VlcuSharedArray<CData/*7:0*/, 36>, 3, 3, 4> t__DOT__memn;  // run-time indexed: shared memory [index][lane]
    VlcuSharedArray<CData/*2:0*/, 2> t__DOT__cstyle;  // run-time indexed: shared memory [index][lane]
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vt_mem_multidim__cuda final {
    typedef Vt_mem_multidim___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 488,
        OFF_t__DOT__narrow = 489,
        OFF_t__DOT__index0 = 490,
        OFF_t__DOT__index1 = 491,
        OFF_t__DOT__index2 = 492,
        OFF___Vtrigprevexpr___TOP__clk__0 = 493,
        OFF___VactPhaseResult = 494,
        OFF___VnbaPhaseResult = 495,
        OFF_t__DOT__cyc = 24,
        OFF_t__DOT__wide = 28,
        OFF_t__DOT__wread = 40,
        OFF___VactIterCount = 52,
        OFF_t__DOT__crc = 0,
        OFF_t__DOT__memw = 56,
        OFF_t__DOT__memn = 496,
        OFF_t__DOT__cstyle = 532,
        OFF___VactTriggered = 8,
        OFF___VnbaTriggered = 16,
        kStateBytes = 536,  // S_persist
        kTouchedReadBytes = 104,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 103,  // S_state_wr
    };
    struct Locals final {
        VlWide<3>/*71:0*/ t__DOT__memw[36];
    };
    struct Shared final {
        CData/*7:0*/ t__DOT__memn[36][VLCU_BLOCK];
        CData/*2:0*/ t__DOT__cstyle[2][VLCU_BLOCK];
    };
    VLCU_D static void bind(Root& s, Locals& l, Shared& sh, const VlcuState& st, uint64_t i) {
        (void)l; (void)sh; (void)st; (void)i;
        s.vlNamep = "TOP";
        s.t__DOT__memw.bind(l.t__DOT__memw);
        s.t__DOT__memn.bind(sh.t__DOT__memn, threadIdx.x);
        s.t__DOT__cstyle.bind(sh.t__DOT__cstyle, threadIdx.x);
    }
    VLCU_D static void load(Root& s, const VlcuState& st, uint64_t i) {
        s.clk = st.ld<VlcuMem8>(OFF_clk, i);
        s.t__DOT__narrow = st.ld<VlcuMem8>(OFF_t__DOT__narrow, i);
        s.t__DOT__index0 = st.ld<VlcuMem8>(OFF_t__DOT__index0, i);
        s.t__DOT__index1 = st.ld<VlcuMem8>(OFF_t__DOT__index1, i);
        s.t__DOT__index2 = st.ld<VlcuMem8>(OFF_t__DOT__index2, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        s.t__DOT__cyc = st.ld<IData>(OFF_t__DOT__cyc, i);
        st.ldW(OFF_t__DOT__wide, i, s.t__DOT__wide);
        st.ldW(OFF_t__DOT__wread, i, s.t__DOT__wread);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        s.t__DOT__crc = st.ld<QData>(OFF_t__DOT__crc, i);
        for (uint32_t e = 0; e < 36; ++e) {
            for (uint32_t w = 0; w < 3; ++w) s.t__DOT__memw.flat()[e][w] = st.ldE<EData>(OFF_t__DOT__memw, e * 3 + w, i);
        }
        for (uint32_t e = 0; e < 36; ++e) s.t__DOT__memn.flat()[e] = st.ldE<VlcuMem8>(OFF_t__DOT__memn, e, i);
        for (uint32_t e = 0; e < 2; ++e) s.t__DOT__cstyle[e] = st.ldE<VlcuMem8>(OFF_t__DOT__cstyle, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.st<VlcuMem8>(OFF_t__DOT__narrow, i, s.t__DOT__narrow);
        st.st<VlcuMem8>(OFF_t__DOT__index0, i, s.t__DOT__index0);
        st.st<VlcuMem8>(OFF_t__DOT__index1, i, s.t__DOT__index1);
        st.st<VlcuMem8>(OFF_t__DOT__index2, i, s.t__DOT__index2);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.st<IData>(OFF_t__DOT__cyc, i, s.t__DOT__cyc);
        st.stW(OFF_t__DOT__wide, i, s.t__DOT__wide);
        st.stW(OFF_t__DOT__wread, i, s.t__DOT__wread);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
        st.st<QData>(OFF_t__DOT__crc, i, s.t__DOT__crc);
        for (uint32_t e = 0; e < 36; ++e) {
            for (uint32_t w = 0; w < 3; ++w) st.stE<EData>(OFF_t__DOT__memw, e * 3 + w, i, s.t__DOT__memw.flat()[e][w]);
        }
        for (uint32_t e = 0; e < 36; ++e) st.stE<VlcuMem8>(OFF_t__DOT__memn, e, i, s.t__DOT__memn.flat()[e]);
        for (uint32_t e = 0; e < 2; ++e) st.stE<VlcuMem8>(OFF_t__DOT__cstyle, e, i, s.t__DOT__cstyle[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VactTriggered, e, i, s.__VactTriggered[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VnbaTriggered, e, i, s.__VnbaTriggered[e]);
    }
    VLCU_D static void load_inputs_row(Root& s, const uint32_t* rowp, uint64_t stride, uint64_t i) {
        s.clk = static_cast<VlcuMem8>(rowp[0 * stride + i] & 0x1U);
    }
    VLCU_D static void store_outputs_row(const Root& s, uint32_t* rowp, uint64_t stride, uint64_t i) {
    }
    VLCU_D static uint64_t outputs_sig(const Root& s, uint64_t sig) {
        uint32_t acc = VLCU_SIG_ACC_INIT;
        return vlcu_sig_mix(sig, acc);
    }
    enum : uint32_t { kInWords = 1, kOutWords = 0, kInBytes = 1, kOutBytes = 0, kTraceWords = 0 };

    VLCU_D static void Vt_mem_multidim___024root___eval_static(Vt_mem_multidim___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
    }

    VLCU_D static void Vt_mem_multidim___024root___eval_initial(Vt_mem_multidim___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        Vt_mem_multidim___024root___eval_initial__TOP(vlSelf, vlCtx);
    }

    VLCU_D static void Vt_mem_multidim___024root___eval_initial__TOP(Vt_mem_multidim___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        VlUnpacked<VlUnpacked<IData/*31:0*/, 3>, 3> t__DOT__imem;
        for (int __Vi0 = 0; __Vi0 < 3; ++__Vi0) {
            for (int __Vi1 = 0; __Vi1 < 3; ++__Vi1) {
                t__DOT__imem[__Vi0][__Vi1] = 0;
            }
        }
        // Body
        vlSelfRef.t__DOT__cyc = 0U;
        t__DOT__imem[0U][0U] = 1U;
        vlSelfRef.t__DOT__memw[0U][0U][0U][0U] = 0xfee00012U;
        vlSelfRef.t__DOT__memw[0U][0U][0U][1U] = 0xfee0fee0U;
        vlSelfRef.t__DOT__memw[0U][0U][0U][2U] = 0x000000feU;
        vlSelfRef.t__DOT__memn[0U][0U][0U] = 0x81U;
        vlSelfRef.t__DOT__memw[0U][0U][1U][0U] = 0xfee00013U;
        vlSelfRef.t__DOT__memw[0U][0U][1U][1U] = 0xfee0fee0U;
        vlSelfRef.t__DOT__memw[0U][0U][1U][2U] = 0x000000feU;
        vlSelfRef.t__DOT__memn[0U][0U][1U] = 0x81U;
        vlSelfRef.t__DOT__memw[0U][0U][2U][0U] = 0xfee00014U;
        vlSelfRef.t__DOT__memw[0U][0U][2U][1U] = 0xfee0fee0U;
        vlSelfRef.t__DOT__memw[0U][0U][2U][2U] = 0x000000feU;
        vlSelfRef.t__DOT__memn[0U][0U][2U] = 0x81U;
        vlSelfRef.t__DOT__memw[0U][0U][3U][0U] = 0xfee00015U;
        vlSelfRef.t__DOT__memw[0U][0U][3U][1U] = 0xfee0fee0U;
        vlSelfRef.t__DOT__memw[0U][0U][3U][2U] = 0x000000feU;
        vlSelfRef.t__DOT__memn[0U][0U][3U] = 0x81U;
        t__DOT__imem[0U][1U] = 2U;
        vlSelfRef.t__DOT__memw[0U][1U][0U][0U] = 0xfee00022U;
        vlSelfRef.t__DOT__memw[0U][1U][0U][1U] = 0xfee0fee0U;
        vlSelfRef.t__DOT__memw[0U][1U][0U][2U] = 0x000000feU;
        vlSelfRef.t__DOT__memn[0U][1U][0U] = 0x81U;
        vlSelfRef.t__DOT__memw[0U][1U][1U][0U] = 0xfee00023U;
        vlSelfRef.t__DOT__memw[0U][1U][1U][1U] = 0xfee0fee0U;
        vlSelfRef.t__DOT__memw[0U][1U][1U][2U] = 0x000000feU;
        vlSelfRef.t__DOT__memn[0U][1U][1U] = 0x81U;
        vlSelfRef.t__DOT__memw[0U][1U][2U][0U] = 0xfee00024U;
        vlSelfRef.t__DOT__memw[0U][1U][2U][1U] = 0xfee0fee0U;
        vlSelfRef.t__DOT__memw[0U][1U][2U][2U] = 0x000000feU;
        vlSelfRef.t__DOT__memn[0U][1U][2U] = 0x81U;
        vlSelfRef.t__DOT__memw[0U][1U][3U][0U] = 0xfee00025U;
        vlSelfRef.t__DOT__memw[0U][1U][3U][1U] = 0xfee0fee0U;
        vlSelfRef.t__DOT__memw[0U][1U][3U][2U] = 0x000000feU;
        vlSelfRef.t__DOT__memn[0U][1U][3U] = 0x81U;
        t__DOT__imem[0U][2U] = 3U;
        vlSelfRef.t__DOT__memw[0U][2U][0U][0U] = 0xfee00032U;
        vlSelfRef.t__DOT__memw[0U][2U][0U][1U] = 0xfee0fee0U;
        vlSelfRef.t__DOT__memw[0U][2U][0U][2U] = 0x000000feU;
        vlSelfRef.t__DOT__memn[0U][2U][0U] = 0x81U;
        vlSelfRef.t__DOT__memw[0U][2U][1U][0U] = 0xfee00033U;
        vlSelfRef.t__DOT__memw[0U][2U][1U][1U] = 0xfee0fee0U;
        vlSelfRef.t__DOT__memw[0U][2U][1U][2U] = 0x000000feU;
        vlSelfRef.t__DOT__memn[0U][2U][1U] = 0x81U;
        vlSelfRef.t__DOT__memw[0U][2U][2U][0U] = 0xfee00034U;
        vlSelfRef.t__DOT__memw[0U][2U][2U][1U] = 0xfee0fee0U;
        vlSelfRef.t__DOT__memw[0U][2U][2U][2U] = 0x000000feU;
        vlSelfRef.t__DOT__memn[0U][2U][2U] = 0x81U;
        vlSelfRef.t__DOT__memw[0U][2U][3U][0U] = 0xfee00035U;
        vlSelfRef.t__DOT__memw[0U][2U][3U][1U] = 0xfee0fee0U;
        vlSelfRef.t__DOT__memw[0U][2U][3U][2U] = 0x000000feU;
        vlSelfRef.t__DOT__memn[0U][2U][3U] = 0x81U;
        t__DOT__imem[1U][0U] = 1U;
        vlSelfRef.t__DOT__memw[1U][0U][0U][0U] = 0xfee00112U;
        vlSelfRef.t__DOT__memw[1U][0U][0U][1U] = 0xfee0fee0U;
        vlSelfRef.t__DOT__memw[1U][0U][0U][2U] = 0x000000feU;
        vlSelfRef.t__DOT__memn[1U][0U][0U] = 0x81U;
        vlSelfRef.t__DOT__memw[1U][0U][1U][0U] = 0xfee00113U;
        vlSelfRef.t__DOT__memw[1U][0U][1U][1U] = 0xfee0fee0U;
        vlSelfRef.t__DOT__memw[1U][0U][1U][2U] = 0x000000feU;
        vlSelfRef.t__DOT__memn[1U][0U][1U] = 0x81U;
        vlSelfRef.t__DOT__memw[1U][0U][2U][0U] = 0xfee00114U;
        vlSelfRef.t__DOT__memw[1U][0U][2U][1U] = 0xfee0fee0U;
        vlSelfRef.t__DOT__memw[1U][0U][2U][2U] = 0x000000feU;
        vlSelfRef.t__DOT__memn[1U][0U][2U] = 0x81U;
        vlSelfRef.t__DOT__memw[1U][0U][3U][0U] = 0xfee00115U;
        vlSelfRef.t__DOT__memw[1U][0U][3U][1U] = 0xfee0fee0U;
        vlSelfRef.t__DOT__memw[1U][0U][3U][2U] = 0x000000feU;
        vlSelfRef.t__DOT__memn[1U][0U][3U] = 0x81U;
        t__DOT__imem[1U][1U] = 2U;
        vlSelfRef.t__DOT__memw[1U][1U][0U][0U] = 0xfee00122U;
        vlSelfRef.t__DOT__memw[1U][1U][0U][1U] = 0xfee0fee0U;
        vlSelfRef.t__DOT__memw[1U][1U][0U][2U] = 0x000000feU;
        vlSelfRef.t__DOT__memn[1U][1U][0U] = 0x81U;
        vlSelfRef.t__DOT__memw[1U][1U][1U][0U] = 0xfee00123U;
        vlSelfRef.t__DOT__memw[1U][1U][1U][1U] = 0xfee0fee0U;
        vlSelfRef.t__DOT__memw[1U][1U][1U][2U] = 0x000000feU;
        vlSelfRef.t__DOT__memn[1U][1U][1U] = 0x81U;
        vlSelfRef.t__DOT__memw[1U][1U][2U][0U] = 0xfee00124U;
        vlSelfRef.t__DOT__memw[1U][1U][2U][1U] = 0xfee0fee0U;
        vlSelfRef.t__DOT__memw[1U][1U][2U][2U] = 0x000000feU;
        vlSelfRef.t__DOT__memn[1U][1U][2U] = 0x81U;
        vlSelfRef.t__DOT__memw[1U][1U][3U][0U] = 0xfee00125U;
        vlSelfRef.t__DOT__memw[1U][1U][3U][1U] = 0xfee0fee0U;
        vlSelfRef.t__DOT__memw[1U][1U][3U][2U] = 0x000000feU;
        vlSelfRef.t__DOT__memn[1U][1U][3U] = 0x81U;
        t__DOT__imem[1U][2U] = 3U;
        vlSelfRef.t__DOT__memw[1U][2U][0U][0U] = 0xfee00132U;
        vlSelfRef.t__DOT__memw[1U][2U][0U][1U] = 0xfee0fee0U;
        vlSelfRef.t__DOT__memw[1U][2U][0U][2U] = 0x000000feU;
        vlSelfRef.t__DOT__memn[1U][2U][0U] = 0x81U;
        vlSelfRef.t__DOT__memw[1U][2U][1U][0U] = 0xfee00133U;
        vlSelfRef.t__DOT__memw[1U][2U][1U][1U] = 0xfee0fee0U;
        vlSelfRef.t__DOT__memw[1U][2U][1U][2U] = 0x000000feU;
        vlSelfRef.t__DOT__memn[1U][2U][1U] = 0x81U;
        vlSelfRef.t__DOT__memw[1U][2U][2U][0U] = 0xfee00134U;
        vlSelfRef.t__DOT__memw[1U][2U][2U][1U] = 0xfee0fee0U;
        vlSelfRef.t__DOT__memw[1U][2U][2U][2U] = 0x000000feU;
        vlSelfRef.t__DOT__memn[1U][2U][2U] = 0x81U;
        vlSelfRef.t__DOT__memw[1U][2U][3U][0U] = 0xfee00135U;
        vlSelfRef.t__DOT__memw[1U][2U][3U][1U] = 0xfee0fee0U;
        vlSelfRef.t__DOT__memw[1U][2U][3U][2U] = 0x000000feU;
        vlSelfRef.t__DOT__memn[1U][2U][3U] = 0x81U;
        t__DOT__imem[2U][0U] = 1U;
        vlSelfRef.t__DOT__memw[2U][0U][0U][0U] = 0xfee00212U;
        vlSelfRef.t__DOT__memw[2U][0U][0U][1U] = 0xfee0fee0U;
        vlSelfRef.t__DOT__memw[2U][0U][0U][2U] = 0x000000feU;
        vlSelfRef.t__DOT__memn[2U][0U][0U] = 0x81U;
        vlSelfRef.t__DOT__memw[2U][0U][1U][0U] = 0xfee00213U;
        vlSelfRef.t__DOT__memw[2U][0U][1U][1U] = 0xfee0fee0U;
        vlSelfRef.t__DOT__memw[2U][0U][1U][2U] = 0x000000feU;
        vlSelfRef.t__DOT__memn[2U][0U][1U] = 0x81U;
        vlSelfRef.t__DOT__memw[2U][0U][2U][0U] = 0xfee00214U;
        vlSelfRef.t__DOT__memw[2U][0U][2U][1U] = 0xfee0fee0U;
        vlSelfRef.t__DOT__memw[2U][0U][2U][2U] = 0x000000feU;
        vlSelfRef.t__DOT__memn[2U][0U][2U] = 0x81U;
        vlSelfRef.t__DOT__memw[2U][0U][3U][0U] = 0xfee00215U;
        vlSelfRef.t__DOT__memw[2U][0U][3U][1U] = 0xfee0fee0U;
        vlSelfRef.t__DOT__memw[2U][0U][3U][2U] = 0x000000feU;
        vlSelfRef.t__DOT__memn[2U][0U][3U] = 0x81U;
        t__DOT__imem[2U][1U] = 2U;
        vlSelfRef.t__DOT__memw[2U][1U][0U][0U] = 0xfee00222U;
        vlSelfRef.t__DOT__memw[2U][1U][0U][1U] = 0xfee0fee0U;
        vlSelfRef.t__DOT__memw[2U][1U][0U][2U] = 0x000000feU;
        vlSelfRef.t__DOT__memn[2U][1U][0U] = 0x81U;
        vlSelfRef.t__DOT__memw[2U][1U][1U][0U] = 0xfee00223U;
        vlSelfRef.t__DOT__memw[2U][1U][1U][1U] = 0xfee0fee0U;
        vlSelfRef.t__DOT__memw[2U][1U][1U][2U] = 0x000000feU;
        vlSelfRef.t__DOT__memn[2U][1U][1U] = 0x81U;
        vlSelfRef.t__DOT__memw[2U][1U][2U][0U] = 0xfee00224U;
        vlSelfRef.t__DOT__memw[2U][1U][2U][1U] = 0xfee0fee0U;
        vlSelfRef.t__DOT__memw[2U][1U][2U][2U] = 0x000000feU;
        vlSelfRef.t__DOT__memn[2U][1U][2U] = 0x81U;
        vlSelfRef.t__DOT__memw[2U][1U][3U][0U] = 0xfee00225U;
        vlSelfRef.t__DOT__memw[2U][1U][3U][1U] = 0xfee0fee0U;
        vlSelfRef.t__DOT__memw[2U][1U][3U][2U] = 0x000000feU;
        vlSelfRef.t__DOT__memn[2U][1U][3U] = 0x81U;
        t__DOT__imem[2U][2U] = 3U;
        vlSelfRef.t__DOT__memw[2U][2U][0U][0U] = 0xfee00232U;
        vlSelfRef.t__DOT__memw[2U][2U][0U][1U] = 0xfee0fee0U;
        vlSelfRef.t__DOT__memw[2U][2U][0U][2U] = 0x000000feU;
        vlSelfRef.t__DOT__memn[2U][2U][0U] = 0x81U;
        vlSelfRef.t__DOT__memw[2U][2U][1U][0U] = 0xfee00233U;
        vlSelfRef.t__DOT__memw[2U][2U][1U][1U] = 0xfee0fee0U;
        vlSelfRef.t__DOT__memw[2U][2U][1U][2U] = 0x000000feU;
        vlSelfRef.t__DOT__memn[2U][2U][1U] = 0x81U;
        vlSelfRef.t__DOT__memw[2U][2U][2U][0U] = 0xfee00234U;
        vlSelfRef.t__DOT__memw[2U][2U][2U][1U] = 0xfee0fee0U;
        vlSelfRef.t__DOT__memw[2U][2U][2U][2U] = 0x000000feU;
        vlSelfRef.t__DOT__memn[2U][2U][2U] = 0x81U;
        vlSelfRef.t__DOT__memw[2U][2U][3U][0U] = 0xfee00235U;
        vlSelfRef.t__DOT__memw[2U][2U][3U][1U] = 0xfee0fee0U;
        vlSelfRef.t__DOT__memw[2U][2U][3U][2U] = 0x000000feU;
        vlSelfRef.t__DOT__memn[2U][2U][3U] = 0x81U;
    }

    VLCU_D static void Vt_mem_multidim___024root___eval_final(Vt_mem_multidim___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_mem_multidim___024root___eval_settle(Vt_mem_multidim___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static bool Vt_mem_multidim___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
        typedef bool VlcuFuncRet;
        // Locals
        IData/*31:0*/ n;
        // Body
        n = 0U;
        do {
            if (in[n]) {
                return (1U);
            }
            n = ((IData)(1U) + n);
        } while ((1U > n));
        return (0U);
    }

    VLCU_D static void Vt_mem_multidim___024root___nba_sequent__TOP__0(Vt_mem_multidim___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        IData/*31:0*/ __Vdly__t__DOT__cyc;
        __Vdly__t__DOT__cyc = 0;
        QData/*63:0*/ __Vdly__t__DOT__crc;
        __Vdly__t__DOT__crc = 0;
        CData/*7:0*/ __Vdly__t__DOT__narrow;
        __Vdly__t__DOT__narrow = 0;
        VlWide<3>/*71:0*/ __Vdly__t__DOT__wide;
        VL_ZERO_W(72, __Vdly__t__DOT__wide);
        CData/*1:0*/ __Vdly__t__DOT__index0;
        __Vdly__t__DOT__index0 = 0;
        CData/*1:0*/ __Vdly__t__DOT__index1;
        __Vdly__t__DOT__index1 = 0;
        CData/*2:0*/ __Vdly__t__DOT__index2;
        __Vdly__t__DOT__index2 = 0;
        CData/*2:0*/ __VdlyVal__t__DOT__cstyle__v0;
        __VdlyVal__t__DOT__cstyle__v0 = 0;
        CData/*0:0*/ __VdlyDim0__t__DOT__cstyle__v0;
        __VdlyDim0__t__DOT__cstyle__v0 = 0;
        CData/*0:0*/ __VdlySet__t__DOT__cstyle__v0;
        __VdlySet__t__DOT__cstyle__v0 = 0;
        CData/*2:0*/ __VdlyVal__t__DOT__memn__v0;
        __VdlyVal__t__DOT__memn__v0 = 0;
        CData/*2:0*/ __VdlyLsb__t__DOT__memn__v0;
        __VdlyLsb__t__DOT__memn__v0 = 0;
        CData/*2:0*/ __VdlyDim0__t__DOT__memn__v0;
        __VdlyDim0__t__DOT__memn__v0 = 0;
        CData/*1:0*/ __VdlyDim1__t__DOT__memn__v0;
        __VdlyDim1__t__DOT__memn__v0 = 0;
        CData/*1:0*/ __VdlyDim2__t__DOT__memn__v0;
        __VdlyDim2__t__DOT__memn__v0 = 0;
        CData/*0:0*/ __VdlySet__t__DOT__memn__v0;
        __VdlySet__t__DOT__memn__v0 = 0;
        CData/*7:0*/ __VdlyVal__t__DOT__memn__v1;
        __VdlyVal__t__DOT__memn__v1 = 0;
        CData/*2:0*/ __VdlyDim0__t__DOT__memn__v1;
        __VdlyDim0__t__DOT__memn__v1 = 0;
        CData/*1:0*/ __VdlyDim1__t__DOT__memn__v1;
        __VdlyDim1__t__DOT__memn__v1 = 0;
        CData/*1:0*/ __VdlyDim2__t__DOT__memn__v1;
        __VdlyDim2__t__DOT__memn__v1 = 0;
        CData/*0:0*/ __VdlySet__t__DOT__memn__v1;
        __VdlySet__t__DOT__memn__v1 = 0;
        VlWide<3>/*71:0*/ __VdlyVal__t__DOT__memw__v0;
        VL_ZERO_W(72, __VdlyVal__t__DOT__memw__v0);
        CData/*2:0*/ __VdlyDim0__t__DOT__memw__v0;
        __VdlyDim0__t__DOT__memw__v0 = 0;
        CData/*1:0*/ __VdlyDim1__t__DOT__memw__v0;
        __VdlyDim1__t__DOT__memw__v0 = 0;
        CData/*1:0*/ __VdlyDim2__t__DOT__memw__v0;
        __VdlyDim2__t__DOT__memw__v0 = 0;
        CData/*0:0*/ __VdlySet__t__DOT__memw__v0;
        __VdlySet__t__DOT__memw__v0 = 0;
        CData/*7:0*/ __VdlyVal__t__DOT__memn__v2;
        __VdlyVal__t__DOT__memn__v2 = 0;
        CData/*0:0*/ __VdlySet__t__DOT__memn__v2;
        __VdlySet__t__DOT__memn__v2 = 0;
        // Body
        __Vdly__t__DOT__cyc = vlSelfRef.t__DOT__cyc;
        __Vdly__t__DOT__crc = vlSelfRef.t__DOT__crc;
        __Vdly__t__DOT__narrow = vlSelfRef.t__DOT__narrow;
        __Vdly__t__DOT__wide[0U] = vlSelfRef.t__DOT__wide[0U];
        __Vdly__t__DOT__wide[1U] = vlSelfRef.t__DOT__wide[1U];
        __Vdly__t__DOT__wide[2U] = vlSelfRef.t__DOT__wide[2U];
        __Vdly__t__DOT__index0 = vlSelfRef.t__DOT__index0;
        __Vdly__t__DOT__index1 = vlSelfRef.t__DOT__index1;
        __Vdly__t__DOT__index2 = vlSelfRef.t__DOT__index2;
        __VdlySet__t__DOT__cstyle__v0 = 0U;
        __VdlySet__t__DOT__memw__v0 = 0U;
        __VdlySet__t__DOT__memn__v0 = 0U;
        __VdlySet__t__DOT__memn__v1 = 0U;
        __VdlySet__t__DOT__memn__v2 = 0U;
        __Vdly__t__DOT__cyc = ((IData)(1U) + vlSelfRef.t__DOT__cyc);
        if ((0U == vlSelfRef.t__DOT__cyc)) {
            __Vdly__t__DOT__crc = 0x5aef0c8dd70a4497ULL;
            __Vdly__t__DOT__narrow = 0U;
            __Vdly__t__DOT__wide[0U] = 0U;
            __Vdly__t__DOT__wide[1U] = 0U;
            __Vdly__t__DOT__wide[2U] = 0U;
            __Vdly__t__DOT__index0 = 0U;
            __Vdly__t__DOT__index1 = 0U;
            __Vdly__t__DOT__index2 = 0U;
        } else if (VL_GTS_III(32, 0x0000005aU, vlSelfRef.t__DOT__cyc)) {
            __VdlyVal__t__DOT__cstyle__v0 = (7U & vlSelfRef.t__DOT__cyc);
            __VdlyDim0__t__DOT__cstyle__v0 = (1U & vlSelfRef.t__DOT__cyc);
            __VdlySet__t__DOT__cstyle__v0 = 1U;
            if (VL_LTS_III(32, 0x00000014U, vlSelfRef.t__DOT__cyc)) {
                if (VL_UNLIKELY(((vlSelfRef.t__DOT__cstyle
                                  [(1U & (~ vlSelfRef.t__DOT__cyc))] 
                                  != (7U & (vlSelfRef.t__DOT__cyc 
                                            - (IData)(1U))))))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_mem_multidim.v", 80, "");
                }
            }
            __Vdly__t__DOT__index0 = (3U & (IData)(vlSelfRef.t__DOT__crc));
            __Vdly__t__DOT__index1 = (3U & (IData)(
                                                   (vlSelfRef.t__DOT__crc 
                                                    >> 2U)));
            __Vdly__t__DOT__index2 = (7U & (IData)(
                                                   (vlSelfRef.t__DOT__crc 
                                                    >> 4U)));
            __Vdly__t__DOT__crc = ((vlSelfRef.t__DOT__crc 
                                    << 1U) | (QData)((IData)(
                                                             (1U 
                                                              & VL_REDXOR_64(
                                                                             (0x8000000000000005ULL 
                                                                              & vlSelfRef.t__DOT__crc))))));
            if ((((((1U <= (IData)(vlSelfRef.t__DOT__index0)) 
                    & (2U >= (IData)(vlSelfRef.t__DOT__index0))) 
                   & (2U <= (IData)(vlSelfRef.t__DOT__index1))) 
                  & (3U <= (IData)(vlSelfRef.t__DOT__index2))) 
                 & (5U >= (IData)(vlSelfRef.t__DOT__index2)))) {
                __Vdly__t__DOT__narrow = (((0x000000feU 
                                            & ((IData)(vlSelfRef.t__DOT__narrow) 
                                               << 1U)) 
                                           | (1U & 
                                              (((IData)(vlSelfRef.t__DOT__narrow) 
                                                >> 7U) 
                                               ^ (IData)(vlSelfRef.t__DOT__narrow)))) 
                                          ^ ((3U >= 
                                              (7U & 
                                               ((IData)(vlSelfRef.t__DOT__index2) 
                                                - (IData)(2U))))
                                              ? vlSelfRef.t__DOT__memn
                                             [((2U 
                                                >= (IData)(vlSelfRef.t__DOT__index0))
                                                ? (IData)(vlSelfRef.t__DOT__index0)
                                                : 0U)]
                                             [((2U 
                                                >= 
                                                (3U 
                                                 & ((IData)(vlSelfRef.t__DOT__index1) 
                                                    - (IData)(1U))))
                                                ? (3U 
                                                   & ((IData)(vlSelfRef.t__DOT__index1) 
                                                      - (IData)(1U)))
                                                : 0U)]
                                             [(7U & 
                                               ((IData)(vlSelfRef.t__DOT__index2) 
                                                - (IData)(2U)))]
                                              : 0U));
                if ((3U >= (7U & ((IData)(vlSelfRef.t__DOT__index2) 
                                  - (IData)(2U))))) {
                    vlSelfRef.t__DOT__wread[0U] = vlSelfRef.t__DOT__memw
                        [((2U >= (IData)(vlSelfRef.t__DOT__index0))
                           ? (IData)(vlSelfRef.t__DOT__index0)
                           : 0U)][((2U >= (3U & ((IData)(vlSelfRef.t__DOT__index1) 
                                                 - (IData)(1U))))
                                    ? (3U & ((IData)(vlSelfRef.t__DOT__index1) 
                                             - (IData)(1U)))
                                    : 0U)][(7U & ((IData)(vlSelfRef.t__DOT__index2) 
                                                  - (IData)(2U)))][0U];
                    vlSelfRef.t__DOT__wread[1U] = vlSelfRef.t__DOT__memw
                        [((2U >= (IData)(vlSelfRef.t__DOT__index0))
                           ? (IData)(vlSelfRef.t__DOT__index0)
                           : 0U)][((2U >= (3U & ((IData)(vlSelfRef.t__DOT__index1) 
                                                 - (IData)(1U))))
                                    ? (3U & ((IData)(vlSelfRef.t__DOT__index1) 
                                             - (IData)(1U)))
                                    : 0U)][(7U & ((IData)(vlSelfRef.t__DOT__index2) 
                                                  - (IData)(2U)))][1U];
                    vlSelfRef.t__DOT__wread[2U] = vlSelfRef.t__DOT__memw
                        [((2U >= (IData)(vlSelfRef.t__DOT__index0))
                           ? (IData)(vlSelfRef.t__DOT__index0)
                           : 0U)][((2U >= (3U & ((IData)(vlSelfRef.t__DOT__index1) 
                                                 - (IData)(1U))))
                                    ? (3U & ((IData)(vlSelfRef.t__DOT__index1) 
                                             - (IData)(1U)))
                                    : 0U)][(7U & ((IData)(vlSelfRef.t__DOT__index2) 
                                                  - (IData)(2U)))][2U];
                } else {
                    vlSelfRef.t__DOT__wread[0U] = 0U;
                    vlSelfRef.t__DOT__wread[1U] = 0U;
                    vlSelfRef.t__DOT__wread[2U] = 0U;
                }
                __Vdly__t__DOT__wide[0U] = (((vlSelfRef.t__DOT__wide[0U] 
                                              << 1U) 
                                             | (1U 
                                                & (((vlSelfRef.t__DOT__wide[2U] 
                                                     >> 7U) 
                                                    ^ 
                                                    (vlSelfRef.t__DOT__wide[0U] 
                                                     >> 2U)) 
                                                   ^ vlSelfRef.t__DOT__wide[0U]))) 
                                            ^ vlSelfRef.t__DOT__wread[0U]);
                __Vdly__t__DOT__wide[1U] = (((vlSelfRef.t__DOT__wide[0U] 
                                              >> 0x0000001fU) 
                                             | (vlSelfRef.t__DOT__wide[1U] 
                                                << 1U)) 
                                            ^ vlSelfRef.t__DOT__wread[1U]);
                __Vdly__t__DOT__wide[2U] = (((vlSelfRef.t__DOT__wide[1U] 
                                              >> 0x0000001fU) 
                                             | (0x000000feU 
                                                & (vlSelfRef.t__DOT__wide[2U] 
                                                   << 1U))) 
                                            ^ vlSelfRef.t__DOT__wread[2U]);
            }
            if (VL_LIKELY((((3U >= (7U & ((IData)(vlSelfRef.t__DOT__index2) 
                                          - (IData)(2U)))) 
                            && ((2U >= (3U & ((IData)(vlSelfRef.t__DOT__index1) 
                                              - (IData)(1U)))) 
                                && (2U >= (IData)(vlSelfRef.t__DOT__index0))))))) {
                __VdlyVal__t__DOT__memn__v0 = (7U & (IData)(vlSelfRef.t__DOT__crc));
                __VdlyLsb__t__DOT__memn__v0 = (7U & (IData)(
                                                            (vlSelfRef.t__DOT__crc 
                                                             >> 8U)));
                __VdlyDim0__t__DOT__memn__v0 = (7U 
                                                & ((IData)(vlSelfRef.t__DOT__index2) 
                                                   - (IData)(2U)));
                __VdlyDim1__t__DOT__memn__v0 = (3U 
                                                & ((IData)(vlSelfRef.t__DOT__index1) 
                                                   - (IData)(1U)));
                __VdlyDim2__t__DOT__memn__v0 = vlSelfRef.t__DOT__index0;
                __VdlySet__t__DOT__memn__v0 = 1U;
                __VdlyVal__t__DOT__memn__v1 = ((0x000000feU 
                                                & ((~ (IData)(vlSelfRef.t__DOT__crc)) 
                                                   << 1U)) 
                                               | (1U 
                                                  & (IData)(
                                                            (vlSelfRef.t__DOT__crc 
                                                             >> 7U))));
                __VdlyDim0__t__DOT__memn__v1 = (7U 
                                                & ((IData)(vlSelfRef.t__DOT__index2) 
                                                   - (IData)(2U)));
                __VdlyDim1__t__DOT__memn__v1 = (3U 
                                                & ((IData)(vlSelfRef.t__DOT__index1) 
                                                   - (IData)(1U)));
                __VdlyDim2__t__DOT__memn__v1 = vlSelfRef.t__DOT__index0;
                __VdlySet__t__DOT__memn__v1 = 1U;
                __VdlyVal__t__DOT__memw__v0[0U] = (IData)(vlSelfRef.t__DOT__crc);
                __VdlyVal__t__DOT__memw__v0[1U] = (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 0x00000020U));
                __VdlyVal__t__DOT__memw__v0[2U] = (0x000000ffU 
                                                   & (~ (IData)(vlSelfRef.t__DOT__crc)));
                __VdlyDim0__t__DOT__memw__v0 = (7U 
                                                & ((IData)(vlSelfRef.t__DOT__index2) 
                                                   - (IData)(2U)));
                __VdlyDim1__t__DOT__memw__v0 = (3U 
                                                & ((IData)(vlSelfRef.t__DOT__index1) 
                                                   - (IData)(1U)));
                __VdlyDim2__t__DOT__memw__v0 = vlSelfRef.t__DOT__index0;
                __VdlySet__t__DOT__memw__v0 = 1U;
            }
        } else if ((0x0000005aU == vlSelfRef.t__DOT__cyc)) {
            __VdlyVal__t__DOT__memn__v2 = (0xa8U ^ vlSelfRef.t__DOT__memn[0U][0U][1U]);
            __VdlySet__t__DOT__memn__v2 = 1U;
        } else if ((0x0000005bU != vlSelfRef.t__DOT__cyc)) {
            if (VL_UNLIKELY(((0x00000063U == vlSelfRef.t__DOT__cyc)))) {
                vlCtx.writef(0U, 5U, vlcu_msg_arg(VL_TIME_UNITED_Q(1)), vlcu_msg_arg(vlSelfRef.t__DOT__cyc), vlcu_msg_arg(vlSelfRef.t__DOT__crc), vlcu_msg_arg((IData)(vlSelfRef.t__DOT__narrow)), vlcu_msg_arg(vlSelfRef.t__DOT__wide));
                if (VL_UNLIKELY(((0x65e3bddcd9bc2750ULL 
                                  != vlSelfRef.t__DOT__crc)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_mem_multidim.v", 89, "");
                }
                if (VL_UNLIKELY(((0xcaU != (IData)(vlSelfRef.t__DOT__narrow))))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_mem_multidim.v", 90, "");
                }
                if (VL_UNLIKELY(((0U != (((0xa6873f73U 
                                           ^ vlSelfRef.t__DOT__wide[0U]) 
                                          | (0xdafed31bU 
                                             ^ vlSelfRef.t__DOT__wide[1U])) 
                                         | (0x0000004eU 
                                            ^ vlSelfRef.t__DOT__wide[2U])))))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_mem_multidim.v", 91, "");
                }
                vlCtx.writef(1U, 0U);
                VL_FINISH_MT("/root/reference/test_regress/t/t_mem_multidim.v", 93, "");
            }
        }
        vlSelfRef.t__DOT__cyc = __Vdly__t__DOT__cyc;
        vlSelfRef.t__DOT__crc = __Vdly__t__DOT__crc;
        vlSelfRef.t__DOT__narrow = __Vdly__t__DOT__narrow;
        vlSelfRef.t__DOT__wide[0U] = __Vdly__t__DOT__wide[0U];
        vlSelfRef.t__DOT__wide[1U] = __Vdly__t__DOT__wide[1U];
        vlSelfRef.t__DOT__wide[2U] = __Vdly__t__DOT__wide[2U];
        vlSelfRef.t__DOT__index0 = __Vdly__t__DOT__index0;
        vlSelfRef.t__DOT__index1 = __Vdly__t__DOT__index1;
        vlSelfRef.t__DOT__index2 = __Vdly__t__DOT__index2;
        if (__VdlySet__t__DOT__cstyle__v0) {
            vlSelfRef.t__DOT__cstyle[__VdlyDim0__t__DOT__cstyle__v0] 
                = __VdlyVal__t__DOT__cstyle__v0;
        }
        if (__VdlySet__t__DOT__memw__v0) {
            vlSelfRef.t__DOT__memw[__VdlyDim2__t__DOT__memw__v0][__VdlyDim1__t__DOT__memw__v0][__VdlyDim0__t__DOT__memw__v0][0U] 
                = __VdlyVal__t__DOT__memw__v0[0U];
            vlSelfRef.t__DOT__memw[__VdlyDim2__t__DOT__memw__v0][__VdlyDim1__t__DOT__memw__v0][__VdlyDim0__t__DOT__memw__v0][1U] 
                = __VdlyVal__t__DOT__memw__v0[1U];
            vlSelfRef.t__DOT__memw[__VdlyDim2__t__DOT__memw__v0][__VdlyDim1__t__DOT__memw__v0][__VdlyDim0__t__DOT__memw__v0][2U] 
                = __VdlyVal__t__DOT__memw__v0[2U];
        }
        if (__VdlySet__t__DOT__memn__v0) {
            vlSelfRef.t__DOT__memn[__VdlyDim2__t__DOT__memn__v0][__VdlyDim1__t__DOT__memn__v0][__VdlyDim0__t__DOT__memn__v0] 
                = (((~ ((CData)(7U) << (IData)(__VdlyLsb__t__DOT__memn__v0))) 
                    & vlSelfRef.t__DOT__memn[__VdlyDim2__t__DOT__memn__v0]
                    [__VdlyDim1__t__DOT__memn__v0][__VdlyDim0__t__DOT__memn__v0]) 
                   | (0x00ffU & ((CData)((IData)(__VdlyVal__t__DOT__memn__v0)) 
                                 << (IData)(__VdlyLsb__t__DOT__memn__v0))));
        }
        if (__VdlySet__t__DOT__memn__v1) {
            vlSelfRef.t__DOT__memn[__VdlyDim2__t__DOT__memn__v1][__VdlyDim1__t__DOT__memn__v1][__VdlyDim0__t__DOT__memn__v1] 
                = __VdlyVal__t__DOT__memn__v1;
        }
        if (__VdlySet__t__DOT__memn__v2) {
            vlSelfRef.t__DOT__memn[0U][0U][1U] = __VdlyVal__t__DOT__memn__v2;
        }
    }

    VLCU_D static void Vt_mem_multidim___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
        typedef void VlcuFuncRet;
        // Locals
        IData/*31:0*/ n;
        // Body
        n = 0U;
        do {
            out[n] = (out[n] | in[n]);
            n = ((IData)(1U) + n);
        } while ((0U >= n));
    }

    VLCU_D static bool Vt_mem_multidim___024root___eval_phase__act(Vt_mem_multidim___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        {
            // Inlined CFunc: _eval_triggers_vec__act
            vlSelfRef.__VactTriggered[0U] = (QData)((IData)(
                                                            ((IData)(vlSelfRef.clk) 
                                                             & (~ (IData)(vlSelfRef.__Vtrigprevexpr___TOP__clk__0)))));
            vlSelfRef.__Vtrigprevexpr___TOP__clk__0 
                = vlSelfRef.clk;
        }
#ifdef VL_DEBUG
        if (VL_UNLIKELY(vlSymsp->_vm_contextp__->debug())) {
            Vt_mem_multidim___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vt_mem_multidim___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vt_mem_multidim___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
        typedef void VlcuFuncRet;
        // Locals
        IData/*31:0*/ n;
        // Body
        n = 0U;
        do {
            out[n] = 0ULL;
            n = ((IData)(1U) + n);
        } while ((1U > n));
    }

    VLCU_D static bool Vt_mem_multidim___024root___eval_phase__nba(Vt_mem_multidim___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vt_mem_multidim___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    Vt_mem_multidim___024root___nba_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
            Vt_mem_multidim___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vt_mem_multidim___024root___eval(Vt_mem_multidim___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        IData/*31:0*/ __VnbaIterCount;
        // Body
        __VnbaIterCount = 0U;
        do {
            if (VL_UNLIKELY(((0x00002710U < __VnbaIterCount)))) {
#ifdef VL_DEBUG
                Vt_mem_multidim___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("/root/reference/test_regress/t/t_mem_multidim.v", 7, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vt_mem_multidim___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("/root/reference/test_regress/t/t_mem_multidim.v", 7, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vt_mem_multidim___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vt_mem_multidim___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vt_mem_multidim___024root___ctor_var_reset(Vt_mem_multidim___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        for (int __Vi0 = 0; __Vi0 < 3; ++__Vi0) {
            for (int __Vi1 = 0; __Vi1 < 3; ++__Vi1) {
                for (int __Vi2 = 0; __Vi2 < 4; ++__Vi2) {
                    VL_SCOPED_RAND_RESET_W(72, vlSelf->t__DOT__memw[__Vi0][__Vi1][__Vi2], __VscopeHash, 13565296381900315573ull);
                }
            }
        }
        for (int __Vi0 = 0; __Vi0 < 3; ++__Vi0) {
            for (int __Vi1 = 0; __Vi1 < 3; ++__Vi1) {
                for (int __Vi2 = 0; __Vi2 < 4; ++__Vi2) {
                    vlSelf->t__DOT__memn[__Vi0][__Vi1][__Vi2] = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 11813441450269430454ull);
                }
            }
        }
        vlSelf->t__DOT__cyc = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 15508623621340971111ull);
        vlSelf->t__DOT__crc = VL_SCOPED_RAND_RESET_Q(64, __VscopeHash, 17680824324711171631ull);
        VL_SCOPED_RAND_RESET_W(72, vlSelf->t__DOT__wide, __VscopeHash, 13798169374321348451ull);
        vlSelf->t__DOT__narrow = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 16957994623349195476ull);
        vlSelf->t__DOT__index0 = VL_SCOPED_RAND_RESET_I(2, __VscopeHash, 15644075809843603140ull);
        vlSelf->t__DOT__index1 = VL_SCOPED_RAND_RESET_I(2, __VscopeHash, 11710272881616179728ull);
        vlSelf->t__DOT__index2 = VL_SCOPED_RAND_RESET_I(3, __VscopeHash, 9692872366076411579ull);
        for (int __Vi0 = 0; __Vi0 < 2; ++__Vi0) {
            vlSelf->t__DOT__cstyle[__Vi0] = VL_SCOPED_RAND_RESET_I(3, __VscopeHash, 7934909833828868595ull);
        }
        VL_SCOPED_RAND_RESET_W(72, vlSelf->t__DOT__wread, __VscopeHash, 5369758569704348851ull);
        for (int __Vi0 = 0; __Vi0 < 1; ++__Vi0) {
            vlSelf->__VactTriggered[__Vi0] = 0;
        }
        vlSelf->__Vtrigprevexpr___TOP__clk__0 = 0;
        for (int __Vi0 = 0; __Vi0 < 1; ++__Vi0) {
            vlSelf->__VnbaTriggered[__Vi0] = 0;
        }
    }

    // tfp->dump(): every traced value stored at its trace code.  Full functions, then the
    // constants, as VerilatedTrace::dump runs them the first time (verilated_trace_imp.h:348-358)
    VLCU_D static void trace_sample(Root* vlSelf, VlcuContext& vlCtx, uint32_t* valsp) {
        (void)vlSelf; (void)vlCtx; (void)valsp;
    }

    VLCU_D static void ctor_var_reset(Root* vlSelf, VlcuContext& vlCtx) {
        Vt_mem_multidim___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt_mem_multidim___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt_mem_multidim___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt_mem_multidim___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vt_mem_multidim___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt_mem_multidim___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt_mem_multidim__cuda::info() {
    typedef Vt_mem_multidim__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"t__DOT__narrow", "t.narrow", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF_t__DOT__narrow},
        {"t__DOT__index0", "t.index0", VLCU_MEMBER_STATE, 1, 0, 1, 2, 1, 0, T::OFF_t__DOT__index0},
        {"t__DOT__index1", "t.index1", VLCU_MEMBER_STATE, 1, 0, 1, 2, 1, 0, T::OFF_t__DOT__index1},
        {"t__DOT__index2", "t.index2", VLCU_MEMBER_STATE, 1, 0, 1, 3, 2, 0, T::OFF_t__DOT__index2},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"t__DOT__cyc", "t.cyc", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__cyc},
        {"t__DOT__wide", "t.wide", VLCU_MEMBER_STATE, 4, 0, 3, 72, 71, 0, T::OFF_t__DOT__wide},
        {"t__DOT__wread", "t.wread", VLCU_MEMBER_STATE, 4, 0, 3, 72, 71, 0, T::OFF_t__DOT__wread},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"t__DOT__crc", "t.crc", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF_t__DOT__crc},
        {"t__DOT__memw", "t.memw", VLCU_MEMBER_STATE, 4, 0, 108, 72, 71, 0, T::OFF_t__DOT__memw},
        {"t__DOT__memn", "t.memn", VLCU_MEMBER_STATE, 1, 0, 36, 8, 7, 0, T::OFF_t__DOT__memn},
        {"t__DOT__cstyle", "t.cstyle", VLCU_MEMBER_STATE, 1, 0, 2, 3, 2, 0, T::OFF_t__DOT__cstyle},
        {"__VactTriggered", "__VactTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VactTriggered},
        {"__VnbaTriggered", "__VnbaTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VnbaTriggered},
    };
    static const char* const messages[] = {"[%0t] cyc==%0d crc=%x nar=%x wide=%x\n", "*-* All Finished *-*\n", nullptr};
    static const char* const plusargs[] = {nullptr};
    static const VlcuTraceDecl traceDecls[] = {
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0}};  // (terminator, not counted)
    static const uint32_t traceOrder[] = {0};  // codes in dump order (+ terminator)
    static const uint32_t saveOrder[] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "t", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 2, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 18, 0, 0x87da0bcdd9ba2f6dULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt_mem_multidim__cuda VLCU_TESTBENCHES);
