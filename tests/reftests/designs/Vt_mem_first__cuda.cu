// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt_mem_first"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt_mem_first___024root final {
    CData/*0:0*/ clk;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    IData/*31:0*/ t__DOT___mode;
    IData/*31:0*/ __VactIterCount;
    QData/*34:0*/ t__DOT__vec_wide;
    VlcuLocalArray<CData/*7:0*/, 32> t__DOT__mem_narrow;  // thread-local memory
    VlcuGlobalArray<VlWide<3>/*77:0*/, 1025> t__DOT__mem_wide;  // stays in HBM, instance-major
    VlUnpacked<CData/*7:0*/, 2> t__DOT__mem_dly_narrow;
    VlUnpacked<VlWide<3>/*77:0*/, 2> t__DOT__mem_dly_wide;
    VlUnpacked<IData/*31:0*/, 16> t__DOT__wrd0;
    VlUnpacked<CData/*2:0*/, 2> t__DOT__np2_guard;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vt_mem_first__cuda final {
    typedef Vt_mem_first___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 12420,
        OFF___Vtrigprevexpr___TOP__clk__0 = 12421,
        OFF___VactPhaseResult = 12422,
        OFF___VnbaPhaseResult = 12423,
        OFF_t__DOT___mode = 24,
        OFF___VactIterCount = 28,
        OFF_t__DOT__vec_wide = 0,
        OFF_t__DOT__mem_narrow = 12424,
        OFF_t__DOT__mem_wide = 32,
        OFF_t__DOT__mem_dly_narrow = 12456,
        OFF_t__DOT__mem_dly_wide = 12332,
        OFF_t__DOT__wrd0 = 12356,
        OFF_t__DOT__np2_guard = 12458,
        OFF___VactTriggered = 8,
        OFF___VnbaTriggered = 16,
        kStateBytes = 12464,  // S_persist
        kTouchedReadBytes = 1252,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 1251,  // S_state_wr
    };
    struct Locals final {
        CData/*7:0*/ t__DOT__mem_narrow[32];
    };
    struct Shared final {
    };
    VLCU_D static void bind(Root& s, Locals& l, Shared& sh, const VlcuState& st, uint64_t i) {
        (void)l; (void)sh; (void)st; (void)i;
        s.vlNamep = "TOP";
        s.t__DOT__mem_narrow.bind(l.t__DOT__mem_narrow);
        s.t__DOT__mem_wide.bind(st, OFF_t__DOT__mem_wide, i);
    }
    VLCU_D static void load(Root& s, const VlcuState& st, uint64_t i) {
        s.clk = st.ld<VlcuMem8>(OFF_clk, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        s.t__DOT___mode = st.ld<IData>(OFF_t__DOT___mode, i);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        s.t__DOT__vec_wide = st.ld<QData>(OFF_t__DOT__vec_wide, i);
        for (uint32_t e = 0; e < 32; ++e) s.t__DOT__mem_narrow[e] = st.ldE<VlcuMem8>(OFF_t__DOT__mem_narrow, e, i);
        for (uint32_t e = 0; e < 2; ++e) s.t__DOT__mem_dly_narrow[e] = st.ldE<VlcuMem8>(OFF_t__DOT__mem_dly_narrow, e, i);
        for (uint32_t e = 0; e < 2; ++e) {
            for (uint32_t w = 0; w < 3; ++w) s.t__DOT__mem_dly_wide[e][w] = st.ldE<EData>(OFF_t__DOT__mem_dly_wide, e * 3 + w, i);
        }
        for (uint32_t e = 0; e < 16; ++e) s.t__DOT__wrd0[e] = st.ldE<IData>(OFF_t__DOT__wrd0, e, i);
        for (uint32_t e = 0; e < 2; ++e) s.t__DOT__np2_guard[e] = st.ldE<VlcuMem8>(OFF_t__DOT__np2_guard, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.st<IData>(OFF_t__DOT___mode, i, s.t__DOT___mode);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
        st.st<QData>(OFF_t__DOT__vec_wide, i, s.t__DOT__vec_wide);
        for (uint32_t e = 0; e < 32; ++e) st.stE<VlcuMem8>(OFF_t__DOT__mem_narrow, e, i, s.t__DOT__mem_narrow[e]);
        for (uint32_t e = 0; e < 2; ++e) st.stE<VlcuMem8>(OFF_t__DOT__mem_dly_narrow, e, i, s.t__DOT__mem_dly_narrow[e]);
        for (uint32_t e = 0; e < 2; ++e) {
            for (uint32_t w = 0; w < 3; ++w) st.stE<EData>(OFF_t__DOT__mem_dly_wide, e * 3 + w, i, s.t__DOT__mem_dly_wide[e][w]);
        }
        for (uint32_t e = 0; e < 16; ++e) st.stE<IData>(OFF_t__DOT__wrd0, e, i, s.t__DOT__wrd0[e]);
        for (uint32_t e = 0; e < 2; ++e) st.stE<VlcuMem8>(OFF_t__DOT__np2_guard, e, i, s.t__DOT__np2_guard[e]);
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

    VLCU_D static void Vt_mem_first___024root___eval_static(Vt_mem_first___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
    }

    VLCU_D static void Vt_mem_first___024root___eval_initial(Vt_mem_first___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        {
            // Inlined CFunc: _eval_initial__TOP
            vlSelfRef.t__DOT___mode = 0U;
        }
    }

    VLCU_D static void Vt_mem_first___024root___eval_final(Vt_mem_first___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_mem_first___024root___eval_settle(Vt_mem_first___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static bool Vt_mem_first___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt_mem_first___024root___nba_sequent__TOP__0(Vt_mem_first___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        VlUnpacked<CData/*2:0*/, 6> t__DOT__np2_mem;
        for (int __Vi0 = 0; __Vi0 < 6; ++__Vi0) {
            t__DOT__np2_mem[__Vi0] = 0;
        }
        QData/*34:0*/ __Vdly__t__DOT__vec_wide;
        __Vdly__t__DOT__vec_wide = 0;
        IData/*31:0*/ __Vdly__t__DOT___mode;
        __Vdly__t__DOT___mode = 0;
        CData/*0:0*/ __VdlySet__t__DOT__mem_dly_narrow__v0;
        __VdlySet__t__DOT__mem_dly_narrow__v0 = 0;
        CData/*0:0*/ __VdlySet__t__DOT__mem_dly_narrow__v1;
        __VdlySet__t__DOT__mem_dly_narrow__v1 = 0;
        // Body
        __Vdly__t__DOT__vec_wide = vlSelfRef.t__DOT__vec_wide;
        __Vdly__t__DOT___mode = vlSelfRef.t__DOT___mode;
        __VdlySet__t__DOT__mem_dly_narrow__v0 = 0U;
        __VdlySet__t__DOT__mem_dly_narrow__v1 = 0U;
        if (VL_UNLIKELY(((0U != vlSelfRef.t__DOT___mode)))) {
            vlSelfRef.t__DOT__wrd0[0U] = 1U;
            vlSelfRef.t__DOT__mem_narrow[0U] = 0U;
            vlSelfRef.t__DOT__np2_guard[0U] = 0U;
            vlSelfRef.t__DOT__mem_wide[0U][0U] = 0x0000feedU;
            vlSelfRef.t__DOT__mem_wide[0U][1U] = 0U;
            vlSelfRef.t__DOT__mem_wide[0U][2U] = 0U;
            __Vdly__t__DOT__vec_wide = 0ULL;
            vlSelfRef.t__DOT__np2_guard[1U] = 0U;
            vlSelfRef.t__DOT__mem_narrow[1U] = 1U;
            vlSelfRef.t__DOT__mem_wide[1U][0U] = 0x0000feedU;
            vlSelfRef.t__DOT__mem_wide[1U][1U] = 0U;
            vlSelfRef.t__DOT__mem_wide[1U][2U] = 0x00000040U;
            vlSelfRef.t__DOT__mem_narrow[2U] = 2U;
            vlSelfRef.t__DOT__mem_wide[2U][0U] = 0x0000feedU;
            vlSelfRef.t__DOT__mem_wide[2U][1U] = 0U;
            vlSelfRef.t__DOT__mem_wide[2U][2U] = 0x00000080U;
            vlSelfRef.t__DOT__mem_narrow[3U] = 3U;
            vlSelfRef.t__DOT__mem_wide[3U][0U] = 0x0000feedU;
            vlSelfRef.t__DOT__mem_wide[3U][1U] = 0U;
            vlSelfRef.t__DOT__mem_wide[3U][2U] = 0x000000c0U;
            vlSelfRef.t__DOT__mem_narrow[4U] = 4U;
            vlSelfRef.t__DOT__mem_wide[4U][0U] = 0x0000feedU;
            vlSelfRef.t__DOT__mem_wide[4U][1U] = 0U;
            vlSelfRef.t__DOT__mem_wide[4U][2U] = 0x00000100U;
            vlSelfRef.t__DOT__mem_narrow[5U] = 5U;
            vlSelfRef.t__DOT__mem_wide[5U][0U] = 0x0000feedU;
            vlSelfRef.t__DOT__mem_wide[5U][1U] = 0U;
            vlSelfRef.t__DOT__mem_wide[5U][2U] = 0x00000140U;
            vlSelfRef.t__DOT__mem_narrow[6U] = 6U;
            vlSelfRef.t__DOT__mem_wide[6U][0U] = 0x0000feedU;
            vlSelfRef.t__DOT__mem_wide[6U][1U] = 0U;
            vlSelfRef.t__DOT__mem_wide[6U][2U] = 0x00000180U;
            vlSelfRef.t__DOT__mem_narrow[7U] = 7U;
            vlSelfRef.t__DOT__mem_wide[7U][0U] = 0x0000feedU;
            vlSelfRef.t__DOT__mem_wide[7U][1U] = 0U;
            vlSelfRef.t__DOT__mem_wide[7U][2U] = 0x000001c0U;
            vlSelfRef.t__DOT__mem_narrow[8U] = 8U;
            vlSelfRef.t__DOT__mem_wide[8U][0U] = 0x0000feedU;
            vlSelfRef.t__DOT__mem_wide[8U][1U] = 0U;
            vlSelfRef.t__DOT__mem_wide[8U][2U] = 0x00000200U;
            vlSelfRef.t__DOT__mem_narrow[9U] = 9U;
            vlSelfRef.t__DOT__mem_wide[9U][0U] = 0x0000feedU;
            vlSelfRef.t__DOT__mem_wide[9U][1U] = 0U;
            vlSelfRef.t__DOT__mem_wide[9U][2U] = 0x00000240U;
            vlSelfRef.t__DOT__mem_narrow[10U] = 0x0aU;
            vlSelfRef.t__DOT__mem_wide[10U][0U] = 0x0000feedU;
            vlSelfRef.t__DOT__mem_wide[10U][1U] = 0U;
            vlSelfRef.t__DOT__mem_wide[10U][2U] = 0x00000280U;
            vlSelfRef.t__DOT__mem_narrow[11U] = 0x0bU;
            vlSelfRef.t__DOT__mem_wide[11U][0U] = 0x0000feedU;
            vlSelfRef.t__DOT__mem_wide[11U][1U] = 0U;
            vlSelfRef.t__DOT__mem_wide[11U][2U] = 0x000002c0U;
            vlSelfRef.t__DOT__mem_narrow[12U] = 0x0cU;
            vlSelfRef.t__DOT__mem_wide[12U][0U] = 0x0000feedU;
            vlSelfRef.t__DOT__mem_wide[12U][1U] = 0U;
            vlSelfRef.t__DOT__mem_wide[12U][2U] = 0x00000300U;
            vlSelfRef.t__DOT__mem_narrow[13U] = 0x0dU;
            vlSelfRef.t__DOT__mem_wide[13U][0U] = 0x0000feedU;
            vlSelfRef.t__DOT__mem_wide[13U][1U] = 0U;
            vlSelfRef.t__DOT__mem_wide[13U][2U] = 0x00000340U;
            vlSelfRef.t__DOT__mem_narrow[14U] = 0x0eU;
            vlSelfRef.t__DOT__mem_wide[14U][0U] = 0x0000feedU;
            vlSelfRef.t__DOT__mem_wide[14U][1U] = 0U;
            vlSelfRef.t__DOT__mem_wide[14U][2U] = 0x00000380U;
            vlSelfRef.t__DOT__mem_narrow[15U] = 0x0fU;
            vlSelfRef.t__DOT__mem_wide[15U][0U] = 0x0000feedU;
            vlSelfRef.t__DOT__mem_wide[15U][1U] = 0U;
            vlSelfRef.t__DOT__mem_wide[15U][2U] = 0x000003c0U;
            vlSelfRef.t__DOT__mem_narrow[16U] = 0x10U;
            vlSelfRef.t__DOT__mem_wide[16U][0U] = 0x0000feedU;
            vlSelfRef.t__DOT__mem_wide[16U][1U] = 0U;
            vlSelfRef.t__DOT__mem_wide[16U][2U] = 0x00000400U;
            vlSelfRef.t__DOT__mem_narrow[17U] = 0x11U;
            vlSelfRef.t__DOT__mem_wide[17U][0U] = 0x0000feedU;
            vlSelfRef.t__DOT__mem_wide[17U][1U] = 0U;
            vlSelfRef.t__DOT__mem_wide[17U][2U] = 0x00000440U;
            vlSelfRef.t__DOT__mem_narrow[18U] = 0x12U;
            vlSelfRef.t__DOT__mem_wide[18U][0U] = 0x0000feedU;
            vlSelfRef.t__DOT__mem_wide[18U][1U] = 0U;
            vlSelfRef.t__DOT__mem_wide[18U][2U] = 0x00000480U;
            vlSelfRef.t__DOT__mem_narrow[19U] = 0x13U;
            vlSelfRef.t__DOT__mem_wide[19U][0U] = 0x0000feedU;
            vlSelfRef.t__DOT__mem_wide[19U][1U] = 0U;
            vlSelfRef.t__DOT__mem_wide[19U][2U] = 0x000004c0U;
            vlSelfRef.t__DOT__mem_narrow[20U] = 0x14U;
            vlSelfRef.t__DOT__mem_wide[20U][0U] = 0x0000feedU;
            vlSelfRef.t__DOT__mem_wide[20U][1U] = 0U;
            vlSelfRef.t__DOT__mem_wide[20U][2U] = 0x00000500U;
            vlSelfRef.t__DOT__mem_narrow[21U] = 0x15U;
            vlSelfRef.t__DOT__mem_wide[21U][0U] = 0x0000feedU;
            vlSelfRef.t__DOT__mem_wide[21U][1U] = 0U;
            vlSelfRef.t__DOT__mem_wide[21U][2U] = 0x00000540U;
            vlSelfRef.t__DOT__mem_narrow[22U] = 0x16U;
            vlSelfRef.t__DOT__mem_wide[22U][0U] = 0x0000feedU;
            vlSelfRef.t__DOT__mem_wide[22U][1U] = 0U;
            vlSelfRef.t__DOT__mem_wide[22U][2U] = 0x00000580U;
            vlSelfRef.t__DOT__mem_narrow[23U] = 0x17U;
            vlSelfRef.t__DOT__mem_wide[23U][0U] = 0x0000feedU;
            vlSelfRef.t__DOT__mem_wide[23U][1U] = 0U;
            vlSelfRef.t__DOT__mem_wide[23U][2U] = 0x000005c0U;
            vlSelfRef.t__DOT__mem_narrow[24U] = 0x18U;
            vlSelfRef.t__DOT__mem_wide[24U][0U] = 0x0000feedU;
            vlSelfRef.t__DOT__mem_wide[24U][1U] = 0U;
            vlSelfRef.t__DOT__mem_wide[24U][2U] = 0x00000600U;
            vlSelfRef.t__DOT__mem_narrow[25U] = 0x19U;
            vlSelfRef.t__DOT__mem_wide[25U][0U] = 0x0000feedU;
            vlSelfRef.t__DOT__mem_wide[25U][1U] = 0U;
            vlSelfRef.t__DOT__mem_wide[25U][2U] = 0x00000640U;
            vlSelfRef.t__DOT__mem_narrow[26U] = 0x1aU;
            vlSelfRef.t__DOT__mem_wide[26U][0U] = 0x0000feedU;
            vlSelfRef.t__DOT__mem_wide[26U][1U] = 0U;
            vlSelfRef.t__DOT__mem_wide[26U][2U] = 0x00000680U;
            vlSelfRef.t__DOT__mem_narrow[27U] = 0x1bU;
            vlSelfRef.t__DOT__mem_wide[27U][0U] = 0x0000feedU;
            vlSelfRef.t__DOT__mem_wide[27U][1U] = 0U;
            vlSelfRef.t__DOT__mem_wide[27U][2U] = 0x000006c0U;
            vlSelfRef.t__DOT__mem_narrow[28U] = 0x1cU;
            vlSelfRef.t__DOT__mem_wide[28U][0U] = 0x0000feedU;
            vlSelfRef.t__DOT__mem_wide[28U][1U] = 0U;
            vlSelfRef.t__DOT__mem_wide[28U][2U] = 0x00000700U;
            vlSelfRef.t__DOT__mem_narrow[29U] = 0x1dU;
            vlSelfRef.t__DOT__mem_wide[29U][0U] = 0x0000feedU;
            vlSelfRef.t__DOT__mem_wide[29U][1U] = 0U;
            vlSelfRef.t__DOT__mem_wide[29U][2U] = 0x00000740U;
            vlSelfRef.t__DOT__mem_narrow[30U] = 0x1eU;
            vlSelfRef.t__DOT__mem_wide[30U][0U] = 0x0000feedU;
            vlSelfRef.t__DOT__mem_wide[30U][1U] = 0U;
            vlSelfRef.t__DOT__mem_wide[30U][2U] = 0x00000780U;
            vlSelfRef.t__DOT__mem_narrow[31U] = 0x1fU;
            vlSelfRef.t__DOT__mem_wide[31U][0U] = 0x0000feedU;
            vlSelfRef.t__DOT__mem_wide[31U][1U] = 0U;
            vlSelfRef.t__DOT__mem_wide[31U][2U] = 0x000007c0U;
            if (VL_UNLIKELY(((0U != vlSelfRef.t__DOT__mem_narrow[0U])))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 50, "");
            }
            if (VL_UNLIKELY(((0U != (((0x0000feedU 
                                       ^ vlSelfRef.t__DOT__mem_wide[0U][0U]) 
                                      | vlSelfRef.t__DOT__mem_wide[0U][1U]) 
                                     | vlSelfRef.t__DOT__mem_wide[0U][2U]))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 51, "");
            }
            if (VL_UNLIKELY(((1U != vlSelfRef.t__DOT__mem_narrow[1U])))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 50, "");
            }
            if (VL_UNLIKELY(((0U != (((0x0000feedU 
                                       ^ vlSelfRef.t__DOT__mem_wide[1U][0U]) 
                                      | vlSelfRef.t__DOT__mem_wide[1U][1U]) 
                                     | (0x00000040U 
                                        ^ vlSelfRef.t__DOT__mem_wide[1U][2U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 51, "");
            }
            if (VL_UNLIKELY(((2U != vlSelfRef.t__DOT__mem_narrow[2U])))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 50, "");
            }
            if (VL_UNLIKELY(((0U != (((0x0000feedU 
                                       ^ vlSelfRef.t__DOT__mem_wide[2U][0U]) 
                                      | vlSelfRef.t__DOT__mem_wide[2U][1U]) 
                                     | (0x00000080U 
                                        ^ vlSelfRef.t__DOT__mem_wide[2U][2U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 51, "");
            }
            if (VL_UNLIKELY(((3U != vlSelfRef.t__DOT__mem_narrow[3U])))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 50, "");
            }
            if (VL_UNLIKELY(((0U != (((0x0000feedU 
                                       ^ vlSelfRef.t__DOT__mem_wide[3U][0U]) 
                                      | vlSelfRef.t__DOT__mem_wide[3U][1U]) 
                                     | (0x000000c0U 
                                        ^ vlSelfRef.t__DOT__mem_wide[3U][2U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 51, "");
            }
            if (VL_UNLIKELY(((4U != vlSelfRef.t__DOT__mem_narrow[4U])))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 50, "");
            }
            if (VL_UNLIKELY(((0U != (((0x0000feedU 
                                       ^ vlSelfRef.t__DOT__mem_wide[4U][0U]) 
                                      | vlSelfRef.t__DOT__mem_wide[4U][1U]) 
                                     | (0x00000100U 
                                        ^ vlSelfRef.t__DOT__mem_wide[4U][2U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 51, "");
            }
            if (VL_UNLIKELY(((5U != vlSelfRef.t__DOT__mem_narrow[5U])))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 50, "");
            }
            if (VL_UNLIKELY(((0U != (((0x0000feedU 
                                       ^ vlSelfRef.t__DOT__mem_wide[5U][0U]) 
                                      | vlSelfRef.t__DOT__mem_wide[5U][1U]) 
                                     | (0x00000140U 
                                        ^ vlSelfRef.t__DOT__mem_wide[5U][2U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 51, "");
            }
            if (VL_UNLIKELY(((6U != vlSelfRef.t__DOT__mem_narrow[6U])))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 50, "");
            }
            if (VL_UNLIKELY(((0U != (((0x0000feedU 
                                       ^ vlSelfRef.t__DOT__mem_wide[6U][0U]) 
                                      | vlSelfRef.t__DOT__mem_wide[6U][1U]) 
                                     | (0x00000180U 
                                        ^ vlSelfRef.t__DOT__mem_wide[6U][2U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 51, "");
            }
            if (VL_UNLIKELY(((7U != vlSelfRef.t__DOT__mem_narrow[7U])))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 50, "");
            }
            if (VL_UNLIKELY(((0U != (((0x0000feedU 
                                       ^ vlSelfRef.t__DOT__mem_wide[7U][0U]) 
                                      | vlSelfRef.t__DOT__mem_wide[7U][1U]) 
                                     | (0x000001c0U 
                                        ^ vlSelfRef.t__DOT__mem_wide[7U][2U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 51, "");
            }
            if (VL_UNLIKELY(((8U != vlSelfRef.t__DOT__mem_narrow[8U])))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 50, "");
            }
            if (VL_UNLIKELY(((0U != (((0x0000feedU 
                                       ^ vlSelfRef.t__DOT__mem_wide[8U][0U]) 
                                      | vlSelfRef.t__DOT__mem_wide[8U][1U]) 
                                     | (0x00000200U 
                                        ^ vlSelfRef.t__DOT__mem_wide[8U][2U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 51, "");
            }
            if (VL_UNLIKELY(((9U != vlSelfRef.t__DOT__mem_narrow[9U])))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 50, "");
            }
            if (VL_UNLIKELY(((0U != (((0x0000feedU 
                                       ^ vlSelfRef.t__DOT__mem_wide[9U][0U]) 
                                      | vlSelfRef.t__DOT__mem_wide[9U][1U]) 
                                     | (0x00000240U 
                                        ^ vlSelfRef.t__DOT__mem_wide[9U][2U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 51, "");
            }
            if (VL_UNLIKELY(((0x0aU != vlSelfRef.t__DOT__mem_narrow[10U])))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 50, "");
            }
            if (VL_UNLIKELY(((0U != (((0x0000feedU 
                                       ^ vlSelfRef.t__DOT__mem_wide[10U][0U]) 
                                      | vlSelfRef.t__DOT__mem_wide[10U][1U]) 
                                     | (0x00000280U 
                                        ^ vlSelfRef.t__DOT__mem_wide[10U][2U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 51, "");
            }
            if (VL_UNLIKELY(((0x0bU != vlSelfRef.t__DOT__mem_narrow[11U])))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 50, "");
            }
            if (VL_UNLIKELY(((0U != (((0x0000feedU 
                                       ^ vlSelfRef.t__DOT__mem_wide[11U][0U]) 
                                      | vlSelfRef.t__DOT__mem_wide[11U][1U]) 
                                     | (0x000002c0U 
                                        ^ vlSelfRef.t__DOT__mem_wide[11U][2U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 51, "");
            }
            if (VL_UNLIKELY(((0x0cU != vlSelfRef.t__DOT__mem_narrow[12U])))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 50, "");
            }
            if (VL_UNLIKELY(((0U != (((0x0000feedU 
                                       ^ vlSelfRef.t__DOT__mem_wide[12U][0U]) 
                                      | vlSelfRef.t__DOT__mem_wide[12U][1U]) 
                                     | (0x00000300U 
                                        ^ vlSelfRef.t__DOT__mem_wide[12U][2U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 51, "");
            }
            if (VL_UNLIKELY(((0x0dU != vlSelfRef.t__DOT__mem_narrow[13U])))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 50, "");
            }
            if (VL_UNLIKELY(((0U != (((0x0000feedU 
                                       ^ vlSelfRef.t__DOT__mem_wide[13U][0U]) 
                                      | vlSelfRef.t__DOT__mem_wide[13U][1U]) 
                                     | (0x00000340U 
                                        ^ vlSelfRef.t__DOT__mem_wide[13U][2U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 51, "");
            }
            if (VL_UNLIKELY(((0x0eU != vlSelfRef.t__DOT__mem_narrow[14U])))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 50, "");
            }
            if (VL_UNLIKELY(((0U != (((0x0000feedU 
                                       ^ vlSelfRef.t__DOT__mem_wide[14U][0U]) 
                                      | vlSelfRef.t__DOT__mem_wide[14U][1U]) 
                                     | (0x00000380U 
                                        ^ vlSelfRef.t__DOT__mem_wide[14U][2U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 51, "");
            }
            if (VL_UNLIKELY(((0x0fU != vlSelfRef.t__DOT__mem_narrow[15U])))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 50, "");
            }
            if (VL_UNLIKELY(((0U != (((0x0000feedU 
                                       ^ vlSelfRef.t__DOT__mem_wide[15U][0U]) 
                                      | vlSelfRef.t__DOT__mem_wide[15U][1U]) 
                                     | (0x000003c0U 
                                        ^ vlSelfRef.t__DOT__mem_wide[15U][2U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 51, "");
            }
            if (VL_UNLIKELY(((0x10U != vlSelfRef.t__DOT__mem_narrow[16U])))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 50, "");
            }
            if (VL_UNLIKELY(((0U != (((0x0000feedU 
                                       ^ vlSelfRef.t__DOT__mem_wide[16U][0U]) 
                                      | vlSelfRef.t__DOT__mem_wide[16U][1U]) 
                                     | (0x00000400U 
                                        ^ vlSelfRef.t__DOT__mem_wide[16U][2U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 51, "");
            }
            if (VL_UNLIKELY(((0x11U != vlSelfRef.t__DOT__mem_narrow[17U])))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 50, "");
            }
            if (VL_UNLIKELY(((0U != (((0x0000feedU 
                                       ^ vlSelfRef.t__DOT__mem_wide[17U][0U]) 
                                      | vlSelfRef.t__DOT__mem_wide[17U][1U]) 
                                     | (0x00000440U 
                                        ^ vlSelfRef.t__DOT__mem_wide[17U][2U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 51, "");
            }
            if (VL_UNLIKELY(((0x12U != vlSelfRef.t__DOT__mem_narrow[18U])))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 50, "");
            }
            if (VL_UNLIKELY(((0U != (((0x0000feedU 
                                       ^ vlSelfRef.t__DOT__mem_wide[18U][0U]) 
                                      | vlSelfRef.t__DOT__mem_wide[18U][1U]) 
                                     | (0x00000480U 
                                        ^ vlSelfRef.t__DOT__mem_wide[18U][2U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 51, "");
            }
            if (VL_UNLIKELY(((0x13U != vlSelfRef.t__DOT__mem_narrow[19U])))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 50, "");
            }
            if (VL_UNLIKELY(((0U != (((0x0000feedU 
                                       ^ vlSelfRef.t__DOT__mem_wide[19U][0U]) 
                                      | vlSelfRef.t__DOT__mem_wide[19U][1U]) 
                                     | (0x000004c0U 
                                        ^ vlSelfRef.t__DOT__mem_wide[19U][2U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 51, "");
            }
            if (VL_UNLIKELY(((0x14U != vlSelfRef.t__DOT__mem_narrow[20U])))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 50, "");
            }
            if (VL_UNLIKELY(((0U != (((0x0000feedU 
                                       ^ vlSelfRef.t__DOT__mem_wide[20U][0U]) 
                                      | vlSelfRef.t__DOT__mem_wide[20U][1U]) 
                                     | (0x00000500U 
                                        ^ vlSelfRef.t__DOT__mem_wide[20U][2U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 51, "");
            }
            if (VL_UNLIKELY(((0x15U != vlSelfRef.t__DOT__mem_narrow[21U])))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 50, "");
            }
            if (VL_UNLIKELY(((0U != (((0x0000feedU 
                                       ^ vlSelfRef.t__DOT__mem_wide[21U][0U]) 
                                      | vlSelfRef.t__DOT__mem_wide[21U][1U]) 
                                     | (0x00000540U 
                                        ^ vlSelfRef.t__DOT__mem_wide[21U][2U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 51, "");
            }
            if (VL_UNLIKELY(((0x16U != vlSelfRef.t__DOT__mem_narrow[22U])))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 50, "");
            }
            if (VL_UNLIKELY(((0U != (((0x0000feedU 
                                       ^ vlSelfRef.t__DOT__mem_wide[22U][0U]) 
                                      | vlSelfRef.t__DOT__mem_wide[22U][1U]) 
                                     | (0x00000580U 
                                        ^ vlSelfRef.t__DOT__mem_wide[22U][2U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 51, "");
            }
            if (VL_UNLIKELY(((0x17U != vlSelfRef.t__DOT__mem_narrow[23U])))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 50, "");
            }
            if (VL_UNLIKELY(((0U != (((0x0000feedU 
                                       ^ vlSelfRef.t__DOT__mem_wide[23U][0U]) 
                                      | vlSelfRef.t__DOT__mem_wide[23U][1U]) 
                                     | (0x000005c0U 
                                        ^ vlSelfRef.t__DOT__mem_wide[23U][2U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 51, "");
            }
            if (VL_UNLIKELY(((0x18U != vlSelfRef.t__DOT__mem_narrow[24U])))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 50, "");
            }
            if (VL_UNLIKELY(((0U != (((0x0000feedU 
                                       ^ vlSelfRef.t__DOT__mem_wide[24U][0U]) 
                                      | vlSelfRef.t__DOT__mem_wide[24U][1U]) 
                                     | (0x00000600U 
                                        ^ vlSelfRef.t__DOT__mem_wide[24U][2U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 51, "");
            }
            if (VL_UNLIKELY(((0x19U != vlSelfRef.t__DOT__mem_narrow[25U])))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 50, "");
            }
            if (VL_UNLIKELY(((0U != (((0x0000feedU 
                                       ^ vlSelfRef.t__DOT__mem_wide[25U][0U]) 
                                      | vlSelfRef.t__DOT__mem_wide[25U][1U]) 
                                     | (0x00000640U 
                                        ^ vlSelfRef.t__DOT__mem_wide[25U][2U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 51, "");
            }
            if (VL_UNLIKELY(((0x1aU != vlSelfRef.t__DOT__mem_narrow[26U])))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 50, "");
            }
            if (VL_UNLIKELY(((0U != (((0x0000feedU 
                                       ^ vlSelfRef.t__DOT__mem_wide[26U][0U]) 
                                      | vlSelfRef.t__DOT__mem_wide[26U][1U]) 
                                     | (0x00000680U 
                                        ^ vlSelfRef.t__DOT__mem_wide[26U][2U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 51, "");
            }
            if (VL_UNLIKELY(((0x1bU != vlSelfRef.t__DOT__mem_narrow[27U])))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 50, "");
            }
            if (VL_UNLIKELY(((0U != (((0x0000feedU 
                                       ^ vlSelfRef.t__DOT__mem_wide[27U][0U]) 
                                      | vlSelfRef.t__DOT__mem_wide[27U][1U]) 
                                     | (0x000006c0U 
                                        ^ vlSelfRef.t__DOT__mem_wide[27U][2U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 51, "");
            }
            if (VL_UNLIKELY(((0x1cU != vlSelfRef.t__DOT__mem_narrow[28U])))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 50, "");
            }
            if (VL_UNLIKELY(((0U != (((0x0000feedU 
                                       ^ vlSelfRef.t__DOT__mem_wide[28U][0U]) 
                                      | vlSelfRef.t__DOT__mem_wide[28U][1U]) 
                                     | (0x00000700U 
                                        ^ vlSelfRef.t__DOT__mem_wide[28U][2U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 51, "");
            }
            if (VL_UNLIKELY(((0x1dU != vlSelfRef.t__DOT__mem_narrow[29U])))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 50, "");
            }
            if (VL_UNLIKELY(((0U != (((0x0000feedU 
                                       ^ vlSelfRef.t__DOT__mem_wide[29U][0U]) 
                                      | vlSelfRef.t__DOT__mem_wide[29U][1U]) 
                                     | (0x00000740U 
                                        ^ vlSelfRef.t__DOT__mem_wide[29U][2U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 51, "");
            }
            if (VL_UNLIKELY(((0x1eU != vlSelfRef.t__DOT__mem_narrow[30U])))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 50, "");
            }
            if (VL_UNLIKELY(((0U != (((0x0000feedU 
                                       ^ vlSelfRef.t__DOT__mem_wide[30U][0U]) 
                                      | vlSelfRef.t__DOT__mem_wide[30U][1U]) 
                                     | (0x00000780U 
                                        ^ vlSelfRef.t__DOT__mem_wide[30U][2U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 51, "");
            }
            if (VL_UNLIKELY(((0x1fU != vlSelfRef.t__DOT__mem_narrow[31U])))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 50, "");
            }
            if (VL_UNLIKELY(((0U != (((0x0000feedU 
                                       ^ vlSelfRef.t__DOT__mem_wide[31U][0U]) 
                                      | vlSelfRef.t__DOT__mem_wide[31U][1U]) 
                                     | (0x000007c0U 
                                        ^ vlSelfRef.t__DOT__mem_wide[31U][2U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 51, "");
            }
            vlCtx.writef(0U, 1U, vlcu_msg_arg(vlSelfRef.t__DOT__wrd0[3U]));
        }
        if (VL_UNLIKELY(((1U == vlSelfRef.t__DOT___mode)))) {
            t__DOT__np2_mem[0U] = 0U;
            if (VL_UNLIKELY((((0U != vlSelfRef.t__DOT__np2_guard[0U]) 
                              | (0U != vlSelfRef.t__DOT__np2_guard[1U]))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 75, "");
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 75, "");
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 75, "");
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 75, "");
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 75, "");
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 75, "");
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 75, "");
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 75, "");
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 75, "");
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 75, "");
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 75, "");
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 75, "");
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 75, "");
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 75, "");
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 75, "");
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 75, "");
            }
            __Vdly__t__DOT___mode = 2U;
            __VdlySet__t__DOT__mem_dly_narrow__v0 = 1U;
            t__DOT__np2_mem[1U] = 1U;
            __VdlySet__t__DOT__mem_dly_narrow__v1 = 1U;
            t__DOT__np2_mem[2U] = 2U;
            t__DOT__np2_mem[3U] = 3U;
            t__DOT__np2_mem[4U] = 4U;
            t__DOT__np2_mem[5U] = 5U;
            t__DOT__np2_mem[0U] = 0U;
            t__DOT__np2_mem[1U] = 1U;
            t__DOT__np2_mem[2U] = 2U;
            t__DOT__np2_mem[3U] = 3U;
            t__DOT__np2_mem[4U] = 4U;
            t__DOT__np2_mem[5U] = 5U;
            vlCtx.writef(1U, 1U, vlcu_msg_arg(vlSelfRef.t__DOT__vec_wide));
            __Vdly__t__DOT__vec_wide = (0x0000000180000000ULL 
                                        | __Vdly__t__DOT__vec_wide);
            __Vdly__t__DOT__vec_wide = (0x0000000400000000ULL 
                                        | __Vdly__t__DOT__vec_wide);
        }
        if (VL_UNLIKELY(((2U == vlSelfRef.t__DOT___mode)))) {
            if (VL_UNLIKELY(((0xffU != vlSelfRef.t__DOT__mem_dly_narrow[0U])))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 93, "");
            }
            __Vdly__t__DOT___mode = 3U;
            if (VL_UNLIKELY(((0U != (((0x0000faceU 
                                       ^ vlSelfRef.t__DOT__mem_dly_wide[0U][0U]) 
                                      | vlSelfRef.t__DOT__mem_dly_wide[0U][1U]) 
                                     | (0x00003fc0U 
                                        ^ vlSelfRef.t__DOT__mem_dly_wide[0U][2U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 94, "");
            }
            if (VL_UNLIKELY(((0xfeU != vlSelfRef.t__DOT__mem_dly_narrow[1U])))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 93, "");
            }
            if (VL_UNLIKELY(((0U != (((0x0000faceU 
                                       ^ vlSelfRef.t__DOT__mem_dly_wide[1U][0U]) 
                                      | vlSelfRef.t__DOT__mem_dly_wide[1U][1U]) 
                                     | (0x00003f80U 
                                        ^ vlSelfRef.t__DOT__mem_dly_wide[1U][2U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 94, "");
            }
            if (VL_UNLIKELY(((0x0000000580000000ULL 
                              != vlSelfRef.t__DOT__vec_wide)))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_mem_first.v", 98, "");
            }
            vlCtx.writef(2U, 0U);
            VL_FINISH_MT("/root/reference/test_regress/t/t_mem_first.v", 101, "");
        }
        __Vdly__t__DOT___mode = ((IData)(1U) + vlSelfRef.t__DOT___mode);
        vlSelfRef.t__DOT__vec_wide = __Vdly__t__DOT__vec_wide;
        vlSelfRef.t__DOT___mode = __Vdly__t__DOT___mode;
        if (__VdlySet__t__DOT__mem_dly_narrow__v0) {
            vlSelfRef.t__DOT__mem_dly_narrow[0U] = 0xffU;
            vlSelfRef.t__DOT__mem_dly_wide[0U][0U] = 0x0000faceU;
            vlSelfRef.t__DOT__mem_dly_wide[0U][1U] = 0U;
            vlSelfRef.t__DOT__mem_dly_wide[0U][2U] = 0x00003fc0U;
        }
        if (__VdlySet__t__DOT__mem_dly_narrow__v1) {
            vlSelfRef.t__DOT__mem_dly_narrow[1U] = 0xfeU;
            vlSelfRef.t__DOT__mem_dly_wide[1U][0U] = 0x0000faceU;
            vlSelfRef.t__DOT__mem_dly_wide[1U][1U] = 0U;
            vlSelfRef.t__DOT__mem_dly_wide[1U][2U] = 0x00003f80U;
        }
    }

    VLCU_D static void Vt_mem_first___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt_mem_first___024root___eval_phase__act(Vt_mem_first___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt_mem_first___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vt_mem_first___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vt_mem_first___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vt_mem_first___024root___eval_phase__nba(Vt_mem_first___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vt_mem_first___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    Vt_mem_first___024root___nba_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
            Vt_mem_first___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vt_mem_first___024root___eval(Vt_mem_first___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt_mem_first___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("/root/reference/test_regress/t/t_mem_first.v", 7, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vt_mem_first___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("/root/reference/test_regress/t/t_mem_first.v", 7, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vt_mem_first___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vt_mem_first___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vt_mem_first___024root___ctor_var_reset(Vt_mem_first___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        vlSelf->t__DOT___mode = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 11256059270696350879ull);
        for (int __Vi0 = 0; __Vi0 < 32; ++__Vi0) {
            vlSelf->t__DOT__mem_narrow[__Vi0] = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 8970802931191437655ull);
        }
        for (int __Vi0 = 0; __Vi0 < 1025; ++__Vi0) {
            VL_SCOPED_RAND_RESET_W(78, vlSelf->t__DOT__mem_wide[__Vi0], __VscopeHash, 17545074035767859410ull);
        }
        for (int __Vi0 = 0; __Vi0 < 2; ++__Vi0) {
            vlSelf->t__DOT__mem_dly_narrow[__Vi0] = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 2794652068125219960ull);
        }
        for (int __Vi0 = 0; __Vi0 < 2; ++__Vi0) {
            VL_SCOPED_RAND_RESET_W(78, vlSelf->t__DOT__mem_dly_wide[__Vi0], __VscopeHash, 11864380938093956087ull);
        }
        vlSelf->t__DOT__vec_wide = VL_SCOPED_RAND_RESET_Q(35, __VscopeHash, 8554776647279785015ull);
        for (int __Vi0 = 0; __Vi0 < 16; ++__Vi0) {
            vlSelf->t__DOT__wrd0[__Vi0] = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 17564878676930088912ull);
        }
        for (int __Vi0 = 0; __Vi0 < 2; ++__Vi0) {
            vlSelf->t__DOT__np2_guard[__Vi0] = VL_SCOPED_RAND_RESET_I(3, __VscopeHash, 15495349570875049838ull);
        }
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
        Vt_mem_first___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt_mem_first___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt_mem_first___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt_mem_first___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vt_mem_first___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt_mem_first___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt_mem_first__cuda::info() {
    typedef Vt_mem_first__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"t__DOT___mode", "t._mode", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT___mode},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"t__DOT__vec_wide", "t.vec_wide", VLCU_MEMBER_STATE, 8, 0, 1, 35, 34, 0, T::OFF_t__DOT__vec_wide},
        {"t__DOT__mem_narrow", "t.mem_narrow", VLCU_MEMBER_STATE, 1, 0, 32, 8, 7, 0, T::OFF_t__DOT__mem_narrow},
        {"t__DOT__mem_wide", "t.mem_wide", VLCU_MEMBER_MEMORY, 4, 0, 3075, 78, 77, 0, T::OFF_t__DOT__mem_wide},
        {"t__DOT__mem_dly_narrow", "t.mem_dly_narrow", VLCU_MEMBER_STATE, 1, 0, 2, 8, 7, 0, T::OFF_t__DOT__mem_dly_narrow},
        {"t__DOT__mem_dly_wide", "t.mem_dly_wide", VLCU_MEMBER_STATE, 4, 0, 6, 78, 77, 0, T::OFF_t__DOT__mem_dly_wide},
        {"t__DOT__wrd0", "t.wrd0", VLCU_MEMBER_STATE, 4, 0, 16, 32, 31, 0, T::OFF_t__DOT__wrd0},
        {"t__DOT__np2_guard", "t.np2_guard", VLCU_MEMBER_STATE, 1, 0, 2, 3, 2, 0, T::OFF_t__DOT__np2_guard},
        {"__VactTriggered", "__VactTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VactTriggered},
        {"__VnbaTriggered", "__VnbaTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VnbaTriggered},
    };
    static const char* const messages[] = {"selout %b 0 1\n", "%x\n", "*-* All Finished *-*\n", nullptr};
    static const char* const plusargs[] = {nullptr};
    static const VlcuTraceDecl traceDecls[] = {
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0}};  // (terminator, not counted)
    static const uint32_t traceOrder[] = {0};  // codes in dump order (+ terminator)
    static const uint32_t saveOrder[] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "t", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 3, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 15, 0, 0x7425f85357fa3c58ULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt_mem_first__cuda VLCU_TESTBENCHES);
