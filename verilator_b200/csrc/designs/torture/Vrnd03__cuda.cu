// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vrnd03"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vrnd03___024root final {
    CData/*0:0*/ clk;
    QData/*39:0*/ i0;
    CData/*6:0*/ i1;
    VlWide<5>/*129:0*/ i2;
    VlWide<4>/*127:0*/ i3;
    VlWide<3>/*95:0*/ i4;
    QData/*62:0*/ i5;
    CData/*0:0*/ o0;
    SData/*15:0*/ o1;
    IData/*31:0*/ o2;
    QData/*39:0*/ o3;
    VlWide<3>/*64:0*/ o4;
    CData/*2:0*/ rnd03__DOT__r0;
    CData/*2:0*/ rnd03__DOT__r2;
    CData/*6:0*/ rnd03__DOT__r3;
    CData/*4:0*/ rnd03__DOT__w4;
    CData/*0:0*/ __VstlFirstIteration;
    CData/*0:0*/ __VstlPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*6:0*/ __Vtrigprevexpr___TOP__i1__0;
    CData/*0:0*/ __VicoDidInit;
    CData/*0:0*/ __VicoPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__1;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    SData/*8:0*/ rnd03__DOT__w6;
    VlWide<3>/*71:0*/ rnd03__DOT__r1;
    IData/*23:0*/ rnd03__DOT__w1;
    IData/*23:0*/ rnd03__DOT__w3;
    VlWide<3>/*64:0*/ rnd03__DOT__w5;
    IData/*23:0*/ rnd03__DOT__w7;
    IData/*23:0*/ rnd03__DOT__w10;
    IData/*23:0*/ rnd03__DOT__w11;
    VlWide<5>/*129:0*/ __Vtrigprevexpr___TOP__i2__0;
    VlWide<4>/*127:0*/ __Vtrigprevexpr___TOP__i3__0;
    VlWide<3>/*95:0*/ __Vtrigprevexpr___TOP__i4__0;
    IData/*31:0*/ __VactIterCount;
    QData/*32:0*/ rnd03__DOT__w8;
    QData/*62:0*/ __VdfgRegularize_h6e95ff9d_0_0;
    QData/*39:0*/ __Vtrigprevexpr___TOP__i0__0;
    QData/*62:0*/ __Vtrigprevexpr___TOP__i5__0;
    VlcuSharedArray<IData/*23:0*/, 16> rnd03__DOT__mem;  // run-time indexed: shared memory [index][lane]
    VlUnpacked<QData/*63:0*/, 1> __VstlTriggered;
    VlUnpacked<QData/*63:0*/, 2> __VicoTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vrnd03__cuda final {
    typedef Vrnd03___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 324,
        OFF_i0 = 0,
        OFF_i1 = 325,
        OFF_i2 = 96,
        OFF_i3 = 116,
        OFF_i4 = 132,
        OFF_i5 = 8,
        OFF_o0 = 326,
        OFF_o1 = 320,
        OFF_o2 = 144,
        OFF_o3 = 16,
        OFF_o4 = 148,
        OFF_rnd03__DOT__r0 = 327,
        OFF_rnd03__DOT__r2 = 328,
        OFF_rnd03__DOT__r3 = 329,
        OFF_rnd03__DOT__w4 = 330,
        OFF___VstlFirstIteration = 331,
        OFF___VstlPhaseResult = 332,
        OFF___Vtrigprevexpr___TOP__clk__0 = 333,
        OFF___Vtrigprevexpr___TOP__i1__0 = 334,
        OFF___VicoDidInit = 335,
        OFF___VicoPhaseResult = 336,
        OFF___Vtrigprevexpr___TOP__clk__1 = 337,
        OFF___VactPhaseResult = 338,
        OFF___VnbaPhaseResult = 339,
        OFF_rnd03__DOT__w6 = 322,
        OFF_rnd03__DOT__r1 = 160,
        OFF_rnd03__DOT__w1 = 172,
        OFF_rnd03__DOT__w3 = 176,
        OFF_rnd03__DOT__w5 = 180,
        OFF_rnd03__DOT__w7 = 192,
        OFF_rnd03__DOT__w10 = 196,
        OFF_rnd03__DOT__w11 = 200,
        OFF___Vtrigprevexpr___TOP__i2__0 = 204,
        OFF___Vtrigprevexpr___TOP__i3__0 = 224,
        OFF___Vtrigprevexpr___TOP__i4__0 = 240,
        OFF___VactIterCount = 252,
        OFF_rnd03__DOT__w8 = 24,
        OFF___VdfgRegularize_h6e95ff9d_0_0 = 32,
        OFF___Vtrigprevexpr___TOP__i0__0 = 40,
        OFF___Vtrigprevexpr___TOP__i5__0 = 48,
        OFF_rnd03__DOT__mem = 256,
        OFF___VstlTriggered = 56,
        OFF___VicoTriggered = 64,
        OFF___VactTriggered = 80,
        OFF___VnbaTriggered = 88,
        kStateBytes = 344,  // S_persist
        kTouchedReadBytes = 263,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 204,  // S_state_wr
    };
    struct Locals final {
    };
    struct Shared final {
        IData/*23:0*/ rnd03__DOT__mem[16][VLCU_BLOCK];
    };
    VLCU_D static void bind(Root& s, Locals& l, Shared& sh, const VlcuState& st, uint64_t i) {
        (void)l; (void)sh; (void)st; (void)i;
        s.vlNamep = "TOP";
        s.rnd03__DOT__mem.bind(sh.rnd03__DOT__mem, threadIdx.x);
    }
    VLCU_D static void load(Root& s, const VlcuState& st, uint64_t i) {
        s.clk = st.ld<VlcuMem8>(OFF_clk, i);
        s.i0 = st.ld<QData>(OFF_i0, i);
        s.i1 = st.ld<VlcuMem8>(OFF_i1, i);
        st.ldW(OFF_i2, i, s.i2);
        st.ldW(OFF_i3, i, s.i3);
        st.ldW(OFF_i4, i, s.i4);
        s.i5 = st.ld<QData>(OFF_i5, i);
        s.o0 = st.ld<VlcuMem8>(OFF_o0, i);
        s.o1 = st.ld<VlcuMem16>(OFF_o1, i);
        s.o2 = st.ld<IData>(OFF_o2, i);
        s.o3 = st.ld<QData>(OFF_o3, i);
        st.ldW(OFF_o4, i, s.o4);
        s.rnd03__DOT__r0 = st.ld<VlcuMem8>(OFF_rnd03__DOT__r0, i);
        s.rnd03__DOT__r2 = st.ld<VlcuMem8>(OFF_rnd03__DOT__r2, i);
        s.rnd03__DOT__r3 = st.ld<VlcuMem8>(OFF_rnd03__DOT__r3, i);
        s.rnd03__DOT__w4 = st.ld<VlcuMem8>(OFF_rnd03__DOT__w4, i);
        s.__VstlFirstIteration = st.ld<VlcuMem8>(OFF___VstlFirstIteration, i);
        s.__VstlPhaseResult = st.ld<VlcuMem8>(OFF___VstlPhaseResult, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__Vtrigprevexpr___TOP__i1__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__i1__0, i);
        s.__VicoDidInit = st.ld<VlcuMem8>(OFF___VicoDidInit, i);
        s.__VicoPhaseResult = st.ld<VlcuMem8>(OFF___VicoPhaseResult, i);
        s.__Vtrigprevexpr___TOP__clk__1 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__1, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        s.rnd03__DOT__w6 = st.ld<VlcuMem16>(OFF_rnd03__DOT__w6, i);
        st.ldW(OFF_rnd03__DOT__r1, i, s.rnd03__DOT__r1);
        s.rnd03__DOT__w1 = st.ld<IData>(OFF_rnd03__DOT__w1, i);
        s.rnd03__DOT__w3 = st.ld<IData>(OFF_rnd03__DOT__w3, i);
        st.ldW(OFF_rnd03__DOT__w5, i, s.rnd03__DOT__w5);
        s.rnd03__DOT__w7 = st.ld<IData>(OFF_rnd03__DOT__w7, i);
        s.rnd03__DOT__w10 = st.ld<IData>(OFF_rnd03__DOT__w10, i);
        s.rnd03__DOT__w11 = st.ld<IData>(OFF_rnd03__DOT__w11, i);
        st.ldW(OFF___Vtrigprevexpr___TOP__i2__0, i, s.__Vtrigprevexpr___TOP__i2__0);
        st.ldW(OFF___Vtrigprevexpr___TOP__i3__0, i, s.__Vtrigprevexpr___TOP__i3__0);
        st.ldW(OFF___Vtrigprevexpr___TOP__i4__0, i, s.__Vtrigprevexpr___TOP__i4__0);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        s.rnd03__DOT__w8 = st.ld<QData>(OFF_rnd03__DOT__w8, i);
        s.__VdfgRegularize_h6e95ff9d_0_0 = st.ld<QData>(OFF___VdfgRegularize_h6e95ff9d_0_0, i);
        s.__Vtrigprevexpr___TOP__i0__0 = st.ld<QData>(OFF___Vtrigprevexpr___TOP__i0__0, i);
        s.__Vtrigprevexpr___TOP__i5__0 = st.ld<QData>(OFF___Vtrigprevexpr___TOP__i5__0, i);
        for (uint32_t e = 0; e < 16; ++e) s.rnd03__DOT__mem[e] = st.ldE<IData>(OFF_rnd03__DOT__mem, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VstlTriggered[e] = st.ldE<QData>(OFF___VstlTriggered, e, i);
        for (uint32_t e = 0; e < 2; ++e) s.__VicoTriggered[e] = st.ldE<QData>(OFF___VicoTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.st<QData>(OFF_i0, i, s.i0);
        st.st<VlcuMem8>(OFF_i1, i, s.i1);
        st.stW(OFF_i2, i, s.i2);
        st.stW(OFF_i3, i, s.i3);
        st.stW(OFF_i4, i, s.i4);
        st.st<QData>(OFF_i5, i, s.i5);
        st.st<VlcuMem8>(OFF_o0, i, s.o0);
        st.st<VlcuMem16>(OFF_o1, i, s.o1);
        st.st<IData>(OFF_o2, i, s.o2);
        st.st<QData>(OFF_o3, i, s.o3);
        st.stW(OFF_o4, i, s.o4);
        st.st<VlcuMem8>(OFF_rnd03__DOT__r0, i, s.rnd03__DOT__r0);
        st.st<VlcuMem8>(OFF_rnd03__DOT__r2, i, s.rnd03__DOT__r2);
        st.st<VlcuMem8>(OFF_rnd03__DOT__r3, i, s.rnd03__DOT__r3);
        st.st<VlcuMem8>(OFF_rnd03__DOT__w4, i, s.rnd03__DOT__w4);
        st.st<VlcuMem8>(OFF___VstlFirstIteration, i, s.__VstlFirstIteration);
        st.st<VlcuMem8>(OFF___VstlPhaseResult, i, s.__VstlPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__i1__0, i, s.__Vtrigprevexpr___TOP__i1__0);
        st.st<VlcuMem8>(OFF___VicoDidInit, i, s.__VicoDidInit);
        st.st<VlcuMem8>(OFF___VicoPhaseResult, i, s.__VicoPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__1, i, s.__Vtrigprevexpr___TOP__clk__1);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.st<VlcuMem16>(OFF_rnd03__DOT__w6, i, s.rnd03__DOT__w6);
        st.stW(OFF_rnd03__DOT__r1, i, s.rnd03__DOT__r1);
        st.st<IData>(OFF_rnd03__DOT__w1, i, s.rnd03__DOT__w1);
        st.st<IData>(OFF_rnd03__DOT__w3, i, s.rnd03__DOT__w3);
        st.stW(OFF_rnd03__DOT__w5, i, s.rnd03__DOT__w5);
        st.st<IData>(OFF_rnd03__DOT__w7, i, s.rnd03__DOT__w7);
        st.st<IData>(OFF_rnd03__DOT__w10, i, s.rnd03__DOT__w10);
        st.st<IData>(OFF_rnd03__DOT__w11, i, s.rnd03__DOT__w11);
        st.stW(OFF___Vtrigprevexpr___TOP__i2__0, i, s.__Vtrigprevexpr___TOP__i2__0);
        st.stW(OFF___Vtrigprevexpr___TOP__i3__0, i, s.__Vtrigprevexpr___TOP__i3__0);
        st.stW(OFF___Vtrigprevexpr___TOP__i4__0, i, s.__Vtrigprevexpr___TOP__i4__0);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
        st.st<QData>(OFF_rnd03__DOT__w8, i, s.rnd03__DOT__w8);
        st.st<QData>(OFF___VdfgRegularize_h6e95ff9d_0_0, i, s.__VdfgRegularize_h6e95ff9d_0_0);
        st.st<QData>(OFF___Vtrigprevexpr___TOP__i0__0, i, s.__Vtrigprevexpr___TOP__i0__0);
        st.st<QData>(OFF___Vtrigprevexpr___TOP__i5__0, i, s.__Vtrigprevexpr___TOP__i5__0);
        for (uint32_t e = 0; e < 16; ++e) st.stE<IData>(OFF_rnd03__DOT__mem, e, i, s.rnd03__DOT__mem[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VstlTriggered, e, i, s.__VstlTriggered[e]);
        for (uint32_t e = 0; e < 2; ++e) st.stE<QData>(OFF___VicoTriggered, e, i, s.__VicoTriggered[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VactTriggered, e, i, s.__VactTriggered[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VnbaTriggered, e, i, s.__VnbaTriggered[e]);
    }
    VLCU_D static void load_inputs_row(Root& s, const uint32_t* rowp, uint64_t stride, uint64_t i) {
        s.clk = static_cast<VlcuMem8>(rowp[0 * stride + i] & 0x1U);
        s.i0 = (static_cast<QData>(rowp[2 * stride + i] & 0xffU) << 32) | rowp[1 * stride + i];
        s.i1 = static_cast<VlcuMem8>(rowp[3 * stride + i] & 0x7fU);
        s.i2[0] = rowp[4 * stride + i] & 0xffffffffU;
        s.i2[1] = rowp[5 * stride + i] & 0xffffffffU;
        s.i2[2] = rowp[6 * stride + i] & 0xffffffffU;
        s.i2[3] = rowp[7 * stride + i] & 0xffffffffU;
        s.i2[4] = rowp[8 * stride + i] & 0x3U;
        s.i3[0] = rowp[9 * stride + i] & 0xffffffffU;
        s.i3[1] = rowp[10 * stride + i] & 0xffffffffU;
        s.i3[2] = rowp[11 * stride + i] & 0xffffffffU;
        s.i3[3] = rowp[12 * stride + i] & 0xffffffffU;
        s.i4[0] = rowp[13 * stride + i] & 0xffffffffU;
        s.i4[1] = rowp[14 * stride + i] & 0xffffffffU;
        s.i4[2] = rowp[15 * stride + i] & 0xffffffffU;
        s.i5 = (static_cast<QData>(rowp[17 * stride + i] & 0x7fffffffU) << 32) | rowp[16 * stride + i];
    }
    VLCU_D static void store_outputs_row(const Root& s, uint32_t* rowp, uint64_t stride, uint64_t i) {
        rowp[0 * stride + i] = s.o0;
        rowp[1 * stride + i] = s.o1;
        rowp[2 * stride + i] = s.o2;
        rowp[3 * stride + i] = static_cast<uint32_t>(s.o3);
        rowp[4 * stride + i] = static_cast<uint32_t>(s.o3 >> 32);
        rowp[5 * stride + i] = s.o4[0];
        rowp[6 * stride + i] = s.o4[1];
        rowp[7 * stride + i] = s.o4[2];
    }
    VLCU_D static uint64_t outputs_sig(const Root& s, uint64_t sig) {
        uint32_t acc = VLCU_SIG_ACC_INIT;
        acc = vlcu_sig_word(acc, s.o0);
        acc = vlcu_sig_word(acc, s.o1);
        acc = vlcu_sig_word(acc, s.o2);
        acc = vlcu_sig_word(acc, static_cast<uint32_t>(s.o3));
        acc = vlcu_sig_word(acc, static_cast<uint32_t>(s.o3 >> 32));
        acc = vlcu_sig_word(acc, s.o4[0]);
        acc = vlcu_sig_word(acc, s.o4[1]);
        acc = vlcu_sig_word(acc, s.o4[2]);
        return vlcu_sig_mix(sig, acc);
    }
    enum : uint32_t { kInWords = 18, kOutWords = 8, kInBytes = 66, kOutBytes = 27, kTraceWords = 0 };

    VLCU_D static void Vrnd03___024root___eval_static(Vrnd03___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
        vlSelfRef.__Vtrigprevexpr___TOP__i0__0 = vlSelfRef.i0;
        vlSelfRef.__Vtrigprevexpr___TOP__i1__0 = vlSelfRef.i1;
        vlSelfRef.__Vtrigprevexpr___TOP__i2__0[0U] 
            = vlSelfRef.i2[0U];
        vlSelfRef.__Vtrigprevexpr___TOP__i2__0[1U] 
            = vlSelfRef.i2[1U];
        vlSelfRef.__Vtrigprevexpr___TOP__i2__0[2U] 
            = vlSelfRef.i2[2U];
        vlSelfRef.__Vtrigprevexpr___TOP__i2__0[3U] 
            = vlSelfRef.i2[3U];
        vlSelfRef.__Vtrigprevexpr___TOP__i2__0[4U] 
            = vlSelfRef.i2[4U];
        vlSelfRef.__Vtrigprevexpr___TOP__i3__0[0U] 
            = vlSelfRef.i3[0U];
        vlSelfRef.__Vtrigprevexpr___TOP__i3__0[1U] 
            = vlSelfRef.i3[1U];
        vlSelfRef.__Vtrigprevexpr___TOP__i3__0[2U] 
            = vlSelfRef.i3[2U];
        vlSelfRef.__Vtrigprevexpr___TOP__i3__0[3U] 
            = vlSelfRef.i3[3U];
        vlSelfRef.__Vtrigprevexpr___TOP__i4__0[0U] 
            = vlSelfRef.i4[0U];
        vlSelfRef.__Vtrigprevexpr___TOP__i4__0[1U] 
            = vlSelfRef.i4[1U];
        vlSelfRef.__Vtrigprevexpr___TOP__i4__0[2U] 
            = vlSelfRef.i4[2U];
        vlSelfRef.__Vtrigprevexpr___TOP__i5__0 = vlSelfRef.i5;
        vlSelfRef.__Vtrigprevexpr___TOP__clk__1 = vlSelfRef.clk;
    }

    VLCU_D static void Vrnd03___024root___eval_initial(Vrnd03___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vrnd03___024root___eval_final(Vrnd03___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vrnd03___024root___eval_settle(Vrnd03___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        IData/*31:0*/ __VstlIterCount;
        // Body
        __VstlIterCount = 0U;
        vlSelfRef.__VstlFirstIteration = 1U;
        do {
            if (VL_UNLIKELY(((0x00002710U < __VstlIterCount)))) {
#ifdef VL_DEBUG
                Vrnd03___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/rnd03.v", 2, "", "DIDNOTCONVERGE: Settle region did not converge after '--converge-limit' of 10000 tries");
            }
            __VstlIterCount = ((IData)(1U) + __VstlIterCount);
            vlSelfRef.__VstlPhaseResult = Vrnd03___024root___eval_phase__stl(vlSelf, vlCtx);
            vlSelfRef.__VstlFirstIteration = 0U;
        } while (vlSelfRef.__VstlPhaseResult);
    }

    VLCU_D static bool Vrnd03___024root___trigger_anySet__stl(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vrnd03___024root___stl_sequent__TOP__0(Vrnd03___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        VlWide<5>/*159:0*/ __Vtemp_3;
        VlWide<5>/*159:0*/ __Vtemp_4;
        VlWide<5>/*159:0*/ __Vtemp_5;
        VlWide<5>/*159:0*/ __Vtemp_6;
        VlWide<4>/*127:0*/ __Vtemp_10;
        VlWide<4>/*127:0*/ __Vtemp_11;
        VlWide<4>/*127:0*/ __Vtemp_12;
        VlWide<3>/*95:0*/ __Vtemp_14;
        VlWide<3>/*95:0*/ __Vtemp_15;
        VlWide<3>/*95:0*/ __Vtemp_16;
        VlWide<3>/*95:0*/ __Vtemp_17;
        VlWide<3>/*95:0*/ __Vtemp_18;
        VlWide<3>/*95:0*/ __Vtemp_19;
        VlWide<3>/*95:0*/ __Vtemp_21;
        VlWide<3>/*95:0*/ __Vtemp_22;
        VlWide<3>/*95:0*/ __Vtemp_23;
        VlWide<5>/*159:0*/ __Vtemp_25;
        VlWide<5>/*159:0*/ __Vtemp_26;
        // Body
        vlSelfRef.rnd03__DOT__w1 = (0x00ffffffU & VL_DIV_III(24, (IData)(vlSelfRef.i1), (IData)(vlSelfRef.rnd03__DOT__r3)));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0 = 
            (0x7fffffffffffffffULL & VL_DIV_QQQ(63, vlSelfRef.i0, vlSelfRef.i5));
        vlSelfRef.o1 = VL_SHIFTR_IIQ(16,16,63, 0xffffU, 
                                     (0x7fffffffffffffffULL 
                                      & ((QData)((IData)(
                                                         (1U 
                                                          & (~ 
                                                             VL_REDXOR_32(vlSelfRef.rnd03__DOT__mem
                                                                          [
                                                                          (0x0000000fU 
                                                                           & vlSelfRef.rnd03__DOT__w1)]))))) 
                                         - vlSelfRef.i5)));
        vlSelfRef.rnd03__DOT__w6 = (0x000001ffU & (~ 
                                                   ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0) 
                                                    ^ (IData)(vlSelfRef.rnd03__DOT__r0))));
        vlSelfRef.rnd03__DOT__w3 = (0x00ffffffU & (
                                                   VL_DIV_III(24, (IData)(vlSelfRef.rnd03__DOT__r2), 
                                                              (0x00001fffU 
                                                               & (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0))) 
                                                   ^ 
                                                   ((vlSelfRef.i0 
                                                     >> (IData)(vlSelfRef.rnd03__DOT__r2)) 
                                                    != (QData)((IData)(
                                                                       (1U 
                                                                        & (IData)(vlSelfRef.rnd03__DOT__r0)))))));
        vlSelfRef.rnd03__DOT__w10 = (0x00ffffffU & 
                                     (VL_LTS_III(1, 
                                                 VL_GTES_III(7, 0U, 
                                                             (0x0000007fU 
                                                              & VL_EXTENDS_II(7,1, 
                                                                              (1U 
                                                                               & (vlSelfRef.i3
                                                                                [
                                                                                ((1U 
                                                                                < (IData)(vlSelfRef.rnd03__DOT__r0)) 
                                                                                >> 5U)] 
                                                                                >> 
                                                                                (0x0000001fU 
                                                                                & (1U 
                                                                                < (IData)(vlSelfRef.rnd03__DOT__r0)))))))), 
                                                 (1U 
                                                  & (IData)(
                                                            (vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0 
                                                             >> (IData)(vlSelfRef.rnd03__DOT__r2))))) 
                                      ^ ((~ (IData)(vlSelfRef.i0)) 
                                         ^ VL_SHIFTL_III(24,24,8, vlSelfRef.i2[0U], 
                                                         (0x000000ffU 
                                                          & (IData)(vlSelfRef.rnd03__DOT__w6))))));
        __Vtemp_3[0U] = (~ (IData)(vlSelfRef.rnd03__DOT__w6));
        __Vtemp_3[1U] = 0xffffffffU;
        __Vtemp_3[2U] = 0xffffffffU;
        __Vtemp_3[3U] = 0xffffffffU;
        __Vtemp_3[4U] = 0xffffffffU;
        VL_SHIFTL_WWI(150,150,8, __Vtemp_4, __Vtemp_3, 
                      (0x000000ffU & (IData)(vlSelfRef.rnd03__DOT__w6)));
        __Vtemp_5[0U] = __Vtemp_4[0U];
        __Vtemp_5[1U] = __Vtemp_4[1U];
        __Vtemp_5[2U] = __Vtemp_4[2U];
        __Vtemp_5[3U] = __Vtemp_4[3U];
        __Vtemp_5[4U] = (0x003fffffU & __Vtemp_4[4U]);
        VL_SHIFTR_WWI(150,150,8, __Vtemp_6, __Vtemp_5, 
                      (0x000000ffU & (IData)(vlSelfRef.i0)));
        vlSelfRef.rnd03__DOT__w11 = (0x00ffffffU & __Vtemp_6[0U]);
        vlSelfRef.rnd03__DOT__w8 = (0x00000001ffffffffULL 
                                    & VL_SHIFTL_QQI(33,33,8, (QData)((IData)(vlSelfRef.rnd03__DOT__w3)), 
                                                    (0x000000ffU 
                                                     & (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0))));
        __Vtemp_10[0U] = (~ ((IData)((((QData)((IData)(vlSelfRef.i4[1U])) 
                                       << 0x00000020U) 
                                      | (QData)((IData)(vlSelfRef.i4[0U])))) 
                             ^ (IData)((0x0000001fffffffffULL 
                                        & (((QData)((IData)(vlSelfRef.i4[1U])) 
                                            << 0x0000001eU) 
                                           | ((QData)((IData)(vlSelfRef.i4[0U])) 
                                              >> 2U))))));
        __Vtemp_10[1U] = (~ ((IData)(((((QData)((IData)(vlSelfRef.i4[1U])) 
                                        << 0x00000020U) 
                                       | (QData)((IData)(vlSelfRef.i4[0U]))) 
                                      >> 0x00000020U)) 
                             ^ (IData)(((0x0000001fffffffffULL 
                                         & (((QData)((IData)(vlSelfRef.i4[1U])) 
                                             << 0x0000001eU) 
                                            | ((QData)((IData)(vlSelfRef.i4[0U])) 
                                               >> 2U))) 
                                        >> 0x00000020U))));
        __Vtemp_10[2U] = 0xffffffffU;
        __Vtemp_10[3U] = 0xffffffffU;
        __Vtemp_11[0U] = vlSelfRef.rnd03__DOT__w1;
        __Vtemp_11[1U] = 0U;
        __Vtemp_11[2U] = 0U;
        __Vtemp_11[3U] = 0U;
        VL_MUL_W(4, __Vtemp_12, __Vtemp_10, __Vtemp_11);
        vlSelfRef.rnd03__DOT__w4 = (0x0000001fU & (~ 
                                                   (__Vtemp_12[0U] 
                                                    ^ 
                                                    (vlSelfRef.rnd03__DOT__w3 
                                                     & vlSelfRef.i3[0U]))));
        __Vtemp_14[0U] = (IData)(vlSelfRef.i5);
        __Vtemp_14[1U] = ((vlSelfRef.rnd03__DOT__w11 
                           << 0x0000001fU) | (IData)(
                                                     (vlSelfRef.i5 
                                                      >> 0x00000020U)));
        __Vtemp_14[2U] = (vlSelfRef.rnd03__DOT__w11 
                          >> 1U);
        VL_SHIFTR_WWI(87,87,8, __Vtemp_15, __Vtemp_14, 
                      (0x000000ffU & (IData)(vlSelfRef.rnd03__DOT__w6)));
        __Vtemp_16[0U] = __Vtemp_15[0U];
        __Vtemp_16[1U] = __Vtemp_15[1U];
        __Vtemp_16[2U] = __Vtemp_15[2U];
        if ((0U != (VL_COUNTONES_I((0x000000ffU & (vlSelfRef.rnd03__DOT__mem
                                                   [
                                                   (0x0000000fU 
                                                    & vlSelfRef.rnd03__DOT__w1)] 
                                                   >> 
                                                   (0x0000000fU 
                                                    & VL_SHIFTL_III(4,4,8, vlSelfRef.rnd03__DOT__w3, 
                                                                    (0x000000ffU 
                                                                     & (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0))))))) 
                    + VL_ONEHOT_I(vlSelfRef.rnd03__DOT__w3)))) {
            __Vtemp_17[0U] = 1U;
            __Vtemp_17[1U] = 0U;
            __Vtemp_17[2U] = 0U;
        } else {
            __Vtemp_18[0U] = (3U & (IData)((vlSelfRef.i0 
                                            >> 4U)));
            __Vtemp_18[1U] = 0U;
            __Vtemp_18[2U] = 0U;
            VL_SHIFTL_WWI(65,65,3, __Vtemp_17, __Vtemp_18, (IData)(vlSelfRef.rnd03__DOT__r0));
        }
        VL_SUB_W(3, __Vtemp_19, __Vtemp_16, __Vtemp_17);
        vlSelfRef.o4[0U] = __Vtemp_19[0U];
        vlSelfRef.o4[1U] = __Vtemp_19[1U];
        vlSelfRef.o4[2U] = (1U & __Vtemp_19[2U]);
        vlSelfRef.o0 = (1U & (IData)((0x00000001ffffffffULL 
                                      & VL_DIVS_QQQ(33, 
                                                    (0x00000001ffffffffULL 
                                                     & VL_EXTENDS_QI(33,5, (IData)(vlSelfRef.rnd03__DOT__w4))), vlSelfRef.rnd03__DOT__w8))));
        vlSelfRef.rnd03__DOT__w7 = (0x00ffffffU & (
                                                   (1U 
                                                    & ((~ 
                                                        (0U 
                                                         != (IData)(vlSelfRef.rnd03__DOT__w4))) 
                                                       | (0U 
                                                          != 
                                                          (0x000001ffU 
                                                           & VL_MODDIV_III(9, (IData)(vlSelfRef.rnd03__DOT__w4), (IData)(vlSelfRef.rnd03__DOT__w6))))))
                                                    ? (IData)(vlSelfRef.i5)
                                                    : (IData)(
                                                              (0x7fffffffffffffffULL 
                                                               & VL_DIV_QQQ(63, (QData)((IData)(vlSelfRef.rnd03__DOT__r2)), (QData)((IData)(vlSelfRef.rnd03__DOT__r3)))))));
        vlSelfRef.rnd03__DOT__w5[0U] = ((0xfffffffeU 
                                         & vlSelfRef.rnd03__DOT__w5[0U]) 
                                        | (((0x7fffffffffffffffULL 
                                             & VL_MODDIV_QQQ(63, (QData)((IData)(vlSelfRef.rnd03__DOT__r0)), vlSelfRef.i5)) 
                                            >= (QData)((IData)(
                                                               (0x000003ffU 
                                                                & (((IData)(vlSelfRef.rnd03__DOT__w4) 
                                                                    << 5U) 
                                                                   | (IData)(vlSelfRef.rnd03__DOT__w4)))))) 
                                           | (0U != 
                                              ((IData)(1U) 
                                               + (vlSelfRef.i5 
                                                  > (QData)((IData)(
                                                                    (1U 
                                                                     & (IData)(vlSelfRef.i5)))))))));
        vlSelfRef.rnd03__DOT__w5[0U] = (1U & vlSelfRef.rnd03__DOT__w5[0U]);
        vlSelfRef.rnd03__DOT__w5[1U] = 0U;
        vlSelfRef.rnd03__DOT__w5[2U] = 0U;
        vlSelfRef.o3 = (QData)((IData)(VL_LTES_III(24, vlSelfRef.rnd03__DOT__w11, 
                                                   (0x00ffffffU 
                                                    & (~ 
                                                       ((0x00000044U 
                                                         | vlSelfRef.rnd03__DOT__w3) 
                                                        ^ 
                                                        (0x0000000fU 
                                                         & (1U 
                                                            >> 
                                                            (0x0000000fU 
                                                             & vlSelfRef.rnd03__DOT__w5[0U])))))))));
        __Vtemp_21[0U] = (1U & (IData)((vlSelfRef.rnd03__DOT__w8 
                                        >> (0x0000001fU 
                                            & vlSelfRef.rnd03__DOT__w5[0U]))));
        __Vtemp_21[1U] = 0U;
        __Vtemp_21[2U] = 0U;
        VL_MUL_W(3, __Vtemp_22, vlSelfRef.rnd03__DOT__r1, __Vtemp_21);
        __Vtemp_23[0U] = __Vtemp_22[0U];
        __Vtemp_23[1U] = __Vtemp_22[1U];
        __Vtemp_23[2U] = (0x000000ffU & __Vtemp_22[2U]);
        VL_EXTENDS_WI(130,24, __Vtemp_25, vlSelfRef.rnd03__DOT__mem
                      [(0x0000000fU & vlSelfRef.rnd03__DOT__w1)]);
        __Vtemp_26[0U] = __Vtemp_25[0U];
        __Vtemp_26[1U] = __Vtemp_25[1U];
        __Vtemp_26[2U] = __Vtemp_25[2U];
        __Vtemp_26[3U] = __Vtemp_25[3U];
        __Vtemp_26[4U] = (3U & __Vtemp_25[4U]);
        vlSelfRef.o2 = ((VL_COUNTONES_I((IData)(vlSelfRef.rnd03__DOT__w4)) 
                         + VL_ONEHOT_W(3, __Vtemp_23)) 
                        <= (VL_LTES_IWW(130, vlSelfRef.i2, __Vtemp_26) 
                            == (0x00ffffffU & ((1U 
                                                & ((IData)(vlSelfRef.rnd03__DOT__r2) 
                                                   >> 
                                                   (1U 
                                                    & (IData)(vlSelfRef.rnd03__DOT__r2)))) 
                                               + vlSelfRef.rnd03__DOT__w7))));
    }

    VLCU_D static bool Vrnd03___024root___eval_phase__stl(Vrnd03___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VstlExecute;
        // Body
        {
            // Inlined CFunc: _eval_triggers_vec__stl
            vlSelfRef.__VstlTriggered[0U] = ((0xfffffffffffffffeULL 
                                              & vlSelfRef.__VstlTriggered[0U]) 
                                             | (IData)((IData)(vlSelfRef.__VstlFirstIteration)));
        }
#ifdef VL_DEBUG
        if (VL_UNLIKELY(vlSymsp->_vm_contextp__->debug())) {
            Vrnd03___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
        }
#endif
        __VstlExecute = Vrnd03___024root___trigger_anySet__stl(vlSelfRef.__VstlTriggered);
        if (__VstlExecute) {
            {
                // Inlined CFunc: _eval_stl
                if ((1ULL & vlSelfRef.__VstlTriggered[0U])) {
                    Vrnd03___024root___stl_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
        }
        return (__VstlExecute);
    }

    VLCU_D static void Vrnd03___024root___eval_triggers_vec__ico(Vrnd03___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlSelfRef.__VicoTriggered[0U] = (QData)((IData)(
                                                        ((((vlSelfRef.i5 
                                                            != vlSelfRef.__Vtrigprevexpr___TOP__i5__0) 
                                                           << 6U) 
                                                          | (((0U 
                                                               != 
                                                               (((vlSelfRef.i4[0U] 
                                                                  ^ vlSelfRef.__Vtrigprevexpr___TOP__i4__0[0U]) 
                                                                 | (vlSelfRef.i4[1U] 
                                                                    ^ vlSelfRef.__Vtrigprevexpr___TOP__i4__0[1U])) 
                                                                | (vlSelfRef.i4[2U] 
                                                                   ^ vlSelfRef.__Vtrigprevexpr___TOP__i4__0[2U]))) 
                                                              << 5U) 
                                                             | ((0U 
                                                                 != 
                                                                 ((((vlSelfRef.i3[0U] 
                                                                     ^ vlSelfRef.__Vtrigprevexpr___TOP__i3__0[0U]) 
                                                                    | (vlSelfRef.i3[1U] 
                                                                       ^ vlSelfRef.__Vtrigprevexpr___TOP__i3__0[1U])) 
                                                                   | (vlSelfRef.i3[2U] 
                                                                      ^ vlSelfRef.__Vtrigprevexpr___TOP__i3__0[2U])) 
                                                                  | (vlSelfRef.i3[3U] 
                                                                     ^ vlSelfRef.__Vtrigprevexpr___TOP__i3__0[3U]))) 
                                                                << 4U))) 
                                                         | ((((0U 
                                                               != 
                                                               (((((vlSelfRef.i2[0U] 
                                                                    ^ vlSelfRef.__Vtrigprevexpr___TOP__i2__0[0U]) 
                                                                   | (vlSelfRef.i2[1U] 
                                                                      ^ vlSelfRef.__Vtrigprevexpr___TOP__i2__0[1U])) 
                                                                  | (vlSelfRef.i2[2U] 
                                                                     ^ vlSelfRef.__Vtrigprevexpr___TOP__i2__0[2U])) 
                                                                 | (vlSelfRef.i2[3U] 
                                                                    ^ vlSelfRef.__Vtrigprevexpr___TOP__i2__0[3U])) 
                                                                | (vlSelfRef.i2[4U] 
                                                                   ^ vlSelfRef.__Vtrigprevexpr___TOP__i2__0[4U]))) 
                                                              << 3U) 
                                                             | (((IData)(vlSelfRef.i1) 
                                                                 != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__i1__0)) 
                                                                << 2U)) 
                                                            | (((vlSelfRef.i0 
                                                                 != vlSelfRef.__Vtrigprevexpr___TOP__i0__0) 
                                                                << 1U) 
                                                               | ((IData)(vlSelfRef.clk) 
                                                                  != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__clk__0)))))));
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
        vlSelfRef.__Vtrigprevexpr___TOP__i0__0 = vlSelfRef.i0;
        vlSelfRef.__Vtrigprevexpr___TOP__i1__0 = vlSelfRef.i1;
        vlSelfRef.__Vtrigprevexpr___TOP__i2__0[0U] 
            = vlSelfRef.i2[0U];
        vlSelfRef.__Vtrigprevexpr___TOP__i2__0[1U] 
            = vlSelfRef.i2[1U];
        vlSelfRef.__Vtrigprevexpr___TOP__i2__0[2U] 
            = vlSelfRef.i2[2U];
        vlSelfRef.__Vtrigprevexpr___TOP__i2__0[3U] 
            = vlSelfRef.i2[3U];
        vlSelfRef.__Vtrigprevexpr___TOP__i2__0[4U] 
            = vlSelfRef.i2[4U];
        vlSelfRef.__Vtrigprevexpr___TOP__i3__0[0U] 
            = vlSelfRef.i3[0U];
        vlSelfRef.__Vtrigprevexpr___TOP__i3__0[1U] 
            = vlSelfRef.i3[1U];
        vlSelfRef.__Vtrigprevexpr___TOP__i3__0[2U] 
            = vlSelfRef.i3[2U];
        vlSelfRef.__Vtrigprevexpr___TOP__i3__0[3U] 
            = vlSelfRef.i3[3U];
        vlSelfRef.__Vtrigprevexpr___TOP__i4__0[0U] 
            = vlSelfRef.i4[0U];
        vlSelfRef.__Vtrigprevexpr___TOP__i4__0[1U] 
            = vlSelfRef.i4[1U];
        vlSelfRef.__Vtrigprevexpr___TOP__i4__0[2U] 
            = vlSelfRef.i4[2U];
        vlSelfRef.__Vtrigprevexpr___TOP__i5__0 = vlSelfRef.i5;
        if (VL_UNLIKELY(((1U & (~ (IData)(vlSelfRef.__VicoDidInit)))))) {
            vlSelfRef.__VicoDidInit = 1U;
            vlSelfRef.__VicoTriggered[0U] = (1ULL | vlSelfRef.__VicoTriggered[0U]);
            vlSelfRef.__VicoTriggered[0U] = (2ULL | vlSelfRef.__VicoTriggered[0U]);
            vlSelfRef.__VicoTriggered[0U] = (4ULL | vlSelfRef.__VicoTriggered[0U]);
            vlSelfRef.__VicoTriggered[0U] = (8ULL | vlSelfRef.__VicoTriggered[0U]);
            vlSelfRef.__VicoTriggered[0U] = (0x0000000000000010ULL 
                                             | vlSelfRef.__VicoTriggered[0U]);
            vlSelfRef.__VicoTriggered[0U] = (0x0000000000000020ULL 
                                             | vlSelfRef.__VicoTriggered[0U]);
            vlSelfRef.__VicoTriggered[0U] = (0x0000000000000040ULL 
                                             | vlSelfRef.__VicoTriggered[0U]);
        }
    }

    VLCU_D static bool Vrnd03___024root___trigger_anySet__ico(const VlUnpacked<QData/*63:0*/, 2> &in) {
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
        } while ((2U > n));
        return (0U);
    }

    VLCU_D static void Vrnd03___024root___ico_comb__TOP__3(Vrnd03___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        VlWide<4>/*127:0*/ __Vtemp_4;
        VlWide<4>/*127:0*/ __Vtemp_5;
        VlWide<4>/*127:0*/ __Vtemp_6;
        // Body
        __Vtemp_4[0U] = (~ ((IData)((((QData)((IData)(vlSelfRef.i4[1U])) 
                                      << 0x00000020U) 
                                     | (QData)((IData)(vlSelfRef.i4[0U])))) 
                            ^ (IData)((0x0000001fffffffffULL 
                                       & (((QData)((IData)(vlSelfRef.i4[1U])) 
                                           << 0x0000001eU) 
                                          | ((QData)((IData)(vlSelfRef.i4[0U])) 
                                             >> 2U))))));
        __Vtemp_4[1U] = (~ ((IData)(((((QData)((IData)(vlSelfRef.i4[1U])) 
                                       << 0x00000020U) 
                                      | (QData)((IData)(vlSelfRef.i4[0U]))) 
                                     >> 0x00000020U)) 
                            ^ (IData)(((0x0000001fffffffffULL 
                                        & (((QData)((IData)(vlSelfRef.i4[1U])) 
                                            << 0x0000001eU) 
                                           | ((QData)((IData)(vlSelfRef.i4[0U])) 
                                              >> 2U))) 
                                       >> 0x00000020U))));
        __Vtemp_4[2U] = 0xffffffffU;
        __Vtemp_4[3U] = 0xffffffffU;
        __Vtemp_5[0U] = vlSelfRef.rnd03__DOT__w1;
        __Vtemp_5[1U] = 0U;
        __Vtemp_5[2U] = 0U;
        __Vtemp_5[3U] = 0U;
        VL_MUL_W(4, __Vtemp_6, __Vtemp_4, __Vtemp_5);
        vlSelfRef.rnd03__DOT__w4 = (0x0000001fU & (~ 
                                                   (__Vtemp_6[0U] 
                                                    ^ 
                                                    (vlSelfRef.rnd03__DOT__w3 
                                                     & vlSelfRef.i3[0U]))));
        vlSelfRef.o0 = (1U & (IData)((0x00000001ffffffffULL 
                                      & VL_DIVS_QQQ(33, 
                                                    (0x00000001ffffffffULL 
                                                     & VL_EXTENDS_QI(33,5, (IData)(vlSelfRef.rnd03__DOT__w4))), vlSelfRef.rnd03__DOT__w8))));
        vlSelfRef.rnd03__DOT__w7 = (0x00ffffffU & (
                                                   (1U 
                                                    & ((~ 
                                                        (0U 
                                                         != (IData)(vlSelfRef.rnd03__DOT__w4))) 
                                                       | (0U 
                                                          != 
                                                          (0x000001ffU 
                                                           & VL_MODDIV_III(9, (IData)(vlSelfRef.rnd03__DOT__w4), (IData)(vlSelfRef.rnd03__DOT__w6))))))
                                                    ? (IData)(vlSelfRef.i5)
                                                    : (IData)(
                                                              (0x7fffffffffffffffULL 
                                                               & VL_DIV_QQQ(63, (QData)((IData)(vlSelfRef.rnd03__DOT__r2)), (QData)((IData)(vlSelfRef.rnd03__DOT__r3)))))));
        vlSelfRef.rnd03__DOT__w5[0U] = ((0xfffffffeU 
                                         & vlSelfRef.rnd03__DOT__w5[0U]) 
                                        | (((0x7fffffffffffffffULL 
                                             & VL_MODDIV_QQQ(63, (QData)((IData)(vlSelfRef.rnd03__DOT__r0)), vlSelfRef.i5)) 
                                            >= (QData)((IData)(
                                                               (0x000003ffU 
                                                                & (((IData)(vlSelfRef.rnd03__DOT__w4) 
                                                                    << 5U) 
                                                                   | (IData)(vlSelfRef.rnd03__DOT__w4)))))) 
                                           | (0U != 
                                              ((IData)(1U) 
                                               + (vlSelfRef.i5 
                                                  > (QData)((IData)(
                                                                    (1U 
                                                                     & (IData)(vlSelfRef.i5)))))))));
        vlSelfRef.rnd03__DOT__w5[0U] = (1U & vlSelfRef.rnd03__DOT__w5[0U]);
        vlSelfRef.rnd03__DOT__w5[1U] = 0U;
        vlSelfRef.rnd03__DOT__w5[2U] = 0U;
        vlSelfRef.o3 = (QData)((IData)(VL_LTES_III(24, vlSelfRef.rnd03__DOT__w11, 
                                                   (0x00ffffffU 
                                                    & (~ 
                                                       ((0x00000044U 
                                                         | vlSelfRef.rnd03__DOT__w3) 
                                                        ^ 
                                                        (0x0000000fU 
                                                         & (1U 
                                                            >> 
                                                            (0x0000000fU 
                                                             & vlSelfRef.rnd03__DOT__w5[0U])))))))));
    }

    VLCU_D static void Vrnd03___024root___eval_ico(Vrnd03___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        if ((4ULL & vlSelfRef.__VicoTriggered[0U])) {
            {
                // Inlined CFunc: _ico_sequent__TOP__0
                vlSelfRef.rnd03__DOT__w1 = (0x00ffffffU 
                                            & VL_DIV_III(24, (IData)(vlSelfRef.i1), (IData)(vlSelfRef.rnd03__DOT__r3)));
            }
        }
        if ((0x0000000000000042ULL & vlSelfRef.__VicoTriggered[0U])) {
            {
                // Inlined CFunc: _ico_comb__TOP__0
                VlWide<5>/*159:0*/ __Vinline_0__ico_comb__TOP__0___Vtemp_3;
                VlWide<5>/*159:0*/ __Vinline_0__ico_comb__TOP__0___Vtemp_4;
                VlWide<5>/*159:0*/ __Vinline_0__ico_comb__TOP__0___Vtemp_5;
                VlWide<5>/*159:0*/ __Vinline_0__ico_comb__TOP__0___Vtemp_6;
                vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0 
                    = (0x7fffffffffffffffULL & VL_DIV_QQQ(63, vlSelfRef.i0, vlSelfRef.i5));
                vlSelfRef.rnd03__DOT__w6 = (0x000001ffU 
                                            & (~ ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0) 
                                                  ^ (IData)(vlSelfRef.rnd03__DOT__r0))));
                vlSelfRef.rnd03__DOT__w3 = (0x00ffffffU 
                                            & (VL_DIV_III(24, (IData)(vlSelfRef.rnd03__DOT__r2), 
                                                          (0x00001fffU 
                                                           & (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0))) 
                                               ^ ((vlSelfRef.i0 
                                                   >> (IData)(vlSelfRef.rnd03__DOT__r2)) 
                                                  != (QData)((IData)(
                                                                     (1U 
                                                                      & (IData)(vlSelfRef.rnd03__DOT__r0)))))));
                __Vinline_0__ico_comb__TOP__0___Vtemp_3[0U] 
                    = (~ (IData)(vlSelfRef.rnd03__DOT__w6));
                __Vinline_0__ico_comb__TOP__0___Vtemp_3[1U] = 0xffffffffU;
                __Vinline_0__ico_comb__TOP__0___Vtemp_3[2U] = 0xffffffffU;
                __Vinline_0__ico_comb__TOP__0___Vtemp_3[3U] = 0xffffffffU;
                __Vinline_0__ico_comb__TOP__0___Vtemp_3[4U] = 0xffffffffU;
                VL_SHIFTL_WWI(150,150,8, __Vinline_0__ico_comb__TOP__0___Vtemp_4, __Vinline_0__ico_comb__TOP__0___Vtemp_3, 
                              (0x000000ffU & (IData)(vlSelfRef.rnd03__DOT__w6)));
                __Vinline_0__ico_comb__TOP__0___Vtemp_5[0U] 
                    = __Vinline_0__ico_comb__TOP__0___Vtemp_4[0U];
                __Vinline_0__ico_comb__TOP__0___Vtemp_5[1U] 
                    = __Vinline_0__ico_comb__TOP__0___Vtemp_4[1U];
                __Vinline_0__ico_comb__TOP__0___Vtemp_5[2U] 
                    = __Vinline_0__ico_comb__TOP__0___Vtemp_4[2U];
                __Vinline_0__ico_comb__TOP__0___Vtemp_5[3U] 
                    = __Vinline_0__ico_comb__TOP__0___Vtemp_4[3U];
                __Vinline_0__ico_comb__TOP__0___Vtemp_5[4U] 
                    = (0x003fffffU & __Vinline_0__ico_comb__TOP__0___Vtemp_4[4U]);
                VL_SHIFTR_WWI(150,150,8, __Vinline_0__ico_comb__TOP__0___Vtemp_6, __Vinline_0__ico_comb__TOP__0___Vtemp_5, 
                              (0x000000ffU & (IData)(vlSelfRef.i0)));
                vlSelfRef.rnd03__DOT__w11 = (0x00ffffffU 
                                             & __Vinline_0__ico_comb__TOP__0___Vtemp_6[0U]);
                vlSelfRef.rnd03__DOT__w8 = (0x00000001ffffffffULL 
                                            & VL_SHIFTL_QQI(33,33,8, (QData)((IData)(vlSelfRef.rnd03__DOT__w3)), 
                                                            (0x000000ffU 
                                                             & (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0))));
            }
        }
        if ((0x0000000000000044ULL & vlSelfRef.__VicoTriggered[0U])) {
            {
                // Inlined CFunc: _ico_comb__TOP__1
                vlSelfRef.o1 = VL_SHIFTR_IIQ(16,16,63, 0xffffU, 
                                             (0x7fffffffffffffffULL 
                                              & ((QData)((IData)(
                                                                 (1U 
                                                                  & (~ 
                                                                     VL_REDXOR_32(vlSelfRef.rnd03__DOT__mem
                                                                                [
                                                                                (0x0000000fU 
                                                                                & vlSelfRef.rnd03__DOT__w1)]))))) 
                                                 - vlSelfRef.i5)));
            }
        }
        if ((0x000000000000005aULL & vlSelfRef.__VicoTriggered[0U])) {
            {
                // Inlined CFunc: _ico_comb__TOP__2
                vlSelfRef.rnd03__DOT__w10 = (0x00ffffffU 
                                             & (VL_LTS_III(1, 
                                                           VL_GTES_III(7, 0U, 
                                                                       (0x0000007fU 
                                                                        & VL_EXTENDS_II(7,1, 
                                                                                (1U 
                                                                                & (vlSelfRef.i3
                                                                                [
                                                                                ((1U 
                                                                                < (IData)(vlSelfRef.rnd03__DOT__r0)) 
                                                                                >> 5U)] 
                                                                                >> 
                                                                                (0x0000001fU 
                                                                                & (1U 
                                                                                < (IData)(vlSelfRef.rnd03__DOT__r0)))))))), 
                                                           (1U 
                                                            & (IData)(
                                                                      (vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0 
                                                                       >> (IData)(vlSelfRef.rnd03__DOT__r2))))) 
                                                ^ (
                                                   (~ (IData)(vlSelfRef.i0)) 
                                                   ^ 
                                                   VL_SHIFTL_III(24,24,8, vlSelfRef.i2[0U], 
                                                                 (0x000000ffU 
                                                                  & (IData)(vlSelfRef.rnd03__DOT__w6))))));
            }
        }
        if ((0x0000000000000076ULL & vlSelfRef.__VicoTriggered[0U])) {
            Vrnd03___024root___ico_comb__TOP__3(vlSelf, vlCtx);
        }
        if ((0x0000000000000046ULL & vlSelfRef.__VicoTriggered[0U])) {
            {
                // Inlined CFunc: _ico_comb__TOP__4
                VlWide<3>/*95:0*/ __Vinline_0__ico_comb__TOP__4___Vtemp_2;
                VlWide<3>/*95:0*/ __Vinline_0__ico_comb__TOP__4___Vtemp_3;
                VlWide<3>/*95:0*/ __Vinline_0__ico_comb__TOP__4___Vtemp_4;
                VlWide<3>/*95:0*/ __Vinline_0__ico_comb__TOP__4___Vtemp_5;
                VlWide<3>/*95:0*/ __Vinline_0__ico_comb__TOP__4___Vtemp_6;
                VlWide<3>/*95:0*/ __Vinline_0__ico_comb__TOP__4___Vtemp_7;
                __Vinline_0__ico_comb__TOP__4___Vtemp_2[0U] 
                    = (IData)(vlSelfRef.i5);
                __Vinline_0__ico_comb__TOP__4___Vtemp_2[1U] 
                    = ((vlSelfRef.rnd03__DOT__w11 << 0x0000001fU) 
                       | (IData)((vlSelfRef.i5 >> 0x00000020U)));
                __Vinline_0__ico_comb__TOP__4___Vtemp_2[2U] 
                    = (vlSelfRef.rnd03__DOT__w11 >> 1U);
                VL_SHIFTR_WWI(87,87,8, __Vinline_0__ico_comb__TOP__4___Vtemp_3, __Vinline_0__ico_comb__TOP__4___Vtemp_2, 
                              (0x000000ffU & (IData)(vlSelfRef.rnd03__DOT__w6)));
                __Vinline_0__ico_comb__TOP__4___Vtemp_4[0U] 
                    = __Vinline_0__ico_comb__TOP__4___Vtemp_3[0U];
                __Vinline_0__ico_comb__TOP__4___Vtemp_4[1U] 
                    = __Vinline_0__ico_comb__TOP__4___Vtemp_3[1U];
                __Vinline_0__ico_comb__TOP__4___Vtemp_4[2U] 
                    = __Vinline_0__ico_comb__TOP__4___Vtemp_3[2U];
                if ((0U != (VL_COUNTONES_I((0x000000ffU 
                                            & (vlSelfRef.rnd03__DOT__mem
                                               [(0x0000000fU 
                                                 & vlSelfRef.rnd03__DOT__w1)] 
                                               >> (0x0000000fU 
                                                   & VL_SHIFTL_III(4,4,8, vlSelfRef.rnd03__DOT__w3, 
                                                                   (0x000000ffU 
                                                                    & (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0))))))) 
                            + VL_ONEHOT_I(vlSelfRef.rnd03__DOT__w3)))) {
                    __Vinline_0__ico_comb__TOP__4___Vtemp_5[0U] = 1U;
                    __Vinline_0__ico_comb__TOP__4___Vtemp_5[1U] = 0U;
                    __Vinline_0__ico_comb__TOP__4___Vtemp_5[2U] = 0U;
                } else {
                    __Vinline_0__ico_comb__TOP__4___Vtemp_6[0U] 
                        = (3U & (IData)((vlSelfRef.i0 
                                         >> 4U)));
                    __Vinline_0__ico_comb__TOP__4___Vtemp_6[1U] = 0U;
                    __Vinline_0__ico_comb__TOP__4___Vtemp_6[2U] = 0U;
                    VL_SHIFTL_WWI(65,65,3, __Vinline_0__ico_comb__TOP__4___Vtemp_5, __Vinline_0__ico_comb__TOP__4___Vtemp_6, (IData)(vlSelfRef.rnd03__DOT__r0));
                }
                VL_SUB_W(3, __Vinline_0__ico_comb__TOP__4___Vtemp_7, __Vinline_0__ico_comb__TOP__4___Vtemp_4, __Vinline_0__ico_comb__TOP__4___Vtemp_5);
                vlSelfRef.o4[0U] = __Vinline_0__ico_comb__TOP__4___Vtemp_7[0U];
                vlSelfRef.o4[1U] = __Vinline_0__ico_comb__TOP__4___Vtemp_7[1U];
                vlSelfRef.o4[2U] = (1U & __Vinline_0__ico_comb__TOP__4___Vtemp_7[2U]);
            }
        }
        if ((0x000000000000007eULL & vlSelfRef.__VicoTriggered[0U])) {
            {
                // Inlined CFunc: _ico_comb__TOP__5
                VlWide<3>/*95:0*/ __Vinline_0__ico_comb__TOP__5___Vtemp_2;
                VlWide<3>/*95:0*/ __Vinline_0__ico_comb__TOP__5___Vtemp_3;
                VlWide<3>/*95:0*/ __Vinline_0__ico_comb__TOP__5___Vtemp_4;
                VlWide<5>/*159:0*/ __Vinline_0__ico_comb__TOP__5___Vtemp_6;
                VlWide<5>/*159:0*/ __Vinline_0__ico_comb__TOP__5___Vtemp_7;
                __Vinline_0__ico_comb__TOP__5___Vtemp_2[0U] 
                    = (1U & (IData)((vlSelfRef.rnd03__DOT__w8 
                                     >> (0x0000001fU 
                                         & vlSelfRef.rnd03__DOT__w5[0U]))));
                __Vinline_0__ico_comb__TOP__5___Vtemp_2[1U] = 0U;
                __Vinline_0__ico_comb__TOP__5___Vtemp_2[2U] = 0U;
                VL_MUL_W(3, __Vinline_0__ico_comb__TOP__5___Vtemp_3, vlSelfRef.rnd03__DOT__r1, __Vinline_0__ico_comb__TOP__5___Vtemp_2);
                __Vinline_0__ico_comb__TOP__5___Vtemp_4[0U] 
                    = __Vinline_0__ico_comb__TOP__5___Vtemp_3[0U];
                __Vinline_0__ico_comb__TOP__5___Vtemp_4[1U] 
                    = __Vinline_0__ico_comb__TOP__5___Vtemp_3[1U];
                __Vinline_0__ico_comb__TOP__5___Vtemp_4[2U] 
                    = (0x000000ffU & __Vinline_0__ico_comb__TOP__5___Vtemp_3[2U]);
                VL_EXTENDS_WI(130,24, __Vinline_0__ico_comb__TOP__5___Vtemp_6, vlSelfRef.rnd03__DOT__mem
                              [(0x0000000fU & vlSelfRef.rnd03__DOT__w1)]);
                __Vinline_0__ico_comb__TOP__5___Vtemp_7[0U] 
                    = __Vinline_0__ico_comb__TOP__5___Vtemp_6[0U];
                __Vinline_0__ico_comb__TOP__5___Vtemp_7[1U] 
                    = __Vinline_0__ico_comb__TOP__5___Vtemp_6[1U];
                __Vinline_0__ico_comb__TOP__5___Vtemp_7[2U] 
                    = __Vinline_0__ico_comb__TOP__5___Vtemp_6[2U];
                __Vinline_0__ico_comb__TOP__5___Vtemp_7[3U] 
                    = __Vinline_0__ico_comb__TOP__5___Vtemp_6[3U];
                __Vinline_0__ico_comb__TOP__5___Vtemp_7[4U] 
                    = (3U & __Vinline_0__ico_comb__TOP__5___Vtemp_6[4U]);
                vlSelfRef.o2 = ((VL_COUNTONES_I((IData)(vlSelfRef.rnd03__DOT__w4)) 
                                 + VL_ONEHOT_W(3, __Vinline_0__ico_comb__TOP__5___Vtemp_4)) 
                                <= (VL_LTES_IWW(130, vlSelfRef.i2, __Vinline_0__ico_comb__TOP__5___Vtemp_7) 
                                    == (0x00ffffffU 
                                        & ((1U & ((IData)(vlSelfRef.rnd03__DOT__r2) 
                                                  >> 
                                                  (1U 
                                                   & (IData)(vlSelfRef.rnd03__DOT__r2)))) 
                                           + vlSelfRef.rnd03__DOT__w7))));
            }
        }
    }

    VLCU_D static bool Vrnd03___024root___eval_phase__ico(Vrnd03___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VicoExecute;
        // Body
        Vrnd03___024root___eval_triggers_vec__ico(vlSelf, vlCtx);
#ifdef VL_DEBUG
        if (VL_UNLIKELY(vlSymsp->_vm_contextp__->debug())) {
            Vrnd03___024root___dump_triggers__ico(vlSelfRef.__VicoTriggered, "ico"s);
        }
#endif
        __VicoExecute = Vrnd03___024root___trigger_anySet__ico(vlSelfRef.__VicoTriggered);
        if (__VicoExecute) {
            Vrnd03___024root___eval_ico(vlSelf, vlCtx);
        }
        return (__VicoExecute);
    }

    VLCU_D static bool Vrnd03___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vrnd03___024root___nba_sequent__TOP__0(Vrnd03___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        VlWide<3>/*71:0*/ __Vdly__rnd03__DOT__r1;
        VL_ZERO_W(72, __Vdly__rnd03__DOT__r1);
        IData/*23:0*/ __VdlyVal__rnd03__DOT__mem__v0;
        __VdlyVal__rnd03__DOT__mem__v0 = 0;
        CData/*3:0*/ __VdlyDim0__rnd03__DOT__mem__v0;
        __VdlyDim0__rnd03__DOT__mem__v0 = 0;
        VlWide<4>/*127:0*/ __Vtemp_2;
        VlWide<4>/*127:0*/ __Vtemp_5;
        VlWide<4>/*127:0*/ __Vtemp_6;
        VlWide<4>/*127:0*/ __Vtemp_7;
        VlWide<3>/*95:0*/ __Vtemp_14;
        VlWide<3>/*95:0*/ __Vtemp_18;
        VlWide<3>/*95:0*/ __Vtemp_19;
        VlWide<3>/*95:0*/ __Vtemp_23;
        VlWide<3>/*95:0*/ __Vtemp_25;
        VlWide<3>/*95:0*/ __Vtemp_26;
        VlWide<3>/*95:0*/ __Vtemp_27;
        VlWide<3>/*95:0*/ __Vtemp_29;
        VlWide<3>/*95:0*/ __Vtemp_30;
        VlWide<3>/*95:0*/ __Vtemp_31;
        VlWide<3>/*95:0*/ __Vtemp_33;
        VlWide<3>/*95:0*/ __Vtemp_34;
        VlWide<5>/*159:0*/ __Vtemp_37;
        VlWide<5>/*159:0*/ __Vtemp_38;
        VlWide<5>/*159:0*/ __Vtemp_39;
        VlWide<5>/*159:0*/ __Vtemp_40;
        VlWide<4>/*127:0*/ __Vtemp_44;
        VlWide<4>/*127:0*/ __Vtemp_45;
        VlWide<4>/*127:0*/ __Vtemp_46;
        VlWide<3>/*95:0*/ __Vtemp_48;
        VlWide<3>/*95:0*/ __Vtemp_49;
        VlWide<3>/*95:0*/ __Vtemp_50;
        VlWide<3>/*95:0*/ __Vtemp_51;
        VlWide<3>/*95:0*/ __Vtemp_52;
        VlWide<3>/*95:0*/ __Vtemp_53;
        VlWide<3>/*95:0*/ __Vtemp_55;
        VlWide<3>/*95:0*/ __Vtemp_56;
        VlWide<3>/*95:0*/ __Vtemp_57;
        VlWide<5>/*159:0*/ __Vtemp_59;
        VlWide<5>/*159:0*/ __Vtemp_60;
        // Body
        __Vdly__rnd03__DOT__r1[0U] = vlSelfRef.rnd03__DOT__r1[0U];
        __Vdly__rnd03__DOT__r1[1U] = vlSelfRef.rnd03__DOT__r1[1U];
        __Vdly__rnd03__DOT__r1[2U] = vlSelfRef.rnd03__DOT__r1[2U];
        __VdlyVal__rnd03__DOT__mem__v0 = (0x00ffffffU 
                                          & ((IData)(1U) 
                                             - (1U 
                                                & (~ 
                                                   (0x7fffffffffffffffULL 
                                                    == vlSelfRef.i5)))));
        __VdlyDim0__rnd03__DOT__mem__v0 = (0x0000000fU 
                                           & vlSelfRef.rnd03__DOT__w11);
        __Vtemp_2[0U] = (1U & (vlSelfRef.rnd03__DOT__w10 
                               >> (0x0000000fU & (IData)(vlSelfRef.rnd03__DOT__r3))));
        __Vtemp_2[1U] = 0U;
        __Vtemp_2[2U] = 0U;
        __Vtemp_2[3U] = 0U;
        __Vtemp_5[0U] = (1U | vlSelfRef.rnd03__DOT__w7);
        __Vtemp_5[1U] = 0U;
        __Vtemp_5[2U] = 0U;
        __Vtemp_5[3U] = 0U;
        VL_MUL_W(4, __Vtemp_6, __Vtemp_2, __Vtemp_5);
        VL_SHIFTL_WWI(100,100,8, __Vtemp_7, __Vtemp_6, 
                      (0x000000ffU & vlSelfRef.i2[0U]));
        vlSelfRef.rnd03__DOT__r0 = (7U & __Vtemp_7[0U]);
        vlSelfRef.rnd03__DOT__mem[__VdlyDim0__rnd03__DOT__mem__v0] 
            = __VdlyVal__rnd03__DOT__mem__v0;
        vlSelfRef.rnd03__DOT__w6 = (0x000001ffU & (~ 
                                                   ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0) 
                                                    ^ (IData)(vlSelfRef.rnd03__DOT__r0))));
        __Vtemp_14[0U] = (vlSelfRef.rnd03__DOT__r1[0U] 
                          ^ vlSelfRef.rnd03__DOT__w1);
        __Vtemp_14[1U] = vlSelfRef.rnd03__DOT__r1[1U];
        __Vtemp_14[2U] = vlSelfRef.rnd03__DOT__r1[2U];
        __Vtemp_18[0U] = (~ (0x0000007fU & (vlSelfRef.i4[0U] 
                                            >> 0x00000016U)));
        __Vtemp_18[1U] = 0xffffffffU;
        __Vtemp_18[2U] = 0x000000ffU;
        VL_MUL_W(3, __Vtemp_19, __Vtemp_14, __Vtemp_18);
        __Vdly__rnd03__DOT__r1[0U] = (~ (((7U & (- (IData)(
                                                           (1U 
                                                            & (vlSelfRef.rnd03__DOT__w1 
                                                               >> 
                                                               (0x0000000fU 
                                                                & (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0))))))) 
                                          ^ (1U & (~ (IData)(
                                                             (vlSelfRef.i0 
                                                              >> 1U))))) 
                                         ^ __Vtemp_19[0U]));
        __Vdly__rnd03__DOT__r1[1U] = (~ __Vtemp_19[1U]);
        __Vdly__rnd03__DOT__r1[2U] = (0x000000ffU & 
                                      (~ __Vtemp_19[2U]));
        __Vtemp_23[0U] = (0x2bc755d3U & vlSelfRef.rnd03__DOT__r1[0U]);
        __Vtemp_23[1U] = (0x0000b8c5U & vlSelfRef.rnd03__DOT__r1[1U]);
        __Vtemp_23[2U] = 0U;
        VL_SHIFTR_WWI(65,65,32, __Vtemp_25, vlSelfRef.rnd03__DOT__w5, 1U);
        VL_EXTENDS_WW(72,65, __Vtemp_26, __Vtemp_25);
        __Vtemp_27[0U] = __Vtemp_26[0U];
        __Vtemp_27[1U] = __Vtemp_26[1U];
        __Vtemp_27[2U] = (0x000000ffU & __Vtemp_26[2U]);
        vlSelfRef.rnd03__DOT__r2 = (7U & (- VL_LTS_IWW(72, __Vtemp_23, __Vtemp_27)));
        __Vtemp_29[0U] = vlSelfRef.rnd03__DOT__w10;
        __Vtemp_29[1U] = 0U;
        __Vtemp_29[2U] = 0U;
        VL_SUB_W(3, __Vtemp_30, vlSelfRef.rnd03__DOT__w5, __Vtemp_29);
        __Vtemp_31[0U] = __Vtemp_30[0U];
        __Vtemp_31[1U] = __Vtemp_30[1U];
        __Vtemp_31[2U] = (1U & __Vtemp_30[2U]);
        VL_EXTENDS_WI(65,4, __Vtemp_33, (0x0000000fU 
                                         & (vlSelfRef.rnd03__DOT__w7 
                                            >> 2U)));
        __Vtemp_34[0U] = __Vtemp_33[0U];
        __Vtemp_34[1U] = __Vtemp_33[1U];
        __Vtemp_34[2U] = (1U & __Vtemp_33[2U]);
        vlSelfRef.rnd03__DOT__r3 = (0x0000007fU & (~ 
                                                   (VL_SHIFTR_III(7,24,8, 
                                                                  (0x00ffffffU 
                                                                   & (vlSelfRef.rnd03__DOT__w3 
                                                                      + vlSelfRef.rnd03__DOT__w7)), 
                                                                  (0x000000ffU 
                                                                   & vlSelfRef.rnd03__DOT__r1[0U])) 
                                                    ^ 
                                                    VL_GTES_IWW(65, __Vtemp_31, __Vtemp_34))));
        __Vtemp_37[0U] = (~ (IData)(vlSelfRef.rnd03__DOT__w6));
        __Vtemp_37[1U] = 0xffffffffU;
        __Vtemp_37[2U] = 0xffffffffU;
        __Vtemp_37[3U] = 0xffffffffU;
        __Vtemp_37[4U] = 0xffffffffU;
        VL_SHIFTL_WWI(150,150,8, __Vtemp_38, __Vtemp_37, 
                      (0x000000ffU & (IData)(vlSelfRef.rnd03__DOT__w6)));
        __Vtemp_39[0U] = __Vtemp_38[0U];
        __Vtemp_39[1U] = __Vtemp_38[1U];
        __Vtemp_39[2U] = __Vtemp_38[2U];
        __Vtemp_39[3U] = __Vtemp_38[3U];
        __Vtemp_39[4U] = (0x003fffffU & __Vtemp_38[4U]);
        VL_SHIFTR_WWI(150,150,8, __Vtemp_40, __Vtemp_39, 
                      (0x000000ffU & (IData)(vlSelfRef.i0)));
        vlSelfRef.rnd03__DOT__w11 = (0x00ffffffU & __Vtemp_40[0U]);
        vlSelfRef.rnd03__DOT__r1[0U] = __Vdly__rnd03__DOT__r1[0U];
        vlSelfRef.rnd03__DOT__r1[1U] = __Vdly__rnd03__DOT__r1[1U];
        vlSelfRef.rnd03__DOT__r1[2U] = __Vdly__rnd03__DOT__r1[2U];
        vlSelfRef.rnd03__DOT__w1 = (0x00ffffffU & VL_DIV_III(24, (IData)(vlSelfRef.i1), (IData)(vlSelfRef.rnd03__DOT__r3)));
        vlSelfRef.rnd03__DOT__w3 = (0x00ffffffU & (
                                                   VL_DIV_III(24, (IData)(vlSelfRef.rnd03__DOT__r2), 
                                                              (0x00001fffU 
                                                               & (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0))) 
                                                   ^ 
                                                   ((vlSelfRef.i0 
                                                     >> (IData)(vlSelfRef.rnd03__DOT__r2)) 
                                                    != (QData)((IData)(
                                                                       (1U 
                                                                        & (IData)(vlSelfRef.rnd03__DOT__r0)))))));
        vlSelfRef.rnd03__DOT__w10 = (0x00ffffffU & 
                                     (VL_LTS_III(1, 
                                                 VL_GTES_III(7, 0U, 
                                                             (0x0000007fU 
                                                              & VL_EXTENDS_II(7,1, 
                                                                              (1U 
                                                                               & (vlSelfRef.i3
                                                                                [
                                                                                ((1U 
                                                                                < (IData)(vlSelfRef.rnd03__DOT__r0)) 
                                                                                >> 5U)] 
                                                                                >> 
                                                                                (0x0000001fU 
                                                                                & (1U 
                                                                                < (IData)(vlSelfRef.rnd03__DOT__r0)))))))), 
                                                 (1U 
                                                  & (IData)(
                                                            (vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0 
                                                             >> (IData)(vlSelfRef.rnd03__DOT__r2))))) 
                                      ^ ((~ (IData)(vlSelfRef.i0)) 
                                         ^ VL_SHIFTL_III(24,24,8, vlSelfRef.i2[0U], 
                                                         (0x000000ffU 
                                                          & (IData)(vlSelfRef.rnd03__DOT__w6))))));
        vlSelfRef.o1 = VL_SHIFTR_IIQ(16,16,63, 0xffffU, 
                                     (0x7fffffffffffffffULL 
                                      & ((QData)((IData)(
                                                         (1U 
                                                          & (~ 
                                                             VL_REDXOR_32(vlSelfRef.rnd03__DOT__mem
                                                                          [
                                                                          (0x0000000fU 
                                                                           & vlSelfRef.rnd03__DOT__w1)]))))) 
                                         - vlSelfRef.i5)));
        vlSelfRef.rnd03__DOT__w8 = (0x00000001ffffffffULL 
                                    & VL_SHIFTL_QQI(33,33,8, (QData)((IData)(vlSelfRef.rnd03__DOT__w3)), 
                                                    (0x000000ffU 
                                                     & (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0))));
        __Vtemp_44[0U] = (~ ((IData)((((QData)((IData)(vlSelfRef.i4[1U])) 
                                       << 0x00000020U) 
                                      | (QData)((IData)(vlSelfRef.i4[0U])))) 
                             ^ (IData)((0x0000001fffffffffULL 
                                        & (((QData)((IData)(vlSelfRef.i4[1U])) 
                                            << 0x0000001eU) 
                                           | ((QData)((IData)(vlSelfRef.i4[0U])) 
                                              >> 2U))))));
        __Vtemp_44[1U] = (~ ((IData)(((((QData)((IData)(vlSelfRef.i4[1U])) 
                                        << 0x00000020U) 
                                       | (QData)((IData)(vlSelfRef.i4[0U]))) 
                                      >> 0x00000020U)) 
                             ^ (IData)(((0x0000001fffffffffULL 
                                         & (((QData)((IData)(vlSelfRef.i4[1U])) 
                                             << 0x0000001eU) 
                                            | ((QData)((IData)(vlSelfRef.i4[0U])) 
                                               >> 2U))) 
                                        >> 0x00000020U))));
        __Vtemp_44[2U] = 0xffffffffU;
        __Vtemp_44[3U] = 0xffffffffU;
        __Vtemp_45[0U] = vlSelfRef.rnd03__DOT__w1;
        __Vtemp_45[1U] = 0U;
        __Vtemp_45[2U] = 0U;
        __Vtemp_45[3U] = 0U;
        VL_MUL_W(4, __Vtemp_46, __Vtemp_44, __Vtemp_45);
        vlSelfRef.rnd03__DOT__w4 = (0x0000001fU & (~ 
                                                   (__Vtemp_46[0U] 
                                                    ^ 
                                                    (vlSelfRef.rnd03__DOT__w3 
                                                     & vlSelfRef.i3[0U]))));
        __Vtemp_48[0U] = (IData)(vlSelfRef.i5);
        __Vtemp_48[1U] = ((vlSelfRef.rnd03__DOT__w11 
                           << 0x0000001fU) | (IData)(
                                                     (vlSelfRef.i5 
                                                      >> 0x00000020U)));
        __Vtemp_48[2U] = (vlSelfRef.rnd03__DOT__w11 
                          >> 1U);
        VL_SHIFTR_WWI(87,87,8, __Vtemp_49, __Vtemp_48, 
                      (0x000000ffU & (IData)(vlSelfRef.rnd03__DOT__w6)));
        __Vtemp_50[0U] = __Vtemp_49[0U];
        __Vtemp_50[1U] = __Vtemp_49[1U];
        __Vtemp_50[2U] = __Vtemp_49[2U];
        if ((0U != (VL_COUNTONES_I((0x000000ffU & (vlSelfRef.rnd03__DOT__mem
                                                   [
                                                   (0x0000000fU 
                                                    & vlSelfRef.rnd03__DOT__w1)] 
                                                   >> 
                                                   (0x0000000fU 
                                                    & VL_SHIFTL_III(4,4,8, vlSelfRef.rnd03__DOT__w3, 
                                                                    (0x000000ffU 
                                                                     & (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0))))))) 
                    + VL_ONEHOT_I(vlSelfRef.rnd03__DOT__w3)))) {
            __Vtemp_51[0U] = 1U;
            __Vtemp_51[1U] = 0U;
            __Vtemp_51[2U] = 0U;
        } else {
            __Vtemp_52[0U] = (3U & (IData)((vlSelfRef.i0 
                                            >> 4U)));
            __Vtemp_52[1U] = 0U;
            __Vtemp_52[2U] = 0U;
            VL_SHIFTL_WWI(65,65,3, __Vtemp_51, __Vtemp_52, (IData)(vlSelfRef.rnd03__DOT__r0));
        }
        VL_SUB_W(3, __Vtemp_53, __Vtemp_50, __Vtemp_51);
        vlSelfRef.o4[0U] = __Vtemp_53[0U];
        vlSelfRef.o4[1U] = __Vtemp_53[1U];
        vlSelfRef.o4[2U] = (1U & __Vtemp_53[2U]);
        vlSelfRef.o0 = (1U & (IData)((0x00000001ffffffffULL 
                                      & VL_DIVS_QQQ(33, 
                                                    (0x00000001ffffffffULL 
                                                     & VL_EXTENDS_QI(33,5, (IData)(vlSelfRef.rnd03__DOT__w4))), vlSelfRef.rnd03__DOT__w8))));
        vlSelfRef.rnd03__DOT__w7 = (0x00ffffffU & (
                                                   (1U 
                                                    & ((~ 
                                                        (0U 
                                                         != (IData)(vlSelfRef.rnd03__DOT__w4))) 
                                                       | (0U 
                                                          != 
                                                          (0x000001ffU 
                                                           & VL_MODDIV_III(9, (IData)(vlSelfRef.rnd03__DOT__w4), (IData)(vlSelfRef.rnd03__DOT__w6))))))
                                                    ? (IData)(vlSelfRef.i5)
                                                    : (IData)(
                                                              (0x7fffffffffffffffULL 
                                                               & VL_DIV_QQQ(63, (QData)((IData)(vlSelfRef.rnd03__DOT__r2)), (QData)((IData)(vlSelfRef.rnd03__DOT__r3)))))));
        vlSelfRef.rnd03__DOT__w5[0U] = ((0xfffffffeU 
                                         & vlSelfRef.rnd03__DOT__w5[0U]) 
                                        | (((0x7fffffffffffffffULL 
                                             & VL_MODDIV_QQQ(63, (QData)((IData)(vlSelfRef.rnd03__DOT__r0)), vlSelfRef.i5)) 
                                            >= (QData)((IData)(
                                                               (0x000003ffU 
                                                                & (((IData)(vlSelfRef.rnd03__DOT__w4) 
                                                                    << 5U) 
                                                                   | (IData)(vlSelfRef.rnd03__DOT__w4)))))) 
                                           | (0U != 
                                              ((IData)(1U) 
                                               + (vlSelfRef.i5 
                                                  > (QData)((IData)(
                                                                    (1U 
                                                                     & (IData)(vlSelfRef.i5)))))))));
        vlSelfRef.rnd03__DOT__w5[0U] = (1U & vlSelfRef.rnd03__DOT__w5[0U]);
        vlSelfRef.rnd03__DOT__w5[1U] = 0U;
        vlSelfRef.rnd03__DOT__w5[2U] = 0U;
        vlSelfRef.o3 = (QData)((IData)(VL_LTES_III(24, vlSelfRef.rnd03__DOT__w11, 
                                                   (0x00ffffffU 
                                                    & (~ 
                                                       ((0x00000044U 
                                                         | vlSelfRef.rnd03__DOT__w3) 
                                                        ^ 
                                                        (0x0000000fU 
                                                         & (1U 
                                                            >> 
                                                            (0x0000000fU 
                                                             & vlSelfRef.rnd03__DOT__w5[0U])))))))));
        __Vtemp_55[0U] = (1U & (IData)((vlSelfRef.rnd03__DOT__w8 
                                        >> (0x0000001fU 
                                            & vlSelfRef.rnd03__DOT__w5[0U]))));
        __Vtemp_55[1U] = 0U;
        __Vtemp_55[2U] = 0U;
        VL_MUL_W(3, __Vtemp_56, vlSelfRef.rnd03__DOT__r1, __Vtemp_55);
        __Vtemp_57[0U] = __Vtemp_56[0U];
        __Vtemp_57[1U] = __Vtemp_56[1U];
        __Vtemp_57[2U] = (0x000000ffU & __Vtemp_56[2U]);
        VL_EXTENDS_WI(130,24, __Vtemp_59, vlSelfRef.rnd03__DOT__mem
                      [(0x0000000fU & vlSelfRef.rnd03__DOT__w1)]);
        __Vtemp_60[0U] = __Vtemp_59[0U];
        __Vtemp_60[1U] = __Vtemp_59[1U];
        __Vtemp_60[2U] = __Vtemp_59[2U];
        __Vtemp_60[3U] = __Vtemp_59[3U];
        __Vtemp_60[4U] = (3U & __Vtemp_59[4U]);
        vlSelfRef.o2 = ((VL_COUNTONES_I((IData)(vlSelfRef.rnd03__DOT__w4)) 
                         + VL_ONEHOT_W(3, __Vtemp_57)) 
                        <= (VL_LTES_IWW(130, vlSelfRef.i2, __Vtemp_60) 
                            == (0x00ffffffU & ((1U 
                                                & ((IData)(vlSelfRef.rnd03__DOT__r2) 
                                                   >> 
                                                   (1U 
                                                    & (IData)(vlSelfRef.rnd03__DOT__r2)))) 
                                               + vlSelfRef.rnd03__DOT__w7))));
    }

    VLCU_D static void Vrnd03___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vrnd03___024root___eval_phase__act(Vrnd03___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        {
            // Inlined CFunc: _eval_triggers_vec__act
            vlSelfRef.__VactTriggered[0U] = (QData)((IData)(
                                                            ((IData)(vlSelfRef.clk) 
                                                             & (~ (IData)(vlSelfRef.__Vtrigprevexpr___TOP__clk__1)))));
            vlSelfRef.__Vtrigprevexpr___TOP__clk__1 
                = vlSelfRef.clk;
        }
#ifdef VL_DEBUG
        if (VL_UNLIKELY(vlSymsp->_vm_contextp__->debug())) {
            Vrnd03___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vrnd03___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vrnd03___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vrnd03___024root___eval_phase__nba(Vrnd03___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vrnd03___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    Vrnd03___024root___nba_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
            Vrnd03___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vrnd03___024root___eval(Vrnd03___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        IData/*31:0*/ __VicoIterCount;
        IData/*31:0*/ __VnbaIterCount;
        // Body
        __VicoIterCount = 0U;
        do {
            if (VL_UNLIKELY(((0x00002710U < __VicoIterCount)))) {
#ifdef VL_DEBUG
                Vrnd03___024root___dump_triggers__ico(vlSelfRef.__VicoTriggered, "ico"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/rnd03.v", 2, "", "DIDNOTCONVERGE: Input combinational region did not converge after '--converge-limit' of 10000 tries");
            }
            __VicoIterCount = ((IData)(1U) + __VicoIterCount);
            vlSelfRef.__VicoPhaseResult = Vrnd03___024root___eval_phase__ico(vlSelf, vlCtx);
        } while (vlSelfRef.__VicoPhaseResult);
        __VnbaIterCount = 0U;
        do {
            if (VL_UNLIKELY(((0x00002710U < __VnbaIterCount)))) {
#ifdef VL_DEBUG
                Vrnd03___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/rnd03.v", 2, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vrnd03___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("../verilator_b200/csrc/designs/torture/rnd03.v", 2, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vrnd03___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vrnd03___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vrnd03___024root___ctor_var_reset(Vrnd03___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        vlSelf->i0 = VL_SCOPED_RAND_RESET_Q(40, __VscopeHash, 8471475882151657015ull);
        vlSelf->i1 = VL_SCOPED_RAND_RESET_I(7, __VscopeHash, 2165040029148572678ull);
        VL_SCOPED_RAND_RESET_W(130, vlSelf->i2, __VscopeHash, 12628911202397369737ull);
        VL_SCOPED_RAND_RESET_W(128, vlSelf->i3, __VscopeHash, 6886176223853197501ull);
        VL_SCOPED_RAND_RESET_W(96, vlSelf->i4, __VscopeHash, 5103504462683158156ull);
        vlSelf->i5 = VL_SCOPED_RAND_RESET_Q(63, __VscopeHash, 10954576514065331241ull);
        vlSelf->o0 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 4709259436872190324ull);
        vlSelf->o1 = VL_SCOPED_RAND_RESET_I(16, __VscopeHash, 16871796271195187576ull);
        vlSelf->o2 = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 6221000330064061286ull);
        vlSelf->o3 = VL_SCOPED_RAND_RESET_Q(40, __VscopeHash, 17472377812510545679ull);
        VL_SCOPED_RAND_RESET_W(65, vlSelf->o4, __VscopeHash, 12832999154864337794ull);
        vlSelf->rnd03__DOT__r0 = VL_SCOPED_RAND_RESET_I(3, __VscopeHash, 17334884956985768054ull);
        VL_SCOPED_RAND_RESET_W(72, vlSelf->rnd03__DOT__r1, __VscopeHash, 12146097306228259504ull);
        vlSelf->rnd03__DOT__r2 = VL_SCOPED_RAND_RESET_I(3, __VscopeHash, 157248113326855794ull);
        vlSelf->rnd03__DOT__r3 = VL_SCOPED_RAND_RESET_I(7, __VscopeHash, 4167754743492137210ull);
        for (int __Vi0 = 0; __Vi0 < 16; ++__Vi0) {
            vlSelf->rnd03__DOT__mem[__Vi0] = VL_SCOPED_RAND_RESET_I(24, __VscopeHash, 3206042359269133279ull);
        }
        vlSelf->rnd03__DOT__w1 = VL_SCOPED_RAND_RESET_I(24, __VscopeHash, 12158361160283354487ull);
        vlSelf->rnd03__DOT__w3 = VL_SCOPED_RAND_RESET_I(24, __VscopeHash, 9824255813927239098ull);
        vlSelf->rnd03__DOT__w4 = VL_SCOPED_RAND_RESET_I(5, __VscopeHash, 15085981996569672358ull);
        VL_SCOPED_RAND_RESET_W(65, vlSelf->rnd03__DOT__w5, __VscopeHash, 2473363798540618616ull);
        vlSelf->rnd03__DOT__w6 = VL_SCOPED_RAND_RESET_I(9, __VscopeHash, 2392629202140256747ull);
        vlSelf->rnd03__DOT__w7 = VL_SCOPED_RAND_RESET_I(24, __VscopeHash, 7063168962096622264ull);
        vlSelf->rnd03__DOT__w8 = VL_SCOPED_RAND_RESET_Q(33, __VscopeHash, 10710339675531088524ull);
        vlSelf->rnd03__DOT__w10 = VL_SCOPED_RAND_RESET_I(24, __VscopeHash, 10757228718266317569ull);
        vlSelf->rnd03__DOT__w11 = VL_SCOPED_RAND_RESET_I(24, __VscopeHash, 4110676668828103217ull);
        vlSelf->__VdfgRegularize_h6e95ff9d_0_0 = 0;
        for (int __Vi0 = 0; __Vi0 < 1; ++__Vi0) {
            vlSelf->__VstlTriggered[__Vi0] = 0;
        }
        for (int __Vi0 = 0; __Vi0 < 2; ++__Vi0) {
            vlSelf->__VicoTriggered[__Vi0] = 0;
        }
        vlSelf->__Vtrigprevexpr___TOP__clk__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__i0__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__i1__0 = 0;
        VL_ZERO_RESET_W(130, vlSelf->__Vtrigprevexpr___TOP__i2__0);
        VL_ZERO_RESET_W(128, vlSelf->__Vtrigprevexpr___TOP__i3__0);
        VL_ZERO_RESET_W(96, vlSelf->__Vtrigprevexpr___TOP__i4__0);
        vlSelf->__Vtrigprevexpr___TOP__i5__0 = 0;
        vlSelf->__VicoDidInit = 0;
        for (int __Vi0 = 0; __Vi0 < 1; ++__Vi0) {
            vlSelf->__VactTriggered[__Vi0] = 0;
        }
        vlSelf->__Vtrigprevexpr___TOP__clk__1 = 0;
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
        Vrnd03___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vrnd03___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vrnd03___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vrnd03___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vrnd03___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vrnd03___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vrnd03__cuda::info() {
    typedef Vrnd03__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"i0", "i0", VLCU_MEMBER_IN, 8, 0, 1, 40, 39, 0, T::OFF_i0},
        {"i1", "i1", VLCU_MEMBER_IN, 1, 0, 1, 7, 6, 0, T::OFF_i1},
        {"i2", "i2", VLCU_MEMBER_IN, 4, 0, 5, 130, 129, 0, T::OFF_i2},
        {"i3", "i3", VLCU_MEMBER_IN, 4, 0, 4, 128, 127, 0, T::OFF_i3},
        {"i4", "i4", VLCU_MEMBER_IN, 4, 0, 3, 96, 95, 0, T::OFF_i4},
        {"i5", "i5", VLCU_MEMBER_IN, 8, 0, 1, 63, 62, 0, T::OFF_i5},
        {"o0", "o0", VLCU_MEMBER_OUT, 1, 0, 1, 1, 0, 0, T::OFF_o0},
        {"o1", "o1", VLCU_MEMBER_OUT, 2, 0, 1, 16, 15, 0, T::OFF_o1},
        {"o2", "o2", VLCU_MEMBER_OUT, 4, 0, 1, 32, 31, 0, T::OFF_o2},
        {"o3", "o3", VLCU_MEMBER_OUT, 8, 0, 1, 40, 39, 0, T::OFF_o3},
        {"o4", "o4", VLCU_MEMBER_OUT, 4, 0, 3, 65, 64, 0, T::OFF_o4},
        {"rnd03__DOT__r0", "rnd03.r0", VLCU_MEMBER_STATE, 1, 0, 1, 3, 2, 0, T::OFF_rnd03__DOT__r0},
        {"rnd03__DOT__r2", "rnd03.r2", VLCU_MEMBER_STATE, 1, 0, 1, 3, 2, 0, T::OFF_rnd03__DOT__r2},
        {"rnd03__DOT__r3", "rnd03.r3", VLCU_MEMBER_STATE, 1, 0, 1, 7, 6, 0, T::OFF_rnd03__DOT__r3},
        {"rnd03__DOT__w4", "rnd03.w4", VLCU_MEMBER_STATE, 1, 0, 1, 5, 4, 0, T::OFF_rnd03__DOT__w4},
        {"__VstlFirstIteration", "__VstlFirstIteration", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlFirstIteration},
        {"__VstlPhaseResult", "__VstlPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__Vtrigprevexpr___TOP__i1__0", "__Vtrigprevexpr___TOP__i1__0", VLCU_MEMBER_STATE, 1, 0, 1, 7, 6, 0, T::OFF___Vtrigprevexpr___TOP__i1__0},
        {"__VicoDidInit", "__VicoDidInit", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VicoDidInit},
        {"__VicoPhaseResult", "__VicoPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VicoPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__1", "__Vtrigprevexpr___TOP__clk__1", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__1},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"rnd03__DOT__w6", "rnd03.w6", VLCU_MEMBER_STATE, 2, 0, 1, 9, 8, 0, T::OFF_rnd03__DOT__w6},
        {"rnd03__DOT__r1", "rnd03.r1", VLCU_MEMBER_STATE, 4, 0, 3, 72, 71, 0, T::OFF_rnd03__DOT__r1},
        {"rnd03__DOT__w1", "rnd03.w1", VLCU_MEMBER_STATE, 4, 0, 1, 24, 23, 0, T::OFF_rnd03__DOT__w1},
        {"rnd03__DOT__w3", "rnd03.w3", VLCU_MEMBER_STATE, 4, 0, 1, 24, 23, 0, T::OFF_rnd03__DOT__w3},
        {"rnd03__DOT__w5", "rnd03.w5", VLCU_MEMBER_STATE, 4, 0, 3, 65, 64, 0, T::OFF_rnd03__DOT__w5},
        {"rnd03__DOT__w7", "rnd03.w7", VLCU_MEMBER_STATE, 4, 0, 1, 24, 23, 0, T::OFF_rnd03__DOT__w7},
        {"rnd03__DOT__w10", "rnd03.w10", VLCU_MEMBER_STATE, 4, 0, 1, 24, 23, 0, T::OFF_rnd03__DOT__w10},
        {"rnd03__DOT__w11", "rnd03.w11", VLCU_MEMBER_STATE, 4, 0, 1, 24, 23, 0, T::OFF_rnd03__DOT__w11},
        {"__Vtrigprevexpr___TOP__i2__0", "__Vtrigprevexpr___TOP__i2__0", VLCU_MEMBER_STATE, 4, 0, 5, 130, 129, 0, T::OFF___Vtrigprevexpr___TOP__i2__0},
        {"__Vtrigprevexpr___TOP__i3__0", "__Vtrigprevexpr___TOP__i3__0", VLCU_MEMBER_STATE, 4, 0, 4, 128, 127, 0, T::OFF___Vtrigprevexpr___TOP__i3__0},
        {"__Vtrigprevexpr___TOP__i4__0", "__Vtrigprevexpr___TOP__i4__0", VLCU_MEMBER_STATE, 4, 0, 3, 96, 95, 0, T::OFF___Vtrigprevexpr___TOP__i4__0},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"rnd03__DOT__w8", "rnd03.w8", VLCU_MEMBER_STATE, 8, 0, 1, 33, 32, 0, T::OFF_rnd03__DOT__w8},
        {"__VdfgRegularize_h6e95ff9d_0_0", "__VdfgRegularize_h6e95ff9d_0_0", VLCU_MEMBER_STATE, 8, 0, 1, 63, 62, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_0},
        {"__Vtrigprevexpr___TOP__i0__0", "__Vtrigprevexpr___TOP__i0__0", VLCU_MEMBER_STATE, 8, 0, 1, 40, 39, 0, T::OFF___Vtrigprevexpr___TOP__i0__0},
        {"__Vtrigprevexpr___TOP__i5__0", "__Vtrigprevexpr___TOP__i5__0", VLCU_MEMBER_STATE, 8, 0, 1, 63, 62, 0, T::OFF___Vtrigprevexpr___TOP__i5__0},
        {"rnd03__DOT__mem", "rnd03.mem", VLCU_MEMBER_STATE, 4, 0, 16, 24, 23, 0, T::OFF_rnd03__DOT__mem},
        {"__VstlTriggered", "__VstlTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VstlTriggered},
        {"__VicoTriggered", "__VicoTriggered", VLCU_MEMBER_STATE, 8, 0, 2, 64, 63, 0, T::OFF___VicoTriggered},
        {"__VactTriggered", "__VactTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VactTriggered},
        {"__VnbaTriggered", "__VnbaTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VnbaTriggered},
    };
    static const char* const messages[] = {nullptr};
    static const char* const plusargs[] = {nullptr};
    static const VlcuTraceDecl traceDecls[] = {
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0}};  // (terminator, not counted)
    static const uint32_t traceOrder[] = {0};  // codes in dump order (+ terminator)
    static const uint32_t saveOrder[] = {0, 2, 7, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 8, 25, 3, 4, 5, 9, 11, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 36, 1, 6, 10, 37, 38, 39, 40, 41, 42, 43, 44, 45, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "rnd03", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 0, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 46, 0, 0x16d69c6230dce010ULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vrnd03__cuda VLCU_TESTBENCHES);
