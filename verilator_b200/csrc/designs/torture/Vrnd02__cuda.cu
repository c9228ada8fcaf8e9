// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vrnd02"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vrnd02___024root final {
    CData/*0:0*/ clk;
    CData/*6:0*/ i0;
    QData/*63:0*/ i1;
    VlWide<3>/*95:0*/ i2;
    VlWide<3>/*71:0*/ i3;
    QData/*63:0*/ i4;
    SData/*8:0*/ i5;
    CData/*0:0*/ o0;
    SData/*15:0*/ o1;
    IData/*31:0*/ o2;
    QData/*39:0*/ o3;
    VlWide<3>/*95:0*/ o4;
    CData/*6:0*/ rnd02__DOT__w5;
    CData/*6:0*/ rnd02__DOT__w6;
    CData/*0:0*/ __VdfgRegularize_h6e95ff9d_0_1;
    CData/*0:0*/ __VstlFirstIteration;
    CData/*0:0*/ __VstlPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*6:0*/ __Vtrigprevexpr___TOP__i0__0;
    CData/*0:0*/ __VicoDidInit;
    CData/*0:0*/ __VicoPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__1;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    SData/*8:0*/ rnd02__DOT__r2;
    SData/*15:0*/ rnd02__DOT__w10;
    SData/*8:0*/ __Vtrigprevexpr___TOP__i5__0;
    IData/*23:0*/ rnd02__DOT__r0;
    IData/*31:0*/ rnd02__DOT__r3;
    VlWide<3>/*64:0*/ rnd02__DOT__w0;
    IData/*16:0*/ rnd02__DOT__w9;
    VlWide<4>/*102:0*/ __VdfgRegularize_h6e95ff9d_0_0;
    VlWide<3>/*95:0*/ __Vtrigprevexpr___TOP__i2__0;
    VlWide<3>/*71:0*/ __Vtrigprevexpr___TOP__i3__0;
    IData/*31:0*/ __VactIterCount;
    QData/*39:0*/ rnd02__DOT__r1;
    QData/*63:0*/ __Vtrigprevexpr___TOP__i1__0;
    QData/*63:0*/ __Vtrigprevexpr___TOP__i4__0;
    VlUnpacked<QData/*63:0*/, 1> __VstlTriggered;
    VlUnpacked<QData/*63:0*/, 2> __VicoTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vrnd02__cuda final {
    typedef Vrnd02___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 206,
        OFF_i0 = 207,
        OFF_i1 = 0,
        OFF_i2 = 88,
        OFF_i3 = 100,
        OFF_i4 = 8,
        OFF_i5 = 196,
        OFF_o0 = 208,
        OFF_o1 = 198,
        OFF_o2 = 112,
        OFF_o3 = 16,
        OFF_o4 = 116,
        OFF_rnd02__DOT__w5 = 209,
        OFF_rnd02__DOT__w6 = 210,
        OFF___VdfgRegularize_h6e95ff9d_0_1 = 211,
        OFF___VstlFirstIteration = 212,
        OFF___VstlPhaseResult = 213,
        OFF___Vtrigprevexpr___TOP__clk__0 = 214,
        OFF___Vtrigprevexpr___TOP__i0__0 = 215,
        OFF___VicoDidInit = 216,
        OFF___VicoPhaseResult = 217,
        OFF___Vtrigprevexpr___TOP__clk__1 = 218,
        OFF___VactPhaseResult = 219,
        OFF___VnbaPhaseResult = 220,
        OFF_rnd02__DOT__r2 = 200,
        OFF_rnd02__DOT__w10 = 202,
        OFF___Vtrigprevexpr___TOP__i5__0 = 204,
        OFF_rnd02__DOT__r0 = 128,
        OFF_rnd02__DOT__r3 = 132,
        OFF_rnd02__DOT__w0 = 136,
        OFF_rnd02__DOT__w9 = 148,
        OFF___VdfgRegularize_h6e95ff9d_0_0 = 152,
        OFF___Vtrigprevexpr___TOP__i2__0 = 168,
        OFF___Vtrigprevexpr___TOP__i3__0 = 180,
        OFF___VactIterCount = 192,
        OFF_rnd02__DOT__r1 = 24,
        OFF___Vtrigprevexpr___TOP__i1__0 = 32,
        OFF___Vtrigprevexpr___TOP__i4__0 = 40,
        OFF___VstlTriggered = 48,
        OFF___VicoTriggered = 56,
        OFF___VactTriggered = 72,
        OFF___VnbaTriggered = 80,
        kStateBytes = 224,  // S_persist
        kTouchedReadBytes = 184,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 167,  // S_state_wr
    };
    struct Locals final {
    };
    struct Shared final {
    };
    VLCU_D static void bind(Root& s, Locals& l, Shared& sh, const VlcuState& st, uint64_t i) {
        (void)l; (void)sh; (void)st; (void)i;
        s.vlNamep = "TOP";
    }
    VLCU_D static void load(Root& s, const VlcuState& st, uint64_t i) {
        s.clk = st.ld<VlcuMem8>(OFF_clk, i);
        s.i0 = st.ld<VlcuMem8>(OFF_i0, i);
        s.i1 = st.ld<QData>(OFF_i1, i);
        st.ldW(OFF_i2, i, s.i2);
        st.ldW(OFF_i3, i, s.i3);
        s.i4 = st.ld<QData>(OFF_i4, i);
        s.i5 = st.ld<VlcuMem16>(OFF_i5, i);
        s.o0 = st.ld<VlcuMem8>(OFF_o0, i);
        s.o1 = st.ld<VlcuMem16>(OFF_o1, i);
        s.o2 = st.ld<IData>(OFF_o2, i);
        s.o3 = st.ld<QData>(OFF_o3, i);
        st.ldW(OFF_o4, i, s.o4);
        s.rnd02__DOT__w5 = st.ld<VlcuMem8>(OFF_rnd02__DOT__w5, i);
        s.rnd02__DOT__w6 = st.ld<VlcuMem8>(OFF_rnd02__DOT__w6, i);
        s.__VdfgRegularize_h6e95ff9d_0_1 = st.ld<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_1, i);
        s.__VstlFirstIteration = st.ld<VlcuMem8>(OFF___VstlFirstIteration, i);
        s.__VstlPhaseResult = st.ld<VlcuMem8>(OFF___VstlPhaseResult, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__Vtrigprevexpr___TOP__i0__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__i0__0, i);
        s.__VicoDidInit = st.ld<VlcuMem8>(OFF___VicoDidInit, i);
        s.__VicoPhaseResult = st.ld<VlcuMem8>(OFF___VicoPhaseResult, i);
        s.__Vtrigprevexpr___TOP__clk__1 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__1, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        s.rnd02__DOT__r2 = st.ld<VlcuMem16>(OFF_rnd02__DOT__r2, i);
        s.rnd02__DOT__w10 = st.ld<VlcuMem16>(OFF_rnd02__DOT__w10, i);
        s.__Vtrigprevexpr___TOP__i5__0 = st.ld<VlcuMem16>(OFF___Vtrigprevexpr___TOP__i5__0, i);
        s.rnd02__DOT__r0 = st.ld<IData>(OFF_rnd02__DOT__r0, i);
        s.rnd02__DOT__r3 = st.ld<IData>(OFF_rnd02__DOT__r3, i);
        st.ldW(OFF_rnd02__DOT__w0, i, s.rnd02__DOT__w0);
        s.rnd02__DOT__w9 = st.ld<IData>(OFF_rnd02__DOT__w9, i);
        st.ldW(OFF___VdfgRegularize_h6e95ff9d_0_0, i, s.__VdfgRegularize_h6e95ff9d_0_0);
        st.ldW(OFF___Vtrigprevexpr___TOP__i2__0, i, s.__Vtrigprevexpr___TOP__i2__0);
        st.ldW(OFF___Vtrigprevexpr___TOP__i3__0, i, s.__Vtrigprevexpr___TOP__i3__0);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        s.rnd02__DOT__r1 = st.ld<QData>(OFF_rnd02__DOT__r1, i);
        s.__Vtrigprevexpr___TOP__i1__0 = st.ld<QData>(OFF___Vtrigprevexpr___TOP__i1__0, i);
        s.__Vtrigprevexpr___TOP__i4__0 = st.ld<QData>(OFF___Vtrigprevexpr___TOP__i4__0, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VstlTriggered[e] = st.ldE<QData>(OFF___VstlTriggered, e, i);
        for (uint32_t e = 0; e < 2; ++e) s.__VicoTriggered[e] = st.ldE<QData>(OFF___VicoTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.st<VlcuMem8>(OFF_i0, i, s.i0);
        st.st<QData>(OFF_i1, i, s.i1);
        st.stW(OFF_i2, i, s.i2);
        st.stW(OFF_i3, i, s.i3);
        st.st<QData>(OFF_i4, i, s.i4);
        st.st<VlcuMem16>(OFF_i5, i, s.i5);
        st.st<VlcuMem8>(OFF_o0, i, s.o0);
        st.st<VlcuMem16>(OFF_o1, i, s.o1);
        st.st<IData>(OFF_o2, i, s.o2);
        st.st<QData>(OFF_o3, i, s.o3);
        st.stW(OFF_o4, i, s.o4);
        st.st<VlcuMem8>(OFF_rnd02__DOT__w5, i, s.rnd02__DOT__w5);
        st.st<VlcuMem8>(OFF_rnd02__DOT__w6, i, s.rnd02__DOT__w6);
        st.st<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_1, i, s.__VdfgRegularize_h6e95ff9d_0_1);
        st.st<VlcuMem8>(OFF___VstlFirstIteration, i, s.__VstlFirstIteration);
        st.st<VlcuMem8>(OFF___VstlPhaseResult, i, s.__VstlPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__i0__0, i, s.__Vtrigprevexpr___TOP__i0__0);
        st.st<VlcuMem8>(OFF___VicoDidInit, i, s.__VicoDidInit);
        st.st<VlcuMem8>(OFF___VicoPhaseResult, i, s.__VicoPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__1, i, s.__Vtrigprevexpr___TOP__clk__1);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.st<VlcuMem16>(OFF_rnd02__DOT__r2, i, s.rnd02__DOT__r2);
        st.st<VlcuMem16>(OFF_rnd02__DOT__w10, i, s.rnd02__DOT__w10);
        st.st<VlcuMem16>(OFF___Vtrigprevexpr___TOP__i5__0, i, s.__Vtrigprevexpr___TOP__i5__0);
        st.st<IData>(OFF_rnd02__DOT__r0, i, s.rnd02__DOT__r0);
        st.st<IData>(OFF_rnd02__DOT__r3, i, s.rnd02__DOT__r3);
        st.stW(OFF_rnd02__DOT__w0, i, s.rnd02__DOT__w0);
        st.st<IData>(OFF_rnd02__DOT__w9, i, s.rnd02__DOT__w9);
        st.stW(OFF___VdfgRegularize_h6e95ff9d_0_0, i, s.__VdfgRegularize_h6e95ff9d_0_0);
        st.stW(OFF___Vtrigprevexpr___TOP__i2__0, i, s.__Vtrigprevexpr___TOP__i2__0);
        st.stW(OFF___Vtrigprevexpr___TOP__i3__0, i, s.__Vtrigprevexpr___TOP__i3__0);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
        st.st<QData>(OFF_rnd02__DOT__r1, i, s.rnd02__DOT__r1);
        st.st<QData>(OFF___Vtrigprevexpr___TOP__i1__0, i, s.__Vtrigprevexpr___TOP__i1__0);
        st.st<QData>(OFF___Vtrigprevexpr___TOP__i4__0, i, s.__Vtrigprevexpr___TOP__i4__0);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VstlTriggered, e, i, s.__VstlTriggered[e]);
        for (uint32_t e = 0; e < 2; ++e) st.stE<QData>(OFF___VicoTriggered, e, i, s.__VicoTriggered[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VactTriggered, e, i, s.__VactTriggered[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VnbaTriggered, e, i, s.__VnbaTriggered[e]);
    }
    VLCU_D static void load_inputs_row(Root& s, const uint32_t* rowp, uint64_t stride, uint64_t i) {
        s.clk = static_cast<VlcuMem8>(rowp[0 * stride + i] & 0x1U);
        s.i0 = static_cast<VlcuMem8>(rowp[1 * stride + i] & 0x7fU);
        s.i1 = (static_cast<QData>(rowp[3 * stride + i] & 0xffffffffU) << 32) | rowp[2 * stride + i];
        s.i2[0] = rowp[4 * stride + i] & 0xffffffffU;
        s.i2[1] = rowp[5 * stride + i] & 0xffffffffU;
        s.i2[2] = rowp[6 * stride + i] & 0xffffffffU;
        s.i3[0] = rowp[7 * stride + i] & 0xffffffffU;
        s.i3[1] = rowp[8 * stride + i] & 0xffffffffU;
        s.i3[2] = rowp[9 * stride + i] & 0xffU;
        s.i4 = (static_cast<QData>(rowp[11 * stride + i] & 0xffffffffU) << 32) | rowp[10 * stride + i];
        s.i5 = static_cast<VlcuMem16>(rowp[12 * stride + i] & 0x1ffU);
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
    enum : uint32_t { kInWords = 13, kOutWords = 8, kInBytes = 44, kOutBytes = 27, kTraceWords = 0 };

    VLCU_D static void Vrnd02___024root___eval_static(Vrnd02___024root* vlSelf, VlcuContext& vlCtx) {
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
        vlSelfRef.__Vtrigprevexpr___TOP__i3__0[0U] 
            = vlSelfRef.i3[0U];
        vlSelfRef.__Vtrigprevexpr___TOP__i3__0[1U] 
            = vlSelfRef.i3[1U];
        vlSelfRef.__Vtrigprevexpr___TOP__i3__0[2U] 
            = vlSelfRef.i3[2U];
        vlSelfRef.__Vtrigprevexpr___TOP__i4__0 = vlSelfRef.i4;
        vlSelfRef.__Vtrigprevexpr___TOP__i5__0 = vlSelfRef.i5;
        vlSelfRef.__Vtrigprevexpr___TOP__clk__1 = vlSelfRef.clk;
    }

    VLCU_D static void Vrnd02___024root___eval_initial(Vrnd02___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vrnd02___024root___eval_final(Vrnd02___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vrnd02___024root___eval_settle(Vrnd02___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vrnd02___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/rnd02.v", 2, "", "DIDNOTCONVERGE: Settle region did not converge after '--converge-limit' of 10000 tries");
            }
            __VstlIterCount = ((IData)(1U) + __VstlIterCount);
            vlSelfRef.__VstlPhaseResult = Vrnd02___024root___eval_phase__stl(vlSelf, vlCtx);
            vlSelfRef.__VstlFirstIteration = 0U;
        } while (vlSelfRef.__VstlPhaseResult);
    }

    VLCU_D static bool Vrnd02___024root___trigger_anySet__stl(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vrnd02___024root___stl_sequent__TOP__0(Vrnd02___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        VlWide<3>/*95:0*/ __Vtemp_2;
        VlWide<3>/*95:0*/ __Vtemp_3;
        VlWide<3>/*95:0*/ __Vtemp_5;
        VlWide<3>/*95:0*/ __Vtemp_6;
        VlWide<3>/*95:0*/ __Vtemp_7;
        VlWide<3>/*95:0*/ __Vtemp_8;
        VlWide<3>/*95:0*/ __Vtemp_9;
        VlWide<3>/*95:0*/ __Vtemp_12;
        VlWide<3>/*95:0*/ __Vtemp_16;
        VlWide<3>/*95:0*/ __Vtemp_18;
        // Body
        vlSelfRef.o2 = (IData)((0x000000ffffffffffULL 
                                & VL_MODDIV_QQQ(40, (QData)((IData)(
                                                                    (0x0000007fU 
                                                                     & (IData)(
                                                                               (0x000000ffffffffffULL 
                                                                                & VL_DIVS_QQQ(40, 
                                                                                (0x000000ffffffffffULL 
                                                                                & VL_EXTENDS_QI(40,24, vlSelfRef.rnd02__DOT__r0)), 0ULL)))))), 0ULL)));
        vlSelfRef.rnd02__DOT__w6 = (0x0000007fU & (IData)(
                                                          (0x000000ffffffffffULL 
                                                           & VL_DIVS_QQQ(40, 
                                                                         (0x000000ffffffffffULL 
                                                                          & VL_EXTENDS_QI(40,24, vlSelfRef.rnd02__DOT__r0)), 0ULL))));
        __Vtemp_2[0U] = (0x01ffffffU & (IData)((vlSelfRef.i4 
                                                >> 0x0000000dU)));
        __Vtemp_2[1U] = 0U;
        __Vtemp_2[2U] = 0U;
        __Vtemp_3[0U] = 0U;
        __Vtemp_3[1U] = 0U;
        __Vtemp_3[2U] = 0U;
        __Vtemp_5[0U] = vlSelfRef.rnd02__DOT__r3;
        __Vtemp_5[1U] = 0U;
        __Vtemp_5[2U] = 0U;
        VL_SUB_W(3, __Vtemp_6, vlSelfRef.i3, __Vtemp_5);
        __Vtemp_7[0U] = __Vtemp_6[0U];
        __Vtemp_7[1U] = __Vtemp_6[1U];
        __Vtemp_7[2U] = (0x000000ffU & __Vtemp_6[2U]);
        __Vtemp_8[0U] = VL_LTS_IWW(72, __Vtemp_3, __Vtemp_7);
        __Vtemp_8[1U] = 0U;
        __Vtemp_8[2U] = 0U;
        VL_SUB_W(3, __Vtemp_9, __Vtemp_2, __Vtemp_8);
        vlSelfRef.rnd02__DOT__w0[0U] = __Vtemp_9[0U];
        vlSelfRef.rnd02__DOT__w0[1U] = __Vtemp_9[1U];
        vlSelfRef.rnd02__DOT__w0[2U] = (1U & __Vtemp_9[2U]);
        if ((0U != (0x000000ffffffffffULL & VL_SHIFTRS_QQI(40,40,8, 
                                                           (0x000000ffffffffffULL 
                                                            & VL_DIV_QQQ(40, (QData)((IData)(vlSelfRef.i5)), vlSelfRef.rnd02__DOT__r1)), 
                                                           (0x000000ffU 
                                                            & vlSelfRef.rnd02__DOT__r3))))) {
            vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[0U] 
                = (vlSelfRef.i2[0U] ^ (IData)(vlSelfRef.i5));
            vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[1U] 
                = vlSelfRef.i2[1U];
            vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[2U] 
                = vlSelfRef.i2[2U];
            vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[3U] 
                = vlSelfRef.i0;
        } else {
            vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[0U] 
                = (IData)(vlSelfRef.i1);
            vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[1U] 
                = (IData)((vlSelfRef.i1 >> 0x00000020U));
            vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[2U] = 0U;
            vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[3U] = 0U;
        }
        vlSelfRef.rnd02__DOT__w9 = (0x0001ffffU & VL_SHIFTL_III(17,17,8, vlSelfRef.rnd02__DOT__w0[0U], 
                                                                (0x000000ffU 
                                                                 & (IData)(vlSelfRef.i4))));
        VL_EXTENDS_WI(96,24, __Vtemp_12, (0x00ffffffU 
                                          & (~ (0x00000040U 
                                                | (1U 
                                                   != (IData)(vlSelfRef.rnd02__DOT__r2))))));
        VL_SHIFTRS_WWI(96,96,8, vlSelfRef.o4, __Vtemp_12, 
                       (0x000000ffU & vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[0U]));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1 = 
            (VL_SHIFTR_QQI(49,49,8, (QData)((IData)(
                                                    ((0ULL 
                                                      != 
                                                      (0x00000000000ff000ULL 
                                                       & vlSelfRef.i4)) 
                                                     | (0U 
                                                        != 
                                                        ((vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[0U] 
                                                          | vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[1U]) 
                                                         | (0x000000ffU 
                                                            & vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[2U])))))), 
                           (0x000000ffU & vlSelfRef.rnd02__DOT__r0)) 
             >= ((0x0001fffffffffe00ULL & (VL_DIVS_QQQ(40, 
                                                       (0x000000ffffffffffULL 
                                                        & VL_EXTENDS_QI(40,9, (IData)(vlSelfRef.i5))), vlSelfRef.rnd02__DOT__r1) 
                                           << 9U)) 
                 | (QData)((IData)(vlSelfRef.i5))));
        VL_EXTENDS_WI(96,24, __Vtemp_16, vlSelfRef.rnd02__DOT__r0);
        vlSelfRef.o3 = VL_SHIFTR_QQI(40,40,8, (((QData)((IData)(
                                                                ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1) 
                                                                 - vlSelfRef.rnd02__DOT__r3))) 
                                                << 1U) 
                                               | (QData)((IData)(
                                                                 VL_LTS_IWW(96, vlSelfRef.i2, __Vtemp_16)))), 
                                     (0x000000ffU & (IData)(vlSelfRef.rnd02__DOT__r2)));
        vlSelfRef.rnd02__DOT__w5 = (0x0000007fU & (IData)(
                                                          VL_MODDIVS_QQQ(64, vlSelfRef.i1, 
                                                                         VL_EXTENDS_QI(64,8, (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1)))));
        __Vtemp_18[0U] = (0x2f6f6b26U | (1U & (IData)(
                                                      (vlSelfRef.rnd02__DOT__r1 
                                                       >> (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1)))));
        __Vtemp_18[1U] = 0x2e00a8aeU;
        __Vtemp_18[2U] = 0x000000e6U;
        vlSelfRef.rnd02__DOT__w10 = (0x0000ffffU & 
                                     (((IData)(1U) 
                                       + VL_ONEHOT_W(3, __Vtemp_18)) 
                                      - VL_SHIFTR_III(16,32,8, 
                                                      (1U 
                                                       & ((IData)(vlSelfRef.i5) 
                                                          >> 
                                                          (7U 
                                                           & vlSelfRef.rnd02__DOT__w0[0U]))), 
                                                      (0x000000ffU 
                                                       & vlSelfRef.rnd02__DOT__r0))));
        vlSelfRef.o0 = (1U & VL_DIV_III(16, (IData)(vlSelfRef.rnd02__DOT__w5), 
                                        VL_SHIFTR_III(16,16,8, 
                                                      ((0U 
                                                        != 
                                                        (0xffdd7dfeU 
                                                         | ((~ (IData)(
                                                                       (0xee181ce8a7a97d65ULL 
                                                                        ^ vlSelfRef.i1))) 
                                                            | (~ (IData)(
                                                                         ((0xee181ce8a7a97d65ULL 
                                                                           ^ vlSelfRef.i1) 
                                                                          >> 0x00000020U)))))) 
                                                       & (0U 
                                                          != 
                                                          ((vlSelfRef.rnd02__DOT__w0[0U] 
                                                            | vlSelfRef.rnd02__DOT__w0[1U]) 
                                                           | vlSelfRef.rnd02__DOT__w0[2U]))), 
                                                      (0x000000ffU 
                                                       & vlSelfRef.i3[0U]))));
        vlSelfRef.o1 = (0x0000ffffU & (IData)(VL_DIVS_QQQ(64, 
                                                          VL_EXTENDS_QI(64,16, (IData)(vlSelfRef.rnd02__DOT__w10)), vlSelfRef.i1)));
    }

    VLCU_D static bool Vrnd02___024root___eval_phase__stl(Vrnd02___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vrnd02___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
        }
#endif
        __VstlExecute = Vrnd02___024root___trigger_anySet__stl(vlSelfRef.__VstlTriggered);
        if (__VstlExecute) {
            {
                // Inlined CFunc: _eval_stl
                if ((1ULL & vlSelfRef.__VstlTriggered[0U])) {
                    Vrnd02___024root___stl_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
        }
        return (__VstlExecute);
    }

    VLCU_D static void Vrnd02___024root___eval_triggers_vec__ico(Vrnd02___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlSelfRef.__VicoTriggered[0U] = (QData)((IData)(
                                                        (((((IData)(vlSelfRef.i5) 
                                                            != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__i5__0)) 
                                                           << 6U) 
                                                          | (((vlSelfRef.i4 
                                                               != vlSelfRef.__Vtrigprevexpr___TOP__i4__0) 
                                                              << 5U) 
                                                             | ((0U 
                                                                 != 
                                                                 (((vlSelfRef.i3[0U] 
                                                                    ^ vlSelfRef.__Vtrigprevexpr___TOP__i3__0[0U]) 
                                                                   | (vlSelfRef.i3[1U] 
                                                                      ^ vlSelfRef.__Vtrigprevexpr___TOP__i3__0[1U])) 
                                                                  | (vlSelfRef.i3[2U] 
                                                                     ^ vlSelfRef.__Vtrigprevexpr___TOP__i3__0[2U]))) 
                                                                << 4U))) 
                                                         | ((((0U 
                                                               != 
                                                               (((vlSelfRef.i2[0U] 
                                                                  ^ vlSelfRef.__Vtrigprevexpr___TOP__i2__0[0U]) 
                                                                 | (vlSelfRef.i2[1U] 
                                                                    ^ vlSelfRef.__Vtrigprevexpr___TOP__i2__0[1U])) 
                                                                | (vlSelfRef.i2[2U] 
                                                                   ^ vlSelfRef.__Vtrigprevexpr___TOP__i2__0[2U]))) 
                                                              << 3U) 
                                                             | ((vlSelfRef.i1 
                                                                 != vlSelfRef.__Vtrigprevexpr___TOP__i1__0) 
                                                                << 2U)) 
                                                            | ((((IData)(vlSelfRef.i0) 
                                                                 != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__i0__0)) 
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
        vlSelfRef.__Vtrigprevexpr___TOP__i3__0[0U] 
            = vlSelfRef.i3[0U];
        vlSelfRef.__Vtrigprevexpr___TOP__i3__0[1U] 
            = vlSelfRef.i3[1U];
        vlSelfRef.__Vtrigprevexpr___TOP__i3__0[2U] 
            = vlSelfRef.i3[2U];
        vlSelfRef.__Vtrigprevexpr___TOP__i4__0 = vlSelfRef.i4;
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

    VLCU_D static bool Vrnd02___024root___trigger_anySet__ico(const VlUnpacked<QData/*63:0*/, 2> &in) {
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

    VLCU_D static void Vrnd02___024root___eval_ico(Vrnd02___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        if ((0x0000000000000030ULL & vlSelfRef.__VicoTriggered[0U])) {
            {
                // Inlined CFunc: _ico_comb__TOP__0
                VlWide<3>/*95:0*/ __Vinline_0__ico_comb__TOP__0___Vtemp_2;
                VlWide<3>/*95:0*/ __Vinline_0__ico_comb__TOP__0___Vtemp_3;
                VlWide<3>/*95:0*/ __Vinline_0__ico_comb__TOP__0___Vtemp_5;
                VlWide<3>/*95:0*/ __Vinline_0__ico_comb__TOP__0___Vtemp_6;
                VlWide<3>/*95:0*/ __Vinline_0__ico_comb__TOP__0___Vtemp_7;
                VlWide<3>/*95:0*/ __Vinline_0__ico_comb__TOP__0___Vtemp_8;
                VlWide<3>/*95:0*/ __Vinline_0__ico_comb__TOP__0___Vtemp_9;
                __Vinline_0__ico_comb__TOP__0___Vtemp_2[0U] 
                    = (0x01ffffffU & (IData)((vlSelfRef.i4 
                                              >> 0x0000000dU)));
                __Vinline_0__ico_comb__TOP__0___Vtemp_2[1U] = 0U;
                __Vinline_0__ico_comb__TOP__0___Vtemp_2[2U] = 0U;
                __Vinline_0__ico_comb__TOP__0___Vtemp_3[0U] = 0U;
                __Vinline_0__ico_comb__TOP__0___Vtemp_3[1U] = 0U;
                __Vinline_0__ico_comb__TOP__0___Vtemp_3[2U] = 0U;
                __Vinline_0__ico_comb__TOP__0___Vtemp_5[0U] 
                    = vlSelfRef.rnd02__DOT__r3;
                __Vinline_0__ico_comb__TOP__0___Vtemp_5[1U] = 0U;
                __Vinline_0__ico_comb__TOP__0___Vtemp_5[2U] = 0U;
                VL_SUB_W(3, __Vinline_0__ico_comb__TOP__0___Vtemp_6, vlSelfRef.i3, __Vinline_0__ico_comb__TOP__0___Vtemp_5);
                __Vinline_0__ico_comb__TOP__0___Vtemp_7[0U] 
                    = __Vinline_0__ico_comb__TOP__0___Vtemp_6[0U];
                __Vinline_0__ico_comb__TOP__0___Vtemp_7[1U] 
                    = __Vinline_0__ico_comb__TOP__0___Vtemp_6[1U];
                __Vinline_0__ico_comb__TOP__0___Vtemp_7[2U] 
                    = (0x000000ffU & __Vinline_0__ico_comb__TOP__0___Vtemp_6[2U]);
                __Vinline_0__ico_comb__TOP__0___Vtemp_8[0U] 
                    = VL_LTS_IWW(72, __Vinline_0__ico_comb__TOP__0___Vtemp_3, __Vinline_0__ico_comb__TOP__0___Vtemp_7);
                __Vinline_0__ico_comb__TOP__0___Vtemp_8[1U] = 0U;
                __Vinline_0__ico_comb__TOP__0___Vtemp_8[2U] = 0U;
                VL_SUB_W(3, __Vinline_0__ico_comb__TOP__0___Vtemp_9, __Vinline_0__ico_comb__TOP__0___Vtemp_2, __Vinline_0__ico_comb__TOP__0___Vtemp_8);
                vlSelfRef.rnd02__DOT__w0[0U] = __Vinline_0__ico_comb__TOP__0___Vtemp_9[0U];
                vlSelfRef.rnd02__DOT__w0[1U] = __Vinline_0__ico_comb__TOP__0___Vtemp_9[1U];
                vlSelfRef.rnd02__DOT__w0[2U] = (1U 
                                                & __Vinline_0__ico_comb__TOP__0___Vtemp_9[2U]);
                vlSelfRef.rnd02__DOT__w9 = (0x0001ffffU 
                                            & VL_SHIFTL_III(17,17,8, vlSelfRef.rnd02__DOT__w0[0U], 
                                                            (0x000000ffU 
                                                             & (IData)(vlSelfRef.i4))));
            }
        }
        if ((0x000000000000004eULL & vlSelfRef.__VicoTriggered[0U])) {
            {
                // Inlined CFunc: _ico_comb__TOP__1
                VlWide<3>/*95:0*/ __Vinline_0__ico_comb__TOP__1___Vtemp_3;
                if ((0U != (0x000000ffffffffffULL & 
                            VL_SHIFTRS_QQI(40,40,8, 
                                           (0x000000ffffffffffULL 
                                            & VL_DIV_QQQ(40, (QData)((IData)(vlSelfRef.i5)), vlSelfRef.rnd02__DOT__r1)), 
                                           (0x000000ffU 
                                            & vlSelfRef.rnd02__DOT__r3))))) {
                    vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[0U] 
                        = (vlSelfRef.i2[0U] ^ (IData)(vlSelfRef.i5));
                    vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[1U] 
                        = vlSelfRef.i2[1U];
                    vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[2U] 
                        = vlSelfRef.i2[2U];
                    vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[3U] 
                        = vlSelfRef.i0;
                } else {
                    vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[0U] 
                        = (IData)(vlSelfRef.i1);
                    vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[1U] 
                        = (IData)((vlSelfRef.i1 >> 0x00000020U));
                    vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[2U] = 0U;
                    vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[3U] = 0U;
                }
                VL_EXTENDS_WI(96,24, __Vinline_0__ico_comb__TOP__1___Vtemp_3, 
                              (0x00ffffffU & (~ (0x00000040U 
                                                 | (1U 
                                                    != (IData)(vlSelfRef.rnd02__DOT__r2))))));
                VL_SHIFTRS_WWI(96,96,8, vlSelfRef.o4, __Vinline_0__ico_comb__TOP__1___Vtemp_3, 
                               (0x000000ffU & vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[0U]));
            }
        }
        if ((0x000000000000006eULL & vlSelfRef.__VicoTriggered[0U])) {
            {
                // Inlined CFunc: _ico_comb__TOP__2
                VlWide<3>/*95:0*/ __Vinline_0__ico_comb__TOP__2___Vtemp_4;
                vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1 
                    = (VL_SHIFTR_QQI(49,49,8, (QData)((IData)(
                                                              ((0ULL 
                                                                != 
                                                                (0x00000000000ff000ULL 
                                                                 & vlSelfRef.i4)) 
                                                               | (0U 
                                                                  != 
                                                                  ((vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[0U] 
                                                                    | vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[1U]) 
                                                                   | (0x000000ffU 
                                                                      & vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[2U])))))), 
                                     (0x000000ffU & vlSelfRef.rnd02__DOT__r0)) 
                       >= ((0x0001fffffffffe00ULL & 
                            (VL_DIVS_QQQ(40, (0x000000ffffffffffULL 
                                              & VL_EXTENDS_QI(40,9, (IData)(vlSelfRef.i5))), vlSelfRef.rnd02__DOT__r1) 
                             << 9U)) | (QData)((IData)(vlSelfRef.i5))));
                VL_EXTENDS_WI(96,24, __Vinline_0__ico_comb__TOP__2___Vtemp_4, vlSelfRef.rnd02__DOT__r0);
                vlSelfRef.o3 = VL_SHIFTR_QQI(40,40,8, 
                                             (((QData)((IData)(
                                                               ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1) 
                                                                - vlSelfRef.rnd02__DOT__r3))) 
                                               << 1U) 
                                              | (QData)((IData)(
                                                                VL_LTS_IWW(96, vlSelfRef.i2, __Vinline_0__ico_comb__TOP__2___Vtemp_4)))), 
                                             (0x000000ffU 
                                              & (IData)(vlSelfRef.rnd02__DOT__r2)));
                vlSelfRef.rnd02__DOT__w5 = (0x0000007fU 
                                            & (IData)(
                                                      VL_MODDIVS_QQQ(64, vlSelfRef.i1, 
                                                                     VL_EXTENDS_QI(64,8, (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1)))));
            }
        }
        if ((0x000000000000007eULL & vlSelfRef.__VicoTriggered[0U])) {
            {
                // Inlined CFunc: _ico_comb__TOP__3
                VlWide<3>/*95:0*/ __Vinline_0__ico_comb__TOP__3___Vtemp_2;
                __Vinline_0__ico_comb__TOP__3___Vtemp_2[0U] 
                    = (0x2f6f6b26U | (1U & (IData)(
                                                   (vlSelfRef.rnd02__DOT__r1 
                                                    >> (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1)))));
                __Vinline_0__ico_comb__TOP__3___Vtemp_2[1U] = 0x2e00a8aeU;
                __Vinline_0__ico_comb__TOP__3___Vtemp_2[2U] = 0x000000e6U;
                vlSelfRef.rnd02__DOT__w10 = (0x0000ffffU 
                                             & (((IData)(1U) 
                                                 + 
                                                 VL_ONEHOT_W(3, __Vinline_0__ico_comb__TOP__3___Vtemp_2)) 
                                                - VL_SHIFTR_III(16,32,8, 
                                                                (1U 
                                                                 & ((IData)(vlSelfRef.i5) 
                                                                    >> 
                                                                    (7U 
                                                                     & vlSelfRef.rnd02__DOT__w0[0U]))), 
                                                                (0x000000ffU 
                                                                 & vlSelfRef.rnd02__DOT__r0))));
                vlSelfRef.o0 = (1U & VL_DIV_III(16, (IData)(vlSelfRef.rnd02__DOT__w5), 
                                                VL_SHIFTR_III(16,16,8, 
                                                              ((0U 
                                                                != 
                                                                (0xffdd7dfeU 
                                                                 | ((~ (IData)(
                                                                               (0xee181ce8a7a97d65ULL 
                                                                                ^ vlSelfRef.i1))) 
                                                                    | (~ (IData)(
                                                                                ((0xee181ce8a7a97d65ULL 
                                                                                ^ vlSelfRef.i1) 
                                                                                >> 0x00000020U)))))) 
                                                               & (0U 
                                                                  != 
                                                                  ((vlSelfRef.rnd02__DOT__w0[0U] 
                                                                    | vlSelfRef.rnd02__DOT__w0[1U]) 
                                                                   | vlSelfRef.rnd02__DOT__w0[2U]))), 
                                                              (0x000000ffU 
                                                               & vlSelfRef.i3[0U]))));
                vlSelfRef.o1 = (0x0000ffffU & (IData)(
                                                      VL_DIVS_QQQ(64, 
                                                                  VL_EXTENDS_QI(64,16, (IData)(vlSelfRef.rnd02__DOT__w10)), vlSelfRef.i1)));
            }
        }
    }

    VLCU_D static bool Vrnd02___024root___eval_phase__ico(Vrnd02___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VicoExecute;
        // Body
        Vrnd02___024root___eval_triggers_vec__ico(vlSelf, vlCtx);
#ifdef VL_DEBUG
        if (VL_UNLIKELY(vlSymsp->_vm_contextp__->debug())) {
            Vrnd02___024root___dump_triggers__ico(vlSelfRef.__VicoTriggered, "ico"s);
        }
#endif
        __VicoExecute = Vrnd02___024root___trigger_anySet__ico(vlSelfRef.__VicoTriggered);
        if (__VicoExecute) {
            Vrnd02___024root___eval_ico(vlSelf, vlCtx);
        }
        return (__VicoExecute);
    }

    VLCU_D static bool Vrnd02___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vrnd02___024root___nba_sequent__TOP__0(Vrnd02___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        SData/*8:0*/ __Vdly__rnd02__DOT__r2;
        __Vdly__rnd02__DOT__r2 = 0;
        VlWide<7>/*223:0*/ __Vtemp_2;
        VlWide<7>/*223:0*/ __Vtemp_3;
        VlWide<7>/*223:0*/ __Vtemp_4;
        VlWide<3>/*95:0*/ __Vtemp_14;
        VlWide<3>/*95:0*/ __Vtemp_18;
        VlWide<3>/*95:0*/ __Vtemp_19;
        VlWide<5>/*159:0*/ __Vtemp_20;
        VlWide<5>/*159:0*/ __Vtemp_21;
        VlWide<5>/*159:0*/ __Vtemp_22;
        VlWide<3>/*95:0*/ __Vtemp_25;
        VlWide<3>/*95:0*/ __Vtemp_26;
        VlWide<3>/*95:0*/ __Vtemp_28;
        VlWide<3>/*95:0*/ __Vtemp_29;
        VlWide<3>/*95:0*/ __Vtemp_30;
        VlWide<3>/*95:0*/ __Vtemp_31;
        VlWide<3>/*95:0*/ __Vtemp_32;
        VlWide<3>/*95:0*/ __Vtemp_35;
        VlWide<3>/*95:0*/ __Vtemp_39;
        VlWide<3>/*95:0*/ __Vtemp_41;
        // Body
        __Vdly__rnd02__DOT__r2 = vlSelfRef.rnd02__DOT__r2;
        __Vtemp_2[0U] = vlSelfRef.rnd02__DOT__r0;
        __Vtemp_2[1U] = 0U;
        __Vtemp_2[2U] = 0U;
        __Vtemp_2[3U] = 0U;
        __Vtemp_2[4U] = 0U;
        __Vtemp_2[5U] = 0U;
        __Vtemp_2[6U] = 0U;
        __Vtemp_3[0U] = vlSelfRef.i0;
        __Vtemp_3[1U] = 0U;
        __Vtemp_3[2U] = 0U;
        __Vtemp_3[3U] = 0U;
        __Vtemp_3[4U] = 0U;
        __Vtemp_3[5U] = 0U;
        __Vtemp_3[6U] = 0U;
        VL_MODDIV_WWW(216, __Vtemp_4, __Vtemp_2, __Vtemp_3);
        __Vdly__rnd02__DOT__r2 = ((0x0000007fU & (IData)(
                                                         (0x000000ffffffffffULL 
                                                          & VL_DIVS_QQQ(40, 
                                                                        (0x000000ffffffffffULL 
                                                                         & VL_EXTENDS_QI(40,24, vlSelfRef.rnd02__DOT__r0)), 0ULL)))) 
                                  >= (VL_COUNTONES_Q(
                                                     (0x000000ffffffffffULL 
                                                      & VL_SHIFTRS_QQI(40,40,8, vlSelfRef.rnd02__DOT__r1, 
                                                                       (0x000000ffU 
                                                                        & vlSelfRef.rnd02__DOT__w9)))) 
                                      + VL_ONEHOT_I(
                                                    ((IData)(vlSelfRef.i0) 
                                                     | (1U 
                                                        & ((IData)(vlSelfRef.rnd02__DOT__w5) 
                                                           >> 
                                                           ((0U 
                                                             != 
                                                             VL_DIV_QQQ(64, (QData)((IData)(vlSelfRef.i0)), vlSelfRef.i1)) 
                                                            & (0U 
                                                               != 
                                                               (((((((~ 
                                                                      (__Vtemp_4[0U] 
                                                                       ^ vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[0U])) 
                                                                     | (~ 
                                                                        (__Vtemp_4[1U] 
                                                                         ^ vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[1U]))) 
                                                                    | (~ 
                                                                       (__Vtemp_4[2U] 
                                                                        ^ 
                                                                        ((vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[0U] 
                                                                          << 8U) 
                                                                         | (0x000000ffU 
                                                                            & vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[2U]))))) 
                                                                   | (~ 
                                                                      (__Vtemp_4[3U] 
                                                                       ^ 
                                                                       ((vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[0U] 
                                                                         >> 0x00000018U) 
                                                                        | (vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[1U] 
                                                                           << 8U))))) 
                                                                  | (~ 
                                                                     (__Vtemp_4[4U] 
                                                                      ^ 
                                                                      ((vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[0U] 
                                                                        << 0x00000010U) 
                                                                       | ((vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[1U] 
                                                                           >> 0x00000018U) 
                                                                          | (0x0000ff00U 
                                                                             & (vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[2U] 
                                                                                << 8U))))))) 
                                                                 | (~ 
                                                                    (__Vtemp_4[5U] 
                                                                     ^ 
                                                                     ((vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[0U] 
                                                                       >> 0x00000010U) 
                                                                      | (vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[1U] 
                                                                         << 0x00000010U))))) 
                                                                | (0x00ffffffU 
                                                                   & (~ 
                                                                      (__Vtemp_4[6U] 
                                                                       ^ 
                                                                       ((vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[1U] 
                                                                         >> 0x00000010U) 
                                                                        | (0x00ff0000U 
                                                                           & (vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[2U] 
                                                                              << 0x00000010U)))))))))))))));
        vlSelfRef.rnd02__DOT__r3 = (IData)(VL_MODDIV_QQQ(64, vlSelfRef.i4, (QData)((IData)(vlSelfRef.rnd02__DOT__w9))));
        vlSelfRef.rnd02__DOT__r0 = (((IData)(vlSelfRef.rnd02__DOT__w10) 
                                     << 2U) | (3U & (IData)(vlSelfRef.rnd02__DOT__w6)));
        if ((0U != vlSelfRef.rnd02__DOT__w9)) {
            __Vtemp_14[0U] = vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1;
            __Vtemp_14[1U] = 0U;
            __Vtemp_14[2U] = 0U;
        } else {
            __Vtemp_14[0U] = vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[0U];
            __Vtemp_14[1U] = vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[1U];
            __Vtemp_14[2U] = vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[2U];
        }
        __Vtemp_18[0U] = vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[0U];
        __Vtemp_18[1U] = vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[1U];
        __Vtemp_18[2U] = (0x000000ffU & vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[2U]);
        VL_SHIFTR_WWI(72,72,8, __Vtemp_19, __Vtemp_18, 
                      (0x000000ffU & (IData)(vlSelfRef.rnd02__DOT__r2)));
        __Vtemp_20[0U] = __Vtemp_19[0U];
        __Vtemp_20[1U] = __Vtemp_19[1U];
        __Vtemp_20[2U] = ((__Vtemp_14[0U] << 8U) | __Vtemp_19[2U]);
        __Vtemp_20[3U] = ((__Vtemp_14[0U] >> 0x00000018U) 
                          | (__Vtemp_14[1U] << 8U));
        __Vtemp_20[4U] = ((__Vtemp_14[1U] >> 0x00000018U) 
                          | (0x0000ff00U & (__Vtemp_14[2U] 
                                            << 8U)));
        __Vtemp_21[0U] = (1U & (IData)(vlSelfRef.rnd02__DOT__r2));
        __Vtemp_21[1U] = 0U;
        __Vtemp_21[2U] = 0U;
        __Vtemp_21[3U] = 0U;
        __Vtemp_21[4U] = 0U;
        VL_MUL_W(5, __Vtemp_22, __Vtemp_20, __Vtemp_21);
        vlSelfRef.rnd02__DOT__r1 = (0x000000ffffffffffULL 
                                    & (((QData)((IData)(__Vtemp_22[1U])) 
                                        << 0x00000020U) 
                                       | (QData)((IData)(__Vtemp_22[0U]))));
        __Vtemp_25[0U] = (0x01ffffffU & (IData)((vlSelfRef.i4 
                                                 >> 0x0000000dU)));
        __Vtemp_25[1U] = 0U;
        __Vtemp_25[2U] = 0U;
        __Vtemp_26[0U] = 0U;
        __Vtemp_26[1U] = 0U;
        __Vtemp_26[2U] = 0U;
        __Vtemp_28[0U] = vlSelfRef.rnd02__DOT__r3;
        __Vtemp_28[1U] = 0U;
        __Vtemp_28[2U] = 0U;
        VL_SUB_W(3, __Vtemp_29, vlSelfRef.i3, __Vtemp_28);
        __Vtemp_30[0U] = __Vtemp_29[0U];
        __Vtemp_30[1U] = __Vtemp_29[1U];
        __Vtemp_30[2U] = (0x000000ffU & __Vtemp_29[2U]);
        __Vtemp_31[0U] = VL_LTS_IWW(72, __Vtemp_26, __Vtemp_30);
        __Vtemp_31[1U] = 0U;
        __Vtemp_31[2U] = 0U;
        VL_SUB_W(3, __Vtemp_32, __Vtemp_25, __Vtemp_31);
        vlSelfRef.rnd02__DOT__w0[0U] = __Vtemp_32[0U];
        vlSelfRef.rnd02__DOT__w0[1U] = __Vtemp_32[1U];
        vlSelfRef.rnd02__DOT__w0[2U] = (1U & __Vtemp_32[2U]);
        vlSelfRef.rnd02__DOT__r2 = __Vdly__rnd02__DOT__r2;
        vlSelfRef.rnd02__DOT__w9 = (0x0001ffffU & VL_SHIFTL_III(17,17,8, vlSelfRef.rnd02__DOT__w0[0U], 
                                                                (0x000000ffU 
                                                                 & (IData)(vlSelfRef.i4))));
        vlSelfRef.o2 = (IData)((0x000000ffffffffffULL 
                                & VL_MODDIV_QQQ(40, (QData)((IData)(
                                                                    (0x0000007fU 
                                                                     & (IData)(
                                                                               (0x000000ffffffffffULL 
                                                                                & VL_DIVS_QQQ(40, 
                                                                                (0x000000ffffffffffULL 
                                                                                & VL_EXTENDS_QI(40,24, vlSelfRef.rnd02__DOT__r0)), 0ULL)))))), 0ULL)));
        vlSelfRef.rnd02__DOT__w6 = (0x0000007fU & (IData)(
                                                          (0x000000ffffffffffULL 
                                                           & VL_DIVS_QQQ(40, 
                                                                         (0x000000ffffffffffULL 
                                                                          & VL_EXTENDS_QI(40,24, vlSelfRef.rnd02__DOT__r0)), 0ULL))));
        if ((0U != (0x000000ffffffffffULL & VL_SHIFTRS_QQI(40,40,8, 
                                                           (0x000000ffffffffffULL 
                                                            & VL_DIV_QQQ(40, (QData)((IData)(vlSelfRef.i5)), vlSelfRef.rnd02__DOT__r1)), 
                                                           (0x000000ffU 
                                                            & vlSelfRef.rnd02__DOT__r3))))) {
            vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[0U] 
                = (vlSelfRef.i2[0U] ^ (IData)(vlSelfRef.i5));
            vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[1U] 
                = vlSelfRef.i2[1U];
            vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[2U] 
                = vlSelfRef.i2[2U];
            vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[3U] 
                = vlSelfRef.i0;
        } else {
            vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[0U] 
                = (IData)(vlSelfRef.i1);
            vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[1U] 
                = (IData)((vlSelfRef.i1 >> 0x00000020U));
            vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[2U] = 0U;
            vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[3U] = 0U;
        }
        VL_EXTENDS_WI(96,24, __Vtemp_35, (0x00ffffffU 
                                          & (~ (0x00000040U 
                                                | (1U 
                                                   != (IData)(vlSelfRef.rnd02__DOT__r2))))));
        VL_SHIFTRS_WWI(96,96,8, vlSelfRef.o4, __Vtemp_35, 
                       (0x000000ffU & vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[0U]));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1 = 
            (VL_SHIFTR_QQI(49,49,8, (QData)((IData)(
                                                    ((0ULL 
                                                      != 
                                                      (0x00000000000ff000ULL 
                                                       & vlSelfRef.i4)) 
                                                     | (0U 
                                                        != 
                                                        ((vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[0U] 
                                                          | vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[1U]) 
                                                         | (0x000000ffU 
                                                            & vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[2U])))))), 
                           (0x000000ffU & vlSelfRef.rnd02__DOT__r0)) 
             >= ((0x0001fffffffffe00ULL & (VL_DIVS_QQQ(40, 
                                                       (0x000000ffffffffffULL 
                                                        & VL_EXTENDS_QI(40,9, (IData)(vlSelfRef.i5))), vlSelfRef.rnd02__DOT__r1) 
                                           << 9U)) 
                 | (QData)((IData)(vlSelfRef.i5))));
        VL_EXTENDS_WI(96,24, __Vtemp_39, vlSelfRef.rnd02__DOT__r0);
        vlSelfRef.o3 = VL_SHIFTR_QQI(40,40,8, (((QData)((IData)(
                                                                ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1) 
                                                                 - vlSelfRef.rnd02__DOT__r3))) 
                                                << 1U) 
                                               | (QData)((IData)(
                                                                 VL_LTS_IWW(96, vlSelfRef.i2, __Vtemp_39)))), 
                                     (0x000000ffU & (IData)(vlSelfRef.rnd02__DOT__r2)));
        vlSelfRef.rnd02__DOT__w5 = (0x0000007fU & (IData)(
                                                          VL_MODDIVS_QQQ(64, vlSelfRef.i1, 
                                                                         VL_EXTENDS_QI(64,8, (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1)))));
        __Vtemp_41[0U] = (0x2f6f6b26U | (1U & (IData)(
                                                      (vlSelfRef.rnd02__DOT__r1 
                                                       >> (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1)))));
        __Vtemp_41[1U] = 0x2e00a8aeU;
        __Vtemp_41[2U] = 0x000000e6U;
        vlSelfRef.rnd02__DOT__w10 = (0x0000ffffU & 
                                     (((IData)(1U) 
                                       + VL_ONEHOT_W(3, __Vtemp_41)) 
                                      - VL_SHIFTR_III(16,32,8, 
                                                      (1U 
                                                       & ((IData)(vlSelfRef.i5) 
                                                          >> 
                                                          (7U 
                                                           & vlSelfRef.rnd02__DOT__w0[0U]))), 
                                                      (0x000000ffU 
                                                       & vlSelfRef.rnd02__DOT__r0))));
        vlSelfRef.o0 = (1U & VL_DIV_III(16, (IData)(vlSelfRef.rnd02__DOT__w5), 
                                        VL_SHIFTR_III(16,16,8, 
                                                      ((0U 
                                                        != 
                                                        (0xffdd7dfeU 
                                                         | ((~ (IData)(
                                                                       (0xee181ce8a7a97d65ULL 
                                                                        ^ vlSelfRef.i1))) 
                                                            | (~ (IData)(
                                                                         ((0xee181ce8a7a97d65ULL 
                                                                           ^ vlSelfRef.i1) 
                                                                          >> 0x00000020U)))))) 
                                                       & (0U 
                                                          != 
                                                          ((vlSelfRef.rnd02__DOT__w0[0U] 
                                                            | vlSelfRef.rnd02__DOT__w0[1U]) 
                                                           | vlSelfRef.rnd02__DOT__w0[2U]))), 
                                                      (0x000000ffU 
                                                       & vlSelfRef.i3[0U]))));
        vlSelfRef.o1 = (0x0000ffffU & (IData)(VL_DIVS_QQQ(64, 
                                                          VL_EXTENDS_QI(64,16, (IData)(vlSelfRef.rnd02__DOT__w10)), vlSelfRef.i1)));
    }

    VLCU_D static void Vrnd02___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vrnd02___024root___eval_phase__act(Vrnd02___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vrnd02___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vrnd02___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vrnd02___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vrnd02___024root___eval_phase__nba(Vrnd02___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vrnd02___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    Vrnd02___024root___nba_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
            Vrnd02___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vrnd02___024root___eval(Vrnd02___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vrnd02___024root___dump_triggers__ico(vlSelfRef.__VicoTriggered, "ico"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/rnd02.v", 2, "", "DIDNOTCONVERGE: Input combinational region did not converge after '--converge-limit' of 10000 tries");
            }
            __VicoIterCount = ((IData)(1U) + __VicoIterCount);
            vlSelfRef.__VicoPhaseResult = Vrnd02___024root___eval_phase__ico(vlSelf, vlCtx);
        } while (vlSelfRef.__VicoPhaseResult);
        __VnbaIterCount = 0U;
        do {
            if (VL_UNLIKELY(((0x00002710U < __VnbaIterCount)))) {
#ifdef VL_DEBUG
                Vrnd02___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/rnd02.v", 2, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vrnd02___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("../verilator_b200/csrc/designs/torture/rnd02.v", 2, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vrnd02___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vrnd02___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vrnd02___024root___ctor_var_reset(Vrnd02___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        vlSelf->i0 = VL_SCOPED_RAND_RESET_I(7, __VscopeHash, 8471475882151657015ull);
        vlSelf->i1 = VL_SCOPED_RAND_RESET_Q(64, __VscopeHash, 2165040029148572678ull);
        VL_SCOPED_RAND_RESET_W(96, vlSelf->i2, __VscopeHash, 12628911202397369737ull);
        VL_SCOPED_RAND_RESET_W(72, vlSelf->i3, __VscopeHash, 6886176223853197501ull);
        vlSelf->i4 = VL_SCOPED_RAND_RESET_Q(64, __VscopeHash, 5103504462683158156ull);
        vlSelf->i5 = VL_SCOPED_RAND_RESET_I(9, __VscopeHash, 10954576514065331241ull);
        vlSelf->o0 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 4709259436872190324ull);
        vlSelf->o1 = VL_SCOPED_RAND_RESET_I(16, __VscopeHash, 16871796271195187576ull);
        vlSelf->o2 = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 6221000330064061286ull);
        vlSelf->o3 = VL_SCOPED_RAND_RESET_Q(40, __VscopeHash, 17472377812510545679ull);
        VL_SCOPED_RAND_RESET_W(96, vlSelf->o4, __VscopeHash, 12832999154864337794ull);
        vlSelf->rnd02__DOT__r0 = VL_SCOPED_RAND_RESET_I(24, __VscopeHash, 12295341987098237027ull);
        vlSelf->rnd02__DOT__r1 = VL_SCOPED_RAND_RESET_Q(40, __VscopeHash, 3512675062919955528ull);
        vlSelf->rnd02__DOT__r2 = VL_SCOPED_RAND_RESET_I(9, __VscopeHash, 5324002034989625537ull);
        vlSelf->rnd02__DOT__r3 = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 10294965215296926166ull);
        VL_SCOPED_RAND_RESET_W(65, vlSelf->rnd02__DOT__w0, __VscopeHash, 5421601015822508166ull);
        vlSelf->rnd02__DOT__w5 = VL_SCOPED_RAND_RESET_I(7, __VscopeHash, 6658781111111404229ull);
        vlSelf->rnd02__DOT__w6 = VL_SCOPED_RAND_RESET_I(7, __VscopeHash, 3929921750067491446ull);
        vlSelf->rnd02__DOT__w9 = VL_SCOPED_RAND_RESET_I(17, __VscopeHash, 4221215844225302770ull);
        vlSelf->rnd02__DOT__w10 = VL_SCOPED_RAND_RESET_I(16, __VscopeHash, 7611965196836561213ull);
        VL_ZERO_RESET_W(103, vlSelf->__VdfgRegularize_h6e95ff9d_0_0);
        vlSelf->__VdfgRegularize_h6e95ff9d_0_1 = 0;
        for (int __Vi0 = 0; __Vi0 < 1; ++__Vi0) {
            vlSelf->__VstlTriggered[__Vi0] = 0;
        }
        for (int __Vi0 = 0; __Vi0 < 2; ++__Vi0) {
            vlSelf->__VicoTriggered[__Vi0] = 0;
        }
        vlSelf->__Vtrigprevexpr___TOP__clk__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__i0__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__i1__0 = 0;
        VL_ZERO_RESET_W(96, vlSelf->__Vtrigprevexpr___TOP__i2__0);
        VL_ZERO_RESET_W(72, vlSelf->__Vtrigprevexpr___TOP__i3__0);
        vlSelf->__Vtrigprevexpr___TOP__i4__0 = 0;
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
        Vrnd02___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vrnd02___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vrnd02___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vrnd02___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vrnd02___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vrnd02___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vrnd02__cuda::info() {
    typedef Vrnd02__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"i0", "i0", VLCU_MEMBER_IN, 1, 0, 1, 7, 6, 0, T::OFF_i0},
        {"i1", "i1", VLCU_MEMBER_IN, 8, 0, 1, 64, 63, 0, T::OFF_i1},
        {"i2", "i2", VLCU_MEMBER_IN, 4, 0, 3, 96, 95, 0, T::OFF_i2},
        {"i3", "i3", VLCU_MEMBER_IN, 4, 0, 3, 72, 71, 0, T::OFF_i3},
        {"i4", "i4", VLCU_MEMBER_IN, 8, 0, 1, 64, 63, 0, T::OFF_i4},
        {"i5", "i5", VLCU_MEMBER_IN, 2, 0, 1, 9, 8, 0, T::OFF_i5},
        {"o0", "o0", VLCU_MEMBER_OUT, 1, 0, 1, 1, 0, 0, T::OFF_o0},
        {"o1", "o1", VLCU_MEMBER_OUT, 2, 0, 1, 16, 15, 0, T::OFF_o1},
        {"o2", "o2", VLCU_MEMBER_OUT, 4, 0, 1, 32, 31, 0, T::OFF_o2},
        {"o3", "o3", VLCU_MEMBER_OUT, 8, 0, 1, 40, 39, 0, T::OFF_o3},
        {"o4", "o4", VLCU_MEMBER_OUT, 4, 0, 3, 96, 95, 0, T::OFF_o4},
        {"rnd02__DOT__w5", "rnd02.w5", VLCU_MEMBER_STATE, 1, 0, 1, 7, 6, 0, T::OFF_rnd02__DOT__w5},
        {"rnd02__DOT__w6", "rnd02.w6", VLCU_MEMBER_STATE, 1, 0, 1, 7, 6, 0, T::OFF_rnd02__DOT__w6},
        {"__VdfgRegularize_h6e95ff9d_0_1", "__VdfgRegularize_h6e95ff9d_0_1", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_1},
        {"__VstlFirstIteration", "__VstlFirstIteration", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlFirstIteration},
        {"__VstlPhaseResult", "__VstlPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__Vtrigprevexpr___TOP__i0__0", "__Vtrigprevexpr___TOP__i0__0", VLCU_MEMBER_STATE, 1, 0, 1, 7, 6, 0, T::OFF___Vtrigprevexpr___TOP__i0__0},
        {"__VicoDidInit", "__VicoDidInit", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VicoDidInit},
        {"__VicoPhaseResult", "__VicoPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VicoPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__1", "__Vtrigprevexpr___TOP__clk__1", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__1},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"rnd02__DOT__r2", "rnd02.r2", VLCU_MEMBER_STATE, 2, 0, 1, 9, 8, 0, T::OFF_rnd02__DOT__r2},
        {"rnd02__DOT__w10", "rnd02.w10", VLCU_MEMBER_STATE, 2, 0, 1, 16, 15, 0, T::OFF_rnd02__DOT__w10},
        {"__Vtrigprevexpr___TOP__i5__0", "__Vtrigprevexpr___TOP__i5__0", VLCU_MEMBER_STATE, 2, 0, 1, 9, 8, 0, T::OFF___Vtrigprevexpr___TOP__i5__0},
        {"rnd02__DOT__r0", "rnd02.r0", VLCU_MEMBER_STATE, 4, 0, 1, 24, 23, 0, T::OFF_rnd02__DOT__r0},
        {"rnd02__DOT__r3", "rnd02.r3", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rnd02__DOT__r3},
        {"rnd02__DOT__w0", "rnd02.w0", VLCU_MEMBER_STATE, 4, 0, 3, 65, 64, 0, T::OFF_rnd02__DOT__w0},
        {"rnd02__DOT__w9", "rnd02.w9", VLCU_MEMBER_STATE, 4, 0, 1, 17, 16, 0, T::OFF_rnd02__DOT__w9},
        {"__VdfgRegularize_h6e95ff9d_0_0", "__VdfgRegularize_h6e95ff9d_0_0", VLCU_MEMBER_STATE, 4, 0, 4, 103, 102, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_0},
        {"__Vtrigprevexpr___TOP__i2__0", "__Vtrigprevexpr___TOP__i2__0", VLCU_MEMBER_STATE, 4, 0, 3, 96, 95, 0, T::OFF___Vtrigprevexpr___TOP__i2__0},
        {"__Vtrigprevexpr___TOP__i3__0", "__Vtrigprevexpr___TOP__i3__0", VLCU_MEMBER_STATE, 4, 0, 3, 72, 71, 0, T::OFF___Vtrigprevexpr___TOP__i3__0},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"rnd02__DOT__r1", "rnd02.r1", VLCU_MEMBER_STATE, 8, 0, 1, 40, 39, 0, T::OFF_rnd02__DOT__r1},
        {"__Vtrigprevexpr___TOP__i1__0", "__Vtrigprevexpr___TOP__i1__0", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___Vtrigprevexpr___TOP__i1__0},
        {"__Vtrigprevexpr___TOP__i4__0", "__Vtrigprevexpr___TOP__i4__0", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___Vtrigprevexpr___TOP__i4__0},
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
    static const uint32_t saveOrder[] = {0, 1, 7, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 6, 8, 24, 25, 26, 3, 4, 9, 11, 27, 28, 29, 30, 31, 32, 33, 34, 2, 5, 10, 35, 36, 37, 38, 39, 40, 41, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "rnd02", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 0, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 42, 0, 0x11dcf6eb895aecc9ULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vrnd02__cuda VLCU_TESTBENCHES);
