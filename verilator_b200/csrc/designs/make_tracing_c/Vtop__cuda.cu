// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vtop"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vtop___024root final {
    CData/*0:0*/ clk;
    CData/*0:0*/ reset_l;
    CData/*1:0*/ out_small;
    QData/*39:0*/ out_quad;
    VlWide<3>/*69:0*/ out_wide;
    CData/*1:0*/ in_small;
    QData/*39:0*/ in_quad;
    VlWide<3>/*69:0*/ in_wide;
    CData/*1:0*/ top__DOT____VlemCond_0;
    CData/*0:0*/ top__DOT____Vtogcov__clk;
    CData/*0:0*/ top__DOT____Vtogcov__reset_l;
    CData/*1:0*/ top__DOT____Vtogcov__out_small;
    CData/*1:0*/ top__DOT____Vtogcov__in_small;
    CData/*0:0*/ __VstlFirstIteration;
    CData/*0:0*/ __VstlPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*0:0*/ __Vtrigprevexpr___TOP__reset_l__0;
    CData/*1:0*/ __Vtrigprevexpr___TOP__in_small__0;
    CData/*0:0*/ __VicoDidInit;
    CData/*0:0*/ __VicoPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__1;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    VlWide<3>/*69:0*/ top__DOT____VlemCond_2;
    VlWide<3>/*69:0*/ top__DOT____Vtogcov__out_wide;
    VlWide<3>/*69:0*/ top__DOT____Vtogcov__in_wide;
    IData/*31:0*/ top__DOT__sub__DOT__count_c;
    IData/*31:0*/ top__DOT__sub__DOT____Vtogcov__count_c;
    IData/*31:0*/ __Vsampled_TOP__top__DOT__sub__DOT__count_c;
    VlWide<3>/*69:0*/ __Vtrigprevexpr___TOP__in_wide__0;
    IData/*31:0*/ __VactIterCount;
    QData/*39:0*/ top__DOT____VlemCond_1;
    QData/*39:0*/ top__DOT____Vtogcov__out_quad;
    QData/*39:0*/ top__DOT____Vtogcov__in_quad;
    QData/*39:0*/ __Vtrigprevexpr___TOP__in_quad__0;
    VlUnpacked<QData/*63:0*/, 1> __VstlTriggered;
    VlUnpacked<QData/*63:0*/, 2> __VicoTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    VlcuCovArray<543> __Vcoverage;  // stays in HBM as cov[bin][instance], RED.ADD
    const char* vlNamep;
};

struct Vtop__cuda final {
    typedef Vtop___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 2348,
        OFF_reset_l = 2349,
        OFF_out_small = 2350,
        OFF_out_quad = 0,
        OFF_out_wide = 88,
        OFF_in_small = 2351,
        OFF_in_quad = 8,
        OFF_in_wide = 100,
        OFF_top__DOT____VlemCond_0 = 2352,
        OFF_top__DOT____Vtogcov__clk = 2353,
        OFF_top__DOT____Vtogcov__reset_l = 2354,
        OFF_top__DOT____Vtogcov__out_small = 2355,
        OFF_top__DOT____Vtogcov__in_small = 2356,
        OFF___VstlFirstIteration = 2357,
        OFF___VstlPhaseResult = 2358,
        OFF___Vtrigprevexpr___TOP__clk__0 = 2359,
        OFF___Vtrigprevexpr___TOP__reset_l__0 = 2360,
        OFF___Vtrigprevexpr___TOP__in_small__0 = 2361,
        OFF___VicoDidInit = 2362,
        OFF___VicoPhaseResult = 2363,
        OFF___Vtrigprevexpr___TOP__clk__1 = 2364,
        OFF___VactPhaseResult = 2365,
        OFF___VnbaPhaseResult = 2366,
        OFF_top__DOT____VlemCond_2 = 112,
        OFF_top__DOT____Vtogcov__out_wide = 124,
        OFF_top__DOT____Vtogcov__in_wide = 136,
        OFF_top__DOT__sub__DOT__count_c = 148,
        OFF_top__DOT__sub__DOT____Vtogcov__count_c = 152,
        OFF___Vsampled_TOP__top__DOT__sub__DOT__count_c = 156,
        OFF___Vtrigprevexpr___TOP__in_wide__0 = 160,
        OFF___VactIterCount = 172,
        OFF_top__DOT____VlemCond_1 = 16,
        OFF_top__DOT____Vtogcov__out_quad = 24,
        OFF_top__DOT____Vtogcov__in_quad = 32,
        OFF___Vtrigprevexpr___TOP__in_quad__0 = 40,
        OFF___VstlTriggered = 48,
        OFF___VicoTriggered = 56,
        OFF___VactTriggered = 72,
        OFF___VnbaTriggered = 80,
        OFF___Vcoverage = 176,
        kStateBytes = 2368,  // S_persist
        kTouchedReadBytes = 185,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 162,  // S_state_wr
    };
    struct Locals final {
    };
    struct Shared final {
    };
    VLCU_D static void bind(Root& s, Locals& l, Shared& sh, const VlcuState& st, uint64_t i) {
        (void)l; (void)sh; (void)st; (void)i;
        s.vlNamep = "TOP";
        s.__Vcoverage.bind(st, OFF___Vcoverage, i);
    }
    VLCU_D static void load(Root& s, const VlcuState& st, uint64_t i) {
        s.clk = st.ld<VlcuMem8>(OFF_clk, i);
        s.reset_l = st.ld<VlcuMem8>(OFF_reset_l, i);
        s.out_small = st.ld<VlcuMem8>(OFF_out_small, i);
        s.out_quad = st.ld<QData>(OFF_out_quad, i);
        st.ldW(OFF_out_wide, i, s.out_wide);
        s.in_small = st.ld<VlcuMem8>(OFF_in_small, i);
        s.in_quad = st.ld<QData>(OFF_in_quad, i);
        st.ldW(OFF_in_wide, i, s.in_wide);
        s.top__DOT____VlemCond_0 = st.ld<VlcuMem8>(OFF_top__DOT____VlemCond_0, i);
        s.top__DOT____Vtogcov__clk = st.ld<VlcuMem8>(OFF_top__DOT____Vtogcov__clk, i);
        s.top__DOT____Vtogcov__reset_l = st.ld<VlcuMem8>(OFF_top__DOT____Vtogcov__reset_l, i);
        s.top__DOT____Vtogcov__out_small = st.ld<VlcuMem8>(OFF_top__DOT____Vtogcov__out_small, i);
        s.top__DOT____Vtogcov__in_small = st.ld<VlcuMem8>(OFF_top__DOT____Vtogcov__in_small, i);
        s.__VstlFirstIteration = st.ld<VlcuMem8>(OFF___VstlFirstIteration, i);
        s.__VstlPhaseResult = st.ld<VlcuMem8>(OFF___VstlPhaseResult, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__Vtrigprevexpr___TOP__reset_l__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__reset_l__0, i);
        s.__Vtrigprevexpr___TOP__in_small__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__in_small__0, i);
        s.__VicoDidInit = st.ld<VlcuMem8>(OFF___VicoDidInit, i);
        s.__VicoPhaseResult = st.ld<VlcuMem8>(OFF___VicoPhaseResult, i);
        s.__Vtrigprevexpr___TOP__clk__1 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__1, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        st.ldW(OFF_top__DOT____VlemCond_2, i, s.top__DOT____VlemCond_2);
        st.ldW(OFF_top__DOT____Vtogcov__out_wide, i, s.top__DOT____Vtogcov__out_wide);
        st.ldW(OFF_top__DOT____Vtogcov__in_wide, i, s.top__DOT____Vtogcov__in_wide);
        s.top__DOT__sub__DOT__count_c = st.ld<IData>(OFF_top__DOT__sub__DOT__count_c, i);
        s.top__DOT__sub__DOT____Vtogcov__count_c = st.ld<IData>(OFF_top__DOT__sub__DOT____Vtogcov__count_c, i);
        s.__Vsampled_TOP__top__DOT__sub__DOT__count_c = st.ld<IData>(OFF___Vsampled_TOP__top__DOT__sub__DOT__count_c, i);
        st.ldW(OFF___Vtrigprevexpr___TOP__in_wide__0, i, s.__Vtrigprevexpr___TOP__in_wide__0);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        s.top__DOT____VlemCond_1 = st.ld<QData>(OFF_top__DOT____VlemCond_1, i);
        s.top__DOT____Vtogcov__out_quad = st.ld<QData>(OFF_top__DOT____Vtogcov__out_quad, i);
        s.top__DOT____Vtogcov__in_quad = st.ld<QData>(OFF_top__DOT____Vtogcov__in_quad, i);
        s.__Vtrigprevexpr___TOP__in_quad__0 = st.ld<QData>(OFF___Vtrigprevexpr___TOP__in_quad__0, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VstlTriggered[e] = st.ldE<QData>(OFF___VstlTriggered, e, i);
        for (uint32_t e = 0; e < 2; ++e) s.__VicoTriggered[e] = st.ldE<QData>(OFF___VicoTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.st<VlcuMem8>(OFF_reset_l, i, s.reset_l);
        st.st<VlcuMem8>(OFF_out_small, i, s.out_small);
        st.st<QData>(OFF_out_quad, i, s.out_quad);
        st.stW(OFF_out_wide, i, s.out_wide);
        st.st<VlcuMem8>(OFF_in_small, i, s.in_small);
        st.st<QData>(OFF_in_quad, i, s.in_quad);
        st.stW(OFF_in_wide, i, s.in_wide);
        st.st<VlcuMem8>(OFF_top__DOT____VlemCond_0, i, s.top__DOT____VlemCond_0);
        st.st<VlcuMem8>(OFF_top__DOT____Vtogcov__clk, i, s.top__DOT____Vtogcov__clk);
        st.st<VlcuMem8>(OFF_top__DOT____Vtogcov__reset_l, i, s.top__DOT____Vtogcov__reset_l);
        st.st<VlcuMem8>(OFF_top__DOT____Vtogcov__out_small, i, s.top__DOT____Vtogcov__out_small);
        st.st<VlcuMem8>(OFF_top__DOT____Vtogcov__in_small, i, s.top__DOT____Vtogcov__in_small);
        st.st<VlcuMem8>(OFF___VstlFirstIteration, i, s.__VstlFirstIteration);
        st.st<VlcuMem8>(OFF___VstlPhaseResult, i, s.__VstlPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__reset_l__0, i, s.__Vtrigprevexpr___TOP__reset_l__0);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__in_small__0, i, s.__Vtrigprevexpr___TOP__in_small__0);
        st.st<VlcuMem8>(OFF___VicoDidInit, i, s.__VicoDidInit);
        st.st<VlcuMem8>(OFF___VicoPhaseResult, i, s.__VicoPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__1, i, s.__Vtrigprevexpr___TOP__clk__1);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.stW(OFF_top__DOT____VlemCond_2, i, s.top__DOT____VlemCond_2);
        st.stW(OFF_top__DOT____Vtogcov__out_wide, i, s.top__DOT____Vtogcov__out_wide);
        st.stW(OFF_top__DOT____Vtogcov__in_wide, i, s.top__DOT____Vtogcov__in_wide);
        st.st<IData>(OFF_top__DOT__sub__DOT__count_c, i, s.top__DOT__sub__DOT__count_c);
        st.st<IData>(OFF_top__DOT__sub__DOT____Vtogcov__count_c, i, s.top__DOT__sub__DOT____Vtogcov__count_c);
        st.st<IData>(OFF___Vsampled_TOP__top__DOT__sub__DOT__count_c, i, s.__Vsampled_TOP__top__DOT__sub__DOT__count_c);
        st.stW(OFF___Vtrigprevexpr___TOP__in_wide__0, i, s.__Vtrigprevexpr___TOP__in_wide__0);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
        st.st<QData>(OFF_top__DOT____VlemCond_1, i, s.top__DOT____VlemCond_1);
        st.st<QData>(OFF_top__DOT____Vtogcov__out_quad, i, s.top__DOT____Vtogcov__out_quad);
        st.st<QData>(OFF_top__DOT____Vtogcov__in_quad, i, s.top__DOT____Vtogcov__in_quad);
        st.st<QData>(OFF___Vtrigprevexpr___TOP__in_quad__0, i, s.__Vtrigprevexpr___TOP__in_quad__0);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VstlTriggered, e, i, s.__VstlTriggered[e]);
        for (uint32_t e = 0; e < 2; ++e) st.stE<QData>(OFF___VicoTriggered, e, i, s.__VicoTriggered[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VactTriggered, e, i, s.__VactTriggered[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VnbaTriggered, e, i, s.__VnbaTriggered[e]);
    }
    VLCU_D static void load_inputs_row(Root& s, const uint32_t* rowp, uint64_t stride, uint64_t i) {
        s.clk = static_cast<VlcuMem8>(rowp[0 * stride + i] & 0x1U);
        s.reset_l = static_cast<VlcuMem8>(rowp[1 * stride + i] & 0x1U);
        s.in_small = static_cast<VlcuMem8>(rowp[2 * stride + i] & 0x3U);
        s.in_quad = (static_cast<QData>(rowp[4 * stride + i] & 0xffU) << 32) | rowp[3 * stride + i];
        s.in_wide[0] = rowp[5 * stride + i] & 0xffffffffU;
        s.in_wide[1] = rowp[6 * stride + i] & 0xffffffffU;
        s.in_wide[2] = rowp[7 * stride + i] & 0x3fU;
    }
    VLCU_D static void store_outputs_row(const Root& s, uint32_t* rowp, uint64_t stride, uint64_t i) {
        rowp[0 * stride + i] = s.out_small;
        rowp[1 * stride + i] = static_cast<uint32_t>(s.out_quad);
        rowp[2 * stride + i] = static_cast<uint32_t>(s.out_quad >> 32);
        rowp[3 * stride + i] = s.out_wide[0];
        rowp[4 * stride + i] = s.out_wide[1];
        rowp[5 * stride + i] = s.out_wide[2];
    }
    VLCU_D static uint64_t outputs_sig(const Root& s, uint64_t sig) {
        uint32_t acc = VLCU_SIG_ACC_INIT;
        acc = vlcu_sig_word(acc, s.out_small);
        acc = vlcu_sig_word(acc, static_cast<uint32_t>(s.out_quad));
        acc = vlcu_sig_word(acc, static_cast<uint32_t>(s.out_quad >> 32));
        acc = vlcu_sig_word(acc, s.out_wide[0]);
        acc = vlcu_sig_word(acc, s.out_wide[1]);
        acc = vlcu_sig_word(acc, s.out_wide[2]);
        return vlcu_sig_mix(sig, acc);
    }
    enum : uint32_t { kInWords = 8, kOutWords = 6, kInBytes = 23, kOutBytes = 21, kTraceWords = 15 };

    VLCU_D static void Vtop___024root___eval_static(Vtop___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
        vlSelfRef.__Vtrigprevexpr___TOP__reset_l__0 
            = vlSelfRef.reset_l;
        vlSelfRef.__Vtrigprevexpr___TOP__in_small__0 
            = vlSelfRef.in_small;
        vlSelfRef.__Vtrigprevexpr___TOP__in_quad__0 
            = vlSelfRef.in_quad;
        vlSelfRef.__Vtrigprevexpr___TOP__in_wide__0[0U] 
            = vlSelfRef.in_wide[0U];
        vlSelfRef.__Vtrigprevexpr___TOP__in_wide__0[1U] 
            = vlSelfRef.in_wide[1U];
        vlSelfRef.__Vtrigprevexpr___TOP__in_wide__0[2U] 
            = vlSelfRef.in_wide[2U];
        vlSelfRef.__Vtrigprevexpr___TOP__clk__1 = vlSelfRef.clk;
    }

    VLCU_D static void Vtop___024root___eval_initial(Vtop___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        {
            // Inlined CFunc: _eval_initial__TOP
            if (VL_UNLIKELY(((0U != static_cast<IData>(vlCtx.testPlusArg(0U)))))) {
                vlCtx.writef(0U, 1U, vlcu_msg_arg(VL_TIME_UNITED_Q(1)));
                // $dumpfile/$dumpvars: waveforms are written on the host (vlcu_vcd_write)
                // $dumpfile/$dumpvars: waveforms are written on the host (vlcu_vcd_write)
                ++(vlSelf->__Vcoverage[464]);
            } else {
                ++(vlSelf->__Vcoverage[465]);
            }
            vlCtx.writef(1U, 1U, vlcu_msg_arg(VL_TIME_UNITED_Q(1)));
            ++(vlSelf->__Vcoverage[466]);
        }
    }

    VLCU_D static void Vtop___024root___eval_final(Vtop___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vtop___024root___eval_settle(Vtop___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vtop___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
#endif
                VL_FATAL_MT("/root/reference/examples/make_tracing_c/top.v", 11, "", "DIDNOTCONVERGE: Settle region did not converge after '--converge-limit' of 10000 tries");
            }
            __VstlIterCount = ((IData)(1U) + __VstlIterCount);
            vlSelfRef.__VstlPhaseResult = Vtop___024root___eval_phase__stl(vlSelf, vlCtx);
            vlSelfRef.__VstlFirstIteration = 0U;
        } while (vlSelfRef.__VstlPhaseResult);
    }

    VLCU_D static bool Vtop___024root___trigger_anySet__stl(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vtop___024root___stl_sequent__TOP__0(Vtop___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        VlWide<3>/*95:0*/ __Vtemp_4;
        VlWide<3>/*95:0*/ __Vtemp_5;
        // Body
        if (((IData)(vlSelfRef.clk) ^ (IData)(vlSelfRef.top__DOT____Vtogcov__clk))) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 0, vlSelfRef.clk, vlSelfRef.top__DOT____Vtogcov__clk);
            vlSelfRef.top__DOT____Vtogcov__clk = vlSelfRef.clk;
        }
        if (((IData)(vlSelfRef.reset_l) ^ (IData)(vlSelfRef.top__DOT____Vtogcov__reset_l))) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 2, vlSelfRef.reset_l, vlSelfRef.top__DOT____Vtogcov__reset_l);
            vlSelfRef.top__DOT____Vtogcov__reset_l 
                = vlSelfRef.reset_l;
        }
        if ((0U != ((IData)(vlSelfRef.in_small) ^ (IData)(vlSelfRef.top__DOT____Vtogcov__in_small)))) {
            VL_COV_TOGGLE_CHG_ST_I(2, vlSelf->__Vcoverage + 228, vlSelfRef.in_small, vlSelfRef.top__DOT____Vtogcov__in_small);
            vlSelfRef.top__DOT____Vtogcov__in_small 
                = vlSelfRef.in_small;
        }
        if ((0ULL != (vlSelfRef.in_quad ^ vlSelfRef.top__DOT____Vtogcov__in_quad))) {
            VL_COV_TOGGLE_CHG_ST_Q(40, vlSelf->__Vcoverage + 232, vlSelfRef.in_quad, vlSelfRef.top__DOT____Vtogcov__in_quad);
            vlSelfRef.top__DOT____Vtogcov__in_quad 
                = vlSelfRef.in_quad;
        }
        if ((0U != (((vlSelfRef.in_wide[0U] ^ vlSelfRef.top__DOT____Vtogcov__in_wide[0U]) 
                     | (vlSelfRef.in_wide[1U] ^ vlSelfRef.top__DOT____Vtogcov__in_wide[1U])) 
                    | (vlSelfRef.in_wide[2U] ^ vlSelfRef.top__DOT____Vtogcov__in_wide[2U])))) {
            VL_COV_TOGGLE_CHG_ST_W(70, vlSelf->__Vcoverage + 312, vlSelfRef.in_wide, vlSelfRef.top__DOT____Vtogcov__in_wide);
            vlSelfRef.top__DOT____Vtogcov__in_wide[0U] 
                = vlSelfRef.in_wide[0U];
            vlSelfRef.top__DOT____Vtogcov__in_wide[1U] 
                = vlSelfRef.in_wide[1U];
            vlSelfRef.top__DOT____Vtogcov__in_wide[2U] 
                = vlSelfRef.in_wide[2U];
        }
        if ((0U != (vlSelfRef.top__DOT__sub__DOT__count_c 
                    ^ vlSelfRef.top__DOT__sub__DOT____Vtogcov__count_c))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 467, vlSelfRef.top__DOT__sub__DOT__count_c, vlSelfRef.top__DOT__sub__DOT____Vtogcov__count_c);
            vlSelfRef.top__DOT__sub__DOT____Vtogcov__count_c 
                = vlSelfRef.top__DOT__sub__DOT__count_c;
        }
        if (vlSelfRef.reset_l) {
            ++(vlSelf->__Vcoverage[455]);
            vlSelfRef.top__DOT____VlemCond_0 = (3U 
                                                & ((IData)(1U) 
                                                   + (IData)(vlSelfRef.in_small)));
        } else {
            ++(vlSelf->__Vcoverage[454]);
            vlSelfRef.top__DOT____VlemCond_0 = 0U;
        }
        vlSelfRef.out_small = vlSelfRef.top__DOT____VlemCond_0;
        if ((1U & (~ (IData)(vlSelfRef.reset_l)))) {
            ++(vlSelf->__Vcoverage[452]);
        }
        if (vlSelfRef.reset_l) {
            ++(vlSelf->__Vcoverage[453]);
        }
        if (vlSelfRef.reset_l) {
            ++(vlSelf->__Vcoverage[459]);
            vlSelfRef.top__DOT____VlemCond_1 = (0x000000ffffffffffULL 
                                                & (1ULL 
                                                   + vlSelfRef.in_quad));
        } else {
            ++(vlSelf->__Vcoverage[458]);
            vlSelfRef.top__DOT____VlemCond_1 = 0ULL;
        }
        vlSelfRef.out_quad = vlSelfRef.top__DOT____VlemCond_1;
        if ((1U & (~ (IData)(vlSelfRef.reset_l)))) {
            ++(vlSelf->__Vcoverage[456]);
        }
        if (vlSelfRef.reset_l) {
            ++(vlSelf->__Vcoverage[457]);
        }
        if (vlSelfRef.reset_l) {
            ++(vlSelf->__Vcoverage[463]);
            __Vtemp_4[0U] = 1U;
            __Vtemp_4[1U] = 0U;
            __Vtemp_4[2U] = 0U;
            VL_ADD_W(3, __Vtemp_5, __Vtemp_4, vlSelfRef.in_wide);
            vlSelfRef.top__DOT____VlemCond_2[0U] = __Vtemp_5[0U];
            vlSelfRef.top__DOT____VlemCond_2[1U] = __Vtemp_5[1U];
            vlSelfRef.top__DOT____VlemCond_2[2U] = 
                (0x0000003fU & __Vtemp_5[2U]);
        } else {
            ++(vlSelf->__Vcoverage[462]);
            vlSelfRef.top__DOT____VlemCond_2[0U] = 0U;
            vlSelfRef.top__DOT____VlemCond_2[1U] = 0U;
            vlSelfRef.top__DOT____VlemCond_2[2U] = 0U;
        }
        vlSelfRef.out_wide[0U] = vlSelfRef.top__DOT____VlemCond_2[0U];
        vlSelfRef.out_wide[1U] = vlSelfRef.top__DOT____VlemCond_2[1U];
        vlSelfRef.out_wide[2U] = vlSelfRef.top__DOT____VlemCond_2[2U];
        if ((1U & (~ (IData)(vlSelfRef.reset_l)))) {
            ++(vlSelf->__Vcoverage[460]);
        }
        if (vlSelfRef.reset_l) {
            ++(vlSelf->__Vcoverage[461]);
        }
        if ((0U != ((IData)(vlSelfRef.out_small) ^ (IData)(vlSelfRef.top__DOT____Vtogcov__out_small)))) {
            VL_COV_TOGGLE_CHG_ST_I(2, vlSelf->__Vcoverage + 4, vlSelfRef.out_small, vlSelfRef.top__DOT____Vtogcov__out_small);
            vlSelfRef.top__DOT____Vtogcov__out_small 
                = vlSelfRef.out_small;
        }
        if ((0ULL != (vlSelfRef.out_quad ^ vlSelfRef.top__DOT____Vtogcov__out_quad))) {
            VL_COV_TOGGLE_CHG_ST_Q(40, vlSelf->__Vcoverage + 8, vlSelfRef.out_quad, vlSelfRef.top__DOT____Vtogcov__out_quad);
            vlSelfRef.top__DOT____Vtogcov__out_quad 
                = vlSelfRef.out_quad;
        }
        if ((0U != (((vlSelfRef.out_wide[0U] ^ vlSelfRef.top__DOT____Vtogcov__out_wide[0U]) 
                     | (vlSelfRef.out_wide[1U] ^ vlSelfRef.top__DOT____Vtogcov__out_wide[1U])) 
                    | (vlSelfRef.out_wide[2U] ^ vlSelfRef.top__DOT____Vtogcov__out_wide[2U])))) {
            VL_COV_TOGGLE_CHG_ST_W(70, vlSelf->__Vcoverage + 88, vlSelfRef.out_wide, vlSelfRef.top__DOT____Vtogcov__out_wide);
            vlSelfRef.top__DOT____Vtogcov__out_wide[0U] 
                = vlSelfRef.out_wide[0U];
            vlSelfRef.top__DOT____Vtogcov__out_wide[1U] 
                = vlSelfRef.out_wide[1U];
            vlSelfRef.top__DOT____Vtogcov__out_wide[2U] 
                = vlSelfRef.out_wide[2U];
        }
    }

    VLCU_D static bool Vtop___024root___eval_phase__stl(Vtop___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vtop___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
        }
#endif
        __VstlExecute = Vtop___024root___trigger_anySet__stl(vlSelfRef.__VstlTriggered);
        if (__VstlExecute) {
            {
                // Inlined CFunc: _eval_stl
                if ((1ULL & vlSelfRef.__VstlTriggered[0U])) {
                    Vtop___024root___stl_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
        }
        return (__VstlExecute);
    }

    VLCU_D static bool Vtop___024root___trigger_anySet__ico(const VlUnpacked<QData/*63:0*/, 2> &in) {
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

    VLCU_D static void Vtop___024root___eval_ico(Vtop___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        if ((1ULL & vlSelfRef.__VicoTriggered[0U])) {
            {
                // Inlined CFunc: _ico_sequent__TOP__0
                if (((IData)(vlSelfRef.clk) ^ (IData)(vlSelfRef.top__DOT____Vtogcov__clk))) {
                    VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 0, vlSelfRef.clk, vlSelfRef.top__DOT____Vtogcov__clk);
                    vlSelfRef.top__DOT____Vtogcov__clk 
                        = vlSelfRef.clk;
                }
            }
        }
        if ((2ULL & vlSelfRef.__VicoTriggered[0U])) {
            {
                // Inlined CFunc: _ico_sequent__TOP__1
                if (((IData)(vlSelfRef.reset_l) ^ (IData)(vlSelfRef.top__DOT____Vtogcov__reset_l))) {
                    VL_COV_TOGGLE_CHG_ST_I(1, vlSelf->__Vcoverage + 2, vlSelfRef.reset_l, vlSelfRef.top__DOT____Vtogcov__reset_l);
                    vlSelfRef.top__DOT____Vtogcov__reset_l 
                        = vlSelfRef.reset_l;
                }
            }
        }
        if ((4ULL & vlSelfRef.__VicoTriggered[0U])) {
            {
                // Inlined CFunc: _ico_sequent__TOP__2
                if ((0U != ((IData)(vlSelfRef.in_small) 
                            ^ (IData)(vlSelfRef.top__DOT____Vtogcov__in_small)))) {
                    VL_COV_TOGGLE_CHG_ST_I(2, vlSelf->__Vcoverage + 228, vlSelfRef.in_small, vlSelfRef.top__DOT____Vtogcov__in_small);
                    vlSelfRef.top__DOT____Vtogcov__in_small 
                        = vlSelfRef.in_small;
                }
            }
        }
        if ((8ULL & vlSelfRef.__VicoTriggered[0U])) {
            {
                // Inlined CFunc: _ico_sequent__TOP__3
                if ((0ULL != (vlSelfRef.in_quad ^ vlSelfRef.top__DOT____Vtogcov__in_quad))) {
                    VL_COV_TOGGLE_CHG_ST_Q(40, vlSelf->__Vcoverage + 232, vlSelfRef.in_quad, vlSelfRef.top__DOT____Vtogcov__in_quad);
                    vlSelfRef.top__DOT____Vtogcov__in_quad 
                        = vlSelfRef.in_quad;
                }
            }
        }
        if ((0x0000000000000010ULL & vlSelfRef.__VicoTriggered[0U])) {
            {
                // Inlined CFunc: _ico_sequent__TOP__4
                if ((0U != (((vlSelfRef.in_wide[0U] 
                              ^ vlSelfRef.top__DOT____Vtogcov__in_wide[0U]) 
                             | (vlSelfRef.in_wide[1U] 
                                ^ vlSelfRef.top__DOT____Vtogcov__in_wide[1U])) 
                            | (vlSelfRef.in_wide[2U] 
                               ^ vlSelfRef.top__DOT____Vtogcov__in_wide[2U])))) {
                    VL_COV_TOGGLE_CHG_ST_W(70, vlSelf->__Vcoverage + 312, vlSelfRef.in_wide, vlSelfRef.top__DOT____Vtogcov__in_wide);
                    vlSelfRef.top__DOT____Vtogcov__in_wide[0U] 
                        = vlSelfRef.in_wide[0U];
                    vlSelfRef.top__DOT____Vtogcov__in_wide[1U] 
                        = vlSelfRef.in_wide[1U];
                    vlSelfRef.top__DOT____Vtogcov__in_wide[2U] 
                        = vlSelfRef.in_wide[2U];
                }
            }
        }
        if ((6ULL & vlSelfRef.__VicoTriggered[0U])) {
            {
                // Inlined CFunc: _ico_comb__TOP__0
                if (vlSelfRef.reset_l) {
                    ++(vlSelf->__Vcoverage[455]);
                    vlSelfRef.top__DOT____VlemCond_0 
                        = (3U & ((IData)(1U) + (IData)(vlSelfRef.in_small)));
                } else {
                    ++(vlSelf->__Vcoverage[454]);
                    vlSelfRef.top__DOT____VlemCond_0 = 0U;
                }
                vlSelfRef.out_small = vlSelfRef.top__DOT____VlemCond_0;
                if ((1U & (~ (IData)(vlSelfRef.reset_l)))) {
                    ++(vlSelf->__Vcoverage[452]);
                }
                if (vlSelfRef.reset_l) {
                    ++(vlSelf->__Vcoverage[453]);
                }
                if ((0U != ((IData)(vlSelfRef.out_small) 
                            ^ (IData)(vlSelfRef.top__DOT____Vtogcov__out_small)))) {
                    VL_COV_TOGGLE_CHG_ST_I(2, vlSelf->__Vcoverage + 4, vlSelfRef.out_small, vlSelfRef.top__DOT____Vtogcov__out_small);
                    vlSelfRef.top__DOT____Vtogcov__out_small 
                        = vlSelfRef.out_small;
                }
            }
        }
        if ((0x000000000000000aULL & vlSelfRef.__VicoTriggered[0U])) {
            {
                // Inlined CFunc: _ico_comb__TOP__1
                if (vlSelfRef.reset_l) {
                    ++(vlSelf->__Vcoverage[459]);
                    vlSelfRef.top__DOT____VlemCond_1 
                        = (0x000000ffffffffffULL & 
                           (1ULL + vlSelfRef.in_quad));
                } else {
                    ++(vlSelf->__Vcoverage[458]);
                    vlSelfRef.top__DOT____VlemCond_1 = 0ULL;
                }
                vlSelfRef.out_quad = vlSelfRef.top__DOT____VlemCond_1;
                if ((1U & (~ (IData)(vlSelfRef.reset_l)))) {
                    ++(vlSelf->__Vcoverage[456]);
                }
                if (vlSelfRef.reset_l) {
                    ++(vlSelf->__Vcoverage[457]);
                }
                if ((0ULL != (vlSelfRef.out_quad ^ vlSelfRef.top__DOT____Vtogcov__out_quad))) {
                    VL_COV_TOGGLE_CHG_ST_Q(40, vlSelf->__Vcoverage + 8, vlSelfRef.out_quad, vlSelfRef.top__DOT____Vtogcov__out_quad);
                    vlSelfRef.top__DOT____Vtogcov__out_quad 
                        = vlSelfRef.out_quad;
                }
            }
        }
        if ((0x0000000000000012ULL & vlSelfRef.__VicoTriggered[0U])) {
            {
                // Inlined CFunc: _ico_comb__TOP__2
                VlWide<3>/*95:0*/ __Vinline_0__ico_comb__TOP__2___Vtemp_2;
                VlWide<3>/*95:0*/ __Vinline_0__ico_comb__TOP__2___Vtemp_3;
                if (vlSelfRef.reset_l) {
                    ++(vlSelf->__Vcoverage[463]);
                    __Vinline_0__ico_comb__TOP__2___Vtemp_2[0U] = 1U;
                    __Vinline_0__ico_comb__TOP__2___Vtemp_2[1U] = 0U;
                    __Vinline_0__ico_comb__TOP__2___Vtemp_2[2U] = 0U;
                    VL_ADD_W(3, __Vinline_0__ico_comb__TOP__2___Vtemp_3, __Vinline_0__ico_comb__TOP__2___Vtemp_2, vlSelfRef.in_wide);
                    vlSelfRef.top__DOT____VlemCond_2[0U] 
                        = __Vinline_0__ico_comb__TOP__2___Vtemp_3[0U];
                    vlSelfRef.top__DOT____VlemCond_2[1U] 
                        = __Vinline_0__ico_comb__TOP__2___Vtemp_3[1U];
                    vlSelfRef.top__DOT____VlemCond_2[2U] 
                        = (0x0000003fU & __Vinline_0__ico_comb__TOP__2___Vtemp_3[2U]);
                } else {
                    ++(vlSelf->__Vcoverage[462]);
                    vlSelfRef.top__DOT____VlemCond_2[0U] = 0U;
                    vlSelfRef.top__DOT____VlemCond_2[1U] = 0U;
                    vlSelfRef.top__DOT____VlemCond_2[2U] = 0U;
                }
                vlSelfRef.out_wide[0U] = vlSelfRef.top__DOT____VlemCond_2[0U];
                vlSelfRef.out_wide[1U] = vlSelfRef.top__DOT____VlemCond_2[1U];
                vlSelfRef.out_wide[2U] = vlSelfRef.top__DOT____VlemCond_2[2U];
                if ((1U & (~ (IData)(vlSelfRef.reset_l)))) {
                    ++(vlSelf->__Vcoverage[460]);
                }
                if (vlSelfRef.reset_l) {
                    ++(vlSelf->__Vcoverage[461]);
                }
                if ((0U != (((vlSelfRef.out_wide[0U] 
                              ^ vlSelfRef.top__DOT____Vtogcov__out_wide[0U]) 
                             | (vlSelfRef.out_wide[1U] 
                                ^ vlSelfRef.top__DOT____Vtogcov__out_wide[1U])) 
                            | (vlSelfRef.out_wide[2U] 
                               ^ vlSelfRef.top__DOT____Vtogcov__out_wide[2U])))) {
                    VL_COV_TOGGLE_CHG_ST_W(70, vlSelf->__Vcoverage + 88, vlSelfRef.out_wide, vlSelfRef.top__DOT____Vtogcov__out_wide);
                    vlSelfRef.top__DOT____Vtogcov__out_wide[0U] 
                        = vlSelfRef.out_wide[0U];
                    vlSelfRef.top__DOT____Vtogcov__out_wide[1U] 
                        = vlSelfRef.out_wide[1U];
                    vlSelfRef.top__DOT____Vtogcov__out_wide[2U] 
                        = vlSelfRef.out_wide[2U];
                }
            }
        }
    }

    VLCU_D static bool Vtop___024root___eval_phase__ico(Vtop___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VicoExecute;
        // Body
        {
            // Inlined CFunc: _eval_triggers_vec__ico
            vlSelfRef.__VicoTriggered[0U] = (QData)((IData)(
                                                            (((0U 
                                                               != 
                                                               (((vlSelfRef.in_wide[0U] 
                                                                  ^ vlSelfRef.__Vtrigprevexpr___TOP__in_wide__0[0U]) 
                                                                 | (vlSelfRef.in_wide[1U] 
                                                                    ^ vlSelfRef.__Vtrigprevexpr___TOP__in_wide__0[1U])) 
                                                                | (vlSelfRef.in_wide[2U] 
                                                                   ^ vlSelfRef.__Vtrigprevexpr___TOP__in_wide__0[2U]))) 
                                                              << 4U) 
                                                             | ((((vlSelfRef.in_quad 
                                                                   != vlSelfRef.__Vtrigprevexpr___TOP__in_quad__0) 
                                                                  << 3U) 
                                                                 | (((IData)(vlSelfRef.in_small) 
                                                                     != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__in_small__0)) 
                                                                    << 2U)) 
                                                                | ((((IData)(vlSelfRef.reset_l) 
                                                                     != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__reset_l__0)) 
                                                                    << 1U) 
                                                                   | ((IData)(vlSelfRef.clk) 
                                                                      != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__clk__0)))))));
            vlSelfRef.__Vtrigprevexpr___TOP__clk__0 
                = vlSelfRef.clk;
            vlSelfRef.__Vtrigprevexpr___TOP__reset_l__0 
                = vlSelfRef.reset_l;
            vlSelfRef.__Vtrigprevexpr___TOP__in_small__0 
                = vlSelfRef.in_small;
            vlSelfRef.__Vtrigprevexpr___TOP__in_quad__0 
                = vlSelfRef.in_quad;
            vlSelfRef.__Vtrigprevexpr___TOP__in_wide__0[0U] 
                = vlSelfRef.in_wide[0U];
            vlSelfRef.__Vtrigprevexpr___TOP__in_wide__0[1U] 
                = vlSelfRef.in_wide[1U];
            vlSelfRef.__Vtrigprevexpr___TOP__in_wide__0[2U] 
                = vlSelfRef.in_wide[2U];
            if (VL_UNLIKELY(((1U & (~ (IData)(vlSelfRef.__VicoDidInit)))))) {
                vlSelfRef.__VicoDidInit = 1U;
                vlSelfRef.__VicoTriggered[0U] = (1ULL 
                                                 | vlSelfRef.__VicoTriggered[0U]);
                vlSelfRef.__VicoTriggered[0U] = (2ULL 
                                                 | vlSelfRef.__VicoTriggered[0U]);
                vlSelfRef.__VicoTriggered[0U] = (4ULL 
                                                 | vlSelfRef.__VicoTriggered[0U]);
                vlSelfRef.__VicoTriggered[0U] = (8ULL 
                                                 | vlSelfRef.__VicoTriggered[0U]);
                vlSelfRef.__VicoTriggered[0U] = (0x0000000000000010ULL 
                                                 | vlSelfRef.__VicoTriggered[0U]);
            }
        }
#ifdef VL_DEBUG
        if (VL_UNLIKELY(vlSymsp->_vm_contextp__->debug())) {
            Vtop___024root___dump_triggers__ico(vlSelfRef.__VicoTriggered, "ico"s);
        }
#endif
        __VicoExecute = Vtop___024root___trigger_anySet__ico(vlSelfRef.__VicoTriggered);
        if (__VicoExecute) {
            Vtop___024root___eval_ico(vlSelf, vlCtx);
        }
        return (__VicoExecute);
    }

    VLCU_D static bool Vtop___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vtop___024root___nba_sequent__TOP__0(Vtop___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        IData/*31:0*/ __Vdly__top__DOT__sub__DOT__count_c;
        __Vdly__top__DOT__sub__DOT__count_c = 0;
        // Body
        if (vlSymsp->_vm_contextp__->assertCtlGet(VerilatedAssertCtlQuery::ASSERT_CTL_ON, 1, 2)) {
            if (vlSymsp->_vm_contextp__->assertCtlGet(VerilatedAssertCtlQuery::ASSERT_CTL_PASS_ON_NONVACUOUS, 1, 2)) {
                if (VL_LIKELY(((3U == vlSelfRef.__Vsampled_TOP__top__DOT__sub__DOT__count_c)))) {
                    ++(vlSelf->__Vcoverage[542]);
                }
            }
        }
        if (vlSymsp->_vm_contextp__->assertCtlGet(VerilatedAssertCtlQuery::ASSERT_CTL_ON, 2, 1)) {
            if (vlSymsp->_vm_contextp__->assertCtlGet(VerilatedAssertCtlQuery::ASSERT_CTL_FAIL_ON, 2, 1)) {
                if (VL_UNLIKELY(((1U & (~ ((~ (IData)(vlSelfRef.reset_l)) 
                                           | (0x00000064U 
                                              > vlSelfRef.top__DOT__sub__DOT__count_c))))))) {
                    vlCtx.writef(2U, 1U, vlcu_msg_arg(VL_TIME_UNITED_Q(1)));
                    VL_STOP_MT("/root/reference/examples/make_tracing_c/sub.v", 35, "");
                }
            }
        }
        if ((0x00000064U > vlSelfRef.top__DOT__sub__DOT__count_c)) {
            ++(vlSelf->__Vcoverage[538]);
        }
        if ((1U & (~ (IData)(vlSelfRef.reset_l)))) {
            ++(vlSelf->__Vcoverage[539]);
        }
        if (((IData)(vlSelfRef.reset_l) & (0x00000064U 
                                           <= vlSelfRef.top__DOT__sub__DOT__count_c))) {
            ++(vlSelf->__Vcoverage[540]);
        }
        ++(vlSelf->__Vcoverage[541]);
        __Vdly__top__DOT__sub__DOT__count_c = vlSelfRef.top__DOT__sub__DOT__count_c;
        if (vlSelfRef.reset_l) {
            __Vdly__top__DOT__sub__DOT__count_c = ((IData)(1U) 
                                                   + vlSelfRef.top__DOT__sub__DOT__count_c);
            if (VL_UNLIKELY(((3U <= vlSelfRef.top__DOT__sub__DOT__count_c)))) {
                vlCtx.writef(3U, 0U);
                VL_FINISH_MT("/root/reference/examples/make_tracing_c/sub.v", 28, "");
                ++(vlSelf->__Vcoverage[531]);
            } else {
                ++(vlSelf->__Vcoverage[532]);
            }
            ++(vlSelf->__Vcoverage[534]);
        } else {
            ++(vlSelf->__Vcoverage[533]);
            __Vdly__top__DOT__sub__DOT__count_c = 0U;
        }
        if ((1U & (~ (IData)(vlSelfRef.reset_l)))) {
            ++(vlSelf->__Vcoverage[535]);
        }
        if (vlSelfRef.reset_l) {
            ++(vlSelf->__Vcoverage[536]);
        }
        ++(vlSelf->__Vcoverage[537]);
        vlSelfRef.top__DOT__sub__DOT__count_c = __Vdly__top__DOT__sub__DOT__count_c;
        if ((0U != (vlSelfRef.top__DOT__sub__DOT__count_c 
                    ^ vlSelfRef.top__DOT__sub__DOT____Vtogcov__count_c))) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelf->__Vcoverage + 467, vlSelfRef.top__DOT__sub__DOT__count_c, vlSelfRef.top__DOT__sub__DOT____Vtogcov__count_c);
            vlSelfRef.top__DOT__sub__DOT____Vtogcov__count_c 
                = vlSelfRef.top__DOT__sub__DOT__count_c;
        }
    }

    VLCU_D static void Vtop___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vtop___024root___eval_phase__act(Vtop___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vtop___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vtop___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vtop___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vtop___024root___eval_phase__nba(Vtop___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vtop___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    Vtop___024root___nba_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
            Vtop___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vtop___024root___eval(Vtop___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        IData/*31:0*/ __VicoIterCount;
        IData/*31:0*/ __VnbaIterCount;
        // Body
        {
            // Inlined CFunc: _sample
            vlSelfRef.__Vsampled_TOP__top__DOT__sub__DOT__count_c 
                = vlSelfRef.top__DOT__sub__DOT__count_c;
        }
        __VicoIterCount = 0U;
        do {
            if (VL_UNLIKELY(((0x00002710U < __VicoIterCount)))) {
#ifdef VL_DEBUG
                Vtop___024root___dump_triggers__ico(vlSelfRef.__VicoTriggered, "ico"s);
#endif
                VL_FATAL_MT("/root/reference/examples/make_tracing_c/top.v", 11, "", "DIDNOTCONVERGE: Input combinational region did not converge after '--converge-limit' of 10000 tries");
            }
            __VicoIterCount = ((IData)(1U) + __VicoIterCount);
            vlSelfRef.__VicoPhaseResult = Vtop___024root___eval_phase__ico(vlSelf, vlCtx);
        } while (vlSelfRef.__VicoPhaseResult);
        __VnbaIterCount = 0U;
        do {
            if (VL_UNLIKELY(((0x00002710U < __VnbaIterCount)))) {
#ifdef VL_DEBUG
                Vtop___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("/root/reference/examples/make_tracing_c/top.v", 11, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vtop___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("/root/reference/examples/make_tracing_c/top.v", 11, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vtop___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vtop___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vtop___024root___ctor_var_reset(Vtop___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        vlSelf->reset_l = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 12926602469485068081ull);
        vlSelf->out_small = VL_SCOPED_RAND_RESET_I(2, __VscopeHash, 14267498806251343750ull);
        vlSelf->out_quad = VL_SCOPED_RAND_RESET_Q(40, __VscopeHash, 3771664055402206494ull);
        VL_SCOPED_RAND_RESET_W(70, vlSelf->out_wide, __VscopeHash, 6124341913962757201ull);
        vlSelf->in_small = VL_SCOPED_RAND_RESET_I(2, __VscopeHash, 4925076677430572012ull);
        vlSelf->in_quad = VL_SCOPED_RAND_RESET_Q(40, __VscopeHash, 7333266129396445036ull);
        VL_SCOPED_RAND_RESET_W(70, vlSelf->in_wide, __VscopeHash, 18101672252301227225ull);
        vlSelf->top__DOT____Vtogcov__clk = 0;
        vlSelf->top__DOT____Vtogcov__reset_l = 0;
        vlSelf->top__DOT____Vtogcov__out_small = 0;
        vlSelf->top__DOT____Vtogcov__out_quad = 0;
        VL_ZERO_RESET_W(70, vlSelf->top__DOT____Vtogcov__out_wide);
        vlSelf->top__DOT____Vtogcov__in_small = 0;
        vlSelf->top__DOT____Vtogcov__in_quad = 0;
        VL_ZERO_RESET_W(70, vlSelf->top__DOT____Vtogcov__in_wide);
        vlSelf->top__DOT__sub__DOT__count_c = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 2802051788925154131ull);
        vlSelf->top__DOT__sub__DOT____Vtogcov__count_c = 0;
        vlSelf->__Vsampled_TOP__top__DOT__sub__DOT__count_c = 0;
        for (int __Vi0 = 0; __Vi0 < 1; ++__Vi0) {
            vlSelf->__VstlTriggered[__Vi0] = 0;
        }
        for (int __Vi0 = 0; __Vi0 < 2; ++__Vi0) {
            vlSelf->__VicoTriggered[__Vi0] = 0;
        }
        vlSelf->__Vtrigprevexpr___TOP__clk__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__reset_l__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__in_small__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__in_quad__0 = 0;
        VL_ZERO_RESET_W(70, vlSelf->__Vtrigprevexpr___TOP__in_wide__0);
        vlSelf->__VicoDidInit = 0;
        for (int __Vi0 = 0; __Vi0 < 1; ++__Vi0) {
            vlSelf->__VactTriggered[__Vi0] = 0;
        }
        vlSelf->__Vtrigprevexpr___TOP__clk__1 = 0;
        for (int __Vi0 = 0; __Vi0 < 1; ++__Vi0) {
            vlSelf->__VnbaTriggered[__Vi0] = 0;
        }
    }

    VLCU_D static void Vtop___024root__trace_full_0_sub_0(Vtop___024root* vlSelf, VlcuContext& vlCtx, uint32_t* valsp) {
        auto& vlSelfRef VL_ATTR_UNUSED = *vlSelf;
        vlcu_trace_set(valsp + 0, static_cast<IData>(vlSelfRef.clk));
        vlcu_trace_set(valsp + 1, static_cast<IData>(vlSelfRef.reset_l));
        vlcu_trace_set(valsp + 2, static_cast<IData>(vlSelfRef.out_small));
        vlcu_trace_set(valsp + 3, static_cast<QData>(vlSelfRef.out_quad));
        vlcu_trace_set_w(valsp + 5, 3, (vlSelfRef.out_wide));
        vlcu_trace_set(valsp + 8, static_cast<IData>(vlSelfRef.in_small));
        vlcu_trace_set(valsp + 9, static_cast<QData>(vlSelfRef.in_quad));
        vlcu_trace_set_w(valsp + 11, 3, (vlSelfRef.in_wide));
        vlcu_trace_set(valsp + 14, static_cast<IData>(vlSelfRef.top__DOT__sub__DOT__count_c));
    }

    // tfp->dump(): every traced value stored at its trace code.  Full functions, then the
    // constants, as VerilatedTrace::dump runs them the first time (verilated_trace_imp.h:348-358)
    VLCU_D static void trace_sample(Root* vlSelf, VlcuContext& vlCtx, uint32_t* valsp) {
        (void)vlSelf; (void)vlCtx; (void)valsp;
        Vtop___024root__trace_full_0_sub_0(vlSelf, vlCtx, valsp);
    }

    VLCU_D static void ctor_var_reset(Root* vlSelf, VlcuContext& vlCtx) {
        for (uint32_t b = 0; b < 543U; ++b) vlSelf->__Vcoverage[b] = 0U;
        Vtop___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vtop___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vtop___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vtop___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vtop___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vtop___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vtop__cuda::info() {
    typedef Vtop__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"reset_l", "reset_l", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_reset_l},
        {"out_small", "out_small", VLCU_MEMBER_OUT, 1, 0, 1, 2, 1, 0, T::OFF_out_small},
        {"out_quad", "out_quad", VLCU_MEMBER_OUT, 8, 0, 1, 40, 39, 0, T::OFF_out_quad},
        {"out_wide", "out_wide", VLCU_MEMBER_OUT, 4, 0, 3, 70, 69, 0, T::OFF_out_wide},
        {"in_small", "in_small", VLCU_MEMBER_IN, 1, 0, 1, 2, 1, 0, T::OFF_in_small},
        {"in_quad", "in_quad", VLCU_MEMBER_IN, 8, 0, 1, 40, 39, 0, T::OFF_in_quad},
        {"in_wide", "in_wide", VLCU_MEMBER_IN, 4, 0, 3, 70, 69, 0, T::OFF_in_wide},
        {"top__DOT____VlemCond_0", "top.__VlemCond_0", VLCU_MEMBER_STATE, 1, 0, 1, 2, 1, 0, T::OFF_top__DOT____VlemCond_0},
        {"top__DOT____Vtogcov__clk", "top.__Vtogcov__clk", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_top__DOT____Vtogcov__clk},
        {"top__DOT____Vtogcov__reset_l", "top.__Vtogcov__reset_l", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_top__DOT____Vtogcov__reset_l},
        {"top__DOT____Vtogcov__out_small", "top.__Vtogcov__out_small", VLCU_MEMBER_STATE, 1, 0, 1, 2, 1, 0, T::OFF_top__DOT____Vtogcov__out_small},
        {"top__DOT____Vtogcov__in_small", "top.__Vtogcov__in_small", VLCU_MEMBER_STATE, 1, 0, 1, 2, 1, 0, T::OFF_top__DOT____Vtogcov__in_small},
        {"__VstlFirstIteration", "__VstlFirstIteration", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlFirstIteration},
        {"__VstlPhaseResult", "__VstlPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__Vtrigprevexpr___TOP__reset_l__0", "__Vtrigprevexpr___TOP__reset_l__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__reset_l__0},
        {"__Vtrigprevexpr___TOP__in_small__0", "__Vtrigprevexpr___TOP__in_small__0", VLCU_MEMBER_STATE, 1, 0, 1, 2, 1, 0, T::OFF___Vtrigprevexpr___TOP__in_small__0},
        {"__VicoDidInit", "__VicoDidInit", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VicoDidInit},
        {"__VicoPhaseResult", "__VicoPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VicoPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__1", "__Vtrigprevexpr___TOP__clk__1", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__1},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"top__DOT____VlemCond_2", "top.__VlemCond_2", VLCU_MEMBER_STATE, 4, 0, 3, 70, 69, 0, T::OFF_top__DOT____VlemCond_2},
        {"top__DOT____Vtogcov__out_wide", "top.__Vtogcov__out_wide", VLCU_MEMBER_STATE, 4, 0, 3, 70, 69, 0, T::OFF_top__DOT____Vtogcov__out_wide},
        {"top__DOT____Vtogcov__in_wide", "top.__Vtogcov__in_wide", VLCU_MEMBER_STATE, 4, 0, 3, 70, 69, 0, T::OFF_top__DOT____Vtogcov__in_wide},
        {"top__DOT__sub__DOT__count_c", "top.sub.count_c", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_top__DOT__sub__DOT__count_c},
        {"top__DOT__sub__DOT____Vtogcov__count_c", "top.sub.__Vtogcov__count_c", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_top__DOT__sub__DOT____Vtogcov__count_c},
        {"__Vsampled_TOP__top__DOT__sub__DOT__count_c", "__Vsampled_TOP__top.sub.count_c", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___Vsampled_TOP__top__DOT__sub__DOT__count_c},
        {"__Vtrigprevexpr___TOP__in_wide__0", "__Vtrigprevexpr___TOP__in_wide__0", VLCU_MEMBER_STATE, 4, 0, 3, 70, 69, 0, T::OFF___Vtrigprevexpr___TOP__in_wide__0},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"top__DOT____VlemCond_1", "top.__VlemCond_1", VLCU_MEMBER_STATE, 8, 0, 1, 40, 39, 0, T::OFF_top__DOT____VlemCond_1},
        {"top__DOT____Vtogcov__out_quad", "top.__Vtogcov__out_quad", VLCU_MEMBER_STATE, 8, 0, 1, 40, 39, 0, T::OFF_top__DOT____Vtogcov__out_quad},
        {"top__DOT____Vtogcov__in_quad", "top.__Vtogcov__in_quad", VLCU_MEMBER_STATE, 8, 0, 1, 40, 39, 0, T::OFF_top__DOT____Vtogcov__in_quad},
        {"__Vtrigprevexpr___TOP__in_quad__0", "__Vtrigprevexpr___TOP__in_quad__0", VLCU_MEMBER_STATE, 8, 0, 1, 40, 39, 0, T::OFF___Vtrigprevexpr___TOP__in_quad__0},
        {"__VstlTriggered", "__VstlTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VstlTriggered},
        {"__VicoTriggered", "__VicoTriggered", VLCU_MEMBER_STATE, 8, 0, 2, 64, 63, 0, T::OFF___VicoTriggered},
        {"__VactTriggered", "__VactTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VactTriggered},
        {"__VnbaTriggered", "__VnbaTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VnbaTriggered},
        {"__Vcoverage", nullptr, VLCU_MEMBER_COVERAGE, 4, 0, 543, 32, 31, 0, T::OFF___Vcoverage},
    };
    static VlcuCoverTable t;
    if (t.points.empty()) {
        t.toggle(0, 0, 0, 0, "/root/reference/examples/make_tracing_c/top.v", 13, 11, ".top", "v_toggle/top", "clk");
        t.toggle(2, 0, 0, 0, "/root/reference/examples/make_tracing_c/top.v", 14, 11, ".top", "v_toggle/top", "reset_l");
        t.toggle(4, 0, 1, 1, "/root/reference/examples/make_tracing_c/top.v", 16, 23, ".top", "v_toggle/top", "out_small");
        t.toggle(8, 0, 39, 1, "/root/reference/examples/make_tracing_c/top.v", 17, 24, ".top", "v_toggle/top", "out_quad");
        t.toggle(88, 0, 69, 1, "/root/reference/examples/make_tracing_c/top.v", 18, 24, ".top", "v_toggle/top", "out_wide");
        t.toggle(228, 0, 1, 1, "/root/reference/examples/make_tracing_c/top.v", 19, 17, ".top", "v_toggle/top", "in_small");
        t.toggle(232, 0, 39, 1, "/root/reference/examples/make_tracing_c/top.v", 20, 18, ".top", "v_toggle/top", "in_quad");
        t.toggle(312, 0, 69, 1, "/root/reference/examples/make_tracing_c/top.v", 21, 18, ".top", "v_toggle/top", "in_wide");
        t.other(452, "/root/reference/examples/make_tracing_c/top.v", 25, 22, ".top", "v_expr/top", "(reset_l==0) => 1", "");
        t.other(453, "/root/reference/examples/make_tracing_c/top.v", 25, 22, ".top", "v_expr/top", "(reset_l==1) => 0", "");
        t.other(454, "/root/reference/examples/make_tracing_c/top.v", 25, 33, ".top", "v_branch/top", "cond_then", "25");
        t.other(455, "/root/reference/examples/make_tracing_c/top.v", 25, 34, ".top", "v_branch/top", "cond_else", "25");
        t.other(456, "/root/reference/examples/make_tracing_c/top.v", 26, 21, ".top", "v_expr/top", "(reset_l==0) => 1", "");
        t.other(457, "/root/reference/examples/make_tracing_c/top.v", 26, 21, ".top", "v_expr/top", "(reset_l==1) => 0", "");
        t.other(458, "/root/reference/examples/make_tracing_c/top.v", 26, 32, ".top", "v_branch/top", "cond_then", "26");
        t.other(459, "/root/reference/examples/make_tracing_c/top.v", 26, 33, ".top", "v_branch/top", "cond_else", "26");
        t.other(460, "/root/reference/examples/make_tracing_c/top.v", 27, 21, ".top", "v_expr/top", "(reset_l==0) => 1", "");
        t.other(461, "/root/reference/examples/make_tracing_c/top.v", 27, 21, ".top", "v_expr/top", "(reset_l==1) => 0", "");
        t.other(462, "/root/reference/examples/make_tracing_c/top.v", 27, 32, ".top", "v_branch/top", "cond_then", "27");
        t.other(463, "/root/reference/examples/make_tracing_c/top.v", 27, 33, ".top", "v_branch/top", "cond_else", "27");
        t.other(464, "/root/reference/examples/make_tracing_c/top.v", 37, 5, ".top", "v_branch/top", "if", "37-40");
        t.other(465, "/root/reference/examples/make_tracing_c/top.v", 37, 6, ".top", "v_branch/top", "else", "");
        t.other(466, "/root/reference/examples/make_tracing_c/top.v", 36, 3, ".top", "v_line/top", "block", "36,42");
        t.toggle(0, 0, 0, 0, "/root/reference/examples/make_tracing_c/sub.v", 9, 11, ".top.sub", "v_toggle/sub", "clk");
        t.toggle(2, 0, 0, 0, "/root/reference/examples/make_tracing_c/sub.v", 10, 11, ".top.sub", "v_toggle/sub", "reset_l");
        t.toggle(467, 0, 31, 1, "/root/reference/examples/make_tracing_c/sub.v", 14, 14, ".top.sub", "v_toggle/sub", "count_c");
        t.other(531, "/root/reference/examples/make_tracing_c/sub.v", 24, 7, ".top.sub", "v_branch/sub", "if", "24,27-28");
        t.other(532, "/root/reference/examples/make_tracing_c/sub.v", 24, 8, ".top.sub", "v_branch/sub", "else", "");
        t.other(533, "/root/reference/examples/make_tracing_c/sub.v", 16, 5, ".top.sub", "v_branch/sub", "if", "16,19");
        t.other(534, "/root/reference/examples/make_tracing_c/sub.v", 16, 6, ".top.sub", "v_branch/sub", "else", "22-23");
        t.other(535, "/root/reference/examples/make_tracing_c/sub.v", 16, 9, ".top.sub", "v_expr/sub", "(reset_l==0) => 1", "");
        t.other(536, "/root/reference/examples/make_tracing_c/sub.v", 16, 9, ".top.sub", "v_expr/sub", "(reset_l==1) => 0", "");
        t.other(537, "/root/reference/examples/make_tracing_c/sub.v", 15, 3, ".top.sub", "v_line/sub", "block", "15");
        t.other(538, "/root/reference/examples/make_tracing_c/sub.v", 35, 41, ".top.sub", "v_expr/sub", "((count_c < 32'sh64)==1) => 1", "");
        t.other(539, "/root/reference/examples/make_tracing_c/sub.v", 35, 41, ".top.sub", "v_expr/sub", "(reset_l==0) => 1", "");
        t.other(540, "/root/reference/examples/make_tracing_c/sub.v", 35, 41, ".top.sub", "v_expr/sub", "(reset_l==1 && (count_c < 32'sh64)==0) => 0", "");
        t.other(541, "/root/reference/examples/make_tracing_c/sub.v", 34, 3, ".top.sub", "v_line/sub", "block", "34-35");
        t.other(542, "/root/reference/examples/make_tracing_c/sub.v", 39, 3, ".top.sub", "v_user/sub", "cover", "39");
        t.finish();
    }
    static const char* const messages[] = {"[%0t] Tracing to logs/vlt_dump.vcd...\n\n", "[%0t] Model running...\n\n", "[%0t] %%Error: sub.v:35: Assertion failed in TOP.top.sub.AssertionExample: 'assert' failed.\n", "*-* All Finished *-*\n", nullptr};
    static const char* const plusargs[] = {"trace", nullptr};
    static const VlcuTraceDecl traceDecls[] = {
        {VLCU_TRACE_PUSH, 0, "$rootio", 0, VLCU_PREFIX_SCOPE_MODULE, -1, 0, 0},
        {VLCU_TRACE_SIG, 0, "clk", 0, 1, -1, 0, 0},
        {VLCU_TRACE_SIG, 0, "reset_l", 1, 1, -1, 0, 0},
        {VLCU_TRACE_SIG, 1, "out_small", 2, 2, -1, 1, 0},
        {VLCU_TRACE_SIG, 1, "out_quad", 3, 40, -1, 39, 0},
        {VLCU_TRACE_SIG, 1, "out_wide", 5, 70, -1, 69, 0},
        {VLCU_TRACE_SIG, 1, "in_small", 8, 2, -1, 1, 0},
        {VLCU_TRACE_SIG, 1, "in_quad", 9, 40, -1, 39, 0},
        {VLCU_TRACE_SIG, 1, "in_wide", 11, 70, -1, 69, 0},
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0},
        {VLCU_TRACE_PUSH, 0, "top", 0, VLCU_PREFIX_SCOPE_MODULE, -1, 0, 0},
        {VLCU_TRACE_SIG, 0, "clk", 0, 1, -1, 0, 0},
        {VLCU_TRACE_SIG, 0, "reset_l", 1, 1, -1, 0, 0},
        {VLCU_TRACE_SIG, 1, "out_small", 2, 2, -1, 1, 0},
        {VLCU_TRACE_SIG, 1, "out_quad", 3, 40, -1, 39, 0},
        {VLCU_TRACE_SIG, 1, "out_wide", 5, 70, -1, 69, 0},
        {VLCU_TRACE_SIG, 1, "in_small", 8, 2, -1, 1, 0},
        {VLCU_TRACE_SIG, 1, "in_quad", 9, 40, -1, 39, 0},
        {VLCU_TRACE_SIG, 1, "in_wide", 11, 70, -1, 69, 0},
        {VLCU_TRACE_PUSH, 0, "sub", 0, VLCU_PREFIX_SCOPE_MODULE, -1, 0, 0},
        {VLCU_TRACE_SIG, 0, "clk", 0, 1, -1, 0, 0},
        {VLCU_TRACE_SIG, 0, "reset_l", 1, 1, -1, 0, 0},
        {VLCU_TRACE_SIG, 1, "count_c", 14, 32, -1, 31, 0},
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0},
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0},
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0}};  // (terminator, not counted)
    static const uint32_t traceOrder[] = {0, 1, 2, 3, 5, 8, 9, 11, 14, 0};  // codes in dump order (+ terminator)
    static const uint32_t saveOrder[] = {0, 1, 2, 5, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 4, 7, 23, 24, 25, 26, 27, 28, 29, 30, 3, 6, 31, 32, 33, 34, 35, 36, 37, 38, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "top", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        t.points.data(), static_cast<uint32_t>(t.points.size()), 543, T::OFF___Vcoverage,
        messages, 4, plusargs, 1,
        traceDecls, 25, T::kTraceWords, traceOrder, 9,
        saveOrder, 39, VLCU_SAVE_ACTIVITY, 0x63e469006938c01ULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vtop__cuda VLCU_TESTBENCHES);
