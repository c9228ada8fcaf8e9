// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vtop_nocov"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vtop_nocov___024root final {
    CData/*0:0*/ clk;
    CData/*0:0*/ reset_l;
    CData/*1:0*/ out_small;
    QData/*39:0*/ out_quad;
    VlWide<3>/*69:0*/ out_wide;
    CData/*1:0*/ in_small;
    QData/*39:0*/ in_quad;
    VlWide<3>/*69:0*/ in_wide;
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
    IData/*31:0*/ top__DOT__sub__DOT__count_c;
    VlWide<3>/*69:0*/ __Vtrigprevexpr___TOP__in_wide__0;
    IData/*31:0*/ __VactIterCount;
    QData/*39:0*/ __Vtrigprevexpr___TOP__in_quad__0;
    VlUnpacked<QData/*63:0*/, 1> __VstlTriggered;
    VlUnpacked<QData/*63:0*/, 2> __VicoTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vtop_nocov__cuda final {
    typedef Vtop_nocov___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 108,
        OFF_reset_l = 109,
        OFF_out_small = 110,
        OFF_out_quad = 0,
        OFF_out_wide = 64,
        OFF_in_small = 111,
        OFF_in_quad = 8,
        OFF_in_wide = 76,
        OFF___VstlFirstIteration = 112,
        OFF___VstlPhaseResult = 113,
        OFF___Vtrigprevexpr___TOP__clk__0 = 114,
        OFF___Vtrigprevexpr___TOP__reset_l__0 = 115,
        OFF___Vtrigprevexpr___TOP__in_small__0 = 116,
        OFF___VicoDidInit = 117,
        OFF___VicoPhaseResult = 118,
        OFF___Vtrigprevexpr___TOP__clk__1 = 119,
        OFF___VactPhaseResult = 120,
        OFF___VnbaPhaseResult = 121,
        OFF_top__DOT__sub__DOT__count_c = 88,
        OFF___Vtrigprevexpr___TOP__in_wide__0 = 92,
        OFF___VactIterCount = 104,
        OFF___Vtrigprevexpr___TOP__in_quad__0 = 16,
        OFF___VstlTriggered = 24,
        OFF___VicoTriggered = 32,
        OFF___VactTriggered = 48,
        OFF___VnbaTriggered = 56,
        kStateBytes = 128,  // S_persist
        kTouchedReadBytes = 91,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 89,  // S_state_wr
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
        s.reset_l = st.ld<VlcuMem8>(OFF_reset_l, i);
        s.out_small = st.ld<VlcuMem8>(OFF_out_small, i);
        s.out_quad = st.ld<QData>(OFF_out_quad, i);
        st.ldW(OFF_out_wide, i, s.out_wide);
        s.in_small = st.ld<VlcuMem8>(OFF_in_small, i);
        s.in_quad = st.ld<QData>(OFF_in_quad, i);
        st.ldW(OFF_in_wide, i, s.in_wide);
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
        s.top__DOT__sub__DOT__count_c = st.ld<IData>(OFF_top__DOT__sub__DOT__count_c, i);
        st.ldW(OFF___Vtrigprevexpr___TOP__in_wide__0, i, s.__Vtrigprevexpr___TOP__in_wide__0);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
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
        st.st<IData>(OFF_top__DOT__sub__DOT__count_c, i, s.top__DOT__sub__DOT__count_c);
        st.stW(OFF___Vtrigprevexpr___TOP__in_wide__0, i, s.__Vtrigprevexpr___TOP__in_wide__0);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
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

    VLCU_D static void Vtop_nocov___024root___eval_static(Vtop_nocov___024root* vlSelf, VlcuContext& vlCtx) {
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

    VLCU_D static void Vtop_nocov___024root___eval_initial(Vtop_nocov___024root* vlSelf, VlcuContext& vlCtx) {
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
            }
            vlCtx.writef(1U, 1U, vlcu_msg_arg(VL_TIME_UNITED_Q(1)));
        }
    }

    VLCU_D static void Vtop_nocov___024root___eval_final(Vtop_nocov___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vtop_nocov___024root___eval_settle(Vtop_nocov___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vtop_nocov___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
#endif
                VL_FATAL_MT("/root/reference/examples/make_tracing_c/top.v", 11, "", "DIDNOTCONVERGE: Settle region did not converge after '--converge-limit' of 10000 tries");
            }
            __VstlIterCount = ((IData)(1U) + __VstlIterCount);
            vlSelfRef.__VstlPhaseResult = Vtop_nocov___024root___eval_phase__stl(vlSelf, vlCtx);
            vlSelfRef.__VstlFirstIteration = 0U;
        } while (vlSelfRef.__VstlPhaseResult);
    }

    VLCU_D static bool Vtop_nocov___024root___trigger_anySet__stl(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vtop_nocov___024root___eval_phase__stl(Vtop_nocov___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vtop_nocov___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
        }
#endif
        __VstlExecute = Vtop_nocov___024root___trigger_anySet__stl(vlSelfRef.__VstlTriggered);
        if (__VstlExecute) {
            {
                // Inlined CFunc: _eval_stl
                if ((1ULL & vlSelfRef.__VstlTriggered[0U])) {
                    {
                        // Inlined CFunc: _stl_sequent__TOP__0
                        VlWide<3>/*95:0*/ __Vinline_0__eval_stl___Vinline_0__stl_sequent__TOP__0___Vtemp_2;
                        VlWide<3>/*95:0*/ __Vinline_0__eval_stl___Vinline_0__stl_sequent__TOP__0___Vtemp_3;
                        vlSelfRef.out_small = (3U & 
                                               (((IData)(1U) 
                                                 + (IData)(vlSelfRef.in_small)) 
                                                & (- (IData)((IData)(vlSelfRef.reset_l)))));
                        vlSelfRef.out_quad = (0x000000ffffffffffULL 
                                              & ((1ULL 
                                                  + vlSelfRef.in_quad) 
                                                 & (- (QData)((IData)(vlSelfRef.reset_l)))));
                        if (vlSelfRef.reset_l) {
                            __Vinline_0__eval_stl___Vinline_0__stl_sequent__TOP__0___Vtemp_3[0U] = 1U;
                            __Vinline_0__eval_stl___Vinline_0__stl_sequent__TOP__0___Vtemp_3[1U] = 0U;
                            __Vinline_0__eval_stl___Vinline_0__stl_sequent__TOP__0___Vtemp_3[2U] = 0U;
                            VL_ADD_W(3, __Vinline_0__eval_stl___Vinline_0__stl_sequent__TOP__0___Vtemp_2, __Vinline_0__eval_stl___Vinline_0__stl_sequent__TOP__0___Vtemp_3, vlSelfRef.in_wide);
                        } else {
                            __Vinline_0__eval_stl___Vinline_0__stl_sequent__TOP__0___Vtemp_2[0U] = 0U;
                            __Vinline_0__eval_stl___Vinline_0__stl_sequent__TOP__0___Vtemp_2[1U] = 0U;
                            __Vinline_0__eval_stl___Vinline_0__stl_sequent__TOP__0___Vtemp_2[2U] = 0U;
                        }
                        vlSelfRef.out_wide[0U] = __Vinline_0__eval_stl___Vinline_0__stl_sequent__TOP__0___Vtemp_2[0U];
                        vlSelfRef.out_wide[1U] = __Vinline_0__eval_stl___Vinline_0__stl_sequent__TOP__0___Vtemp_2[1U];
                        vlSelfRef.out_wide[2U] = (0x0000003fU 
                                                  & __Vinline_0__eval_stl___Vinline_0__stl_sequent__TOP__0___Vtemp_2[2U]);
                    }
                }
            }
        }
        return (__VstlExecute);
    }

    VLCU_D static bool Vtop_nocov___024root___trigger_anySet__ico(const VlUnpacked<QData/*63:0*/, 2> &in) {
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

    VLCU_D static bool Vtop_nocov___024root___eval_phase__ico(Vtop_nocov___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vtop_nocov___024root___dump_triggers__ico(vlSelfRef.__VicoTriggered, "ico"s);
        }
#endif
        __VicoExecute = Vtop_nocov___024root___trigger_anySet__ico(vlSelfRef.__VicoTriggered);
        if (__VicoExecute) {
            {
                // Inlined CFunc: _eval_ico
                if ((6ULL & vlSelfRef.__VicoTriggered[0U])) {
                    {
                        // Inlined CFunc: _ico_comb__TOP__0
                        vlSelfRef.out_small = (3U & 
                                               (((IData)(1U) 
                                                 + (IData)(vlSelfRef.in_small)) 
                                                & (- (IData)((IData)(vlSelfRef.reset_l)))));
                    }
                }
                if ((0x000000000000000aULL & vlSelfRef.__VicoTriggered[0U])) {
                    {
                        // Inlined CFunc: _ico_comb__TOP__1
                        vlSelfRef.out_quad = (0x000000ffffffffffULL 
                                              & ((1ULL 
                                                  + vlSelfRef.in_quad) 
                                                 & (- (QData)((IData)(vlSelfRef.reset_l)))));
                    }
                }
                if ((0x0000000000000012ULL & vlSelfRef.__VicoTriggered[0U])) {
                    {
                        // Inlined CFunc: _ico_comb__TOP__2
                        VlWide<3>/*95:0*/ __Vinline_0__eval_ico___Vinline_0__ico_comb__TOP__2___Vtemp_2;
                        VlWide<3>/*95:0*/ __Vinline_0__eval_ico___Vinline_0__ico_comb__TOP__2___Vtemp_3;
                        if (vlSelfRef.reset_l) {
                            __Vinline_0__eval_ico___Vinline_0__ico_comb__TOP__2___Vtemp_3[0U] = 1U;
                            __Vinline_0__eval_ico___Vinline_0__ico_comb__TOP__2___Vtemp_3[1U] = 0U;
                            __Vinline_0__eval_ico___Vinline_0__ico_comb__TOP__2___Vtemp_3[2U] = 0U;
                            VL_ADD_W(3, __Vinline_0__eval_ico___Vinline_0__ico_comb__TOP__2___Vtemp_2, __Vinline_0__eval_ico___Vinline_0__ico_comb__TOP__2___Vtemp_3, vlSelfRef.in_wide);
                        } else {
                            __Vinline_0__eval_ico___Vinline_0__ico_comb__TOP__2___Vtemp_2[0U] = 0U;
                            __Vinline_0__eval_ico___Vinline_0__ico_comb__TOP__2___Vtemp_2[1U] = 0U;
                            __Vinline_0__eval_ico___Vinline_0__ico_comb__TOP__2___Vtemp_2[2U] = 0U;
                        }
                        vlSelfRef.out_wide[0U] = __Vinline_0__eval_ico___Vinline_0__ico_comb__TOP__2___Vtemp_2[0U];
                        vlSelfRef.out_wide[1U] = __Vinline_0__eval_ico___Vinline_0__ico_comb__TOP__2___Vtemp_2[1U];
                        vlSelfRef.out_wide[2U] = (0x0000003fU 
                                                  & __Vinline_0__eval_ico___Vinline_0__ico_comb__TOP__2___Vtemp_2[2U]);
                    }
                }
            }
        }
        return (__VicoExecute);
    }

    VLCU_D static bool Vtop_nocov___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vtop_nocov___024root___nba_sequent__TOP__0(Vtop_nocov___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        IData/*31:0*/ __Vdly__top__DOT__sub__DOT__count_c;
        __Vdly__top__DOT__sub__DOT__count_c = 0;
        // Body
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
        __Vdly__top__DOT__sub__DOT__count_c = vlSelfRef.top__DOT__sub__DOT__count_c;
        if (vlSelfRef.reset_l) {
            __Vdly__top__DOT__sub__DOT__count_c = ((IData)(1U) 
                                                   + vlSelfRef.top__DOT__sub__DOT__count_c);
            if (VL_UNLIKELY(((3U <= vlSelfRef.top__DOT__sub__DOT__count_c)))) {
                vlCtx.writef(3U, 0U);
                VL_FINISH_MT("/root/reference/examples/make_tracing_c/sub.v", 28, "");
            }
        } else {
            __Vdly__top__DOT__sub__DOT__count_c = 0U;
        }
        vlSelfRef.top__DOT__sub__DOT__count_c = __Vdly__top__DOT__sub__DOT__count_c;
    }

    VLCU_D static void Vtop_nocov___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vtop_nocov___024root___eval_phase__act(Vtop_nocov___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vtop_nocov___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vtop_nocov___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vtop_nocov___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vtop_nocov___024root___eval_phase__nba(Vtop_nocov___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vtop_nocov___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    Vtop_nocov___024root___nba_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
            Vtop_nocov___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vtop_nocov___024root___eval(Vtop_nocov___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vtop_nocov___024root___dump_triggers__ico(vlSelfRef.__VicoTriggered, "ico"s);
#endif
                VL_FATAL_MT("/root/reference/examples/make_tracing_c/top.v", 11, "", "DIDNOTCONVERGE: Input combinational region did not converge after '--converge-limit' of 10000 tries");
            }
            __VicoIterCount = ((IData)(1U) + __VicoIterCount);
            vlSelfRef.__VicoPhaseResult = Vtop_nocov___024root___eval_phase__ico(vlSelf, vlCtx);
        } while (vlSelfRef.__VicoPhaseResult);
        __VnbaIterCount = 0U;
        do {
            if (VL_UNLIKELY(((0x00002710U < __VnbaIterCount)))) {
#ifdef VL_DEBUG
                Vtop_nocov___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("/root/reference/examples/make_tracing_c/top.v", 11, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vtop_nocov___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("/root/reference/examples/make_tracing_c/top.v", 11, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vtop_nocov___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vtop_nocov___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vtop_nocov___024root___ctor_var_reset(Vtop_nocov___024root* vlSelf, VlcuContext& vlCtx) {
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
        vlSelf->top__DOT__sub__DOT__count_c = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 2802051788925154131ull);
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

    VLCU_D static void Vtop_nocov___024root__trace_full_0_sub_0(Vtop_nocov___024root* vlSelf, VlcuContext& vlCtx, uint32_t* valsp) {
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
        Vtop_nocov___024root__trace_full_0_sub_0(vlSelf, vlCtx, valsp);
    }

    VLCU_D static void ctor_var_reset(Root* vlSelf, VlcuContext& vlCtx) {
        Vtop_nocov___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vtop_nocov___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vtop_nocov___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vtop_nocov___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vtop_nocov___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vtop_nocov___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vtop_nocov__cuda::info() {
    typedef Vtop_nocov__cuda T;
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
        {"top__DOT__sub__DOT__count_c", "top.sub.count_c", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_top__DOT__sub__DOT__count_c},
        {"__Vtrigprevexpr___TOP__in_wide__0", "__Vtrigprevexpr___TOP__in_wide__0", VLCU_MEMBER_STATE, 4, 0, 3, 70, 69, 0, T::OFF___Vtrigprevexpr___TOP__in_wide__0},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"__Vtrigprevexpr___TOP__in_quad__0", "__Vtrigprevexpr___TOP__in_quad__0", VLCU_MEMBER_STATE, 8, 0, 1, 40, 39, 0, T::OFF___Vtrigprevexpr___TOP__in_quad__0},
        {"__VstlTriggered", "__VstlTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VstlTriggered},
        {"__VicoTriggered", "__VicoTriggered", VLCU_MEMBER_STATE, 8, 0, 2, 64, 63, 0, T::OFF___VicoTriggered},
        {"__VactTriggered", "__VactTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VactTriggered},
        {"__VnbaTriggered", "__VnbaTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VnbaTriggered},
    };
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
    static const uint32_t saveOrder[] = {0, 1, 2, 5, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 4, 7, 18, 19, 20, 3, 6, 21, 22, 23, 24, 25, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "top", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 4, plusargs, 1,
        traceDecls, 25, T::kTraceWords, traceOrder, 9,
        saveOrder, 26, VLCU_SAVE_ACTIVITY, 0xa090add8b4f5237eULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vtop_nocov__cuda VLCU_TESTBENCHES);
