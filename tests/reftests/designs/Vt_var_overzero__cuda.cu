// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt_var_overzero"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt_var_overzero___024root final {
    CData/*0:0*/ clk;
    CData/*0:0*/ rstn;
    CData/*7:0*/ dout;
    CData/*7:0*/ dval0;
    CData/*7:0*/ dval1;
    CData/*7:0*/ t__DOT__tsub__DOT__dbgsel;
    CData/*7:0*/ t__DOT__tsub__DOT__sub0__DOT__dbgsel_d1r;
    CData/*0:0*/ __VstlFirstIteration;
    CData/*0:0*/ __VstlPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*0:0*/ __Vtrigprevexpr___TOP__rstn__0;
    CData/*7:0*/ __Vtrigprevexpr___TOP__dval0__0;
    CData/*7:0*/ __Vtrigprevexpr___TOP__dval1__0;
    CData/*0:0*/ __VicoDidInit;
    CData/*0:0*/ __VicoPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__1;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    IData/*31:0*/ __VactIterCount;
    VlUnpacked<QData/*63:0*/, 1> __VstlTriggered;
    VlUnpacked<QData/*63:0*/, 2> __VicoTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vt_var_overzero__cuda final {
    typedef Vt_var_overzero___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 44,
        OFF_rstn = 46,
        OFF_dout = 45,
        OFF_dval0 = 47,
        OFF_dval1 = 48,
        OFF_t__DOT__tsub__DOT__dbgsel = 49,
        OFF_t__DOT__tsub__DOT__sub0__DOT__dbgsel_d1r = 50,
        OFF___VstlFirstIteration = 51,
        OFF___VstlPhaseResult = 52,
        OFF___Vtrigprevexpr___TOP__clk__0 = 53,
        OFF___Vtrigprevexpr___TOP__rstn__0 = 54,
        OFF___Vtrigprevexpr___TOP__dval0__0 = 55,
        OFF___Vtrigprevexpr___TOP__dval1__0 = 56,
        OFF___VicoDidInit = 57,
        OFF___VicoPhaseResult = 58,
        OFF___Vtrigprevexpr___TOP__clk__1 = 59,
        OFF___VactPhaseResult = 60,
        OFF___VnbaPhaseResult = 61,
        OFF___VactIterCount = 40,
        OFF___VstlTriggered = 0,
        OFF___VicoTriggered = 8,
        OFF___VactTriggered = 24,
        OFF___VnbaTriggered = 32,
        kStateBytes = 64,  // S_persist
        kTouchedReadBytes = 51,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 48,  // S_state_wr
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
        s.rstn = st.ld<VlcuMem8>(OFF_rstn, i);
        s.dout = st.ld<VlcuMem8>(OFF_dout, i);
        s.dval0 = st.ld<VlcuMem8>(OFF_dval0, i);
        s.dval1 = st.ld<VlcuMem8>(OFF_dval1, i);
        s.t__DOT__tsub__DOT__dbgsel = st.ld<VlcuMem8>(OFF_t__DOT__tsub__DOT__dbgsel, i);
        s.t__DOT__tsub__DOT__sub0__DOT__dbgsel_d1r = st.ld<VlcuMem8>(OFF_t__DOT__tsub__DOT__sub0__DOT__dbgsel_d1r, i);
        s.__VstlFirstIteration = st.ld<VlcuMem8>(OFF___VstlFirstIteration, i);
        s.__VstlPhaseResult = st.ld<VlcuMem8>(OFF___VstlPhaseResult, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__Vtrigprevexpr___TOP__rstn__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__rstn__0, i);
        s.__Vtrigprevexpr___TOP__dval0__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__dval0__0, i);
        s.__Vtrigprevexpr___TOP__dval1__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__dval1__0, i);
        s.__VicoDidInit = st.ld<VlcuMem8>(OFF___VicoDidInit, i);
        s.__VicoPhaseResult = st.ld<VlcuMem8>(OFF___VicoPhaseResult, i);
        s.__Vtrigprevexpr___TOP__clk__1 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__1, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VstlTriggered[e] = st.ldE<QData>(OFF___VstlTriggered, e, i);
        for (uint32_t e = 0; e < 2; ++e) s.__VicoTriggered[e] = st.ldE<QData>(OFF___VicoTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.st<VlcuMem8>(OFF_rstn, i, s.rstn);
        st.st<VlcuMem8>(OFF_dout, i, s.dout);
        st.st<VlcuMem8>(OFF_dval0, i, s.dval0);
        st.st<VlcuMem8>(OFF_dval1, i, s.dval1);
        st.st<VlcuMem8>(OFF_t__DOT__tsub__DOT__dbgsel, i, s.t__DOT__tsub__DOT__dbgsel);
        st.st<VlcuMem8>(OFF_t__DOT__tsub__DOT__sub0__DOT__dbgsel_d1r, i, s.t__DOT__tsub__DOT__sub0__DOT__dbgsel_d1r);
        st.st<VlcuMem8>(OFF___VstlFirstIteration, i, s.__VstlFirstIteration);
        st.st<VlcuMem8>(OFF___VstlPhaseResult, i, s.__VstlPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__rstn__0, i, s.__Vtrigprevexpr___TOP__rstn__0);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__dval0__0, i, s.__Vtrigprevexpr___TOP__dval0__0);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__dval1__0, i, s.__Vtrigprevexpr___TOP__dval1__0);
        st.st<VlcuMem8>(OFF___VicoDidInit, i, s.__VicoDidInit);
        st.st<VlcuMem8>(OFF___VicoPhaseResult, i, s.__VicoPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__1, i, s.__Vtrigprevexpr___TOP__clk__1);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VstlTriggered, e, i, s.__VstlTriggered[e]);
        for (uint32_t e = 0; e < 2; ++e) st.stE<QData>(OFF___VicoTriggered, e, i, s.__VicoTriggered[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VactTriggered, e, i, s.__VactTriggered[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VnbaTriggered, e, i, s.__VnbaTriggered[e]);
    }
    VLCU_D static void load_inputs_row(Root& s, const uint32_t* rowp, uint64_t stride, uint64_t i) {
        s.clk = static_cast<VlcuMem8>(rowp[0 * stride + i] & 0x1U);
        s.rstn = static_cast<VlcuMem8>(rowp[1 * stride + i] & 0x1U);
        s.dval0 = static_cast<VlcuMem8>(rowp[2 * stride + i] & 0xffU);
        s.dval1 = static_cast<VlcuMem8>(rowp[3 * stride + i] & 0xffU);
    }
    VLCU_D static void store_outputs_row(const Root& s, uint32_t* rowp, uint64_t stride, uint64_t i) {
        rowp[0 * stride + i] = s.dout;
    }
    VLCU_D static uint64_t outputs_sig(const Root& s, uint64_t sig) {
        uint32_t acc = VLCU_SIG_ACC_INIT;
        acc = vlcu_sig_word(acc, s.dout);
        return vlcu_sig_mix(sig, acc);
    }
    enum : uint32_t { kInWords = 4, kOutWords = 1, kInBytes = 4, kOutBytes = 1, kTraceWords = 0 };

    VLCU_D static void Vt_var_overzero___024root___eval_static(Vt_var_overzero___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
        vlSelfRef.__Vtrigprevexpr___TOP__rstn__0 = vlSelfRef.rstn;
        vlSelfRef.__Vtrigprevexpr___TOP__dval0__0 = vlSelfRef.dval0;
        vlSelfRef.__Vtrigprevexpr___TOP__dval1__0 = vlSelfRef.dval1;
        vlSelfRef.__Vtrigprevexpr___TOP__clk__1 = vlSelfRef.clk;
    }

    VLCU_D static void Vt_var_overzero___024root___eval_initial(Vt_var_overzero___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        {
            // Inlined CFunc: _eval_initial__TOP
            vlCtx.writef(0U, 0U);
            VL_FINISH_MT("/root/reference/test_regress/t/t_var_overzero.v", 66, "");
        }
    }

    VLCU_D static void Vt_var_overzero___024root___eval_final(Vt_var_overzero___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_var_overzero___024root___eval_settle(Vt_var_overzero___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt_var_overzero___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
#endif
                VL_FATAL_MT("/root/reference/test_regress/t/t_var_overzero.v", 7, "", "DIDNOTCONVERGE: Settle region did not converge after '--converge-limit' of 10000 tries");
            }
            __VstlIterCount = ((IData)(1U) + __VstlIterCount);
            vlSelfRef.__VstlPhaseResult = Vt_var_overzero___024root___eval_phase__stl(vlSelf, vlCtx);
            vlSelfRef.__VstlFirstIteration = 0U;
        } while (vlSelfRef.__VstlPhaseResult);
    }

    VLCU_D static bool Vt_var_overzero___024root___trigger_anySet__stl(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt_var_overzero___024root___eval_phase__stl(Vt_var_overzero___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt_var_overzero___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
        }
#endif
        __VstlExecute = Vt_var_overzero___024root___trigger_anySet__stl(vlSelfRef.__VstlTriggered);
        if (__VstlExecute) {
            {
                // Inlined CFunc: _eval_stl
                if ((1ULL & vlSelfRef.__VstlTriggered[0U])) {
                    {
                        // Inlined CFunc: _ico_comb__TOP__0
                        vlSelfRef.dout = (((IData)(vlSelfRef.dval0) 
                                           & (- (IData)(
                                                        ((0x22U 
                                                          <= (IData)(vlSelfRef.t__DOT__tsub__DOT__sub0__DOT__dbgsel_d1r)) 
                                                         & (0x41U 
                                                            > (IData)(vlSelfRef.t__DOT__tsub__DOT__sub0__DOT__dbgsel_d1r)))))) 
                                          | ((IData)(vlSelfRef.dval1) 
                                             & (- (IData)(
                                                          ((0x54U 
                                                            <= (IData)(vlSelfRef.t__DOT__tsub__DOT__sub0__DOT__dbgsel_d1r)) 
                                                           & (0x5fU 
                                                              > (IData)(vlSelfRef.t__DOT__tsub__DOT__sub0__DOT__dbgsel_d1r)))))));
                    }
                }
            }
        }
        return (__VstlExecute);
    }

    VLCU_D static bool Vt_var_overzero___024root___trigger_anySet__ico(const VlUnpacked<QData/*63:0*/, 2> &in) {
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

    VLCU_D static bool Vt_var_overzero___024root___eval_phase__ico(Vt_var_overzero___024root* vlSelf, VlcuContext& vlCtx) {
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
                                                            (((((IData)(vlSelfRef.dval1) 
                                                                != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__dval1__0)) 
                                                               << 3U) 
                                                              | (((IData)(vlSelfRef.dval0) 
                                                                  != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__dval0__0)) 
                                                                 << 2U)) 
                                                             | ((((IData)(vlSelfRef.rstn) 
                                                                  != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__rstn__0)) 
                                                                 << 1U) 
                                                                | ((IData)(vlSelfRef.clk) 
                                                                   != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__clk__0))))));
            vlSelfRef.__Vtrigprevexpr___TOP__clk__0 
                = vlSelfRef.clk;
            vlSelfRef.__Vtrigprevexpr___TOP__rstn__0 
                = vlSelfRef.rstn;
            vlSelfRef.__Vtrigprevexpr___TOP__dval0__0 
                = vlSelfRef.dval0;
            vlSelfRef.__Vtrigprevexpr___TOP__dval1__0 
                = vlSelfRef.dval1;
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
            }
        }
#ifdef VL_DEBUG
        if (VL_UNLIKELY(vlSymsp->_vm_contextp__->debug())) {
            Vt_var_overzero___024root___dump_triggers__ico(vlSelfRef.__VicoTriggered, "ico"s);
        }
#endif
        __VicoExecute = Vt_var_overzero___024root___trigger_anySet__ico(vlSelfRef.__VicoTriggered);
        if (__VicoExecute) {
            {
                // Inlined CFunc: _eval_ico
                if ((0x000000000000000cULL & vlSelfRef.__VicoTriggered[0U])) {
                    {
                        // Inlined CFunc: _ico_comb__TOP__0
                        vlSelfRef.dout = (((IData)(vlSelfRef.dval0) 
                                           & (- (IData)(
                                                        ((0x22U 
                                                          <= (IData)(vlSelfRef.t__DOT__tsub__DOT__sub0__DOT__dbgsel_d1r)) 
                                                         & (0x41U 
                                                            > (IData)(vlSelfRef.t__DOT__tsub__DOT__sub0__DOT__dbgsel_d1r)))))) 
                                          | ((IData)(vlSelfRef.dval1) 
                                             & (- (IData)(
                                                          ((0x54U 
                                                            <= (IData)(vlSelfRef.t__DOT__tsub__DOT__sub0__DOT__dbgsel_d1r)) 
                                                           & (0x5fU 
                                                              > (IData)(vlSelfRef.t__DOT__tsub__DOT__sub0__DOT__dbgsel_d1r)))))));
                    }
                }
            }
        }
        return (__VicoExecute);
    }

    VLCU_D static bool Vt_var_overzero___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt_var_overzero___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt_var_overzero___024root___eval_phase__act(Vt_var_overzero___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt_var_overzero___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vt_var_overzero___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vt_var_overzero___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vt_var_overzero___024root___eval_phase__nba(Vt_var_overzero___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vt_var_overzero___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    {
                        // Inlined CFunc: _nba_sequent__TOP__0
                        vlSelfRef.t__DOT__tsub__DOT__sub0__DOT__dbgsel_d1r 
                            = ((IData)(vlSelfRef.rstn)
                                ? (IData)(vlSelfRef.t__DOT__tsub__DOT__dbgsel)
                                : 0U);
                        vlSelfRef.dout = (((IData)(vlSelfRef.dval0) 
                                           & (- (IData)(
                                                        ((0x22U 
                                                          <= (IData)(vlSelfRef.t__DOT__tsub__DOT__sub0__DOT__dbgsel_d1r)) 
                                                         & (0x41U 
                                                            > (IData)(vlSelfRef.t__DOT__tsub__DOT__sub0__DOT__dbgsel_d1r)))))) 
                                          | ((IData)(vlSelfRef.dval1) 
                                             & (- (IData)(
                                                          ((0x54U 
                                                            <= (IData)(vlSelfRef.t__DOT__tsub__DOT__sub0__DOT__dbgsel_d1r)) 
                                                           & (0x5fU 
                                                              > (IData)(vlSelfRef.t__DOT__tsub__DOT__sub0__DOT__dbgsel_d1r)))))));
                        vlSelfRef.t__DOT__tsub__DOT__dbgsel = 0U;
                    }
                }
            }
            Vt_var_overzero___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vt_var_overzero___024root___eval(Vt_var_overzero___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt_var_overzero___024root___dump_triggers__ico(vlSelfRef.__VicoTriggered, "ico"s);
#endif
                VL_FATAL_MT("/root/reference/test_regress/t/t_var_overzero.v", 7, "", "DIDNOTCONVERGE: Input combinational region did not converge after '--converge-limit' of 10000 tries");
            }
            __VicoIterCount = ((IData)(1U) + __VicoIterCount);
            vlSelfRef.__VicoPhaseResult = Vt_var_overzero___024root___eval_phase__ico(vlSelf, vlCtx);
        } while (vlSelfRef.__VicoPhaseResult);
        __VnbaIterCount = 0U;
        do {
            if (VL_UNLIKELY(((0x00002710U < __VnbaIterCount)))) {
#ifdef VL_DEBUG
                Vt_var_overzero___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("/root/reference/test_regress/t/t_var_overzero.v", 7, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vt_var_overzero___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("/root/reference/test_regress/t/t_var_overzero.v", 7, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vt_var_overzero___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vt_var_overzero___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vt_var_overzero___024root___ctor_var_reset(Vt_var_overzero___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->dout = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 11474705599699299244ull);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        vlSelf->rstn = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 5377340664288042355ull);
        vlSelf->dval0 = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 7254528187089584200ull);
        vlSelf->dval1 = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 16663259379624632164ull);
        vlSelf->t__DOT__tsub__DOT__dbgsel = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 17386657774991404647ull);
        vlSelf->t__DOT__tsub__DOT__sub0__DOT__dbgsel_d1r = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 1124118487888871192ull);
        for (int __Vi0 = 0; __Vi0 < 1; ++__Vi0) {
            vlSelf->__VstlTriggered[__Vi0] = 0;
        }
        for (int __Vi0 = 0; __Vi0 < 2; ++__Vi0) {
            vlSelf->__VicoTriggered[__Vi0] = 0;
        }
        vlSelf->__Vtrigprevexpr___TOP__clk__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__rstn__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__dval0__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__dval1__0 = 0;
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
        Vt_var_overzero___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt_var_overzero___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt_var_overzero___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt_var_overzero___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vt_var_overzero___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt_var_overzero___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt_var_overzero__cuda::info() {
    typedef Vt_var_overzero__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"rstn", "rstn", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_rstn},
        {"dout", "dout", VLCU_MEMBER_OUT, 1, 0, 1, 8, 7, 0, T::OFF_dout},
        {"dval0", "dval0", VLCU_MEMBER_IN, 1, 0, 1, 8, 7, 0, T::OFF_dval0},
        {"dval1", "dval1", VLCU_MEMBER_IN, 1, 0, 1, 8, 7, 0, T::OFF_dval1},
        {"t__DOT__tsub__DOT__dbgsel", "t.tsub.dbgsel", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF_t__DOT__tsub__DOT__dbgsel},
        {"t__DOT__tsub__DOT__sub0__DOT__dbgsel_d1r", "t.tsub.sub0.dbgsel_d1r", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF_t__DOT__tsub__DOT__sub0__DOT__dbgsel_d1r},
        {"__VstlFirstIteration", "__VstlFirstIteration", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlFirstIteration},
        {"__VstlPhaseResult", "__VstlPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__Vtrigprevexpr___TOP__rstn__0", "__Vtrigprevexpr___TOP__rstn__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__rstn__0},
        {"__Vtrigprevexpr___TOP__dval0__0", "__Vtrigprevexpr___TOP__dval0__0", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF___Vtrigprevexpr___TOP__dval0__0},
        {"__Vtrigprevexpr___TOP__dval1__0", "__Vtrigprevexpr___TOP__dval1__0", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF___Vtrigprevexpr___TOP__dval1__0},
        {"__VicoDidInit", "__VicoDidInit", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VicoDidInit},
        {"__VicoPhaseResult", "__VicoPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VicoPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__1", "__Vtrigprevexpr___TOP__clk__1", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__1},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"__VstlTriggered", "__VstlTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VstlTriggered},
        {"__VicoTriggered", "__VicoTriggered", VLCU_MEMBER_STATE, 8, 0, 2, 64, 63, 0, T::OFF___VicoTriggered},
        {"__VactTriggered", "__VactTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VactTriggered},
        {"__VnbaTriggered", "__VnbaTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VnbaTriggered},
    };
    static const char* const messages[] = {"*-* All Finished *-*\n", nullptr};
    static const char* const plusargs[] = {nullptr};
    static const VlcuTraceDecl traceDecls[] = {
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0}};  // (terminator, not counted)
    static const uint32_t traceOrder[] = {0};  // codes in dump order (+ terminator)
    static const uint32_t saveOrder[] = {0, 2, 1, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "t", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 1, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 23, 0, 0xc3119a88b7726f52ULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt_var_overzero__cuda VLCU_TESTBENCHES);
