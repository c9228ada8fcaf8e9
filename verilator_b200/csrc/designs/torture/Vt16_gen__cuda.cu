// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt16_gen"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt16_gen___024root final {
    CData/*0:0*/ clk;
    IData/*31:0*/ d;
    IData/*31:0*/ q;
    CData/*7:0*/ x;
    CData/*7:0*/ t16_gen__DOT__lane__BRA__0__KET____DOT__r;
    CData/*7:0*/ t16_gen__DOT__lane__BRA__1__KET____DOT__r;
    CData/*7:0*/ t16_gen__DOT__lane__BRA__2__KET____DOT__r;
    CData/*7:0*/ t16_gen__DOT__lane__BRA__3__KET____DOT__r;
    CData/*0:0*/ __VstlFirstIteration;
    CData/*0:0*/ __VstlPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    IData/*31:0*/ __VactIterCount;
    VlUnpacked<QData/*63:0*/, 1> __VstlTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vt16_gen__cuda final {
    typedef Vt16_gen___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 36,
        OFF_d = 24,
        OFF_q = 28,
        OFF_x = 37,
        OFF_t16_gen__DOT__lane__BRA__0__KET____DOT__r = 38,
        OFF_t16_gen__DOT__lane__BRA__1__KET____DOT__r = 39,
        OFF_t16_gen__DOT__lane__BRA__2__KET____DOT__r = 40,
        OFF_t16_gen__DOT__lane__BRA__3__KET____DOT__r = 41,
        OFF___VstlFirstIteration = 42,
        OFF___VstlPhaseResult = 43,
        OFF___Vtrigprevexpr___TOP__clk__0 = 44,
        OFF___VactPhaseResult = 45,
        OFF___VnbaPhaseResult = 46,
        OFF___VactIterCount = 32,
        OFF___VstlTriggered = 0,
        OFF___VactTriggered = 8,
        OFF___VnbaTriggered = 16,
        kStateBytes = 48,  // S_persist
        kTouchedReadBytes = 33,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 32,  // S_state_wr
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
        s.d = st.ld<IData>(OFF_d, i);
        s.q = st.ld<IData>(OFF_q, i);
        s.x = st.ld<VlcuMem8>(OFF_x, i);
        s.t16_gen__DOT__lane__BRA__0__KET____DOT__r = st.ld<VlcuMem8>(OFF_t16_gen__DOT__lane__BRA__0__KET____DOT__r, i);
        s.t16_gen__DOT__lane__BRA__1__KET____DOT__r = st.ld<VlcuMem8>(OFF_t16_gen__DOT__lane__BRA__1__KET____DOT__r, i);
        s.t16_gen__DOT__lane__BRA__2__KET____DOT__r = st.ld<VlcuMem8>(OFF_t16_gen__DOT__lane__BRA__2__KET____DOT__r, i);
        s.t16_gen__DOT__lane__BRA__3__KET____DOT__r = st.ld<VlcuMem8>(OFF_t16_gen__DOT__lane__BRA__3__KET____DOT__r, i);
        s.__VstlFirstIteration = st.ld<VlcuMem8>(OFF___VstlFirstIteration, i);
        s.__VstlPhaseResult = st.ld<VlcuMem8>(OFF___VstlPhaseResult, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VstlTriggered[e] = st.ldE<QData>(OFF___VstlTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.st<IData>(OFF_d, i, s.d);
        st.st<IData>(OFF_q, i, s.q);
        st.st<VlcuMem8>(OFF_x, i, s.x);
        st.st<VlcuMem8>(OFF_t16_gen__DOT__lane__BRA__0__KET____DOT__r, i, s.t16_gen__DOT__lane__BRA__0__KET____DOT__r);
        st.st<VlcuMem8>(OFF_t16_gen__DOT__lane__BRA__1__KET____DOT__r, i, s.t16_gen__DOT__lane__BRA__1__KET____DOT__r);
        st.st<VlcuMem8>(OFF_t16_gen__DOT__lane__BRA__2__KET____DOT__r, i, s.t16_gen__DOT__lane__BRA__2__KET____DOT__r);
        st.st<VlcuMem8>(OFF_t16_gen__DOT__lane__BRA__3__KET____DOT__r, i, s.t16_gen__DOT__lane__BRA__3__KET____DOT__r);
        st.st<VlcuMem8>(OFF___VstlFirstIteration, i, s.__VstlFirstIteration);
        st.st<VlcuMem8>(OFF___VstlPhaseResult, i, s.__VstlPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VstlTriggered, e, i, s.__VstlTriggered[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VactTriggered, e, i, s.__VactTriggered[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VnbaTriggered, e, i, s.__VnbaTriggered[e]);
    }
    VLCU_D static void load_inputs_row(Root& s, const uint32_t* rowp, uint64_t stride, uint64_t i) {
        s.clk = static_cast<VlcuMem8>(rowp[0 * stride + i] & 0x1U);
        s.d = static_cast<IData>(rowp[1 * stride + i] & 0xffffffffU);
    }
    VLCU_D static void store_outputs_row(const Root& s, uint32_t* rowp, uint64_t stride, uint64_t i) {
        rowp[0 * stride + i] = s.q;
        rowp[1 * stride + i] = s.x;
    }
    VLCU_D static uint64_t outputs_sig(const Root& s, uint64_t sig) {
        uint32_t acc = VLCU_SIG_ACC_INIT;
        acc = vlcu_sig_word(acc, s.q);
        acc = vlcu_sig_word(acc, s.x);
        return vlcu_sig_mix(sig, acc);
    }
    enum : uint32_t { kInWords = 2, kOutWords = 2, kInBytes = 5, kOutBytes = 5, kTraceWords = 0 };

    VLCU_D static void Vt16_gen___024root___eval_static(Vt16_gen___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
    }

    VLCU_D static void Vt16_gen___024root___eval_initial(Vt16_gen___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt16_gen___024root___eval_final(Vt16_gen___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt16_gen___024root___eval_settle(Vt16_gen___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt16_gen___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t16_gen.v", 3, "", "DIDNOTCONVERGE: Settle region did not converge after '--converge-limit' of 10000 tries");
            }
            __VstlIterCount = ((IData)(1U) + __VstlIterCount);
            vlSelfRef.__VstlPhaseResult = Vt16_gen___024root___eval_phase__stl(vlSelf, vlCtx);
            vlSelfRef.__VstlFirstIteration = 0U;
        } while (vlSelfRef.__VstlPhaseResult);
    }

    VLCU_D static bool Vt16_gen___024root___trigger_anySet__stl(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt16_gen___024root___eval_phase__stl(Vt16_gen___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt16_gen___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
        }
#endif
        __VstlExecute = Vt16_gen___024root___trigger_anySet__stl(vlSelfRef.__VstlTriggered);
        if (__VstlExecute) {
            {
                // Inlined CFunc: _eval_stl
                if ((1ULL & vlSelfRef.__VstlTriggered[0U])) {
                    {
                        // Inlined CFunc: _stl_sequent__TOP__0
                        vlSelfRef.q = ((((0x00000900U 
                                          ^ ((IData)(vlSelfRef.t16_gen__DOT__lane__BRA__3__KET____DOT__r) 
                                             << 8U)) 
                                         | (6U ^ (IData)(vlSelfRef.t16_gen__DOT__lane__BRA__2__KET____DOT__r))) 
                                        << 0x00000010U) 
                                       | ((0x00000300U 
                                           ^ ((IData)(vlSelfRef.t16_gen__DOT__lane__BRA__1__KET____DOT__r) 
                                              << 8U)) 
                                          | (IData)(vlSelfRef.t16_gen__DOT__lane__BRA__0__KET____DOT__r)));
                    }
                }
            }
        }
        return (__VstlExecute);
    }

    VLCU_D static bool Vt16_gen___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt16_gen___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt16_gen___024root___eval_phase__act(Vt16_gen___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt16_gen___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vt16_gen___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vt16_gen___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vt16_gen___024root___eval_phase__nba(Vt16_gen___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vt16_gen___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    {
                        // Inlined CFunc: _nba_sequent__TOP__0
                        vlSelfRef.x = (0x000000ffU 
                                       & ((IData)(vlSelfRef.x) 
                                          + vlSelfRef.d));
                        vlSelfRef.t16_gen__DOT__lane__BRA__0__KET____DOT__r 
                            = (0x000000ffU & vlSelfRef.d);
                        vlSelfRef.t16_gen__DOT__lane__BRA__1__KET____DOT__r 
                            = (0x000000ffU & ((IData)(1U) 
                                              + (vlSelfRef.d 
                                                 >> 8U)));
                        vlSelfRef.t16_gen__DOT__lane__BRA__2__KET____DOT__r 
                            = (0x000000ffU & ((IData)(2U) 
                                              + (vlSelfRef.d 
                                                 >> 0x10U)));
                        vlSelfRef.t16_gen__DOT__lane__BRA__3__KET____DOT__r 
                            = (0x000000ffU & ((IData)(3U) 
                                              + (vlSelfRef.d 
                                                 >> 0x18U)));
                        vlSelfRef.q = ((((0x00000900U 
                                          ^ ((IData)(vlSelfRef.t16_gen__DOT__lane__BRA__3__KET____DOT__r) 
                                             << 8U)) 
                                         | (6U ^ (IData)(vlSelfRef.t16_gen__DOT__lane__BRA__2__KET____DOT__r))) 
                                        << 0x00000010U) 
                                       | ((0x00000300U 
                                           ^ ((IData)(vlSelfRef.t16_gen__DOT__lane__BRA__1__KET____DOT__r) 
                                              << 8U)) 
                                          | (IData)(vlSelfRef.t16_gen__DOT__lane__BRA__0__KET____DOT__r)));
                    }
                }
            }
            Vt16_gen___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vt16_gen___024root___eval(Vt16_gen___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt16_gen___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t16_gen.v", 3, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vt16_gen___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t16_gen.v", 3, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vt16_gen___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vt16_gen___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vt16_gen___024root___ctor_var_reset(Vt16_gen___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        vlSelf->d = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 1720370409040345145ull);
        vlSelf->q = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 8861071527689086543ull);
        vlSelf->x = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 9409450202036847209ull);
        vlSelf->t16_gen__DOT__lane__BRA__0__KET____DOT__r = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 9077712670830015917ull);
        vlSelf->t16_gen__DOT__lane__BRA__1__KET____DOT__r = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 12231700974528023125ull);
        vlSelf->t16_gen__DOT__lane__BRA__2__KET____DOT__r = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 17340953662277234518ull);
        vlSelf->t16_gen__DOT__lane__BRA__3__KET____DOT__r = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 1397936141793950169ull);
        for (int __Vi0 = 0; __Vi0 < 1; ++__Vi0) {
            vlSelf->__VstlTriggered[__Vi0] = 0;
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
        Vt16_gen___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt16_gen___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt16_gen___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt16_gen___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vt16_gen___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt16_gen___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt16_gen__cuda::info() {
    typedef Vt16_gen__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"d", "d", VLCU_MEMBER_IN, 4, 0, 1, 32, 31, 0, T::OFF_d},
        {"q", "q", VLCU_MEMBER_OUT, 4, 0, 1, 32, 31, 0, T::OFF_q},
        {"x", "x", VLCU_MEMBER_OUT, 1, 0, 1, 8, 7, 0, T::OFF_x},
        {"t16_gen__DOT__lane__BRA__0__KET____DOT__r", "t16_gen.lane[0].r", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF_t16_gen__DOT__lane__BRA__0__KET____DOT__r},
        {"t16_gen__DOT__lane__BRA__1__KET____DOT__r", "t16_gen.lane[1].r", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF_t16_gen__DOT__lane__BRA__1__KET____DOT__r},
        {"t16_gen__DOT__lane__BRA__2__KET____DOT__r", "t16_gen.lane[2].r", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF_t16_gen__DOT__lane__BRA__2__KET____DOT__r},
        {"t16_gen__DOT__lane__BRA__3__KET____DOT__r", "t16_gen.lane[3].r", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF_t16_gen__DOT__lane__BRA__3__KET____DOT__r},
        {"__VstlFirstIteration", "__VstlFirstIteration", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlFirstIteration},
        {"__VstlPhaseResult", "__VstlPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"__VstlTriggered", "__VstlTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VstlTriggered},
        {"__VactTriggered", "__VactTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VactTriggered},
        {"__VnbaTriggered", "__VnbaTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VnbaTriggered},
    };
    static const char* const messages[] = {nullptr};
    static const char* const plusargs[] = {nullptr};
    static const VlcuTraceDecl traceDecls[] = {
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0}};  // (terminator, not counted)
    static const uint32_t traceOrder[] = {0};  // codes in dump order (+ terminator)
    static const uint32_t saveOrder[] = {0, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 1, 2, 13, 14, 15, 16, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "t16_gen", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 0, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 17, 0, 0x4455e9470139a5e6ULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt16_gen__cuda VLCU_TESTBENCHES);
