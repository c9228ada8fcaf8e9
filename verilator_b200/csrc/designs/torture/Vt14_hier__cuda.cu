// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt14_hier"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt14_hier___024root final {
    CData/*0:0*/ clk;
    CData/*0:0*/ rst;
    IData/*31:0*/ x;
    CData/*7:0*/ a0;
    SData/*15:0*/ a1;
    QData/*39:0*/ a2;
    CData/*3:0*/ msbs;
    CData/*7:0*/ votes;
    CData/*7:0*/ t14_hier__DOT__a3;
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

struct Vt14_hier__cuda final {
    typedef Vt14_hier___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 42,
        OFF_rst = 43,
        OFF_x = 32,
        OFF_a0 = 44,
        OFF_a1 = 40,
        OFF_a2 = 0,
        OFF_msbs = 45,
        OFF_votes = 46,
        OFF_t14_hier__DOT__a3 = 47,
        OFF___VstlFirstIteration = 48,
        OFF___VstlPhaseResult = 49,
        OFF___Vtrigprevexpr___TOP__clk__0 = 50,
        OFF___VactPhaseResult = 51,
        OFF___VnbaPhaseResult = 52,
        OFF___VactIterCount = 36,
        OFF___VstlTriggered = 8,
        OFF___VactTriggered = 16,
        OFF___VnbaTriggered = 24,
        kStateBytes = 56,  // S_persist
        kTouchedReadBytes = 43,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 37,  // S_state_wr
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
        s.rst = st.ld<VlcuMem8>(OFF_rst, i);
        s.x = st.ld<IData>(OFF_x, i);
        s.a0 = st.ld<VlcuMem8>(OFF_a0, i);
        s.a1 = st.ld<VlcuMem16>(OFF_a1, i);
        s.a2 = st.ld<QData>(OFF_a2, i);
        s.msbs = st.ld<VlcuMem8>(OFF_msbs, i);
        s.votes = st.ld<VlcuMem8>(OFF_votes, i);
        s.t14_hier__DOT__a3 = st.ld<VlcuMem8>(OFF_t14_hier__DOT__a3, i);
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
        st.st<VlcuMem8>(OFF_rst, i, s.rst);
        st.st<IData>(OFF_x, i, s.x);
        st.st<VlcuMem8>(OFF_a0, i, s.a0);
        st.st<VlcuMem16>(OFF_a1, i, s.a1);
        st.st<QData>(OFF_a2, i, s.a2);
        st.st<VlcuMem8>(OFF_msbs, i, s.msbs);
        st.st<VlcuMem8>(OFF_votes, i, s.votes);
        st.st<VlcuMem8>(OFF_t14_hier__DOT__a3, i, s.t14_hier__DOT__a3);
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
        s.rst = static_cast<VlcuMem8>(rowp[1 * stride + i] & 0x1U);
        s.x = static_cast<IData>(rowp[2 * stride + i] & 0xffffffffU);
    }
    VLCU_D static void store_outputs_row(const Root& s, uint32_t* rowp, uint64_t stride, uint64_t i) {
        rowp[0 * stride + i] = s.a0;
        rowp[1 * stride + i] = s.a1;
        rowp[2 * stride + i] = static_cast<uint32_t>(s.a2);
        rowp[3 * stride + i] = static_cast<uint32_t>(s.a2 >> 32);
        rowp[4 * stride + i] = s.msbs;
        rowp[5 * stride + i] = s.votes;
    }
    VLCU_D static uint64_t outputs_sig(const Root& s, uint64_t sig) {
        uint32_t acc = VLCU_SIG_ACC_INIT;
        acc = vlcu_sig_word(acc, s.a0);
        acc = vlcu_sig_word(acc, s.a1);
        acc = vlcu_sig_word(acc, static_cast<uint32_t>(s.a2));
        acc = vlcu_sig_word(acc, static_cast<uint32_t>(s.a2 >> 32));
        acc = vlcu_sig_word(acc, s.msbs);
        acc = vlcu_sig_word(acc, s.votes);
        return vlcu_sig_mix(sig, acc);
    }
    enum : uint32_t { kInWords = 3, kOutWords = 6, kInBytes = 6, kOutBytes = 13, kTraceWords = 0 };

    VLCU_D static void Vt14_hier___024root___eval_static(Vt14_hier___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
    }

    VLCU_D static void Vt14_hier___024root___eval_initial(Vt14_hier___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt14_hier___024root___eval_final(Vt14_hier___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt14_hier___024root___eval_settle(Vt14_hier___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt14_hier___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t14_hier.v", 15, "", "DIDNOTCONVERGE: Settle region did not converge after '--converge-limit' of 10000 tries");
            }
            __VstlIterCount = ((IData)(1U) + __VstlIterCount);
            vlSelfRef.__VstlPhaseResult = Vt14_hier___024root___eval_phase__stl(vlSelf, vlCtx);
            vlSelfRef.__VstlFirstIteration = 0U;
        } while (vlSelfRef.__VstlPhaseResult);
    }

    VLCU_D static bool Vt14_hier___024root___trigger_anySet__stl(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt14_hier___024root___eval_phase__stl(Vt14_hier___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt14_hier___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
        }
#endif
        __VstlExecute = Vt14_hier___024root___trigger_anySet__stl(vlSelfRef.__VstlTriggered);
        if (__VstlExecute) {
            {
                // Inlined CFunc: _eval_stl
                if ((1ULL & vlSelfRef.__VstlTriggered[0U])) {
                    {
                        // Inlined CFunc: _stl_sequent__TOP__0
                        vlSelfRef.msbs = ((((2U & ((IData)(vlSelfRef.t14_hier__DOT__a3) 
                                                   >> 6U)) 
                                            | (1U & (IData)(
                                                            (vlSelfRef.a2 
                                                             >> 0x00000027U)))) 
                                           << 2U) | 
                                          ((2U & ((IData)(vlSelfRef.a1) 
                                                  >> 0x0000000eU)) 
                                           | (1U & 
                                              ((IData)(vlSelfRef.a0) 
                                               >> 7U))));
                    }
                }
            }
        }
        return (__VstlExecute);
    }

    VLCU_D static bool Vt14_hier___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt14_hier___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt14_hier___024root___eval_phase__act(Vt14_hier___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt14_hier___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vt14_hier___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vt14_hier___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vt14_hier___024root___eval_phase__nba(Vt14_hier___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vt14_hier___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    {
                        // Inlined CFunc: _nba_sequent__TOP__0
                        CData/*7:0*/ __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___Vdly__a0;
                        __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___Vdly__a0 = 0;
                        __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___Vdly__a0 
                            = vlSelfRef.a0;
                        vlSelfRef.votes = (0x000000ffU 
                                           & (((IData)(vlSelfRef.votes) 
                                               + (IData)(vlSelfRef.msbs)) 
                                              + (IData)(vlSelfRef.t14_hier__DOT__a3)));
                        if (vlSelfRef.rst) {
                            vlSelfRef.a2 = 0x000000ff00ff00ffULL;
                            vlSelfRef.a1 = 0U;
                            __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___Vdly__a0 = 0x0000005aU;
                            vlSelfRef.t14_hier__DOT__a3 = 0U;
                        } else {
                            vlSelfRef.a2 = (((0x000000fffffffffeULL 
                                              & (vlSelfRef.a2 
                                                 << 1U)) 
                                             | (QData)((IData)(
                                                               (1U 
                                                                & (IData)(
                                                                          (vlSelfRef.a2 
                                                                           >> 0x27U)))))) 
                                            ^ (((QData)((IData)(vlSelfRef.x)) 
                                                << 8U) 
                                               | (QData)((IData)(
                                                                 (0x000000ffU 
                                                                  & vlSelfRef.x)))));
                            vlSelfRef.a1 = (0x0000ffffU 
                                            & (((0x0000fffeU 
                                                 & ((IData)(vlSelfRef.a1) 
                                                    << 1U)) 
                                                | (1U 
                                                   & ((IData)(vlSelfRef.a1) 
                                                      >> 0x0fU))) 
                                               ^ (vlSelfRef.x 
                                                  >> 8U)));
                            __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___Vdly__a0 
                                = (0x000000ffU & ((
                                                   (0x000000feU 
                                                    & ((IData)(vlSelfRef.a0) 
                                                       << 1U)) 
                                                   | (1U 
                                                      & ((IData)(vlSelfRef.a0) 
                                                         >> 7U))) 
                                                  ^ vlSelfRef.x));
                            vlSelfRef.t14_hier__DOT__a3 
                                = (0x000000ffU & ((
                                                   (0x000000feU 
                                                    & ((IData)(vlSelfRef.t14_hier__DOT__a3) 
                                                       << 1U)) 
                                                   | (1U 
                                                      & ((IData)(vlSelfRef.t14_hier__DOT__a3) 
                                                         >> 7U))) 
                                                  ^ 
                                                  ((IData)(vlSelfRef.a0) 
                                                   ^ 
                                                   (vlSelfRef.x 
                                                    >> 0x00000018U))));
                        }
                        vlSelfRef.a0 = __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___Vdly__a0;
                        vlSelfRef.msbs = ((((2U & ((IData)(vlSelfRef.t14_hier__DOT__a3) 
                                                   >> 6U)) 
                                            | (1U & (IData)(
                                                            (vlSelfRef.a2 
                                                             >> 0x00000027U)))) 
                                           << 2U) | 
                                          ((2U & ((IData)(vlSelfRef.a1) 
                                                  >> 0x0000000eU)) 
                                           | (1U & 
                                              ((IData)(vlSelfRef.a0) 
                                               >> 7U))));
                    }
                }
            }
            Vt14_hier___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vt14_hier___024root___eval(Vt14_hier___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt14_hier___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t14_hier.v", 15, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vt14_hier___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t14_hier.v", 15, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vt14_hier___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vt14_hier___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vt14_hier___024root___ctor_var_reset(Vt14_hier___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        vlSelf->rst = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 18209466448985614591ull);
        vlSelf->x = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 9409450202036847209ull);
        vlSelf->a0 = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 17342812819118991936ull);
        vlSelf->a1 = VL_SCOPED_RAND_RESET_I(16, __VscopeHash, 17885462169741112028ull);
        vlSelf->a2 = VL_SCOPED_RAND_RESET_Q(40, __VscopeHash, 2614162857933103900ull);
        vlSelf->msbs = VL_SCOPED_RAND_RESET_I(4, __VscopeHash, 17511002752471744084ull);
        vlSelf->votes = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 12316281325160085181ull);
        vlSelf->t14_hier__DOT__a3 = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 14485463327109504829ull);
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
        Vt14_hier___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt14_hier___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt14_hier___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt14_hier___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vt14_hier___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt14_hier___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt14_hier__cuda::info() {
    typedef Vt14_hier__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"rst", "rst", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_rst},
        {"x", "x", VLCU_MEMBER_IN, 4, 0, 1, 32, 31, 0, T::OFF_x},
        {"a0", "a0", VLCU_MEMBER_OUT, 1, 0, 1, 8, 7, 0, T::OFF_a0},
        {"a1", "a1", VLCU_MEMBER_OUT, 2, 0, 1, 16, 15, 0, T::OFF_a1},
        {"a2", "a2", VLCU_MEMBER_OUT, 8, 0, 1, 40, 39, 0, T::OFF_a2},
        {"msbs", "msbs", VLCU_MEMBER_OUT, 1, 0, 1, 4, 3, 0, T::OFF_msbs},
        {"votes", "votes", VLCU_MEMBER_OUT, 1, 0, 1, 8, 7, 0, T::OFF_votes},
        {"t14_hier__DOT__a3", "t14_hier.a3", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF_t14_hier__DOT__a3},
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
    static const uint32_t saveOrder[] = {0, 1, 3, 6, 7, 8, 9, 10, 11, 12, 13, 4, 2, 14, 5, 15, 16, 17, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "t14_hier", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 0, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 18, 0, 0xbd05e079dc4e69b9ULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt14_hier__cuda VLCU_TESTBENCHES);
