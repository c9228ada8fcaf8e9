// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt07_dyn"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt07_dyn___024root final {
    CData/*0:0*/ clk;
    IData/*31:0*/ a;
    IData/*31:0*/ b;
    CData/*5:0*/ n;
    IData/*31:0*/ gcd;
    CData/*6:0*/ steps;
    CData/*5:0*/ pop;
    CData/*2:0*/ flags;
    QData/*63:0*/ acc;
    CData/*7:0*/ slice;
    CData/*0:0*/ bitsel;
    IData/*31:0*/ pw;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    IData/*31:0*/ t07_dyn__DOT__t;
    IData/*31:0*/ __VactIterCount;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vt07_dyn__cuda final {
    typedef Vt07_dyn___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 48,
        OFF_a = 24,
        OFF_b = 28,
        OFF_n = 49,
        OFF_gcd = 32,
        OFF_steps = 50,
        OFF_pop = 51,
        OFF_flags = 52,
        OFF_acc = 0,
        OFF_slice = 53,
        OFF_bitsel = 54,
        OFF_pw = 36,
        OFF___Vtrigprevexpr___TOP__clk__0 = 55,
        OFF___VactPhaseResult = 56,
        OFF___VnbaPhaseResult = 57,
        OFF_t07_dyn__DOT__t = 40,
        OFF___VactIterCount = 44,
        OFF___VactTriggered = 8,
        OFF___VnbaTriggered = 16,
        kStateBytes = 64,  // S_persist
        kTouchedReadBytes = 45,  // S_state_rd: persistent bytes one eval can read
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
        s.a = st.ld<IData>(OFF_a, i);
        s.b = st.ld<IData>(OFF_b, i);
        s.n = st.ld<VlcuMem8>(OFF_n, i);
        s.gcd = st.ld<IData>(OFF_gcd, i);
        s.steps = st.ld<VlcuMem8>(OFF_steps, i);
        s.pop = st.ld<VlcuMem8>(OFF_pop, i);
        s.flags = st.ld<VlcuMem8>(OFF_flags, i);
        s.acc = st.ld<QData>(OFF_acc, i);
        s.slice = st.ld<VlcuMem8>(OFF_slice, i);
        s.bitsel = st.ld<VlcuMem8>(OFF_bitsel, i);
        s.pw = st.ld<IData>(OFF_pw, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        s.t07_dyn__DOT__t = st.ld<IData>(OFF_t07_dyn__DOT__t, i);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.st<IData>(OFF_a, i, s.a);
        st.st<IData>(OFF_b, i, s.b);
        st.st<VlcuMem8>(OFF_n, i, s.n);
        st.st<IData>(OFF_gcd, i, s.gcd);
        st.st<VlcuMem8>(OFF_steps, i, s.steps);
        st.st<VlcuMem8>(OFF_pop, i, s.pop);
        st.st<VlcuMem8>(OFF_flags, i, s.flags);
        st.st<QData>(OFF_acc, i, s.acc);
        st.st<VlcuMem8>(OFF_slice, i, s.slice);
        st.st<VlcuMem8>(OFF_bitsel, i, s.bitsel);
        st.st<IData>(OFF_pw, i, s.pw);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.st<IData>(OFF_t07_dyn__DOT__t, i, s.t07_dyn__DOT__t);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VactTriggered, e, i, s.__VactTriggered[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VnbaTriggered, e, i, s.__VnbaTriggered[e]);
    }
    VLCU_D static void load_inputs_row(Root& s, const uint32_t* rowp, uint64_t stride, uint64_t i) {
        s.clk = static_cast<VlcuMem8>(rowp[0 * stride + i] & 0x1U);
        s.a = static_cast<IData>(rowp[1 * stride + i] & 0xffffffffU);
        s.b = static_cast<IData>(rowp[2 * stride + i] & 0xffffffffU);
        s.n = static_cast<VlcuMem8>(rowp[3 * stride + i] & 0x3fU);
    }
    VLCU_D static void store_outputs_row(const Root& s, uint32_t* rowp, uint64_t stride, uint64_t i) {
        rowp[0 * stride + i] = s.gcd;
        rowp[1 * stride + i] = s.steps;
        rowp[2 * stride + i] = s.pop;
        rowp[3 * stride + i] = s.flags;
        rowp[4 * stride + i] = static_cast<uint32_t>(s.acc);
        rowp[5 * stride + i] = static_cast<uint32_t>(s.acc >> 32);
        rowp[6 * stride + i] = s.slice;
        rowp[7 * stride + i] = s.bitsel;
        rowp[8 * stride + i] = s.pw;
    }
    VLCU_D static uint64_t outputs_sig(const Root& s, uint64_t sig) {
        uint32_t acc = VLCU_SIG_ACC_INIT;
        acc = vlcu_sig_word(acc, s.gcd);
        acc = vlcu_sig_word(acc, s.steps);
        acc = vlcu_sig_word(acc, s.pop);
        acc = vlcu_sig_word(acc, s.flags);
        acc = vlcu_sig_word(acc, static_cast<uint32_t>(s.acc));
        acc = vlcu_sig_word(acc, static_cast<uint32_t>(s.acc >> 32));
        acc = vlcu_sig_word(acc, s.slice);
        acc = vlcu_sig_word(acc, s.bitsel);
        acc = vlcu_sig_word(acc, s.pw);
        return vlcu_sig_mix(sig, acc);
    }
    enum : uint32_t { kInWords = 4, kOutWords = 9, kInBytes = 10, kOutBytes = 21, kTraceWords = 0 };

    VLCU_D static void Vt07_dyn___024root___eval_static(Vt07_dyn___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
    }

    VLCU_D static void Vt07_dyn___024root___eval_initial(Vt07_dyn___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt07_dyn___024root___eval_final(Vt07_dyn___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt07_dyn___024root___eval_settle(Vt07_dyn___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static bool Vt07_dyn___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt07_dyn___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt07_dyn___024root___eval_phase__act(Vt07_dyn___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt07_dyn___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vt07_dyn___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vt07_dyn___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vt07_dyn___024root___eval_phase__nba(Vt07_dyn___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vt07_dyn___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    {
                        // Inlined CFunc: _nba_sequent__TOP__0
                        IData/*31:0*/ __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0_t07_dyn__DOT__x;
                        __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0_t07_dyn__DOT__x = 0;
                        IData/*31:0*/ __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0_t07_dyn__DOT__y;
                        __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0_t07_dyn__DOT__y = 0;
                        CData/*6:0*/ __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0_t07_dyn__DOT__k;
                        __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0_t07_dyn__DOT__k = 0;
                        IData/*31:0*/ __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0_t07_dyn__DOT__i;
                        __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0_t07_dyn__DOT__i = 0;
                        vlSelfRef.pop = (0x0000003fU 
                                         & VL_COUNTONES_I(
                                                          (vlSelfRef.a 
                                                           ^ vlSelfRef.b)));
                        vlSelfRef.flags = ((VL_ONEHOT_I(
                                                        (0x000000ffU 
                                                         & vlSelfRef.a)) 
                                            << 2U) 
                                           | ((VL_ONEHOT0_I(
                                                            (0x000000ffU 
                                                             & vlSelfRef.b)) 
                                               << 1U) 
                                              | (1U 
                                                 & VL_REDXOR_32(vlSelfRef.a))));
                        vlSelfRef.slice = (0x000000ffU 
                                           & ((vlSelfRef.a 
                                               >> (0x0000001fU 
                                                   & (IData)(vlSelfRef.n))) 
                                              ^ (vlSelfRef.b 
                                                 >> 0x18U)));
                        vlSelfRef.bitsel = (1U & (~ 
                                                  (VL_REDXOR_16(
                                                                (0x0000ff00U 
                                                                 & vlSelfRef.b)) 
                                                   ^ 
                                                   (((vlSelfRef.a 
                                                      >> 
                                                      (0x0000001fU 
                                                       & (IData)(vlSelfRef.n))) 
                                                     ^ 
                                                     (0x0000000fU 
                                                      == 
                                                      (0x0000000fU 
                                                       & vlSelfRef.b))) 
                                                    ^ 
                                                    (0U 
                                                     != 
                                                     (vlSelfRef.a 
                                                      >> 0x1cU))))));
                        vlSelfRef.pw = (VL_POW_III(32,32,3, 
                                                   (0x0000000fU 
                                                    & vlSelfRef.a), 
                                                   (7U 
                                                    & vlSelfRef.b)) 
                                        + VL_POW_III(32,32,4, (IData)(3U), 
                                                     (0x0000000fU 
                                                      & (IData)(vlSelfRef.n))));
                        __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0_t07_dyn__DOT__x 
                            = (1U | vlSelfRef.a);
                        __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0_t07_dyn__DOT__y 
                            = vlSelfRef.b;
                        __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0_t07_dyn__DOT__k = 0U;
                        while (((0U != __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0_t07_dyn__DOT__y) 
                                & (0x40U > __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0_t07_dyn__DOT__k))) {
                            vlSelfRef.t07_dyn__DOT__t 
                                = VL_MODDIV_III(32, __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0_t07_dyn__DOT__x, __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0_t07_dyn__DOT__y);
                            __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0_t07_dyn__DOT__x 
                                = __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0_t07_dyn__DOT__y;
                            __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0_t07_dyn__DOT__y 
                                = vlSelfRef.t07_dyn__DOT__t;
                            __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0_t07_dyn__DOT__k 
                                = (0x0000007fU & ((IData)(1U) 
                                                  + (IData)(__Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0_t07_dyn__DOT__k)));
                        }
                        vlSelfRef.gcd = __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0_t07_dyn__DOT__x;
                        vlSelfRef.steps = __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0_t07_dyn__DOT__k;
                        vlSelfRef.t07_dyn__DOT__t = 0U;
                        __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0_t07_dyn__DOT__i = 0U;
                        while ((__Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0_t07_dyn__DOT__i 
                                < (IData)(vlSelfRef.n))) {
                            vlSelfRef.t07_dyn__DOT__t 
                                = (vlSelfRef.t07_dyn__DOT__t 
                                   + (vlSelfRef.a >> 
                                      (0x0000001fU 
                                       & __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0_t07_dyn__DOT__i)));
                            __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0_t07_dyn__DOT__i 
                                = ((IData)(1U) + __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0_t07_dyn__DOT__i);
                        }
                        vlSelfRef.acc = (6ULL + (((QData)((IData)(
                                                                  ((IData)(vlSelfRef.acc) 
                                                                   ^ vlSelfRef.t07_dyn__DOT__t))) 
                                                  << 0x00000020U) 
                                                 | (QData)((IData)(
                                                                   ((IData)(
                                                                            (vlSelfRef.acc 
                                                                             >> 0x20U)) 
                                                                    + vlSelfRef.b)))));
                    }
                }
            }
            Vt07_dyn___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vt07_dyn___024root___eval(Vt07_dyn___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt07_dyn___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t07_dyn.v", 4, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vt07_dyn___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t07_dyn.v", 4, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vt07_dyn___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vt07_dyn___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vt07_dyn___024root___ctor_var_reset(Vt07_dyn___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        vlSelf->a = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 510903276987443985ull);
        vlSelf->b = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 16900879642891266615ull);
        vlSelf->n = VL_SCOPED_RAND_RESET_I(6, __VscopeHash, 1489474827855109852ull);
        vlSelf->gcd = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 10029818939142169790ull);
        vlSelf->steps = VL_SCOPED_RAND_RESET_I(7, __VscopeHash, 1359535656114925994ull);
        vlSelf->pop = VL_SCOPED_RAND_RESET_I(6, __VscopeHash, 10750680354736310321ull);
        vlSelf->flags = VL_SCOPED_RAND_RESET_I(3, __VscopeHash, 7921719909686688401ull);
        vlSelf->acc = VL_SCOPED_RAND_RESET_Q(64, __VscopeHash, 14651699267163694450ull);
        vlSelf->slice = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 5452921321901628446ull);
        vlSelf->bitsel = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 8396549258805030630ull);
        vlSelf->pw = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 5549622911285396126ull);
        vlSelf->t07_dyn__DOT__t = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 850418292027566636ull);
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
        Vt07_dyn___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt07_dyn___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt07_dyn___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt07_dyn___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vt07_dyn___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt07_dyn___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt07_dyn__cuda::info() {
    typedef Vt07_dyn__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"a", "a", VLCU_MEMBER_IN, 4, 0, 1, 32, 31, 0, T::OFF_a},
        {"b", "b", VLCU_MEMBER_IN, 4, 0, 1, 32, 31, 0, T::OFF_b},
        {"n", "n", VLCU_MEMBER_IN, 1, 0, 1, 6, 5, 0, T::OFF_n},
        {"gcd", "gcd", VLCU_MEMBER_OUT, 4, 0, 1, 32, 31, 0, T::OFF_gcd},
        {"steps", "steps", VLCU_MEMBER_OUT, 1, 0, 1, 7, 6, 0, T::OFF_steps},
        {"pop", "pop", VLCU_MEMBER_OUT, 1, 0, 1, 6, 5, 0, T::OFF_pop},
        {"flags", "flags", VLCU_MEMBER_OUT, 1, 0, 1, 3, 2, 0, T::OFF_flags},
        {"acc", "acc", VLCU_MEMBER_OUT, 8, 0, 1, 64, 63, 0, T::OFF_acc},
        {"slice", "slice", VLCU_MEMBER_OUT, 1, 0, 1, 8, 7, 0, T::OFF_slice},
        {"bitsel", "bitsel", VLCU_MEMBER_OUT, 1, 0, 1, 1, 0, 0, T::OFF_bitsel},
        {"pw", "pw", VLCU_MEMBER_OUT, 4, 0, 1, 32, 31, 0, T::OFF_pw},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"t07_dyn__DOT__t", "t07_dyn.t", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t07_dyn__DOT__t},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"__VactTriggered", "__VactTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VactTriggered},
        {"__VnbaTriggered", "__VnbaTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VnbaTriggered},
    };
    static const char* const messages[] = {nullptr};
    static const char* const plusargs[] = {nullptr};
    static const VlcuTraceDecl traceDecls[] = {
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0}};  // (terminator, not counted)
    static const uint32_t traceOrder[] = {0};  // codes in dump order (+ terminator)
    static const uint32_t saveOrder[] = {0, 3, 5, 6, 7, 9, 10, 12, 13, 14, 1, 2, 4, 11, 15, 16, 8, 17, 18, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "t07_dyn", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 0, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 19, 0, 0xd80e36e69bbb5d54ULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt07_dyn__cuda VLCU_TESTBENCHES);
