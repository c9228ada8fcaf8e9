// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt_func_sel"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt_func_sel___024root final {
    CData/*0:0*/ clk;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    IData/*31:0*/ t__DOT__cyc;
    VlWide<4>/*127:0*/ t__DOT__sum;
    VlWide<4>/*103:0*/ t__DOT__out;
    IData/*31:0*/ __VactIterCount;
    QData/*63:0*/ t__DOT__crc;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vt_func_sel__cuda final {
    typedef Vt_func_sel___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 64,
        OFF___Vtrigprevexpr___TOP__clk__0 = 65,
        OFF___VactPhaseResult = 66,
        OFF___VnbaPhaseResult = 67,
        OFF_t__DOT__cyc = 24,
        OFF_t__DOT__sum = 28,
        OFF_t__DOT__out = 44,
        OFF___VactIterCount = 60,
        OFF_t__DOT__crc = 0,
        OFF___VactTriggered = 8,
        OFF___VnbaTriggered = 16,
        kStateBytes = 72,  // S_persist
        kTouchedReadBytes = 68,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 67,  // S_state_wr
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
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        s.t__DOT__cyc = st.ld<IData>(OFF_t__DOT__cyc, i);
        st.ldW(OFF_t__DOT__sum, i, s.t__DOT__sum);
        st.ldW(OFF_t__DOT__out, i, s.t__DOT__out);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        s.t__DOT__crc = st.ld<QData>(OFF_t__DOT__crc, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.st<IData>(OFF_t__DOT__cyc, i, s.t__DOT__cyc);
        st.stW(OFF_t__DOT__sum, i, s.t__DOT__sum);
        st.stW(OFF_t__DOT__out, i, s.t__DOT__out);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
        st.st<QData>(OFF_t__DOT__crc, i, s.t__DOT__crc);
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

    VLCU_D static void Vt_func_sel___024root___eval_static(Vt_func_sel___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        {
            // Inlined CFunc: _eval_static__TOP
            IData/*31:0*/ __Vinline_0__eval_static__TOP_t__DOT__reverse__Vstatic__i;
            __Vinline_0__eval_static__TOP_t__DOT__reverse__Vstatic__i = 0;
            vlSelfRef.t__DOT__cyc = 0U;
            __Vinline_0__eval_static__TOP_t__DOT__reverse__Vstatic__i = 0;
        }
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
    }

    VLCU_D static void Vt_func_sel___024root___eval_initial(Vt_func_sel___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_func_sel___024root___eval_final(Vt_func_sel___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_func_sel___024root___eval_settle(Vt_func_sel___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static bool Vt_func_sel___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt_func_sel___024root___nba_sequent__TOP__0(Vt_func_sel___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        IData/*31:0*/ t__DOT__reverse__Vstatic__i;
        t__DOT__reverse__Vstatic__i = 0;
        VlWide<4>/*103:0*/ __Vfunc_t__DOT__reverse__0__Vfuncout;
        VL_ZERO_W(104, __Vfunc_t__DOT__reverse__0__Vfuncout);
        VlWide<4>/*103:0*/ __Vfunc_t__DOT__reverse__0__val;
        VL_ZERO_W(104, __Vfunc_t__DOT__reverse__0__val);
        IData/*31:0*/ __Vdly__t__DOT__cyc;
        __Vdly__t__DOT__cyc = 0;
        QData/*63:0*/ __Vdly__t__DOT__crc;
        __Vdly__t__DOT__crc = 0;
        VlWide<4>/*127:0*/ __Vdly__t__DOT__sum;
        VL_ZERO_W(128, __Vdly__t__DOT__sum);
        VlWide<4>/*127:0*/ __Vtemp_4;
        VlWide<4>/*127:0*/ __Vtemp_5;
        // Body
        __Vdly__t__DOT__cyc = vlSelfRef.t__DOT__cyc;
        __Vdly__t__DOT__crc = vlSelfRef.t__DOT__crc;
        __Vdly__t__DOT__sum[0U] = vlSelfRef.t__DOT__sum[0U];
        __Vdly__t__DOT__sum[1U] = vlSelfRef.t__DOT__sum[1U];
        __Vdly__t__DOT__sum[2U] = vlSelfRef.t__DOT__sum[2U];
        __Vdly__t__DOT__sum[3U] = vlSelfRef.t__DOT__sum[3U];
        __Vdly__t__DOT__cyc = ((IData)(1U) + vlSelfRef.t__DOT__cyc);
        __Vdly__t__DOT__crc = ((vlSelfRef.t__DOT__crc 
                                << 1U) | (QData)((IData)(
                                                         (1U 
                                                          & VL_REDXOR_64(
                                                                         (0x8000000000000005ULL 
                                                                          & vlSelfRef.t__DOT__crc))))));
        __Vtemp_4[0U] = (0xfffffffeU & vlSelfRef.t__DOT__sum[0U]);
        __Vtemp_4[1U] = ((1U & vlSelfRef.t__DOT__sum[1U]) 
                         | (0xfffffffeU & vlSelfRef.t__DOT__sum[1U]));
        __Vtemp_4[2U] = ((1U & vlSelfRef.t__DOT__sum[2U]) 
                         | (0xfffffffeU & vlSelfRef.t__DOT__sum[2U]));
        __Vtemp_4[3U] = ((1U & vlSelfRef.t__DOT__sum[3U]) 
                         | (0xfffffffeU & vlSelfRef.t__DOT__sum[3U]));
        __Vtemp_5[0U] = vlSelfRef.t__DOT__out[0U];
        __Vtemp_5[1U] = vlSelfRef.t__DOT__out[1U];
        __Vtemp_5[2U] = vlSelfRef.t__DOT__out[2U];
        __Vtemp_5[3U] = vlSelfRef.t__DOT__out[3U];
        VL_ADD_W(4, __Vdly__t__DOT__sum, __Vtemp_4, __Vtemp_5);
        if (VL_GTS_III(32, 0x0000000aU, vlSelfRef.t__DOT__cyc)) {
            __Vdly__t__DOT__crc = 1ULL;
            __Vdly__t__DOT__sum[0U] = 0U;
            __Vdly__t__DOT__sum[1U] = 0U;
            __Vdly__t__DOT__sum[2U] = 0U;
            __Vdly__t__DOT__sum[3U] = 0U;
        } else if (VL_UNLIKELY((VL_LTES_III(32, 0x0000005aU, vlSelfRef.t__DOT__cyc)))) {
            vlCtx.writef(0U, 4U, vlcu_msg_arg(vlSelfRef.t__DOT__sum[3U]), vlcu_msg_arg(vlSelfRef.t__DOT__sum[2U]), vlcu_msg_arg(vlSelfRef.t__DOT__sum[1U]), vlcu_msg_arg(vlSelfRef.t__DOT__sum[0U]));
            if (VL_UNLIKELY(((0U != ((((0x42d1a319U 
                                        ^ vlSelfRef.t__DOT__sum[0U]) 
                                       | (0x8d1a5936U 
                                          ^ vlSelfRef.t__DOT__sum[1U])) 
                                      | (0x42d1a346U 
                                         ^ vlSelfRef.t__DOT__sum[2U])) 
                                     | (0x00002d36U 
                                        ^ vlSelfRef.t__DOT__sum[3U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_func_sel.v", 47, "");
            }
            vlCtx.writef(1U, 0U);
            VL_FINISH_MT("/root/reference/test_regress/t/t_func_sel.v", 49, "");
        }
        vlSelfRef.t__DOT__cyc = __Vdly__t__DOT__cyc;
        vlSelfRef.t__DOT__sum[0U] = __Vdly__t__DOT__sum[0U];
        vlSelfRef.t__DOT__sum[1U] = __Vdly__t__DOT__sum[1U];
        vlSelfRef.t__DOT__sum[2U] = __Vdly__t__DOT__sum[2U];
        vlSelfRef.t__DOT__sum[3U] = __Vdly__t__DOT__sum[3U];
        __Vfunc_t__DOT__reverse__0__val[0U] = (IData)(vlSelfRef.t__DOT__crc);
        __Vfunc_t__DOT__reverse__0__val[1U] = (IData)(
                                                      (vlSelfRef.t__DOT__crc 
                                                       >> 0x00000020U));
        __Vfunc_t__DOT__reverse__0__val[2U] = (IData)(
                                                      (0x000000ffffffffffULL 
                                                       & vlSelfRef.t__DOT__crc));
        __Vfunc_t__DOT__reverse__0__val[3U] = (IData)(
                                                      ((0x000000ffffffffffULL 
                                                        & vlSelfRef.t__DOT__crc) 
                                                       >> 0x00000020U));
        t__DOT__reverse__Vstatic__i = 0U;
        while (VL_GTS_III(32, 0x00000068U, t__DOT__reverse__Vstatic__i)) {
            if (VL_LIKELY(((0x67U >= (0x0000007fU & 
                                      ((IData)(0x67U) 
                                       - t__DOT__reverse__Vstatic__i)))))) {
                __Vfunc_t__DOT__reverse__0__Vfuncout[(3U 
                                                      & (((IData)(0x67U) 
                                                          - t__DOT__reverse__Vstatic__i) 
                                                         >> 5U))] 
                    = (((~ ((IData)(1U) << (0x0000001fU 
                                            & ((IData)(0x67U) 
                                               - t__DOT__reverse__Vstatic__i)))) 
                        & __Vfunc_t__DOT__reverse__0__Vfuncout
                        [(3U & (((IData)(0x67U) - t__DOT__reverse__Vstatic__i) 
                                >> 5U))]) | (((0x67U 
                                               >= (0x0000007fU 
                                                   & t__DOT__reverse__Vstatic__i)) 
                                              && (1U 
                                                  & (__Vfunc_t__DOT__reverse__0__val
                                                     [
                                                     (3U 
                                                      & (t__DOT__reverse__Vstatic__i 
                                                         >> 5U))] 
                                                     >> 
                                                     (0x0000001fU 
                                                      & t__DOT__reverse__Vstatic__i)))) 
                                             << (0x0000001fU 
                                                 & ((IData)(0x67U) 
                                                    - t__DOT__reverse__Vstatic__i))));
            }
            t__DOT__reverse__Vstatic__i = ((IData)(1U) 
                                           + t__DOT__reverse__Vstatic__i);
        }
        vlSelfRef.t__DOT__out[0U] = __Vfunc_t__DOT__reverse__0__Vfuncout[0U];
        vlSelfRef.t__DOT__out[1U] = __Vfunc_t__DOT__reverse__0__Vfuncout[1U];
        vlSelfRef.t__DOT__out[2U] = __Vfunc_t__DOT__reverse__0__Vfuncout[2U];
        vlSelfRef.t__DOT__out[3U] = __Vfunc_t__DOT__reverse__0__Vfuncout[3U];
        vlSelfRef.t__DOT__crc = __Vdly__t__DOT__crc;
    }

    VLCU_D static void Vt_func_sel___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt_func_sel___024root___eval_phase__act(Vt_func_sel___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt_func_sel___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vt_func_sel___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vt_func_sel___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vt_func_sel___024root___eval_phase__nba(Vt_func_sel___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vt_func_sel___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    Vt_func_sel___024root___nba_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
            Vt_func_sel___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vt_func_sel___024root___eval(Vt_func_sel___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt_func_sel___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("/root/reference/test_regress/t/t_func_sel.v", 10, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vt_func_sel___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("/root/reference/test_regress/t/t_func_sel.v", 10, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vt_func_sel___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vt_func_sel___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vt_func_sel___024root___ctor_var_reset(Vt_func_sel___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        vlSelf->t__DOT__crc = VL_SCOPED_RAND_RESET_Q(64, __VscopeHash, 17680824324711171631ull);
        VL_SCOPED_RAND_RESET_W(128, vlSelf->t__DOT__sum, __VscopeHash, 1514072735762163239ull);
        VL_SCOPED_RAND_RESET_W(104, vlSelf->t__DOT__out, __VscopeHash, 9695058467195139273ull);
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
        Vt_func_sel___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt_func_sel___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt_func_sel___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt_func_sel___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vt_func_sel___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt_func_sel___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt_func_sel__cuda::info() {
    typedef Vt_func_sel__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"t__DOT__cyc", "t.cyc", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__cyc},
        {"t__DOT__sum", "t.sum", VLCU_MEMBER_STATE, 4, 0, 4, 128, 127, 0, T::OFF_t__DOT__sum},
        {"t__DOT__out", "t.out", VLCU_MEMBER_STATE, 4, 0, 4, 104, 103, 0, T::OFF_t__DOT__out},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"t__DOT__crc", "t.crc", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF_t__DOT__crc},
        {"__VactTriggered", "__VactTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VactTriggered},
        {"__VnbaTriggered", "__VnbaTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VnbaTriggered},
    };
    static const char* const messages[] = {"SUM = %x_%x_%x_%x\n", "*-* All Finished *-*\n", nullptr};
    static const char* const plusargs[] = {nullptr};
    static const VlcuTraceDecl traceDecls[] = {
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0}};  // (terminator, not counted)
    static const uint32_t traceOrder[] = {0};  // codes in dump order (+ terminator)
    static const uint32_t saveOrder[] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "t", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 2, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 11, 0, 0xd9788a6d28e3530bULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt_func_sel__cuda VLCU_TESTBENCHES);
