// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

// 64 bytes of register-resident state: latency-bound, compile for full occupancy
#ifndef VLCU_MIN_BLOCKS
#define VLCU_MIN_BLOCKS 32
#endif

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt_array_mda"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt_array_mda___024root final {
    CData/*0:0*/ clk;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    IData/*31:0*/ t__DOT__cyc;
    IData/*31:0*/ __VactIterCount;
    QData/*63:0*/ t__DOT__crc;
    QData/*63:0*/ t__DOT__sum;
    VlcuMdArray<VlcuSharedArray<IData/*31:0*/, 7>, 7, 1> t__DOT__A;  // run-time indexed: shared memory [index][lane]
    VlcuMdArray<VlcuSharedArray<IData/*31:0*/, 8>, 8, 1> t__DOT__B;  // run-time indexed: shared memory [index][lane]
    VlUnpacked<IData/*31:0*/, 1> t__DOT__a;
    VlUnpacked<IData/*31:0*/, 1> t__DOT__b;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vt_array_mda__cuda final {
    typedef Vt_array_mda___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 108,
        OFF___Vtrigprevexpr___TOP__clk__0 = 109,
        OFF___VactPhaseResult = 110,
        OFF___VnbaPhaseResult = 111,
        OFF_t__DOT__cyc = 32,
        OFF___VactIterCount = 36,
        OFF_t__DOT__crc = 0,
        OFF_t__DOT__sum = 8,
        OFF_t__DOT__A = 40,
        OFF_t__DOT__B = 68,
        OFF_t__DOT__a = 100,
        OFF_t__DOT__b = 104,
        OFF___VactTriggered = 16,
        OFF___VnbaTriggered = 24,
        kStateBytes = 112,  // S_persist
        kTouchedReadBytes = 60,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 111,  // S_state_wr
    };
    struct Locals final {
    };
    struct Shared final {
        IData/*31:0*/ t__DOT__A[7][VLCU_BLOCK];
        IData/*31:0*/ t__DOT__B[8][VLCU_BLOCK];
    };
    VLCU_D static void bind(Root& s, Locals& l, Shared& sh, const VlcuState& st, uint64_t i) {
        (void)l; (void)sh; (void)st; (void)i;
        s.vlNamep = "TOP";
        s.t__DOT__A.bind(sh.t__DOT__A, threadIdx.x);
        s.t__DOT__B.bind(sh.t__DOT__B, threadIdx.x);
    }
    VLCU_D static void load(Root& s, const VlcuState& st, uint64_t i) {
        s.clk = st.ld<VlcuMem8>(OFF_clk, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        s.t__DOT__cyc = st.ld<IData>(OFF_t__DOT__cyc, i);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        s.t__DOT__crc = st.ld<QData>(OFF_t__DOT__crc, i);
        s.t__DOT__sum = st.ld<QData>(OFF_t__DOT__sum, i);
        for (uint32_t e = 0; e < 7; ++e) s.t__DOT__A.flat()[e] = st.ldE<IData>(OFF_t__DOT__A, e, i);
        for (uint32_t e = 0; e < 8; ++e) s.t__DOT__B.flat()[e] = st.ldE<IData>(OFF_t__DOT__B, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.t__DOT__a[e] = st.ldE<IData>(OFF_t__DOT__a, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.t__DOT__b[e] = st.ldE<IData>(OFF_t__DOT__b, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.st<IData>(OFF_t__DOT__cyc, i, s.t__DOT__cyc);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
        st.st<QData>(OFF_t__DOT__crc, i, s.t__DOT__crc);
        st.st<QData>(OFF_t__DOT__sum, i, s.t__DOT__sum);
        for (uint32_t e = 0; e < 7; ++e) st.stE<IData>(OFF_t__DOT__A, e, i, s.t__DOT__A.flat()[e]);
        for (uint32_t e = 0; e < 8; ++e) st.stE<IData>(OFF_t__DOT__B, e, i, s.t__DOT__B.flat()[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<IData>(OFF_t__DOT__a, e, i, s.t__DOT__a[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<IData>(OFF_t__DOT__b, e, i, s.t__DOT__b[e]);
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

    VLCU_D static void Vt_array_mda___024root___eval_static(Vt_array_mda___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        {
            // Inlined CFunc: _eval_static__TOP
            vlSelfRef.t__DOT__cyc = 0U;
        }
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
    }

    VLCU_D static void Vt_array_mda___024root___eval_initial(Vt_array_mda___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_array_mda___024root___eval_final(Vt_array_mda___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_array_mda___024root___eval_settle(Vt_array_mda___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static bool Vt_array_mda___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt_array_mda___024root___nba_sequent__TOP__0(Vt_array_mda___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        IData/*31:0*/ __Vdly__t__DOT__cyc;
        __Vdly__t__DOT__cyc = 0;
        QData/*63:0*/ __Vdly__t__DOT__crc;
        __Vdly__t__DOT__crc = 0;
        QData/*63:0*/ __Vdly__t__DOT__sum;
        __Vdly__t__DOT__sum = 0;
        IData/*31:0*/ __VdlyVal__t__DOT__a__v0;
        __VdlyVal__t__DOT__a__v0 = 0;
        IData/*31:0*/ __VdlyVal__t__DOT__b__v0;
        __VdlyVal__t__DOT__b__v0 = 0;
        CData/*0:0*/ __VdlySet__t__DOT__A__v0;
        __VdlySet__t__DOT__A__v0 = 0;
        // Body
        __Vdly__t__DOT__cyc = vlSelfRef.t__DOT__cyc;
        __Vdly__t__DOT__crc = vlSelfRef.t__DOT__crc;
        __Vdly__t__DOT__sum = vlSelfRef.t__DOT__sum;
        __VdlySet__t__DOT__A__v0 = 0U;
        __VdlyVal__t__DOT__a__v0 = vlSelfRef.t__DOT__A
            [((6U >= (7U & (IData)(vlSelfRef.t__DOT__crc)))
               ? (7U & (IData)(vlSelfRef.t__DOT__crc))
               : 0U)][0U];
        __VdlyVal__t__DOT__b__v0 = vlSelfRef.t__DOT__B
            [(7U & (IData)(vlSelfRef.t__DOT__crc))][0U];
        __Vdly__t__DOT__cyc = ((IData)(1U) + vlSelfRef.t__DOT__cyc);
        __Vdly__t__DOT__crc = ((vlSelfRef.t__DOT__crc 
                                << 1U) | (QData)((IData)(
                                                         (1U 
                                                          & VL_REDXOR_64(
                                                                         (0x8000000000000005ULL 
                                                                          & vlSelfRef.t__DOT__crc))))));
        __Vdly__t__DOT__sum = ((((QData)((IData)(vlSelfRef.t__DOT__a[0U])) 
                                 << 0x00000020U) | (QData)((IData)(vlSelfRef.t__DOT__b[0U]))) 
                               ^ ((vlSelfRef.t__DOT__sum 
                                   << 1U) | (QData)((IData)(
                                                            (1U 
                                                             & VL_REDXOR_64(
                                                                            (0x8000000000000005ULL 
                                                                             & vlSelfRef.t__DOT__sum)))))));
        if ((0U == vlSelfRef.t__DOT__cyc)) {
            __Vdly__t__DOT__crc = 0x5aef0c8dd70a4497ULL;
            __Vdly__t__DOT__sum = 0ULL;
            __VdlySet__t__DOT__A__v0 = 1U;
        } else if (VL_GTS_III(32, 0x0000000aU, vlSelfRef.t__DOT__cyc)) {
            __Vdly__t__DOT__sum = 0ULL;
        } else if (VL_LTES_III(32, 0x0000005aU, vlSelfRef.t__DOT__cyc)) {
            if (VL_UNLIKELY(((0x00000063U == vlSelfRef.t__DOT__cyc)))) {
                vlCtx.writef(0U, 4U, vlcu_msg_arg(VL_TIME_UNITED_Q(1)), vlcu_msg_arg(vlSelfRef.t__DOT__cyc), vlcu_msg_arg(vlSelfRef.t__DOT__crc), vlcu_msg_arg(vlSelfRef.t__DOT__sum));
                if (VL_UNLIKELY(((0xc77bb9b3784ea091ULL 
                                  != vlSelfRef.t__DOT__crc)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_array_mda.v", 61, "");
                }
                if (VL_UNLIKELY(((0x619f75c3a6d948bdULL 
                                  != vlSelfRef.t__DOT__sum)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_array_mda.v", 64, "");
                }
                vlCtx.writef(1U, 0U);
                VL_FINISH_MT("/root/reference/test_regress/t/t_array_mda.v", 66, "");
            }
        }
        vlSelfRef.t__DOT__cyc = __Vdly__t__DOT__cyc;
        vlSelfRef.t__DOT__a[0U] = __VdlyVal__t__DOT__a__v0;
        vlSelfRef.t__DOT__b[0U] = __VdlyVal__t__DOT__b__v0;
        vlSelfRef.t__DOT__sum = __Vdly__t__DOT__sum;
        vlSelfRef.t__DOT__crc = __Vdly__t__DOT__crc;
        if (__VdlySet__t__DOT__A__v0) {
            vlSelfRef.t__DOT__A[0U][0U] = 0x00000010U;
            vlSelfRef.t__DOT__A[1U][0U] = 0x00000011U;
            vlSelfRef.t__DOT__A[2U][0U] = 0x00000012U;
            vlSelfRef.t__DOT__A[3U][0U] = 0x00000013U;
            vlSelfRef.t__DOT__A[4U][0U] = 0x00000014U;
            vlSelfRef.t__DOT__A[5U][0U] = 0x00000015U;
            vlSelfRef.t__DOT__A[6U][0U] = 0x00000016U;
            vlSelfRef.t__DOT__B[0U][0U] = 0x00000020U;
            vlSelfRef.t__DOT__B[1U][0U] = 0x00000021U;
            vlSelfRef.t__DOT__B[2U][0U] = 0x00000022U;
            vlSelfRef.t__DOT__B[3U][0U] = 0x00000023U;
            vlSelfRef.t__DOT__B[4U][0U] = 0x00000024U;
            vlSelfRef.t__DOT__B[5U][0U] = 0x00000025U;
            vlSelfRef.t__DOT__B[6U][0U] = 0x00000026U;
            vlSelfRef.t__DOT__B[7U][0U] = 0x00000027U;
        }
    }

    VLCU_D static void Vt_array_mda___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt_array_mda___024root___eval_phase__act(Vt_array_mda___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt_array_mda___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vt_array_mda___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vt_array_mda___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vt_array_mda___024root___eval_phase__nba(Vt_array_mda___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vt_array_mda___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    Vt_array_mda___024root___nba_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
            Vt_array_mda___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vt_array_mda___024root___eval(Vt_array_mda___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt_array_mda___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("/root/reference/test_regress/t/t_array_mda.v", 7, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vt_array_mda___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("/root/reference/test_regress/t/t_array_mda.v", 7, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vt_array_mda___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vt_array_mda___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vt_array_mda___024root___ctor_var_reset(Vt_array_mda___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        vlSelf->t__DOT__crc = VL_SCOPED_RAND_RESET_Q(64, __VscopeHash, 17680824324711171631ull);
        vlSelf->t__DOT__sum = VL_SCOPED_RAND_RESET_Q(64, __VscopeHash, 1514072735762163239ull);
        for (int __Vi0 = 0; __Vi0 < 7; ++__Vi0) {
            for (int __Vi1 = 0; __Vi1 < 1; ++__Vi1) {
                vlSelf->t__DOT__A[__Vi0][__Vi1] = 0;
            }
        }
        for (int __Vi0 = 0; __Vi0 < 8; ++__Vi0) {
            for (int __Vi1 = 0; __Vi1 < 1; ++__Vi1) {
                vlSelf->t__DOT__B[__Vi0][__Vi1] = 0;
            }
        }
        for (int __Vi0 = 0; __Vi0 < 1; ++__Vi0) {
            vlSelf->t__DOT__a[__Vi0] = 0;
        }
        for (int __Vi0 = 0; __Vi0 < 1; ++__Vi0) {
            vlSelf->t__DOT__b[__Vi0] = 0;
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
        Vt_array_mda___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt_array_mda___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt_array_mda___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt_array_mda___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vt_array_mda___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt_array_mda___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt_array_mda__cuda::info() {
    typedef Vt_array_mda__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"t__DOT__cyc", "t.cyc", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__cyc},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"t__DOT__crc", "t.crc", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF_t__DOT__crc},
        {"t__DOT__sum", "t.sum", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF_t__DOT__sum},
        {"t__DOT__A", "t.A", VLCU_MEMBER_STATE, 4, 0, 7, 32, 31, 0, T::OFF_t__DOT__A},
        {"t__DOT__B", "t.B", VLCU_MEMBER_STATE, 4, 0, 8, 32, 31, 0, T::OFF_t__DOT__B},
        {"t__DOT__a", "t.a", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__a},
        {"t__DOT__b", "t.b", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__b},
        {"__VactTriggered", "__VactTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VactTriggered},
        {"__VnbaTriggered", "__VnbaTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VnbaTriggered},
    };
    static const char* const messages[] = {"[%0t] cyc==%0d crc=%x sum=%x\n", "*-* All Finished *-*\n", nullptr};
    static const char* const plusargs[] = {nullptr};
    static const VlcuTraceDecl traceDecls[] = {
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0}};  // (terminator, not counted)
    static const uint32_t traceOrder[] = {0};  // codes in dump order (+ terminator)
    static const uint32_t saveOrder[] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "t", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 2, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 14, 0, 0x396f93b4708e9aabULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt_array_mda__cuda VLCU_TESTBENCHES);
