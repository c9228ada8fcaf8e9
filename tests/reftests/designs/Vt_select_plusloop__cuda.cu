// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt_select_plusloop"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt_select_plusloop___024root final {
    CData/*0:0*/ clk;
    CData/*7:0*/ t__DOT__crc;
    CData/*6:0*/ t__DOT__index;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    IData/*31:0*/ t__DOT__narrow;
    VlWide<4>/*127:0*/ t__DOT__wide;
    IData/*31:0*/ t__DOT__cyc;
    IData/*31:0*/ __VactIterCount;
    QData/*63:0*/ t__DOT__quad;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vt_select_plusloop__cuda final {
    typedef Vt_select_plusloop___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 52,
        OFF_t__DOT__crc = 53,
        OFF_t__DOT__index = 54,
        OFF___Vtrigprevexpr___TOP__clk__0 = 55,
        OFF___VactPhaseResult = 56,
        OFF___VnbaPhaseResult = 57,
        OFF_t__DOT__narrow = 24,
        OFF_t__DOT__wide = 28,
        OFF_t__DOT__cyc = 44,
        OFF___VactIterCount = 48,
        OFF_t__DOT__quad = 0,
        OFF___VactTriggered = 8,
        OFF___VnbaTriggered = 16,
        kStateBytes = 64,  // S_persist
        kTouchedReadBytes = 58,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 57,  // S_state_wr
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
        s.t__DOT__crc = st.ld<VlcuMem8>(OFF_t__DOT__crc, i);
        s.t__DOT__index = st.ld<VlcuMem8>(OFF_t__DOT__index, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        s.t__DOT__narrow = st.ld<IData>(OFF_t__DOT__narrow, i);
        st.ldW(OFF_t__DOT__wide, i, s.t__DOT__wide);
        s.t__DOT__cyc = st.ld<IData>(OFF_t__DOT__cyc, i);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        s.t__DOT__quad = st.ld<QData>(OFF_t__DOT__quad, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.st<VlcuMem8>(OFF_t__DOT__crc, i, s.t__DOT__crc);
        st.st<VlcuMem8>(OFF_t__DOT__index, i, s.t__DOT__index);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.st<IData>(OFF_t__DOT__narrow, i, s.t__DOT__narrow);
        st.stW(OFF_t__DOT__wide, i, s.t__DOT__wide);
        st.st<IData>(OFF_t__DOT__cyc, i, s.t__DOT__cyc);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
        st.st<QData>(OFF_t__DOT__quad, i, s.t__DOT__quad);
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

    VLCU_D static void Vt_select_plusloop___024root___eval_static(Vt_select_plusloop___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
    }

    VLCU_D static void Vt_select_plusloop___024root___eval_initial(Vt_select_plusloop___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        {
            // Inlined CFunc: _eval_initial__TOP
            vlSelfRef.t__DOT__cyc = 0U;
        }
    }

    VLCU_D static void Vt_select_plusloop___024root___eval_final(Vt_select_plusloop___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_select_plusloop___024root___eval_settle(Vt_select_plusloop___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static bool Vt_select_plusloop___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt_select_plusloop___024root___nba_sequent__TOP__0(Vt_select_plusloop___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        IData/*31:0*/ __Vdly__t__DOT__cyc;
        __Vdly__t__DOT__cyc = 0;
        IData/*31:0*/ __Vdly__t__DOT__narrow;
        __Vdly__t__DOT__narrow = 0;
        QData/*63:0*/ __Vdly__t__DOT__quad;
        __Vdly__t__DOT__quad = 0;
        VlWide<4>/*127:0*/ __Vdly__t__DOT__wide;
        VL_ZERO_W(128, __Vdly__t__DOT__wide);
        CData/*7:0*/ __Vdly__t__DOT__crc;
        __Vdly__t__DOT__crc = 0;
        CData/*6:0*/ __Vdly__t__DOT__index;
        __Vdly__t__DOT__index = 0;
        // Body
        __Vdly__t__DOT__cyc = vlSelfRef.t__DOT__cyc;
        __Vdly__t__DOT__narrow = vlSelfRef.t__DOT__narrow;
        __Vdly__t__DOT__quad = vlSelfRef.t__DOT__quad;
        __Vdly__t__DOT__wide[0U] = vlSelfRef.t__DOT__wide[0U];
        __Vdly__t__DOT__wide[1U] = vlSelfRef.t__DOT__wide[1U];
        __Vdly__t__DOT__wide[2U] = vlSelfRef.t__DOT__wide[2U];
        __Vdly__t__DOT__wide[3U] = vlSelfRef.t__DOT__wide[3U];
        __Vdly__t__DOT__crc = vlSelfRef.t__DOT__crc;
        __Vdly__t__DOT__index = vlSelfRef.t__DOT__index;
        __Vdly__t__DOT__cyc = ((IData)(1U) + vlSelfRef.t__DOT__cyc);
        if ((0U == vlSelfRef.t__DOT__cyc)) {
            __Vdly__t__DOT__narrow = 0U;
            __Vdly__t__DOT__quad = 0ULL;
            __Vdly__t__DOT__wide[0U] = 0U;
            __Vdly__t__DOT__wide[1U] = 0U;
            __Vdly__t__DOT__wide[2U] = 0U;
            __Vdly__t__DOT__wide[3U] = 0U;
            __Vdly__t__DOT__crc = 0xedU;
            __Vdly__t__DOT__index = 0U;
        } else if (VL_GTS_III(32, 0x0000005aU, vlSelfRef.t__DOT__cyc)) {
            __Vdly__t__DOT__index = (0x0000007fU & 
                                     ((IData)(2U) + (IData)(vlSelfRef.t__DOT__index)));
            __Vdly__t__DOT__crc = ((0x000000feU & ((IData)(vlSelfRef.t__DOT__crc) 
                                                   << 1U)) 
                                   | (1U & (~ (1U & 
                                               VL_REDXOR_8(
                                                           (0xb8U 
                                                            & (IData)(vlSelfRef.t__DOT__crc)))))));
            if ((0x14U > (IData)(vlSelfRef.t__DOT__index))) {
                __Vdly__t__DOT__narrow = (((~ ((IData)(7U) 
                                               << (0x0000001fU 
                                                   & (IData)(vlSelfRef.t__DOT__index)))) 
                                           & __Vdly__t__DOT__narrow) 
                                          | (0x00000000ffffffffULL 
                                             & ((7U 
                                                 & (IData)(vlSelfRef.t__DOT__crc)) 
                                                << 
                                                (0x0000001fU 
                                                 & (IData)(vlSelfRef.t__DOT__index)))));
            }
            if ((0x3cU > (IData)(vlSelfRef.t__DOT__index))) {
                __Vdly__t__DOT__quad = (((~ (7ULL << 
                                             (0x0000003fU 
                                              & (IData)(vlSelfRef.t__DOT__index)))) 
                                         & __Vdly__t__DOT__quad) 
                                        | ((QData)((IData)(
                                                           (7U 
                                                            & (IData)(vlSelfRef.t__DOT__crc)))) 
                                           << (0x0000003fU 
                                               & (IData)(vlSelfRef.t__DOT__index))));
            }
            if ((0x78U > (IData)(vlSelfRef.t__DOT__index))) {
                VL_ASSIGNSEL_WI(128, 3, (IData)(vlSelfRef.t__DOT__index), __Vdly__t__DOT__wide, 
                                (7U & (IData)(vlSelfRef.t__DOT__crc)));
            }
            __Vdly__t__DOT__narrow = (((~ ((IData)(1U) 
                                           << (0x0000000fU 
                                               & (IData)(vlSelfRef.t__DOT__index)))) 
                                       & __Vdly__t__DOT__narrow) 
                                      | (0x00000000ffffffffULL 
                                         & ((1U & (~ 
                                                   (vlSelfRef.t__DOT__narrow 
                                                    >> 
                                                    (0x0000000fU 
                                                     & (IData)(vlSelfRef.t__DOT__index))))) 
                                            << (0x0000000fU 
                                                & (IData)(vlSelfRef.t__DOT__index)))));
            __Vdly__t__DOT__quad = (((~ (1ULL << (0x0000000fU 
                                                  & (~ (IData)(vlSelfRef.t__DOT__index))))) 
                                     & __Vdly__t__DOT__quad) 
                                    | ((QData)((IData)(
                                                       (1U 
                                                        & (~ (IData)(
                                                                     (vlSelfRef.t__DOT__quad 
                                                                      >> 
                                                                      (0x0000000fU 
                                                                       & (~ (IData)(vlSelfRef.t__DOT__index))))))))) 
                                       << (0x0000000fU 
                                           & (~ (IData)(vlSelfRef.t__DOT__index)))));
            __Vdly__t__DOT__wide[0U] = (((~ ((IData)(1U) 
                                             << (0x0000000fU 
                                                 & (~ (IData)(vlSelfRef.t__DOT__index))))) 
                                         & __Vdly__t__DOT__wide[0U]) 
                                        | ((1U & (~ 
                                                  (vlSelfRef.t__DOT__wide[0U] 
                                                   >> 
                                                   (0x0000000fU 
                                                    & (~ (IData)(vlSelfRef.t__DOT__index)))))) 
                                           << (0x0000000fU 
                                               & (~ (IData)(vlSelfRef.t__DOT__index)))));
        } else if ((0x0000005aU == vlSelfRef.t__DOT__cyc)) {
            __Vdly__t__DOT__wide[0U] = (0x00006000U 
                                        | (0xffff0fffU 
                                           & __Vdly__t__DOT__wide[0U]));
            __Vdly__t__DOT__quad = (0x0000000000006000ULL 
                                    | (0xffffffffffff0fffULL 
                                       & __Vdly__t__DOT__quad));
            __Vdly__t__DOT__narrow = (0x00006000U | 
                                      (0xffff0fffU 
                                       & __Vdly__t__DOT__narrow));
            __Vdly__t__DOT__wide[1U] = (0x00001800U 
                                        | (0xffffc3ffU 
                                           & __Vdly__t__DOT__wide[1U]));
            __Vdly__t__DOT__quad = (0x0000180000000000ULL 
                                    | (0xffffc3ffffffffffULL 
                                       & __Vdly__t__DOT__quad));
            __Vdly__t__DOT__wide[2U] = (0x00180000U 
                                        | (0xffc3ffffU 
                                           & __Vdly__t__DOT__wide[2U]));
        } else if ((0x0000005bU == vlSelfRef.t__DOT__cyc)) {
            __Vdly__t__DOT__wide[0U] = (1U | __Vdly__t__DOT__wide[0U]);
            __Vdly__t__DOT__quad = (1ULL | __Vdly__t__DOT__quad);
            __Vdly__t__DOT__narrow = (1U | __Vdly__t__DOT__narrow);
            __Vdly__t__DOT__wide[1U] = (0x00000200U 
                                        | __Vdly__t__DOT__wide[1U]);
            __Vdly__t__DOT__quad = (0x0000020000000000ULL 
                                    | __Vdly__t__DOT__quad);
            __Vdly__t__DOT__wide[2U] = (0x00020000U 
                                        | __Vdly__t__DOT__wide[2U]);
        } else if (VL_UNLIKELY(((0x00000063U == vlSelfRef.t__DOT__cyc)))) {
            vlCtx.writef(0U, 6U, vlcu_msg_arg(VL_TIME_UNITED_Q(1)), vlcu_msg_arg(vlSelfRef.t__DOT__cyc), vlcu_msg_arg((IData)(vlSelfRef.t__DOT__crc)), vlcu_msg_arg(vlSelfRef.t__DOT__narrow), vlcu_msg_arg(vlSelfRef.t__DOT__quad), vlcu_msg_arg(vlSelfRef.t__DOT__wide));
            if (VL_UNLIKELY(((0x79U != (IData)(vlSelfRef.t__DOT__crc))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_select_plusloop.v", 62, "");
            }
            if (VL_UNLIKELY(((0x001661c7U != vlSelfRef.t__DOT__narrow)))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_select_plusloop.v", 63, "");
            }
            if (VL_UNLIKELY(((0x16d49b6f64266039ULL 
                              != vlSelfRef.t__DOT__quad)))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_select_plusloop.v", 64, "");
            }
            if (VL_UNLIKELY(((0U != ((((0x64266039U 
                                        ^ vlSelfRef.t__DOT__wide[0U]) 
                                       | (0xf6d49b6fU 
                                          ^ vlSelfRef.t__DOT__wide[1U])) 
                                      | (0x265b266fU 
                                         ^ vlSelfRef.t__DOT__wide[2U])) 
                                     | (0x012fd26dU 
                                        ^ vlSelfRef.t__DOT__wide[3U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_select_plusloop.v", 65, "");
            }
            vlCtx.writef(1U, 0U);
            VL_FINISH_MT("/root/reference/test_regress/t/t_select_plusloop.v", 67, "");
        }
        vlSelfRef.t__DOT__cyc = __Vdly__t__DOT__cyc;
        vlSelfRef.t__DOT__narrow = __Vdly__t__DOT__narrow;
        vlSelfRef.t__DOT__quad = __Vdly__t__DOT__quad;
        vlSelfRef.t__DOT__wide[0U] = __Vdly__t__DOT__wide[0U];
        vlSelfRef.t__DOT__wide[1U] = __Vdly__t__DOT__wide[1U];
        vlSelfRef.t__DOT__wide[2U] = __Vdly__t__DOT__wide[2U];
        vlSelfRef.t__DOT__wide[3U] = __Vdly__t__DOT__wide[3U];
        vlSelfRef.t__DOT__crc = __Vdly__t__DOT__crc;
        vlSelfRef.t__DOT__index = __Vdly__t__DOT__index;
    }

    VLCU_D static void Vt_select_plusloop___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt_select_plusloop___024root___eval_phase__act(Vt_select_plusloop___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt_select_plusloop___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vt_select_plusloop___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vt_select_plusloop___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vt_select_plusloop___024root___eval_phase__nba(Vt_select_plusloop___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vt_select_plusloop___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    Vt_select_plusloop___024root___nba_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
            Vt_select_plusloop___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vt_select_plusloop___024root___eval(Vt_select_plusloop___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt_select_plusloop___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("/root/reference/test_regress/t/t_select_plusloop.v", 7, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vt_select_plusloop___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("/root/reference/test_regress/t/t_select_plusloop.v", 7, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vt_select_plusloop___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vt_select_plusloop___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vt_select_plusloop___024root___ctor_var_reset(Vt_select_plusloop___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        vlSelf->t__DOT__narrow = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 16957994623349195476ull);
        vlSelf->t__DOT__quad = VL_SCOPED_RAND_RESET_Q(64, __VscopeHash, 17587652640057386158ull);
        VL_SCOPED_RAND_RESET_W(128, vlSelf->t__DOT__wide, __VscopeHash, 13798169374321348451ull);
        vlSelf->t__DOT__cyc = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 15508623621340971111ull);
        vlSelf->t__DOT__crc = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 17680824324711171631ull);
        vlSelf->t__DOT__index = VL_SCOPED_RAND_RESET_I(7, __VscopeHash, 2147293585595700564ull);
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
        Vt_select_plusloop___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt_select_plusloop___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt_select_plusloop___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt_select_plusloop___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vt_select_plusloop___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt_select_plusloop___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt_select_plusloop__cuda::info() {
    typedef Vt_select_plusloop__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"t__DOT__crc", "t.crc", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF_t__DOT__crc},
        {"t__DOT__index", "t.index", VLCU_MEMBER_STATE, 1, 0, 1, 7, 6, 0, T::OFF_t__DOT__index},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"t__DOT__narrow", "t.narrow", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__narrow},
        {"t__DOT__wide", "t.wide", VLCU_MEMBER_STATE, 4, 0, 4, 128, 127, 0, T::OFF_t__DOT__wide},
        {"t__DOT__cyc", "t.cyc", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__cyc},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"t__DOT__quad", "t.quad", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF_t__DOT__quad},
        {"__VactTriggered", "__VactTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VactTriggered},
        {"__VnbaTriggered", "__VnbaTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VnbaTriggered},
    };
    static const char* const messages[] = {"[%0t] cyc==%0d crc=%b n=%x q=%x w=%x\n", "*-* All Finished *-*\n", nullptr};
    static const char* const plusargs[] = {nullptr};
    static const VlcuTraceDecl traceDecls[] = {
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0}};  // (terminator, not counted)
    static const uint32_t traceOrder[] = {0};  // codes in dump order (+ terminator)
    static const uint32_t saveOrder[] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "t", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 2, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 13, 0, 0xcd8b4f7314e80764ULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt_select_plusloop__cuda VLCU_TESTBENCHES);
