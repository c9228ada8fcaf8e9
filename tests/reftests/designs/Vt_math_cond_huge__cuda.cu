// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt_math_cond_huge"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt_math_cond_huge___024root final {
    CData/*0:0*/ clk;
    CData/*5:0*/ __VdfgRegularize_h6e95ff9d_0_1;
    CData/*0:0*/ __VstlFirstIteration;
    CData/*0:0*/ __VstlPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    IData/*31:0*/ t__DOT__cyc;
    VlWide<8>/*255:0*/ __VdfgRegularize_h6e95ff9d_0_0;
    IData/*31:0*/ __VactIterCount;
    QData/*63:0*/ t__DOT__crc;
    QData/*63:0*/ t__DOT__sum;
    VlUnpacked<QData/*63:0*/, 1> __VstlTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vt_math_cond_huge__cuda final {
    typedef Vt_math_cond_huge___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 80,
        OFF___VdfgRegularize_h6e95ff9d_0_1 = 81,
        OFF___VstlFirstIteration = 82,
        OFF___VstlPhaseResult = 83,
        OFF___Vtrigprevexpr___TOP__clk__0 = 84,
        OFF___VactPhaseResult = 85,
        OFF___VnbaPhaseResult = 86,
        OFF_t__DOT__cyc = 40,
        OFF___VdfgRegularize_h6e95ff9d_0_0 = 44,
        OFF___VactIterCount = 76,
        OFF_t__DOT__crc = 0,
        OFF_t__DOT__sum = 8,
        OFF___VstlTriggered = 16,
        OFF___VactTriggered = 24,
        OFF___VnbaTriggered = 32,
        kStateBytes = 88,  // S_persist
        kTouchedReadBytes = 77,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 76,  // S_state_wr
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
        s.__VdfgRegularize_h6e95ff9d_0_1 = st.ld<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_1, i);
        s.__VstlFirstIteration = st.ld<VlcuMem8>(OFF___VstlFirstIteration, i);
        s.__VstlPhaseResult = st.ld<VlcuMem8>(OFF___VstlPhaseResult, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        s.t__DOT__cyc = st.ld<IData>(OFF_t__DOT__cyc, i);
        st.ldW(OFF___VdfgRegularize_h6e95ff9d_0_0, i, s.__VdfgRegularize_h6e95ff9d_0_0);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        s.t__DOT__crc = st.ld<QData>(OFF_t__DOT__crc, i);
        s.t__DOT__sum = st.ld<QData>(OFF_t__DOT__sum, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VstlTriggered[e] = st.ldE<QData>(OFF___VstlTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.st<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_1, i, s.__VdfgRegularize_h6e95ff9d_0_1);
        st.st<VlcuMem8>(OFF___VstlFirstIteration, i, s.__VstlFirstIteration);
        st.st<VlcuMem8>(OFF___VstlPhaseResult, i, s.__VstlPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.st<IData>(OFF_t__DOT__cyc, i, s.t__DOT__cyc);
        st.stW(OFF___VdfgRegularize_h6e95ff9d_0_0, i, s.__VdfgRegularize_h6e95ff9d_0_0);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
        st.st<QData>(OFF_t__DOT__crc, i, s.t__DOT__crc);
        st.st<QData>(OFF_t__DOT__sum, i, s.t__DOT__sum);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VstlTriggered, e, i, s.__VstlTriggered[e]);
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

    VLCU_D static void Vt_math_cond_huge___024root___eval_static(Vt_math_cond_huge___024root* vlSelf, VlcuContext& vlCtx) {
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

    VLCU_D static void Vt_math_cond_huge___024root___eval_initial(Vt_math_cond_huge___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_math_cond_huge___024root___eval_final(Vt_math_cond_huge___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_math_cond_huge___024root___eval_settle(Vt_math_cond_huge___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt_math_cond_huge___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
#endif
                VL_FATAL_MT("/root/reference/test_regress/t/t_math_cond_huge.v", 7, "", "DIDNOTCONVERGE: Settle region did not converge after '--converge-limit' of 10000 tries");
            }
            __VstlIterCount = ((IData)(1U) + __VstlIterCount);
            vlSelfRef.__VstlPhaseResult = Vt_math_cond_huge___024root___eval_phase__stl(vlSelf, vlCtx);
            vlSelfRef.__VstlFirstIteration = 0U;
        } while (vlSelfRef.__VstlPhaseResult);
    }

    VLCU_D static bool Vt_math_cond_huge___024root___trigger_anySet__stl(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt_math_cond_huge___024root___eval_phase__stl(Vt_math_cond_huge___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt_math_cond_huge___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
        }
#endif
        __VstlExecute = Vt_math_cond_huge___024root___trigger_anySet__stl(vlSelfRef.__VstlTriggered);
        if (__VstlExecute) {
            {
                // Inlined CFunc: _eval_stl
                if ((1ULL & vlSelfRef.__VstlTriggered[0U])) {
                    {
                        // Inlined CFunc: _stl_sequent__TOP__0
                        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[0U] 
                            = (IData)(vlSelfRef.t__DOT__crc);
                        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[1U] 
                            = (IData)((vlSelfRef.t__DOT__crc 
                                       >> 0x00000020U));
                        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[2U] 
                            = (IData)(vlSelfRef.t__DOT__crc);
                        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[3U] 
                            = (IData)((vlSelfRef.t__DOT__crc 
                                       >> 0x00000020U));
                        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[4U] 
                            = (IData)(vlSelfRef.t__DOT__crc);
                        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[5U] 
                            = (IData)((vlSelfRef.t__DOT__crc 
                                       >> 0x00000020U));
                        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[6U] 
                            = (IData)(vlSelfRef.t__DOT__crc);
                        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[7U] 
                            = (IData)((vlSelfRef.t__DOT__crc 
                                       >> 0x00000020U));
                        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1 
                            = ((0x00000038U & ((IData)(vlSelfRef.t__DOT__crc) 
                                               << 3U)) 
                               | (7U & (IData)((vlSelfRef.t__DOT__crc 
                                                >> 0x0000003dU))));
                    }
                }
            }
        }
        return (__VstlExecute);
    }

    VLCU_D static bool Vt_math_cond_huge___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt_math_cond_huge___024root___nba_sequent__TOP__0(Vt_math_cond_huge___024root* vlSelf, VlcuContext& vlCtx) {
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
        CData/*31:0*/ __Vdeeptemp_hb3d546fb__0;
        CData/*31:0*/ __Vdeeptemp_h914a33a1__0;
        CData/*31:0*/ __Vdeeptemp_hb1633d74__0;
        // Body
        __Vdly__t__DOT__cyc = vlSelfRef.t__DOT__cyc;
        __Vdly__t__DOT__sum = vlSelfRef.t__DOT__sum;
        __Vdly__t__DOT__crc = vlSelfRef.t__DOT__crc;
        __Vdly__t__DOT__cyc = ((IData)(1U) + vlSelfRef.t__DOT__cyc);
        __Vdly__t__DOT__crc = ((vlSelfRef.t__DOT__crc 
                                << 1U) | (QData)((IData)(
                                                         (1U 
                                                          & VL_REDXOR_64(
                                                                         (0x8000000000000005ULL 
                                                                          & vlSelfRef.t__DOT__crc))))));
        __Vdeeptemp_hb3d546fb__0 = ((0x8aU == (0x000000ffU 
                                               & (IData)(vlSelfRef.t__DOT__crc)))
                                     ? (IData)((vlSelfRef.t__DOT__crc 
                                                >> 0x0000000aU))
                                     : ((0x8bU == (0x000000ffU 
                                                   & (IData)(vlSelfRef.t__DOT__crc)))
                                         ? (IData)(
                                                   (vlSelfRef.t__DOT__crc 
                                                    >> 0x0000000bU))
                                         : ((0x8cU 
                                             == (0x000000ffU 
                                                 & (IData)(vlSelfRef.t__DOT__crc)))
                                             ? (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 0x0000000cU))
                                             : ((0x8dU 
                                                 == 
                                                 (0x000000ffU 
                                                  & (IData)(vlSelfRef.t__DOT__crc)))
                                                 ? (IData)(
                                                           (vlSelfRef.t__DOT__crc 
                                                            >> 0x0000000dU))
                                                 : 
                                                ((0x8eU 
                                                  == 
                                                  (0x000000ffU 
                                                   & (IData)(vlSelfRef.t__DOT__crc)))
                                                  ? (IData)(
                                                            (vlSelfRef.t__DOT__crc 
                                                             >> 0x0000000eU))
                                                  : 
                                                 ((0x8fU 
                                                   == 
                                                   (0x000000ffU 
                                                    & (IData)(vlSelfRef.t__DOT__crc)))
                                                   ? (IData)(
                                                             (vlSelfRef.t__DOT__crc 
                                                              >> 0x0000000fU))
                                                   : 
                                                  ((0x90U 
                                                    == 
                                                    (0x000000ffU 
                                                     & (IData)(vlSelfRef.t__DOT__crc)))
                                                    ? (IData)(
                                                              (vlSelfRef.t__DOT__crc 
                                                               >> 0x00000010U))
                                                    : 
                                                   ((0x91U 
                                                     == 
                                                     (0x000000ffU 
                                                      & (IData)(vlSelfRef.t__DOT__crc)))
                                                     ? (IData)(
                                                               (vlSelfRef.t__DOT__crc 
                                                                >> 0x00000011U))
                                                     : 
                                                    ((0x92U 
                                                      == 
                                                      (0x000000ffU 
                                                       & (IData)(vlSelfRef.t__DOT__crc)))
                                                      ? (IData)(
                                                                (vlSelfRef.t__DOT__crc 
                                                                 >> 0x00000012U))
                                                      : 
                                                     ((0x93U 
                                                       == 
                                                       (0x000000ffU 
                                                        & (IData)(vlSelfRef.t__DOT__crc)))
                                                       ? (IData)(
                                                                 (vlSelfRef.t__DOT__crc 
                                                                  >> 0x00000013U))
                                                       : 
                                                      ((0x94U 
                                                        == 
                                                        (0x000000ffU 
                                                         & (IData)(vlSelfRef.t__DOT__crc)))
                                                        ? (IData)(
                                                                  (vlSelfRef.t__DOT__crc 
                                                                   >> 0x00000014U))
                                                        : 
                                                       ((0x95U 
                                                         == 
                                                         (0x000000ffU 
                                                          & (IData)(vlSelfRef.t__DOT__crc)))
                                                         ? (IData)(
                                                                   (vlSelfRef.t__DOT__crc 
                                                                    >> 0x00000015U))
                                                         : 
                                                        ((0x96U 
                                                          == 
                                                          (0x000000ffU 
                                                           & (IData)(vlSelfRef.t__DOT__crc)))
                                                          ? (IData)(
                                                                    (vlSelfRef.t__DOT__crc 
                                                                     >> 0x00000016U))
                                                          : 
                                                         ((0x98U 
                                                           == 
                                                           (0x000000ffU 
                                                            & (IData)(vlSelfRef.t__DOT__crc)))
                                                           ? (IData)(
                                                                     (vlSelfRef.t__DOT__crc 
                                                                      >> 0x00000017U))
                                                           : 
                                                          ((0x99U 
                                                            == 
                                                            (0x000000ffU 
                                                             & (IData)(vlSelfRef.t__DOT__crc)))
                                                            ? (IData)(
                                                                      (vlSelfRef.t__DOT__crc 
                                                                       >> 0x00000018U))
                                                            : 
                                                           ((0x9aU 
                                                             == 
                                                             (0x000000ffU 
                                                              & (IData)(vlSelfRef.t__DOT__crc)))
                                                             ? (IData)(
                                                                       (vlSelfRef.t__DOT__crc 
                                                                        >> 0x00000019U))
                                                             : 
                                                            ((0x9bU 
                                                              == 
                                                              (0x000000ffU 
                                                               & (IData)(vlSelfRef.t__DOT__crc)))
                                                              ? (IData)(
                                                                        (vlSelfRef.t__DOT__crc 
                                                                         >> 0x0000001aU))
                                                              : 
                                                             ((0x9cU 
                                                               == 
                                                               (0x000000ffU 
                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                               ? (IData)(
                                                                         (vlSelfRef.t__DOT__crc 
                                                                          >> 0x0000001bU))
                                                               : 
                                                              ((0x9dU 
                                                                == 
                                                                (0x000000ffU 
                                                                 & (IData)(vlSelfRef.t__DOT__crc)))
                                                                ? (IData)(
                                                                          (vlSelfRef.t__DOT__crc 
                                                                           >> 0x0000001cU))
                                                                : 
                                                               ((0x9eU 
                                                                 == 
                                                                 (0x000000ffU 
                                                                  & (IData)(vlSelfRef.t__DOT__crc)))
                                                                 ? (IData)(
                                                                           (vlSelfRef.t__DOT__crc 
                                                                            >> 0x0000001dU))
                                                                 : 
                                                                ((0x9fU 
                                                                  == 
                                                                  (0x000000ffU 
                                                                   & (IData)(vlSelfRef.t__DOT__crc)))
                                                                  ? (IData)(
                                                                            (vlSelfRef.t__DOT__crc 
                                                                             >> 0x0000001eU))
                                                                  : 
                                                                 ((0xa0U 
                                                                   == 
                                                                   (0x000000ffU 
                                                                    & (IData)(vlSelfRef.t__DOT__crc)))
                                                                   ? (IData)(
                                                                             (vlSelfRef.t__DOT__crc 
                                                                              >> 0x0000001fU))
                                                                   : 
                                                                  ((0xa1U 
                                                                    == 
                                                                    (0x000000ffU 
                                                                     & (IData)(vlSelfRef.t__DOT__crc)))
                                                                    ? (IData)(
                                                                              (vlSelfRef.t__DOT__crc 
                                                                               >> 0x00000020U))
                                                                    : 
                                                                   ((0xa2U 
                                                                     == 
                                                                     (0x000000ffU 
                                                                      & (IData)(vlSelfRef.t__DOT__crc)))
                                                                     ? (IData)(
                                                                               (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000021U))
                                                                     : 
                                                                    ((0xa3U 
                                                                      == 
                                                                      (0x000000ffU 
                                                                       & (IData)(vlSelfRef.t__DOT__crc)))
                                                                      ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000022U))
                                                                      : 
                                                                     ((0xa4U 
                                                                       == 
                                                                       (0x000000ffU 
                                                                        & (IData)(vlSelfRef.t__DOT__crc)))
                                                                       ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000023U))
                                                                       : 
                                                                      ((0xa5U 
                                                                        == 
                                                                        (0x000000ffU 
                                                                         & (IData)(vlSelfRef.t__DOT__crc)))
                                                                        ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000024U))
                                                                        : 
                                                                       ((0xa6U 
                                                                         == 
                                                                         (0x000000ffU 
                                                                          & (IData)(vlSelfRef.t__DOT__crc)))
                                                                         ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000025U))
                                                                         : 
                                                                        ((0xa7U 
                                                                          == 
                                                                          (0x000000ffU 
                                                                           & (IData)(vlSelfRef.t__DOT__crc)))
                                                                          ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000026U))
                                                                          : 
                                                                         ((0xa8U 
                                                                           == 
                                                                           (0x000000ffU 
                                                                            & (IData)(vlSelfRef.t__DOT__crc)))
                                                                           ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000027U))
                                                                           : 
                                                                          ((0xa9U 
                                                                            == 
                                                                            (0x000000ffU 
                                                                             & (IData)(vlSelfRef.t__DOT__crc)))
                                                                            ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000028U))
                                                                            : 
                                                                           ((0xaaU 
                                                                             == 
                                                                             (0x000000ffU 
                                                                              & (IData)(vlSelfRef.t__DOT__crc)))
                                                                             ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000029U))
                                                                             : 
                                                                            ((0xabU 
                                                                              == 
                                                                              (0x000000ffU 
                                                                               & (IData)(vlSelfRef.t__DOT__crc)))
                                                                              ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000002aU))
                                                                              : 
                                                                             ((0xacU 
                                                                               == 
                                                                               (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                               ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000002bU))
                                                                               : 
                                                                              ((0xadU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000002cU))
                                                                                : 
                                                                               ((0xaeU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000002dU))
                                                                                 : 
                                                                                ((0xafU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000002eU))
                                                                                 : 
                                                                                ((0xb0U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000002fU))
                                                                                 : 
                                                                                ((0xb1U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000030U))
                                                                                 : 
                                                                                ((0xb2U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000031U))
                                                                                 : 
                                                                                ((0xb3U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000032U))
                                                                                 : 
                                                                                ((0xb4U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000033U))
                                                                                 : 
                                                                                ((0xb5U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000034U))
                                                                                 : 
                                                                                ((0xb6U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000035U))
                                                                                 : 
                                                                                ((0xb7U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000036U))
                                                                                 : 
                                                                                ((0xb8U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000037U))
                                                                                 : 
                                                                                ((0xb9U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000038U))
                                                                                 : 
                                                                                ((0xbaU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000039U))
                                                                                 : 
                                                                                ((0xbbU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000003aU))
                                                                                 : 
                                                                                ((0xbcU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000003bU))
                                                                                 : 
                                                                                ((0xbdU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000003cU))
                                                                                 : 
                                                                                ((0xbeU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? 
                                                                                ((vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[6U] 
                                                                                << 3U) 
                                                                                | (vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[5U] 
                                                                                >> 0x0000001dU))
                                                                                 : 
                                                                                ((0xbfU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? 
                                                                                ((vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[6U] 
                                                                                << 2U) 
                                                                                | (vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[5U] 
                                                                                >> 0x0000001eU))
                                                                                 : 
                                                                                ((0xc0U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? 
                                                                                ((vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[6U] 
                                                                                << 1U) 
                                                                                | (vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[5U] 
                                                                                >> 0x0000001fU))
                                                                                 : 
                                                                                ((0xc1U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(vlSelfRef.t__DOT__crc)
                                                                                 : 
                                                                                ((0xc2U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 1U))
                                                                                 : 
                                                                                ((0xc3U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 2U))
                                                                                 : 
                                                                                ((0xc4U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 3U))
                                                                                 : 
                                                                                ((0xc5U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 4U))
                                                                                 : 
                                                                                ((0xc6U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 5U))
                                                                                 : 
                                                                                ((0xc7U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 6U))
                                                                                 : 
                                                                                ((0xc8U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 7U))
                                                                                 : 
                                                                                ((0xc9U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 8U))
                                                                                 : 
                                                                                ((0xcaU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 9U))
                                                                                 : 
                                                                                ((0xcbU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000000aU))
                                                                                 : 
                                                                                ((0xccU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000000bU))
                                                                                 : 
                                                                                ((0xcdU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000000cU))
                                                                                 : 
                                                                                ((0xceU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000000dU))
                                                                                 : 
                                                                                ((0xcfU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000000eU))
                                                                                 : 
                                                                                ((0xd0U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000000fU))
                                                                                 : 
                                                                                ((0xd1U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000010U))
                                                                                 : 
                                                                                ((0xd2U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000011U))
                                                                                 : 
                                                                                ((0xd3U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000012U))
                                                                                 : 
                                                                                ((0xd4U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000013U))
                                                                                 : 
                                                                                ((0xd5U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000014U))
                                                                                 : 
                                                                                ((0xd6U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000015U))
                                                                                 : 
                                                                                ((0xd7U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000016U))
                                                                                 : 
                                                                                ((0xd8U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000017U))
                                                                                 : 
                                                                                ((0xd9U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000018U))
                                                                                 : 
                                                                                ((0xdaU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000019U))
                                                                                 : 
                                                                                ((0xdbU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000001aU))
                                                                                 : 
                                                                                ((0xdcU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000001bU))
                                                                                 : 
                                                                                ((0xddU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000001cU))
                                                                                 : 
                                                                                ((0xdeU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000001dU))
                                                                                 : 
                                                                                ((0xdfU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000001eU))
                                                                                 : 
                                                                                ((0xe0U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000001fU))
                                                                                 : 
                                                                                ((0xe1U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000020U))
                                                                                 : 
                                                                                ((0xe2U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000021U))
                                                                                 : 
                                                                                ((0xe3U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000022U))
                                                                                 : 
                                                                                ((0xe4U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000023U))
                                                                                 : 
                                                                                ((0xe5U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000024U))
                                                                                 : 
                                                                                ((0xe6U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000025U))
                                                                                 : 
                                                                                ((0xe7U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000026U))
                                                                                 : 
                                                                                ((0xe8U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000027U))
                                                                                 : 
                                                                                ((0xe9U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000028U))
                                                                                 : 
                                                                                ((0xeaU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000029U))
                                                                                 : 
                                                                                ((0xebU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000002aU))
                                                                                 : 
                                                                                ((0xecU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000002bU))
                                                                                 : 
                                                                                ((0xedU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000002cU))
                                                                                 : 
                                                                                ((0xeeU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000002dU))
                                                                                 : 
                                                                                ((0xefU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000002eU))
                                                                                 : 
                                                                                ((0xf0U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000002fU))
                                                                                 : 
                                                                                ((0xf1U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000030U))
                                                                                 : 
                                                                                ((0xf2U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000031U))
                                                                                 : 
                                                                                ((0xf3U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000032U))
                                                                                 : 
                                                                                ((0xf4U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000033U))
                                                                                 : 
                                                                                ((0xf5U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000034U))
                                                                                 : 
                                                                                ((0xf6U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000035U))
                                                                                 : 
                                                                                ((0xf7U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000036U))
                                                                                 : 
                                                                                ((0xf8U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000037U))
                                                                                 : 
                                                                                ((0xf9U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000038U))
                                                                                 : 
                                                                                ((0xfaU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000039U))
                                                                                 : 
                                                                                ((0xfbU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000003aU))
                                                                                 : 
                                                                                ((0xfcU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000003bU))
                                                                                 : 
                                                                                ((0xfdU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000003cU))
                                                                                 : 
                                                                                ((0xfeU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1)
                                                                                 : 
                                                                                ((0xffU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? 
                                                                                ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1) 
                                                                                >> 1U)
                                                                                 : 
                                                                                ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1) 
                                                                                >> 2U))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))));
        __Vdeeptemp_hb1633d74__0 = ((0x15U == (0x000000ffU 
                                               & (IData)(vlSelfRef.t__DOT__crc)))
                                     ? (IData)((vlSelfRef.t__DOT__crc 
                                                >> 0x00000015U))
                                     : ((0x16U == (0x000000ffU 
                                                   & (IData)(vlSelfRef.t__DOT__crc)))
                                         ? (IData)(
                                                   (vlSelfRef.t__DOT__crc 
                                                    >> 0x00000016U))
                                         : ((0x17U 
                                             == (0x000000ffU 
                                                 & (IData)(vlSelfRef.t__DOT__crc)))
                                             ? (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 0x00000017U))
                                             : ((0x18U 
                                                 == 
                                                 (0x000000ffU 
                                                  & (IData)(vlSelfRef.t__DOT__crc)))
                                                 ? (IData)(
                                                           (vlSelfRef.t__DOT__crc 
                                                            >> 0x00000018U))
                                                 : 
                                                ((0x19U 
                                                  == 
                                                  (0x000000ffU 
                                                   & (IData)(vlSelfRef.t__DOT__crc)))
                                                  ? (IData)(
                                                            (vlSelfRef.t__DOT__crc 
                                                             >> 0x00000019U))
                                                  : 
                                                 ((0x1aU 
                                                   == 
                                                   (0x000000ffU 
                                                    & (IData)(vlSelfRef.t__DOT__crc)))
                                                   ? (IData)(
                                                             (vlSelfRef.t__DOT__crc 
                                                              >> 0x0000001aU))
                                                   : 
                                                  ((0x1bU 
                                                    == 
                                                    (0x000000ffU 
                                                     & (IData)(vlSelfRef.t__DOT__crc)))
                                                    ? (IData)(
                                                              (vlSelfRef.t__DOT__crc 
                                                               >> 0x0000001bU))
                                                    : 
                                                   ((0x1cU 
                                                     == 
                                                     (0x000000ffU 
                                                      & (IData)(vlSelfRef.t__DOT__crc)))
                                                     ? (IData)(
                                                               (vlSelfRef.t__DOT__crc 
                                                                >> 0x0000001cU))
                                                     : 
                                                    ((0x1dU 
                                                      == 
                                                      (0x000000ffU 
                                                       & (IData)(vlSelfRef.t__DOT__crc)))
                                                      ? (IData)(
                                                                (vlSelfRef.t__DOT__crc 
                                                                 >> 0x0000001dU))
                                                      : 
                                                     ((0x1eU 
                                                       == 
                                                       (0x000000ffU 
                                                        & (IData)(vlSelfRef.t__DOT__crc)))
                                                       ? (IData)(
                                                                 (vlSelfRef.t__DOT__crc 
                                                                  >> 0x0000001eU))
                                                       : 
                                                      ((0x1fU 
                                                        == 
                                                        (0x000000ffU 
                                                         & (IData)(vlSelfRef.t__DOT__crc)))
                                                        ? (IData)(
                                                                  (vlSelfRef.t__DOT__crc 
                                                                   >> 0x0000001fU))
                                                        : 
                                                       ((0x20U 
                                                         == 
                                                         (0x000000ffU 
                                                          & (IData)(vlSelfRef.t__DOT__crc)))
                                                         ? (IData)(
                                                                   (vlSelfRef.t__DOT__crc 
                                                                    >> 0x00000020U))
                                                         : 
                                                        ((0x21U 
                                                          == 
                                                          (0x000000ffU 
                                                           & (IData)(vlSelfRef.t__DOT__crc)))
                                                          ? (IData)(
                                                                    (vlSelfRef.t__DOT__crc 
                                                                     >> 0x00000021U))
                                                          : 
                                                         ((0x22U 
                                                           == 
                                                           (0x000000ffU 
                                                            & (IData)(vlSelfRef.t__DOT__crc)))
                                                           ? (IData)(
                                                                     (vlSelfRef.t__DOT__crc 
                                                                      >> 0x00000022U))
                                                           : 
                                                          ((0x23U 
                                                            == 
                                                            (0x000000ffU 
                                                             & (IData)(vlSelfRef.t__DOT__crc)))
                                                            ? (IData)(
                                                                      (vlSelfRef.t__DOT__crc 
                                                                       >> 0x00000023U))
                                                            : 
                                                           ((0x24U 
                                                             == 
                                                             (0x000000ffU 
                                                              & (IData)(vlSelfRef.t__DOT__crc)))
                                                             ? (IData)(
                                                                       (vlSelfRef.t__DOT__crc 
                                                                        >> 0x00000024U))
                                                             : 
                                                            ((0x25U 
                                                              == 
                                                              (0x000000ffU 
                                                               & (IData)(vlSelfRef.t__DOT__crc)))
                                                              ? (IData)(
                                                                        (vlSelfRef.t__DOT__crc 
                                                                         >> 0x00000025U))
                                                              : 
                                                             ((0x26U 
                                                               == 
                                                               (0x000000ffU 
                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                               ? (IData)(
                                                                         (vlSelfRef.t__DOT__crc 
                                                                          >> 0x00000026U))
                                                               : 
                                                              ((0x27U 
                                                                == 
                                                                (0x000000ffU 
                                                                 & (IData)(vlSelfRef.t__DOT__crc)))
                                                                ? (IData)(
                                                                          (vlSelfRef.t__DOT__crc 
                                                                           >> 0x00000027U))
                                                                : 
                                                               ((0x28U 
                                                                 == 
                                                                 (0x000000ffU 
                                                                  & (IData)(vlSelfRef.t__DOT__crc)))
                                                                 ? (IData)(
                                                                           (vlSelfRef.t__DOT__crc 
                                                                            >> 0x00000028U))
                                                                 : 
                                                                ((0x29U 
                                                                  == 
                                                                  (0x000000ffU 
                                                                   & (IData)(vlSelfRef.t__DOT__crc)))
                                                                  ? (IData)(
                                                                            (vlSelfRef.t__DOT__crc 
                                                                             >> 0x00000029U))
                                                                  : 
                                                                 ((0x2aU 
                                                                   == 
                                                                   (0x000000ffU 
                                                                    & (IData)(vlSelfRef.t__DOT__crc)))
                                                                   ? (IData)(
                                                                             (vlSelfRef.t__DOT__crc 
                                                                              >> 0x0000002aU))
                                                                   : 
                                                                  ((0x2bU 
                                                                    == 
                                                                    (0x000000ffU 
                                                                     & (IData)(vlSelfRef.t__DOT__crc)))
                                                                    ? (IData)(
                                                                              (vlSelfRef.t__DOT__crc 
                                                                               >> 0x0000002bU))
                                                                    : 
                                                                   ((0x2cU 
                                                                     == 
                                                                     (0x000000ffU 
                                                                      & (IData)(vlSelfRef.t__DOT__crc)))
                                                                     ? (IData)(
                                                                               (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000002cU))
                                                                     : 
                                                                    ((0x2dU 
                                                                      == 
                                                                      (0x000000ffU 
                                                                       & (IData)(vlSelfRef.t__DOT__crc)))
                                                                      ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000002dU))
                                                                      : 
                                                                     ((0x2eU 
                                                                       == 
                                                                       (0x000000ffU 
                                                                        & (IData)(vlSelfRef.t__DOT__crc)))
                                                                       ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000002eU))
                                                                       : 
                                                                      ((0x2fU 
                                                                        == 
                                                                        (0x000000ffU 
                                                                         & (IData)(vlSelfRef.t__DOT__crc)))
                                                                        ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000002fU))
                                                                        : 
                                                                       ((0x30U 
                                                                         == 
                                                                         (0x000000ffU 
                                                                          & (IData)(vlSelfRef.t__DOT__crc)))
                                                                         ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000030U))
                                                                         : 
                                                                        ((0x31U 
                                                                          == 
                                                                          (0x000000ffU 
                                                                           & (IData)(vlSelfRef.t__DOT__crc)))
                                                                          ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000031U))
                                                                          : 
                                                                         ((0x32U 
                                                                           == 
                                                                           (0x000000ffU 
                                                                            & (IData)(vlSelfRef.t__DOT__crc)))
                                                                           ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000032U))
                                                                           : 
                                                                          ((0x33U 
                                                                            == 
                                                                            (0x000000ffU 
                                                                             & (IData)(vlSelfRef.t__DOT__crc)))
                                                                            ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000033U))
                                                                            : 
                                                                           ((0x34U 
                                                                             == 
                                                                             (0x000000ffU 
                                                                              & (IData)(vlSelfRef.t__DOT__crc)))
                                                                             ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000034U))
                                                                             : 
                                                                            ((0x35U 
                                                                              == 
                                                                              (0x000000ffU 
                                                                               & (IData)(vlSelfRef.t__DOT__crc)))
                                                                              ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000035U))
                                                                              : 
                                                                             ((0x36U 
                                                                               == 
                                                                               (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                               ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000036U))
                                                                               : 
                                                                              ((0x37U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000037U))
                                                                                : 
                                                                               ((0x38U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000038U))
                                                                                 : 
                                                                                ((0x39U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000039U))
                                                                                 : 
                                                                                ((0x3aU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000003aU))
                                                                                 : 
                                                                                ((0x3bU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000003bU))
                                                                                 : 
                                                                                ((0x3cU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000003cU))
                                                                                 : 
                                                                                ((0x3dU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? 
                                                                                ((vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[2U] 
                                                                                << 3U) 
                                                                                | (vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[1U] 
                                                                                >> 0x0000001dU))
                                                                                 : 
                                                                                ((0x3eU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? 
                                                                                ((vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[2U] 
                                                                                << 2U) 
                                                                                | (vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[1U] 
                                                                                >> 0x0000001eU))
                                                                                 : 
                                                                                ((0x3fU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? 
                                                                                ((vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[2U] 
                                                                                << 1U) 
                                                                                | (vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[1U] 
                                                                                >> 0x0000001fU))
                                                                                 : 
                                                                                ((0x40U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(vlSelfRef.t__DOT__crc)
                                                                                 : 
                                                                                ((0x41U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 1U))
                                                                                 : 
                                                                                ((0x42U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 2U))
                                                                                 : 
                                                                                ((0x43U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 3U))
                                                                                 : 
                                                                                ((0x44U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 4U))
                                                                                 : 
                                                                                ((0x45U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 5U))
                                                                                 : 
                                                                                ((0x46U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 6U))
                                                                                 : 
                                                                                ((0x47U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 7U))
                                                                                 : 
                                                                                ((0x48U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 8U))
                                                                                 : 
                                                                                ((0x49U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 9U))
                                                                                 : 
                                                                                ((0x4aU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000000aU))
                                                                                 : 
                                                                                ((0x4bU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000000bU))
                                                                                 : 
                                                                                ((0x4cU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000000cU))
                                                                                 : 
                                                                                ((0x4dU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000000dU))
                                                                                 : 
                                                                                ((0x4eU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000000eU))
                                                                                 : 
                                                                                ((0x4fU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000000fU))
                                                                                 : 
                                                                                ((0x50U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000010U))
                                                                                 : 
                                                                                ((0x51U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000011U))
                                                                                 : 
                                                                                ((0x52U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000012U))
                                                                                 : 
                                                                                ((0x53U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000013U))
                                                                                 : 
                                                                                ((0x54U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000014U))
                                                                                 : 
                                                                                ((0x55U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000015U))
                                                                                 : 
                                                                                ((0x56U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000016U))
                                                                                 : 
                                                                                ((0x57U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000017U))
                                                                                 : 
                                                                                ((0x58U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000018U))
                                                                                 : 
                                                                                ((0x59U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000019U))
                                                                                 : 
                                                                                ((0x5aU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000001aU))
                                                                                 : 
                                                                                ((0x5bU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000001bU))
                                                                                 : 
                                                                                ((0x5cU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000001cU))
                                                                                 : 
                                                                                ((0x5dU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000001dU))
                                                                                 : 
                                                                                ((0x5eU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000001eU))
                                                                                 : 
                                                                                ((0x5fU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000001fU))
                                                                                 : 
                                                                                ((0x60U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000020U))
                                                                                 : 
                                                                                ((0x61U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000021U))
                                                                                 : 
                                                                                ((0x62U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000022U))
                                                                                 : 
                                                                                ((0x63U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000023U))
                                                                                 : 
                                                                                ((0x64U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000024U))
                                                                                 : 
                                                                                ((0x65U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000025U))
                                                                                 : 
                                                                                ((0x66U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000026U))
                                                                                 : 
                                                                                ((0x67U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000027U))
                                                                                 : 
                                                                                ((0x68U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000028U))
                                                                                 : 
                                                                                ((0x69U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000029U))
                                                                                 : 
                                                                                ((0x6aU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000002aU))
                                                                                 : 
                                                                                ((0x6bU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000002bU))
                                                                                 : 
                                                                                ((0x6cU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000002cU))
                                                                                 : 
                                                                                ((0x6dU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000002dU))
                                                                                 : 
                                                                                ((0x6eU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000002eU))
                                                                                 : 
                                                                                ((0x6fU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000002fU))
                                                                                 : 
                                                                                ((0x70U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000030U))
                                                                                 : 
                                                                                ((0x71U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000031U))
                                                                                 : 
                                                                                ((0x72U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000032U))
                                                                                 : 
                                                                                ((0x73U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000033U))
                                                                                 : 
                                                                                ((0x74U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000034U))
                                                                                 : 
                                                                                ((0x75U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000035U))
                                                                                 : 
                                                                                ((0x76U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000036U))
                                                                                 : 
                                                                                ((0x77U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000037U))
                                                                                 : 
                                                                                ((0x78U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000038U))
                                                                                 : 
                                                                                ((0x79U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000039U))
                                                                                 : 
                                                                                ((0x7aU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000003aU))
                                                                                 : 
                                                                                ((0x7bU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000003bU))
                                                                                 : 
                                                                                ((0x7cU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000003cU))
                                                                                 : 
                                                                                ((0x7dU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? 
                                                                                ((vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[4U] 
                                                                                << 3U) 
                                                                                | (vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[3U] 
                                                                                >> 0x0000001dU))
                                                                                 : 
                                                                                ((0x7eU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? 
                                                                                ((vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[4U] 
                                                                                << 2U) 
                                                                                | (vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[3U] 
                                                                                >> 0x0000001eU))
                                                                                 : 
                                                                                ((0x7fU 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? 
                                                                                ((vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[4U] 
                                                                                << 1U) 
                                                                                | (vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[3U] 
                                                                                >> 0x0000001fU))
                                                                                 : 
                                                                                ((0x80U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(vlSelfRef.t__DOT__crc)
                                                                                 : 
                                                                                ((0x81U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 1U))
                                                                                 : 
                                                                                ((0x82U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 2U))
                                                                                 : 
                                                                                ((0x83U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 3U))
                                                                                 : 
                                                                                ((0x84U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 4U))
                                                                                 : 
                                                                                ((0x85U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 5U))
                                                                                 : 
                                                                                ((0x86U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 6U))
                                                                                 : 
                                                                                ((0x87U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 7U))
                                                                                 : 
                                                                                ((0x88U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 8U))
                                                                                 : 
                                                                                ((0x89U 
                                                                                == 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 9U))
                                                                                 : __Vdeeptemp_hb3d546fb__0)))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))));
        __Vdeeptemp_h914a33a1__0 = ((0x12U == (0x000000ffU 
                                               & (IData)(vlSelfRef.t__DOT__crc)))
                                     ? (IData)((vlSelfRef.t__DOT__crc 
                                                >> 0x00000012U))
                                     : ((0x13U == (0x000000ffU 
                                                   & (IData)(vlSelfRef.t__DOT__crc)))
                                         ? (IData)(
                                                   (vlSelfRef.t__DOT__crc 
                                                    >> 0x00000013U))
                                         : ((0x14U 
                                             == (0x000000ffU 
                                                 & (IData)(vlSelfRef.t__DOT__crc)))
                                             ? (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 0x00000014U))
                                             : __Vdeeptemp_hb1633d74__0)));
        __Vdly__t__DOT__sum = ((QData)((IData)((0x0000000fU 
                                                & ((0U 
                                                    == 
                                                    (0x000000ffU 
                                                     & (IData)(vlSelfRef.t__DOT__crc)))
                                                    ? (IData)(vlSelfRef.t__DOT__crc)
                                                    : 
                                                   ((1U 
                                                     == 
                                                     (0x000000ffU 
                                                      & (IData)(vlSelfRef.t__DOT__crc)))
                                                     ? (IData)(
                                                               (vlSelfRef.t__DOT__crc 
                                                                >> 1U))
                                                     : 
                                                    ((2U 
                                                      == 
                                                      (0x000000ffU 
                                                       & (IData)(vlSelfRef.t__DOT__crc)))
                                                      ? (IData)(
                                                                (vlSelfRef.t__DOT__crc 
                                                                 >> 2U))
                                                      : 
                                                     ((3U 
                                                       == 
                                                       (0x000000ffU 
                                                        & (IData)(vlSelfRef.t__DOT__crc)))
                                                       ? (IData)(
                                                                 (vlSelfRef.t__DOT__crc 
                                                                  >> 3U))
                                                       : 
                                                      ((4U 
                                                        == 
                                                        (0x000000ffU 
                                                         & (IData)(vlSelfRef.t__DOT__crc)))
                                                        ? (IData)(
                                                                  (vlSelfRef.t__DOT__crc 
                                                                   >> 4U))
                                                        : 
                                                       ((5U 
                                                         == 
                                                         (0x000000ffU 
                                                          & (IData)(vlSelfRef.t__DOT__crc)))
                                                         ? (IData)(
                                                                   (vlSelfRef.t__DOT__crc 
                                                                    >> 5U))
                                                         : 
                                                        ((6U 
                                                          == 
                                                          (0x000000ffU 
                                                           & (IData)(vlSelfRef.t__DOT__crc)))
                                                          ? (IData)(
                                                                    (vlSelfRef.t__DOT__crc 
                                                                     >> 6U))
                                                          : 
                                                         ((7U 
                                                           == 
                                                           (0x000000ffU 
                                                            & (IData)(vlSelfRef.t__DOT__crc)))
                                                           ? (IData)(
                                                                     (vlSelfRef.t__DOT__crc 
                                                                      >> 7U))
                                                           : 
                                                          ((8U 
                                                            == 
                                                            (0x000000ffU 
                                                             & (IData)(vlSelfRef.t__DOT__crc)))
                                                            ? (IData)(
                                                                      (vlSelfRef.t__DOT__crc 
                                                                       >> 8U))
                                                            : 
                                                           ((9U 
                                                             == 
                                                             (0x000000ffU 
                                                              & (IData)(vlSelfRef.t__DOT__crc)))
                                                             ? (IData)(
                                                                       (vlSelfRef.t__DOT__crc 
                                                                        >> 9U))
                                                             : 
                                                            ((0x0aU 
                                                              == 
                                                              (0x000000ffU 
                                                               & (IData)(vlSelfRef.t__DOT__crc)))
                                                              ? (IData)(
                                                                        (vlSelfRef.t__DOT__crc 
                                                                         >> 0x0000000aU))
                                                              : 
                                                             ((0x0bU 
                                                               == 
                                                               (0x000000ffU 
                                                                & (IData)(vlSelfRef.t__DOT__crc)))
                                                               ? (IData)(
                                                                         (vlSelfRef.t__DOT__crc 
                                                                          >> 0x0000000bU))
                                                               : 
                                                              ((0x0cU 
                                                                == 
                                                                (0x000000ffU 
                                                                 & (IData)(vlSelfRef.t__DOT__crc)))
                                                                ? (IData)(
                                                                          (vlSelfRef.t__DOT__crc 
                                                                           >> 0x0000000cU))
                                                                : 
                                                               ((0x0dU 
                                                                 == 
                                                                 (0x000000ffU 
                                                                  & (IData)(vlSelfRef.t__DOT__crc)))
                                                                 ? (IData)(
                                                                           (vlSelfRef.t__DOT__crc 
                                                                            >> 0x0000000dU))
                                                                 : 
                                                                ((0x0eU 
                                                                  == 
                                                                  (0x000000ffU 
                                                                   & (IData)(vlSelfRef.t__DOT__crc)))
                                                                  ? (IData)(
                                                                            (vlSelfRef.t__DOT__crc 
                                                                             >> 0x0000000eU))
                                                                  : 
                                                                 ((0x0fU 
                                                                   == 
                                                                   (0x000000ffU 
                                                                    & (IData)(vlSelfRef.t__DOT__crc)))
                                                                   ? (IData)(
                                                                             (vlSelfRef.t__DOT__crc 
                                                                              >> 0x0000000fU))
                                                                   : 
                                                                  ((0x10U 
                                                                    == 
                                                                    (0x000000ffU 
                                                                     & (IData)(vlSelfRef.t__DOT__crc)))
                                                                    ? (IData)(
                                                                              (vlSelfRef.t__DOT__crc 
                                                                               >> 0x00000010U))
                                                                    : 
                                                                   ((0x11U 
                                                                     == 
                                                                     (0x000000ffU 
                                                                      & (IData)(vlSelfRef.t__DOT__crc)))
                                                                     ? (IData)(
                                                                               (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000011U))
                                                                     : __Vdeeptemp_h914a33a1__0))))))))))))))))))))) 
                               ^ ((vlSelfRef.t__DOT__sum 
                                   << 1U) | (QData)((IData)(
                                                            (1U 
                                                             & VL_REDXOR_64(
                                                                            (0x8000000000000005ULL 
                                                                             & vlSelfRef.t__DOT__sum)))))));
        if ((0U == vlSelfRef.t__DOT__cyc)) {
            __Vdly__t__DOT__crc = 0x5aef0c8dd70a4497ULL;
        } else if (VL_GTS_III(32, 0x0000000aU, vlSelfRef.t__DOT__cyc)) {
            __Vdly__t__DOT__sum = 0ULL;
        } else if (VL_LTES_III(32, 0x0000005aU, vlSelfRef.t__DOT__cyc)) {
            if (VL_UNLIKELY(((0x00000063U == vlSelfRef.t__DOT__cyc)))) {
                vlCtx.writef(0U, 4U, vlcu_msg_arg(VL_TIME_UNITED_Q(1)), vlcu_msg_arg(vlSelfRef.t__DOT__cyc), vlcu_msg_arg(vlSelfRef.t__DOT__crc), vlcu_msg_arg(vlSelfRef.t__DOT__sum));
                if (VL_UNLIKELY(((0xc77bb9b3784ea091ULL 
                                  != vlSelfRef.t__DOT__crc)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_cond_huge.v", 317, "");
                }
                if (VL_UNLIKELY(((0x36f3051d15caf07aULL 
                                  != vlSelfRef.t__DOT__sum)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_cond_huge.v", 318, "");
                }
                vlCtx.writef(1U, 0U);
                VL_FINISH_MT("/root/reference/test_regress/t/t_math_cond_huge.v", 320, "");
            }
        }
        vlSelfRef.t__DOT__cyc = __Vdly__t__DOT__cyc;
        vlSelfRef.t__DOT__sum = __Vdly__t__DOT__sum;
        vlSelfRef.t__DOT__crc = __Vdly__t__DOT__crc;
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[0U] 
            = (IData)(vlSelfRef.t__DOT__crc);
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[1U] 
            = (IData)((vlSelfRef.t__DOT__crc >> 0x00000020U));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[2U] 
            = (IData)(vlSelfRef.t__DOT__crc);
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[3U] 
            = (IData)((vlSelfRef.t__DOT__crc >> 0x00000020U));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[4U] 
            = (IData)(vlSelfRef.t__DOT__crc);
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[5U] 
            = (IData)((vlSelfRef.t__DOT__crc >> 0x00000020U));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[6U] 
            = (IData)(vlSelfRef.t__DOT__crc);
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0[7U] 
            = (IData)((vlSelfRef.t__DOT__crc >> 0x00000020U));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1 = 
            ((0x00000038U & ((IData)(vlSelfRef.t__DOT__crc) 
                             << 3U)) | (7U & (IData)(
                                                     (vlSelfRef.t__DOT__crc 
                                                      >> 0x0000003dU))));
    }

    VLCU_D static void Vt_math_cond_huge___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt_math_cond_huge___024root___eval_phase__act(Vt_math_cond_huge___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt_math_cond_huge___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vt_math_cond_huge___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vt_math_cond_huge___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vt_math_cond_huge___024root___eval_phase__nba(Vt_math_cond_huge___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vt_math_cond_huge___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    Vt_math_cond_huge___024root___nba_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
            Vt_math_cond_huge___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vt_math_cond_huge___024root___eval(Vt_math_cond_huge___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt_math_cond_huge___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("/root/reference/test_regress/t/t_math_cond_huge.v", 7, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vt_math_cond_huge___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("/root/reference/test_regress/t/t_math_cond_huge.v", 7, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vt_math_cond_huge___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vt_math_cond_huge___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vt_math_cond_huge___024root___ctor_var_reset(Vt_math_cond_huge___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        vlSelf->t__DOT__crc = VL_SCOPED_RAND_RESET_Q(64, __VscopeHash, 17680824324711171631ull);
        vlSelf->t__DOT__sum = VL_SCOPED_RAND_RESET_Q(64, __VscopeHash, 1514072735762163239ull);
        VL_ZERO_RESET_W(256, vlSelf->__VdfgRegularize_h6e95ff9d_0_0);
        vlSelf->__VdfgRegularize_h6e95ff9d_0_1 = 0;
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
        Vt_math_cond_huge___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_cond_huge___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_cond_huge___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_cond_huge___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_cond_huge___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_cond_huge___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt_math_cond_huge__cuda::info() {
    typedef Vt_math_cond_huge__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"__VdfgRegularize_h6e95ff9d_0_1", "__VdfgRegularize_h6e95ff9d_0_1", VLCU_MEMBER_STATE, 1, 0, 1, 6, 5, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_1},
        {"__VstlFirstIteration", "__VstlFirstIteration", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlFirstIteration},
        {"__VstlPhaseResult", "__VstlPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"t__DOT__cyc", "t.cyc", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__cyc},
        {"__VdfgRegularize_h6e95ff9d_0_0", "__VdfgRegularize_h6e95ff9d_0_0", VLCU_MEMBER_STATE, 4, 0, 8, 256, 255, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_0},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"t__DOT__crc", "t.crc", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF_t__DOT__crc},
        {"t__DOT__sum", "t.sum", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF_t__DOT__sum},
        {"__VstlTriggered", "__VstlTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VstlTriggered},
        {"__VactTriggered", "__VactTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VactTriggered},
        {"__VnbaTriggered", "__VnbaTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VnbaTriggered},
    };
    static const char* const messages[] = {"[%0t] cyc==%0d crc=%x sum=%x\n", "*-* All Finished *-*\n", nullptr};
    static const char* const plusargs[] = {nullptr};
    static const VlcuTraceDecl traceDecls[] = {
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0}};  // (terminator, not counted)
    static const uint32_t traceOrder[] = {0};  // codes in dump order (+ terminator)
    static const uint32_t saveOrder[] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "t", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 2, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 15, 0, 0xd0a17948079424eULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt_math_cond_huge__cuda VLCU_TESTBENCHES);
