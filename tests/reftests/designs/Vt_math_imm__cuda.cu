// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt_math_imm"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt_math_imm___024root final {
    CData/*0:0*/ clk;
    CData/*7:0*/ t__DOT__crc;
    CData/*5:0*/ t__DOT__HighMaskSel_Top;
    CData/*0:0*/ __VstlFirstIteration;
    CData/*0:0*/ __VstlPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    IData/*31:0*/ t__DOT__cyc;
    IData/*31:0*/ __VactIterCount;
    QData/*63:0*/ t__DOT__sum;
    VlUnpacked<QData/*63:0*/, 1> __VstlTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vt_math_imm__cuda final {
    typedef Vt_math_imm___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 40,
        OFF_t__DOT__crc = 41,
        OFF_t__DOT__HighMaskSel_Top = 42,
        OFF___VstlFirstIteration = 43,
        OFF___VstlPhaseResult = 44,
        OFF___Vtrigprevexpr___TOP__clk__0 = 45,
        OFF___VactPhaseResult = 46,
        OFF___VnbaPhaseResult = 47,
        OFF_t__DOT__cyc = 32,
        OFF___VactIterCount = 36,
        OFF_t__DOT__sum = 0,
        OFF___VstlTriggered = 8,
        OFF___VactTriggered = 16,
        OFF___VnbaTriggered = 24,
        kStateBytes = 48,  // S_persist
        kTouchedReadBytes = 38,  // S_state_rd: persistent bytes one eval can read
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
        s.t__DOT__crc = st.ld<VlcuMem8>(OFF_t__DOT__crc, i);
        s.t__DOT__HighMaskSel_Top = st.ld<VlcuMem8>(OFF_t__DOT__HighMaskSel_Top, i);
        s.__VstlFirstIteration = st.ld<VlcuMem8>(OFF___VstlFirstIteration, i);
        s.__VstlPhaseResult = st.ld<VlcuMem8>(OFF___VstlPhaseResult, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        s.t__DOT__cyc = st.ld<IData>(OFF_t__DOT__cyc, i);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        s.t__DOT__sum = st.ld<QData>(OFF_t__DOT__sum, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VstlTriggered[e] = st.ldE<QData>(OFF___VstlTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.st<VlcuMem8>(OFF_t__DOT__crc, i, s.t__DOT__crc);
        st.st<VlcuMem8>(OFF_t__DOT__HighMaskSel_Top, i, s.t__DOT__HighMaskSel_Top);
        st.st<VlcuMem8>(OFF___VstlFirstIteration, i, s.__VstlFirstIteration);
        st.st<VlcuMem8>(OFF___VstlPhaseResult, i, s.__VstlPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.st<IData>(OFF_t__DOT__cyc, i, s.t__DOT__cyc);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
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

    VLCU_D static void Vt_math_imm___024root___eval_static(Vt_math_imm___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
    }

    VLCU_D static void Vt_math_imm___024root___eval_initial(Vt_math_imm___024root* vlSelf, VlcuContext& vlCtx) {
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

    VLCU_D static void Vt_math_imm___024root___eval_final(Vt_math_imm___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_math_imm___024root___eval_settle(Vt_math_imm___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt_math_imm___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
#endif
                VL_FATAL_MT("/root/reference/test_regress/t/t_math_imm.v", 14, "", "DIDNOTCONVERGE: Settle region did not converge after '--converge-limit' of 10000 tries");
            }
            __VstlIterCount = ((IData)(1U) + __VstlIterCount);
            vlSelfRef.__VstlPhaseResult = Vt_math_imm___024root___eval_phase__stl(vlSelf, vlCtx);
            vlSelfRef.__VstlFirstIteration = 0U;
        } while (vlSelfRef.__VstlPhaseResult);
    }

    VLCU_D static bool Vt_math_imm___024root___trigger_anySet__stl(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt_math_imm___024root___eval_phase__stl(Vt_math_imm___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt_math_imm___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
        }
#endif
        __VstlExecute = Vt_math_imm___024root___trigger_anySet__stl(vlSelfRef.__VstlTriggered);
        if (__VstlExecute) {
            {
                // Inlined CFunc: _eval_stl
                if ((1ULL & vlSelfRef.__VstlTriggered[0U])) {
                    {
                        // Inlined CFunc: _stl_sequent__TOP__0
                        vlSelfRef.t__DOT__HighMaskSel_Top 
                            = (0x0000003fU & ((IData)(vlSelfRef.t__DOT__crc) 
                                              + (3U 
                                                 & ((IData)(vlSelfRef.t__DOT__crc) 
                                                    >> 6U))));
                    }
                }
            }
        }
        return (__VstlExecute);
    }

    VLCU_D static bool Vt_math_imm___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt_math_imm___024root___nba_sequent__TOP__0(Vt_math_imm___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        IData/*31:0*/ __Vdly__t__DOT__cyc;
        __Vdly__t__DOT__cyc = 0;
        CData/*7:0*/ __Vdly__t__DOT__crc;
        __Vdly__t__DOT__crc = 0;
        QData/*63:0*/ __Vdly__t__DOT__sum;
        __Vdly__t__DOT__sum = 0;
        // Body
        __Vdly__t__DOT__cyc = vlSelfRef.t__DOT__cyc;
        __Vdly__t__DOT__sum = vlSelfRef.t__DOT__sum;
        __Vdly__t__DOT__crc = vlSelfRef.t__DOT__crc;
        __Vdly__t__DOT__cyc = ((IData)(1U) + vlSelfRef.t__DOT__cyc);
        __Vdly__t__DOT__crc = ((0x000000feU & ((IData)(vlSelfRef.t__DOT__crc) 
                                               << 1U)) 
                               | (1U & (~ (1U & VL_REDXOR_8(
                                                            (0xb8U 
                                                             & (IData)(vlSelfRef.t__DOT__crc)))))));
        if ((0U == vlSelfRef.t__DOT__cyc)) {
            __Vdly__t__DOT__crc = 0U;
            __Vdly__t__DOT__sum = 0ULL;
        } else if ((1U == vlSelfRef.t__DOT__cyc)) {
            __Vdly__t__DOT__crc = 0xedU;
            __Vdly__t__DOT__sum = 0ULL;
        } else if (VL_GTS_III(32, 0x0000005aU, vlSelfRef.t__DOT__cyc)) {
            __Vdly__t__DOT__sum = (((vlSelfRef.t__DOT__sum 
                                     << 1U) | (QData)((IData)(
                                                              (1U 
                                                               & (IData)(
                                                                         (vlSelfRef.t__DOT__sum 
                                                                          >> 0x3fU)))))) 
                                   ^ ((((QData)((IData)(
                                                        (0x80000000U 
                                                         | (((((((0x3eU 
                                                                  >= 
                                                                  (0x0000003fU 
                                                                   & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                 << 6U) 
                                                                | (((0x3dU 
                                                                     >= 
                                                                     (0x0000003fU 
                                                                      & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                    << 5U) 
                                                                   | ((0x3cU 
                                                                       >= 
                                                                       (0x0000003fU 
                                                                        & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                      << 4U))) 
                                                               | ((((0x3bU 
                                                                     >= 
                                                                     (0x0000003fU 
                                                                      & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                    << 3U) 
                                                                   | ((0x3aU 
                                                                       >= 
                                                                       (0x0000003fU 
                                                                        & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                      << 2U)) 
                                                                  | (((0x39U 
                                                                       >= 
                                                                       (0x0000003fU 
                                                                        & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                      << 1U) 
                                                                     | (0x38U 
                                                                        >= 
                                                                        (0x0000003fU 
                                                                         & (IData)(vlSelfRef.t__DOT__crc)))))) 
                                                              << 0x00000018U) 
                                                             | ((((((0x37U 
                                                                     >= 
                                                                     (0x0000003fU 
                                                                      & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                    << 3U) 
                                                                   | ((0x36U 
                                                                       >= 
                                                                       (0x0000003fU 
                                                                        & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                      << 2U)) 
                                                                  | (((0x35U 
                                                                       >= 
                                                                       (0x0000003fU 
                                                                        & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                      << 1U) 
                                                                     | (0x34U 
                                                                        >= 
                                                                        (0x0000003fU 
                                                                         & (IData)(vlSelfRef.t__DOT__crc))))) 
                                                                 << 0x00000014U) 
                                                                | (((((0x33U 
                                                                       >= 
                                                                       (0x0000003fU 
                                                                        & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                      << 3U) 
                                                                     | ((0x32U 
                                                                         >= 
                                                                         (0x0000003fU 
                                                                          & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                        << 2U)) 
                                                                    | (((0x31U 
                                                                         >= 
                                                                         (0x0000003fU 
                                                                          & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                        << 1U) 
                                                                       | (0x30U 
                                                                          >= 
                                                                          (0x0000003fU 
                                                                           & (IData)(vlSelfRef.t__DOT__crc))))) 
                                                                   << 0x00000010U))) 
                                                            | (((((((0x2fU 
                                                                     >= 
                                                                     (0x0000003fU 
                                                                      & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                    << 3U) 
                                                                   | ((0x2eU 
                                                                       >= 
                                                                       (0x0000003fU 
                                                                        & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                      << 2U)) 
                                                                  | (((0x2dU 
                                                                       >= 
                                                                       (0x0000003fU 
                                                                        & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                      << 1U) 
                                                                     | (0x2cU 
                                                                        >= 
                                                                        (0x0000003fU 
                                                                         & (IData)(vlSelfRef.t__DOT__crc))))) 
                                                                 << 0x0000000cU) 
                                                                | (((((0x2bU 
                                                                       >= 
                                                                       (0x0000003fU 
                                                                        & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                      << 3U) 
                                                                     | ((0x2aU 
                                                                         >= 
                                                                         (0x0000003fU 
                                                                          & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                        << 2U)) 
                                                                    | (((0x29U 
                                                                         >= 
                                                                         (0x0000003fU 
                                                                          & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                        << 1U) 
                                                                       | (0x28U 
                                                                          >= 
                                                                          (0x0000003fU 
                                                                           & (IData)(vlSelfRef.t__DOT__crc))))) 
                                                                   << 8U)) 
                                                               | ((((((0x27U 
                                                                       >= 
                                                                       (0x0000003fU 
                                                                        & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                      << 3U) 
                                                                     | ((0x26U 
                                                                         >= 
                                                                         (0x0000003fU 
                                                                          & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                        << 2U)) 
                                                                    | (((0x25U 
                                                                         >= 
                                                                         (0x0000003fU 
                                                                          & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                        << 1U) 
                                                                       | (0x24U 
                                                                          >= 
                                                                          (0x0000003fU 
                                                                           & (IData)(vlSelfRef.t__DOT__crc))))) 
                                                                   << 4U) 
                                                                  | ((((0x23U 
                                                                        >= 
                                                                        (0x0000003fU 
                                                                         & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                       << 3U) 
                                                                      | ((0x22U 
                                                                          >= 
                                                                          (0x0000003fU 
                                                                           & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                         << 2U)) 
                                                                     | (((0x21U 
                                                                          >= 
                                                                          (0x0000003fU 
                                                                           & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                         << 1U) 
                                                                        | (0x20U 
                                                                           >= 
                                                                           (0x0000003fU 
                                                                            & (IData)(vlSelfRef.t__DOT__crc))))))))))) 
                                        << 0x00000020U) 
                                       | (QData)((IData)(
                                                         (((((((((0x1fU 
                                                                  >= 
                                                                  (0x0000003fU 
                                                                   & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                 << 3U) 
                                                                | ((0x1eU 
                                                                    >= 
                                                                    (0x0000003fU 
                                                                     & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                   << 2U)) 
                                                               | (((0x1dU 
                                                                    >= 
                                                                    (0x0000003fU 
                                                                     & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                   << 1U) 
                                                                  | (0x1cU 
                                                                     >= 
                                                                     (0x0000003fU 
                                                                      & (IData)(vlSelfRef.t__DOT__crc))))) 
                                                              << 0x0000000cU) 
                                                             | (((((0x1bU 
                                                                    >= 
                                                                    (0x0000003fU 
                                                                     & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                   << 3U) 
                                                                  | ((0x1aU 
                                                                      >= 
                                                                      (0x0000003fU 
                                                                       & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                     << 2U)) 
                                                                 | (((0x19U 
                                                                      >= 
                                                                      (0x0000003fU 
                                                                       & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                     << 1U) 
                                                                    | (0x18U 
                                                                       >= 
                                                                       (0x0000003fU 
                                                                        & (IData)(vlSelfRef.t__DOT__crc))))) 
                                                                << 8U)) 
                                                            | ((((((0x17U 
                                                                    >= 
                                                                    (0x0000003fU 
                                                                     & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                   << 3U) 
                                                                  | ((0x16U 
                                                                      >= 
                                                                      (0x0000003fU 
                                                                       & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                     << 2U)) 
                                                                 | (((0x15U 
                                                                      >= 
                                                                      (0x0000003fU 
                                                                       & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                     << 1U) 
                                                                    | (0x14U 
                                                                       >= 
                                                                       (0x0000003fU 
                                                                        & (IData)(vlSelfRef.t__DOT__crc))))) 
                                                                << 4U) 
                                                               | ((((0x13U 
                                                                     >= 
                                                                     (0x0000003fU 
                                                                      & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                    << 3U) 
                                                                   | ((0x12U 
                                                                       >= 
                                                                       (0x0000003fU 
                                                                        & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                      << 2U)) 
                                                                  | (((0x11U 
                                                                       >= 
                                                                       (0x0000003fU 
                                                                        & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                      << 1U) 
                                                                     | (0x10U 
                                                                        >= 
                                                                        (0x0000003fU 
                                                                         & (IData)(vlSelfRef.t__DOT__crc))))))) 
                                                           << 0x00000010U) 
                                                          | (((((((0x0fU 
                                                                   >= 
                                                                   (0x0000003fU 
                                                                    & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                  << 3U) 
                                                                 | ((0x0eU 
                                                                     >= 
                                                                     (0x0000003fU 
                                                                      & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                    << 2U)) 
                                                                | (((0x0dU 
                                                                     >= 
                                                                     (0x0000003fU 
                                                                      & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                    << 1U) 
                                                                   | (0x0cU 
                                                                      >= 
                                                                      (0x0000003fU 
                                                                       & (IData)(vlSelfRef.t__DOT__crc))))) 
                                                               << 0x0000000cU) 
                                                              | (((((0x0bU 
                                                                     >= 
                                                                     (0x0000003fU 
                                                                      & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                    << 3U) 
                                                                   | ((0x0aU 
                                                                       >= 
                                                                       (0x0000003fU 
                                                                        & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                      << 2U)) 
                                                                  | (((9U 
                                                                       >= 
                                                                       (0x0000003fU 
                                                                        & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                      << 1U) 
                                                                     | (8U 
                                                                        >= 
                                                                        (0x0000003fU 
                                                                         & (IData)(vlSelfRef.t__DOT__crc))))) 
                                                                 << 8U)) 
                                                             | ((((((7U 
                                                                     >= 
                                                                     (0x0000003fU 
                                                                      & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                    << 3U) 
                                                                   | ((6U 
                                                                       >= 
                                                                       (0x0000003fU 
                                                                        & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                      << 2U)) 
                                                                  | (((5U 
                                                                       >= 
                                                                       (0x0000003fU 
                                                                        & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                      << 1U) 
                                                                     | (4U 
                                                                        >= 
                                                                        (0x0000003fU 
                                                                         & (IData)(vlSelfRef.t__DOT__crc))))) 
                                                                 << 4U) 
                                                                | ((((3U 
                                                                      >= 
                                                                      (0x0000003fU 
                                                                       & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                     << 3U) 
                                                                    | ((2U 
                                                                        >= 
                                                                        (0x0000003fU 
                                                                         & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                       << 2U)) 
                                                                   | (((1U 
                                                                        >= 
                                                                        (0x0000003fU 
                                                                         & (IData)(vlSelfRef.t__DOT__crc))) 
                                                                       << 1U) 
                                                                      | (0U 
                                                                         >= 
                                                                         (0x0000003fU 
                                                                          & (IData)(vlSelfRef.t__DOT__crc))))))))))) 
                                      & (((QData)((IData)(
                                                          (((((((((0x3fU 
                                                                   <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                  << 3U) 
                                                                 | ((0x3eU 
                                                                     <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                    << 2U)) 
                                                                | (((0x3dU 
                                                                     <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                    << 1U) 
                                                                   | (0x3cU 
                                                                      <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)))) 
                                                               << 0x0000000cU) 
                                                              | (((((0x3bU 
                                                                     <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                    << 3U) 
                                                                   | ((0x3aU 
                                                                       <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                      << 2U)) 
                                                                  | (((0x39U 
                                                                       <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                      << 1U) 
                                                                     | (0x38U 
                                                                        <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)))) 
                                                                 << 8U)) 
                                                             | ((((((0x37U 
                                                                     <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                    << 3U) 
                                                                   | ((0x36U 
                                                                       <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                      << 2U)) 
                                                                  | (((0x35U 
                                                                       <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                      << 1U) 
                                                                     | (0x34U 
                                                                        <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)))) 
                                                                 << 4U) 
                                                                | ((((0x33U 
                                                                      <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                     << 3U) 
                                                                    | ((0x32U 
                                                                        <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                       << 2U)) 
                                                                   | (((0x31U 
                                                                        <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                       << 1U) 
                                                                      | (0x30U 
                                                                         <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)))))) 
                                                            << 0x00000010U) 
                                                           | (((((((0x2fU 
                                                                    <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                   << 3U) 
                                                                  | ((0x2eU 
                                                                      <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                     << 2U)) 
                                                                 | (((0x2dU 
                                                                      <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                     << 1U) 
                                                                    | (0x2cU 
                                                                       <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)))) 
                                                                << 0x0000000cU) 
                                                               | (((((0x2bU 
                                                                      <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                     << 3U) 
                                                                    | ((0x2aU 
                                                                        <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                       << 2U)) 
                                                                   | (((0x29U 
                                                                        <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                       << 1U) 
                                                                      | (0x28U 
                                                                         <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)))) 
                                                                  << 8U)) 
                                                              | ((((((0x27U 
                                                                      <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                     << 3U) 
                                                                    | ((0x26U 
                                                                        <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                       << 2U)) 
                                                                   | (((0x25U 
                                                                        <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                       << 1U) 
                                                                      | (0x24U 
                                                                         <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)))) 
                                                                  << 4U) 
                                                                 | ((((0x23U 
                                                                       <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                      << 3U) 
                                                                     | ((0x22U 
                                                                         <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                        << 2U)) 
                                                                    | (((0x21U 
                                                                         <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                        << 1U) 
                                                                       | (0x20U 
                                                                          <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top))))))))) 
                                          << 0x00000020U) 
                                         | (QData)((IData)(
                                                           (1U 
                                                            | (((((((((0x1fU 
                                                                       <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                      << 3U) 
                                                                     | ((0x1eU 
                                                                         <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                        << 2U)) 
                                                                    | (((0x1dU 
                                                                         <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                        << 1U) 
                                                                       | (0x1cU 
                                                                          <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)))) 
                                                                   << 0x0000000cU) 
                                                                  | (((((0x1bU 
                                                                         <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                        << 3U) 
                                                                       | ((0x1aU 
                                                                           <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                          << 2U)) 
                                                                      | (((0x19U 
                                                                           <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                          << 1U) 
                                                                         | (0x18U 
                                                                            <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)))) 
                                                                     << 8U)) 
                                                                 | ((((((0x17U 
                                                                         <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                        << 3U) 
                                                                       | ((0x16U 
                                                                           <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                          << 2U)) 
                                                                      | (((0x15U 
                                                                           <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                          << 1U) 
                                                                         | (0x14U 
                                                                            <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)))) 
                                                                     << 4U) 
                                                                    | ((((0x13U 
                                                                          <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                         << 3U) 
                                                                        | ((0x12U 
                                                                            <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                           << 2U)) 
                                                                       | (((0x11U 
                                                                            <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                           << 1U) 
                                                                          | (0x10U 
                                                                             <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)))))) 
                                                                << 0x00000010U) 
                                                               | (((((((0x0fU 
                                                                        <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                       << 3U) 
                                                                      | ((0x0eU 
                                                                          <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                         << 2U)) 
                                                                     | (((0x0dU 
                                                                          <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                         << 1U) 
                                                                        | (0x0cU 
                                                                           <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)))) 
                                                                    << 0x0000000cU) 
                                                                   | (((((0x0bU 
                                                                          <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                         << 3U) 
                                                                        | ((0x0aU 
                                                                            <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                           << 2U)) 
                                                                       | (((9U 
                                                                            <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                           << 1U) 
                                                                          | (8U 
                                                                             <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)))) 
                                                                      << 8U)) 
                                                                  | ((((((7U 
                                                                          <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                         << 3U) 
                                                                        | ((6U 
                                                                            <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                           << 2U)) 
                                                                       | (((5U 
                                                                            <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                           << 1U) 
                                                                          | (4U 
                                                                             <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)))) 
                                                                      << 4U) 
                                                                     | ((((3U 
                                                                           <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                          << 3U) 
                                                                         | ((2U 
                                                                             <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                            << 2U)) 
                                                                        | ((1U 
                                                                            <= (IData)(vlSelfRef.t__DOT__HighMaskSel_Top)) 
                                                                           << 1U)))))))))));
        } else if (VL_UNLIKELY(((0x00000063U == vlSelfRef.t__DOT__cyc)))) {
            vlCtx.writef(0U, 4U, vlcu_msg_arg(VL_TIME_UNITED_Q(1)), vlcu_msg_arg(vlSelfRef.t__DOT__cyc), vlcu_msg_arg((IData)(vlSelfRef.t__DOT__crc)), vlcu_msg_arg(vlSelfRef.t__DOT__sum));
            if (VL_UNLIKELY(((0x38U != (IData)(vlSelfRef.t__DOT__crc))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_imm.v", 69, "");
            }
            if (VL_UNLIKELY(((0x58743ffa61e41075ULL 
                              != vlSelfRef.t__DOT__sum)))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_imm.v", 70, "");
            }
            vlCtx.writef(1U, 0U);
            VL_FINISH_MT("/root/reference/test_regress/t/t_math_imm.v", 72, "");
        }
        vlSelfRef.t__DOT__cyc = __Vdly__t__DOT__cyc;
        vlSelfRef.t__DOT__sum = __Vdly__t__DOT__sum;
        vlSelfRef.t__DOT__crc = __Vdly__t__DOT__crc;
        vlSelfRef.t__DOT__HighMaskSel_Top = (0x0000003fU 
                                             & ((IData)(vlSelfRef.t__DOT__crc) 
                                                + (3U 
                                                   & ((IData)(vlSelfRef.t__DOT__crc) 
                                                      >> 6U))));
    }

    VLCU_D static void Vt_math_imm___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt_math_imm___024root___eval_phase__act(Vt_math_imm___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt_math_imm___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vt_math_imm___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vt_math_imm___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vt_math_imm___024root___eval_phase__nba(Vt_math_imm___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vt_math_imm___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    Vt_math_imm___024root___nba_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
            Vt_math_imm___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vt_math_imm___024root___eval(Vt_math_imm___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt_math_imm___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("/root/reference/test_regress/t/t_math_imm.v", 14, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vt_math_imm___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("/root/reference/test_regress/t/t_math_imm.v", 14, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vt_math_imm___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vt_math_imm___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vt_math_imm___024root___ctor_var_reset(Vt_math_imm___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        vlSelf->t__DOT__cyc = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 15508623621340971111ull);
        vlSelf->t__DOT__crc = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 17680824324711171631ull);
        vlSelf->t__DOT__sum = VL_SCOPED_RAND_RESET_Q(64, __VscopeHash, 1514072735762163239ull);
        vlSelf->t__DOT__HighMaskSel_Top = VL_SCOPED_RAND_RESET_I(6, __VscopeHash, 14874357441806581606ull);
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
        Vt_math_imm___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_imm___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_imm___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_imm___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_imm___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_imm___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt_math_imm__cuda::info() {
    typedef Vt_math_imm__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"t__DOT__crc", "t.crc", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF_t__DOT__crc},
        {"t__DOT__HighMaskSel_Top", "t.HighMaskSel_Top", VLCU_MEMBER_STATE, 1, 0, 1, 6, 5, 0, T::OFF_t__DOT__HighMaskSel_Top},
        {"__VstlFirstIteration", "__VstlFirstIteration", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlFirstIteration},
        {"__VstlPhaseResult", "__VstlPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"t__DOT__cyc", "t.cyc", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__cyc},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"t__DOT__sum", "t.sum", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF_t__DOT__sum},
        {"__VstlTriggered", "__VstlTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VstlTriggered},
        {"__VactTriggered", "__VactTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VactTriggered},
        {"__VnbaTriggered", "__VnbaTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VnbaTriggered},
    };
    static const char* const messages[] = {"[%0t] cyc==%0d crc=%b %x\n", "*-* All Finished *-*\n", nullptr};
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
        saveOrder, 14, 0, 0xae60ec923deb3032ULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt_math_imm__cuda VLCU_TESTBENCHES);
