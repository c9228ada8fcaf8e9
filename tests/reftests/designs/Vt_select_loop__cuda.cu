// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt_select_loop"
#endif

// Constant pool
static __device__ const VlWide<8>/*255:0*/ Vt_select_loop__ConstPool__CONST_h6bf974be_0 = VlWide<8>{{
        0xc9f395b8, 0x0a48e039, 0xdddbdad2, 0xe236ddfd,
        0xe2bd68c3, 0x982646c8, 0x3d5fe51a, 0x0e17c88f
}};
static __device__ const VlWide<8>/*255:0*/ Vt_select_loop__ConstPool__CONST_hc278fede_0 = VlWide<8>{{
        0x147e7f26, 0x7491f3da, 0x42b14853, 0x45e0eb76,
        0x2413d770, 0xde489bab, 0x6c843d35, 0xed388e64
}};

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt_select_loop___024root final {
    CData/*0:0*/ clk;
    CData/*0:0*/ __VstlFirstIteration;
    CData/*0:0*/ __VstlPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    IData/*31:0*/ t__DOT__cyc;
    VlWide<8>/*255:0*/ t__DOT__a;
    VlWide<8>/*255:0*/ t__DOT__q;
    IData/*31:0*/ __VactIterCount;
    QData/*63:0*/ t__DOT__qq;
    VlUnpacked<QData/*63:0*/, 1> __VstlTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vt_select_loop__cuda final {
    typedef Vt_select_loop___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 104,
        OFF___VstlFirstIteration = 105,
        OFF___VstlPhaseResult = 106,
        OFF___Vtrigprevexpr___TOP__clk__0 = 107,
        OFF___VactPhaseResult = 108,
        OFF___VnbaPhaseResult = 109,
        OFF_t__DOT__cyc = 32,
        OFF_t__DOT__a = 36,
        OFF_t__DOT__q = 68,
        OFF___VactIterCount = 100,
        OFF_t__DOT__qq = 0,
        OFF___VstlTriggered = 8,
        OFF___VactTriggered = 16,
        OFF___VnbaTriggered = 24,
        kStateBytes = 112,  // S_persist
        kTouchedReadBytes = 100,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 99,  // S_state_wr
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
        s.__VstlFirstIteration = st.ld<VlcuMem8>(OFF___VstlFirstIteration, i);
        s.__VstlPhaseResult = st.ld<VlcuMem8>(OFF___VstlPhaseResult, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        s.t__DOT__cyc = st.ld<IData>(OFF_t__DOT__cyc, i);
        st.ldW(OFF_t__DOT__a, i, s.t__DOT__a);
        st.ldW(OFF_t__DOT__q, i, s.t__DOT__q);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        s.t__DOT__qq = st.ld<QData>(OFF_t__DOT__qq, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VstlTriggered[e] = st.ldE<QData>(OFF___VstlTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.st<VlcuMem8>(OFF___VstlFirstIteration, i, s.__VstlFirstIteration);
        st.st<VlcuMem8>(OFF___VstlPhaseResult, i, s.__VstlPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.st<IData>(OFF_t__DOT__cyc, i, s.t__DOT__cyc);
        st.stW(OFF_t__DOT__a, i, s.t__DOT__a);
        st.stW(OFF_t__DOT__q, i, s.t__DOT__q);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
        st.st<QData>(OFF_t__DOT__qq, i, s.t__DOT__qq);
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

    VLCU_D static void Vt_select_loop___024root___eval_static(Vt_select_loop___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
    }

    VLCU_D static void Vt_select_loop___024root___eval_initial(Vt_select_loop___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        {
            // Inlined CFunc: _eval_initial__TOP
            vlSelfRef.t__DOT__cyc = 1U;
        }
    }

    VLCU_D static void Vt_select_loop___024root___eval_final(Vt_select_loop___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_select_loop___024root___eval_settle(Vt_select_loop___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt_select_loop___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
#endif
                VL_FATAL_MT("/root/reference/test_regress/t/t_select_loop.v", 7, "", "DIDNOTCONVERGE: Settle region did not converge after '--converge-limit' of 10000 tries");
            }
            __VstlIterCount = ((IData)(1U) + __VstlIterCount);
            vlSelfRef.__VstlPhaseResult = Vt_select_loop___024root___eval_phase__stl(vlSelf, vlCtx);
            vlSelfRef.__VstlFirstIteration = 0U;
        } while (vlSelfRef.__VstlPhaseResult);
    }

    VLCU_D static bool Vt_select_loop___024root___trigger_anySet__stl(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt_select_loop___024root___stl_sequent__TOP__0(Vt_select_loop___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        IData/*31:0*/ t__DOT__i;
        t__DOT__i = 0;
        // Body
        vlSelfRef.t__DOT__qq = (((QData)((IData)(((
                                                   ((((((2U 
                                                         & (vlSelfRef.t__DOT__a[0U] 
                                                            << 1U)) 
                                                        | (1U 
                                                           & (vlSelfRef.t__DOT__a[0U] 
                                                              >> 1U))) 
                                                       << 6U) 
                                                      | (((2U 
                                                           & (vlSelfRef.t__DOT__a[0U] 
                                                              >> 1U)) 
                                                          | (1U 
                                                             & (vlSelfRef.t__DOT__a[0U] 
                                                                >> 3U))) 
                                                         << 4U)) 
                                                     | ((((2U 
                                                           & (vlSelfRef.t__DOT__a[0U] 
                                                              >> 3U)) 
                                                          | (1U 
                                                             & (vlSelfRef.t__DOT__a[0U] 
                                                                >> 5U))) 
                                                         << 2U) 
                                                        | ((2U 
                                                            & (vlSelfRef.t__DOT__a[0U] 
                                                               >> 5U)) 
                                                           | (1U 
                                                              & (vlSelfRef.t__DOT__a[0U] 
                                                                 >> 7U))))) 
                                                    << 0x00000018U) 
                                                   | ((((((2U 
                                                           & (vlSelfRef.t__DOT__a[0U] 
                                                              >> 7U)) 
                                                          | (1U 
                                                             & (vlSelfRef.t__DOT__a[0U] 
                                                                >> 9U))) 
                                                         << 6U) 
                                                        | (((2U 
                                                             & (vlSelfRef.t__DOT__a[0U] 
                                                                >> 9U)) 
                                                            | (1U 
                                                               & (vlSelfRef.t__DOT__a[0U] 
                                                                  >> 0x0000000bU))) 
                                                           << 4U)) 
                                                       | ((((2U 
                                                             & (vlSelfRef.t__DOT__a[0U] 
                                                                >> 0x0000000bU)) 
                                                            | (1U 
                                                               & (vlSelfRef.t__DOT__a[0U] 
                                                                  >> 0x0000000dU))) 
                                                           << 2U) 
                                                          | ((2U 
                                                              & (vlSelfRef.t__DOT__a[0U] 
                                                                 >> 0x0000000dU)) 
                                                             | (1U 
                                                                & (vlSelfRef.t__DOT__a[0U] 
                                                                   >> 0x0000000fU))))) 
                                                      << 0x00000010U)) 
                                                  | (((((((2U 
                                                           & (vlSelfRef.t__DOT__a[0U] 
                                                              >> 0x0000000fU)) 
                                                          | (1U 
                                                             & (vlSelfRef.t__DOT__a[0U] 
                                                                >> 0x00000011U))) 
                                                         << 6U) 
                                                        | (((2U 
                                                             & (vlSelfRef.t__DOT__a[0U] 
                                                                >> 0x00000011U)) 
                                                            | (1U 
                                                               & (vlSelfRef.t__DOT__a[0U] 
                                                                  >> 0x00000013U))) 
                                                           << 4U)) 
                                                       | ((((2U 
                                                             & (vlSelfRef.t__DOT__a[0U] 
                                                                >> 0x00000013U)) 
                                                            | (1U 
                                                               & (vlSelfRef.t__DOT__a[0U] 
                                                                  >> 0x00000015U))) 
                                                           << 2U) 
                                                          | ((2U 
                                                              & (vlSelfRef.t__DOT__a[0U] 
                                                                 >> 0x00000015U)) 
                                                             | (1U 
                                                                & (vlSelfRef.t__DOT__a[0U] 
                                                                   >> 0x00000017U))))) 
                                                      << 8U) 
                                                     | (((((2U 
                                                            & (vlSelfRef.t__DOT__a[0U] 
                                                               >> 0x00000017U)) 
                                                           | (1U 
                                                              & (vlSelfRef.t__DOT__a[0U] 
                                                                 >> 0x00000019U))) 
                                                          << 6U) 
                                                         | (((2U 
                                                              & (vlSelfRef.t__DOT__a[0U] 
                                                                 >> 0x00000019U)) 
                                                             | (1U 
                                                                & (vlSelfRef.t__DOT__a[0U] 
                                                                   >> 0x0000001bU))) 
                                                            << 4U)) 
                                                        | ((((2U 
                                                              & (vlSelfRef.t__DOT__a[0U] 
                                                                 >> 0x0000001bU)) 
                                                             | (1U 
                                                                & (vlSelfRef.t__DOT__a[0U] 
                                                                   >> 0x0000001dU))) 
                                                            << 2U) 
                                                           | ((2U 
                                                               & (vlSelfRef.t__DOT__a[0U] 
                                                                  >> 0x0000001dU)) 
                                                              | (vlSelfRef.t__DOT__a[0U] 
                                                                 >> 0x0000001fU)))))))) 
                                 << 0x00000020U) | (QData)((IData)(
                                                                   (0x0fed0000U 
                                                                    | (((((2U 
                                                                           & (vlSelfRef.t__DOT__a[1U] 
                                                                              << 1U)) 
                                                                          | (1U 
                                                                             & (vlSelfRef.t__DOT__a[1U] 
                                                                                >> 1U))) 
                                                                         << 0x0000001eU) 
                                                                        | (((0x00002000U 
                                                                             & (vlSelfRef.t__DOT__a[1U] 
                                                                                << 0x0000000bU)) 
                                                                            | (0x00001000U 
                                                                               & (vlSelfRef.t__DOT__a[1U] 
                                                                                << 9U))) 
                                                                           << 0x00000010U)) 
                                                                       | (((((((2U 
                                                                                & (vlSelfRef.t__DOT__a[1U] 
                                                                                >> 0x0000000fU)) 
                                                                               | (1U 
                                                                                & (vlSelfRef.t__DOT__a[1U] 
                                                                                >> 0x00000011U))) 
                                                                              << 6U) 
                                                                             | (((2U 
                                                                                & (vlSelfRef.t__DOT__a[1U] 
                                                                                >> 0x00000011U)) 
                                                                                | (1U 
                                                                                & (vlSelfRef.t__DOT__a[1U] 
                                                                                >> 0x00000013U))) 
                                                                                << 4U)) 
                                                                            | ((((2U 
                                                                                & (vlSelfRef.t__DOT__a[1U] 
                                                                                >> 0x00000013U)) 
                                                                                | (1U 
                                                                                & (vlSelfRef.t__DOT__a[1U] 
                                                                                >> 0x00000015U))) 
                                                                                << 2U) 
                                                                               | ((2U 
                                                                                & (vlSelfRef.t__DOT__a[1U] 
                                                                                >> 0x00000015U)) 
                                                                                | (1U 
                                                                                & (vlSelfRef.t__DOT__a[1U] 
                                                                                >> 0x00000017U))))) 
                                                                           << 8U) 
                                                                          | (((((2U 
                                                                                & (vlSelfRef.t__DOT__a[1U] 
                                                                                >> 0x00000017U)) 
                                                                                | (1U 
                                                                                & (vlSelfRef.t__DOT__a[1U] 
                                                                                >> 0x00000019U))) 
                                                                               << 6U) 
                                                                              | (((2U 
                                                                                & (vlSelfRef.t__DOT__a[1U] 
                                                                                >> 0x00000019U)) 
                                                                                | (1U 
                                                                                & (vlSelfRef.t__DOT__a[1U] 
                                                                                >> 0x0000001bU))) 
                                                                                << 4U)) 
                                                                             | ((((2U 
                                                                                & (vlSelfRef.t__DOT__a[1U] 
                                                                                >> 0x0000001bU)) 
                                                                                | (1U 
                                                                                & (vlSelfRef.t__DOT__a[1U] 
                                                                                >> 0x0000001dU))) 
                                                                                << 2U) 
                                                                                | ((2U 
                                                                                & (vlSelfRef.t__DOT__a[1U] 
                                                                                >> 0x0000001dU)) 
                                                                                | (vlSelfRef.t__DOT__a[1U] 
                                                                                >> 0x0000001fU))))))))));
        t__DOT__i = 0U;
        while (VL_GTS_III(32, 0x00000100U, t__DOT__i)) {
            vlSelfRef.t__DOT__q[(7U & (((IData)(0xffU) 
                                        - t__DOT__i) 
                                       >> 5U))] = (
                                                   ((~ 
                                                     ((IData)(1U) 
                                                      << 
                                                      (0x0000001fU 
                                                       & ((IData)(0xffU) 
                                                          - t__DOT__i)))) 
                                                    & vlSelfRef.t__DOT__q
                                                    [
                                                    (7U 
                                                     & (((IData)(0xffU) 
                                                         - t__DOT__i) 
                                                        >> 5U))]) 
                                                   | ((1U 
                                                       & (vlSelfRef.t__DOT__a
                                                          [
                                                          (7U 
                                                           & (t__DOT__i 
                                                              >> 5U))] 
                                                          >> 
                                                          (0x0000001fU 
                                                           & t__DOT__i))) 
                                                      << 
                                                      (0x0000001fU 
                                                       & ((IData)(0xffU) 
                                                          - t__DOT__i))));
            t__DOT__i = ((IData)(1U) + t__DOT__i);
        }
        vlSelfRef.t__DOT__q[0U] = (0x0fed0000U | (0xf000ffffU 
                                                  & vlSelfRef.t__DOT__q[0U]));
        t__DOT__i = 0x00000040U;
    }

    VLCU_D static bool Vt_select_loop___024root___eval_phase__stl(Vt_select_loop___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt_select_loop___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
        }
#endif
        __VstlExecute = Vt_select_loop___024root___trigger_anySet__stl(vlSelfRef.__VstlTriggered);
        if (__VstlExecute) {
            {
                // Inlined CFunc: _eval_stl
                if ((1ULL & vlSelfRef.__VstlTriggered[0U])) {
                    Vt_select_loop___024root___stl_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
        }
        return (__VstlExecute);
    }

    VLCU_D static bool Vt_select_loop___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt_select_loop___024root___nba_sequent__TOP__0(Vt_select_loop___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        IData/*31:0*/ t__DOT__i;
        t__DOT__i = 0;
        IData/*31:0*/ __Vdly__t__DOT__cyc;
        __Vdly__t__DOT__cyc = 0;
        // Body
        __Vdly__t__DOT__cyc = vlSelfRef.t__DOT__cyc;
        if ((0U != vlSelfRef.t__DOT__cyc)) {
            __Vdly__t__DOT__cyc = ((IData)(1U) + vlSelfRef.t__DOT__cyc);
            if ((1U == vlSelfRef.t__DOT__cyc)) {
                VL_ASSIGN_W(256, vlSelfRef.t__DOT__a, Vt_select_loop__ConstPool__CONST_hc278fede_0);
            }
            if ((2U == vlSelfRef.t__DOT__cyc)) {
                VL_ASSIGN_W(256, vlSelfRef.t__DOT__a, Vt_select_loop__ConstPool__CONST_h6bf974be_0);
                if (VL_UNLIKELY(((0U != ((((((((0x2fed1cb7U 
                                                ^ vlSelfRef.t__DOT__q[0U]) 
                                               | (0xacbc2136U 
                                                  ^ vlSelfRef.t__DOT__q[1U])) 
                                              | (0xd5d9127bU 
                                                 ^ vlSelfRef.t__DOT__q[2U])) 
                                             | (0x0eebc824U 
                                                ^ vlSelfRef.t__DOT__q[3U])) 
                                            | (0x6ed707a2U 
                                               ^ vlSelfRef.t__DOT__q[4U])) 
                                           | (0xca128d42U 
                                              ^ vlSelfRef.t__DOT__q[5U])) 
                                          | (0x5bcf892eU 
                                             ^ vlSelfRef.t__DOT__q[6U])) 
                                         | (0x64fe7e28U 
                                            ^ vlSelfRef.t__DOT__q[7U])))))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_select_loop.v", 41, "");
                }
                if (VL_UNLIKELY(((0x64fe7e285fed892eULL 
                                  != vlSelfRef.t__DOT__qq)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_select_loop.v", 42, "");
                }
            }
            if ((3U == vlSelfRef.t__DOT__cyc)) {
                if (VL_UNLIKELY(((0U != ((((((((0xffede870U 
                                                ^ vlSelfRef.t__DOT__q[0U]) 
                                               | (0x58a7fabcU 
                                                  ^ vlSelfRef.t__DOT__q[1U])) 
                                              | (0x13626419U 
                                                 ^ vlSelfRef.t__DOT__q[2U])) 
                                             | (0xc316bd47U 
                                                ^ vlSelfRef.t__DOT__q[3U])) 
                                            | (0xbfbb6c47U 
                                               ^ vlSelfRef.t__DOT__q[4U])) 
                                           | (0x4b5bdbbbU 
                                              ^ vlSelfRef.t__DOT__q[5U])) 
                                          | (0x9c071250U 
                                             ^ vlSelfRef.t__DOT__q[6U])) 
                                         | (0x1da9cf93U 
                                            ^ vlSelfRef.t__DOT__q[7U])))))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_select_loop.v", 45, "");
                }
                if (VL_UNLIKELY(((0x1da9cf939fed1250ULL 
                                  != vlSelfRef.t__DOT__qq)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_select_loop.v", 46, "");
                }
            }
            if (VL_UNLIKELY(((4U == vlSelfRef.t__DOT__cyc)))) {
                vlCtx.writef(0U, 0U);
                VL_FINISH_MT("/root/reference/test_regress/t/t_select_loop.v", 50, "");
            }
        }
        vlSelfRef.t__DOT__cyc = __Vdly__t__DOT__cyc;
        vlSelfRef.t__DOT__qq = (((QData)((IData)(((
                                                   ((((((2U 
                                                         & (vlSelfRef.t__DOT__a[0U] 
                                                            << 1U)) 
                                                        | (1U 
                                                           & (vlSelfRef.t__DOT__a[0U] 
                                                              >> 1U))) 
                                                       << 6U) 
                                                      | (((2U 
                                                           & (vlSelfRef.t__DOT__a[0U] 
                                                              >> 1U)) 
                                                          | (1U 
                                                             & (vlSelfRef.t__DOT__a[0U] 
                                                                >> 3U))) 
                                                         << 4U)) 
                                                     | ((((2U 
                                                           & (vlSelfRef.t__DOT__a[0U] 
                                                              >> 3U)) 
                                                          | (1U 
                                                             & (vlSelfRef.t__DOT__a[0U] 
                                                                >> 5U))) 
                                                         << 2U) 
                                                        | ((2U 
                                                            & (vlSelfRef.t__DOT__a[0U] 
                                                               >> 5U)) 
                                                           | (1U 
                                                              & (vlSelfRef.t__DOT__a[0U] 
                                                                 >> 7U))))) 
                                                    << 0x00000018U) 
                                                   | ((((((2U 
                                                           & (vlSelfRef.t__DOT__a[0U] 
                                                              >> 7U)) 
                                                          | (1U 
                                                             & (vlSelfRef.t__DOT__a[0U] 
                                                                >> 9U))) 
                                                         << 6U) 
                                                        | (((2U 
                                                             & (vlSelfRef.t__DOT__a[0U] 
                                                                >> 9U)) 
                                                            | (1U 
                                                               & (vlSelfRef.t__DOT__a[0U] 
                                                                  >> 0x0000000bU))) 
                                                           << 4U)) 
                                                       | ((((2U 
                                                             & (vlSelfRef.t__DOT__a[0U] 
                                                                >> 0x0000000bU)) 
                                                            | (1U 
                                                               & (vlSelfRef.t__DOT__a[0U] 
                                                                  >> 0x0000000dU))) 
                                                           << 2U) 
                                                          | ((2U 
                                                              & (vlSelfRef.t__DOT__a[0U] 
                                                                 >> 0x0000000dU)) 
                                                             | (1U 
                                                                & (vlSelfRef.t__DOT__a[0U] 
                                                                   >> 0x0000000fU))))) 
                                                      << 0x00000010U)) 
                                                  | (((((((2U 
                                                           & (vlSelfRef.t__DOT__a[0U] 
                                                              >> 0x0000000fU)) 
                                                          | (1U 
                                                             & (vlSelfRef.t__DOT__a[0U] 
                                                                >> 0x00000011U))) 
                                                         << 6U) 
                                                        | (((2U 
                                                             & (vlSelfRef.t__DOT__a[0U] 
                                                                >> 0x00000011U)) 
                                                            | (1U 
                                                               & (vlSelfRef.t__DOT__a[0U] 
                                                                  >> 0x00000013U))) 
                                                           << 4U)) 
                                                       | ((((2U 
                                                             & (vlSelfRef.t__DOT__a[0U] 
                                                                >> 0x00000013U)) 
                                                            | (1U 
                                                               & (vlSelfRef.t__DOT__a[0U] 
                                                                  >> 0x00000015U))) 
                                                           << 2U) 
                                                          | ((2U 
                                                              & (vlSelfRef.t__DOT__a[0U] 
                                                                 >> 0x00000015U)) 
                                                             | (1U 
                                                                & (vlSelfRef.t__DOT__a[0U] 
                                                                   >> 0x00000017U))))) 
                                                      << 8U) 
                                                     | (((((2U 
                                                            & (vlSelfRef.t__DOT__a[0U] 
                                                               >> 0x00000017U)) 
                                                           | (1U 
                                                              & (vlSelfRef.t__DOT__a[0U] 
                                                                 >> 0x00000019U))) 
                                                          << 6U) 
                                                         | (((2U 
                                                              & (vlSelfRef.t__DOT__a[0U] 
                                                                 >> 0x00000019U)) 
                                                             | (1U 
                                                                & (vlSelfRef.t__DOT__a[0U] 
                                                                   >> 0x0000001bU))) 
                                                            << 4U)) 
                                                        | ((((2U 
                                                              & (vlSelfRef.t__DOT__a[0U] 
                                                                 >> 0x0000001bU)) 
                                                             | (1U 
                                                                & (vlSelfRef.t__DOT__a[0U] 
                                                                   >> 0x0000001dU))) 
                                                            << 2U) 
                                                           | ((2U 
                                                               & (vlSelfRef.t__DOT__a[0U] 
                                                                  >> 0x0000001dU)) 
                                                              | (vlSelfRef.t__DOT__a[0U] 
                                                                 >> 0x0000001fU)))))))) 
                                 << 0x00000020U) | (QData)((IData)(
                                                                   (0x0fed0000U 
                                                                    | (((((2U 
                                                                           & (vlSelfRef.t__DOT__a[1U] 
                                                                              << 1U)) 
                                                                          | (1U 
                                                                             & (vlSelfRef.t__DOT__a[1U] 
                                                                                >> 1U))) 
                                                                         << 0x0000001eU) 
                                                                        | (((0x00002000U 
                                                                             & (vlSelfRef.t__DOT__a[1U] 
                                                                                << 0x0000000bU)) 
                                                                            | (0x00001000U 
                                                                               & (vlSelfRef.t__DOT__a[1U] 
                                                                                << 9U))) 
                                                                           << 0x00000010U)) 
                                                                       | (((((((2U 
                                                                                & (vlSelfRef.t__DOT__a[1U] 
                                                                                >> 0x0000000fU)) 
                                                                               | (1U 
                                                                                & (vlSelfRef.t__DOT__a[1U] 
                                                                                >> 0x00000011U))) 
                                                                              << 6U) 
                                                                             | (((2U 
                                                                                & (vlSelfRef.t__DOT__a[1U] 
                                                                                >> 0x00000011U)) 
                                                                                | (1U 
                                                                                & (vlSelfRef.t__DOT__a[1U] 
                                                                                >> 0x00000013U))) 
                                                                                << 4U)) 
                                                                            | ((((2U 
                                                                                & (vlSelfRef.t__DOT__a[1U] 
                                                                                >> 0x00000013U)) 
                                                                                | (1U 
                                                                                & (vlSelfRef.t__DOT__a[1U] 
                                                                                >> 0x00000015U))) 
                                                                                << 2U) 
                                                                               | ((2U 
                                                                                & (vlSelfRef.t__DOT__a[1U] 
                                                                                >> 0x00000015U)) 
                                                                                | (1U 
                                                                                & (vlSelfRef.t__DOT__a[1U] 
                                                                                >> 0x00000017U))))) 
                                                                           << 8U) 
                                                                          | (((((2U 
                                                                                & (vlSelfRef.t__DOT__a[1U] 
                                                                                >> 0x00000017U)) 
                                                                                | (1U 
                                                                                & (vlSelfRef.t__DOT__a[1U] 
                                                                                >> 0x00000019U))) 
                                                                               << 6U) 
                                                                              | (((2U 
                                                                                & (vlSelfRef.t__DOT__a[1U] 
                                                                                >> 0x00000019U)) 
                                                                                | (1U 
                                                                                & (vlSelfRef.t__DOT__a[1U] 
                                                                                >> 0x0000001bU))) 
                                                                                << 4U)) 
                                                                             | ((((2U 
                                                                                & (vlSelfRef.t__DOT__a[1U] 
                                                                                >> 0x0000001bU)) 
                                                                                | (1U 
                                                                                & (vlSelfRef.t__DOT__a[1U] 
                                                                                >> 0x0000001dU))) 
                                                                                << 2U) 
                                                                                | ((2U 
                                                                                & (vlSelfRef.t__DOT__a[1U] 
                                                                                >> 0x0000001dU)) 
                                                                                | (vlSelfRef.t__DOT__a[1U] 
                                                                                >> 0x0000001fU))))))))));
        t__DOT__i = 0U;
        while (VL_GTS_III(32, 0x00000100U, t__DOT__i)) {
            vlSelfRef.t__DOT__q[(7U & (((IData)(0xffU) 
                                        - t__DOT__i) 
                                       >> 5U))] = (
                                                   ((~ 
                                                     ((IData)(1U) 
                                                      << 
                                                      (0x0000001fU 
                                                       & ((IData)(0xffU) 
                                                          - t__DOT__i)))) 
                                                    & vlSelfRef.t__DOT__q
                                                    [
                                                    (7U 
                                                     & (((IData)(0xffU) 
                                                         - t__DOT__i) 
                                                        >> 5U))]) 
                                                   | ((1U 
                                                       & (vlSelfRef.t__DOT__a
                                                          [
                                                          (7U 
                                                           & (t__DOT__i 
                                                              >> 5U))] 
                                                          >> 
                                                          (0x0000001fU 
                                                           & t__DOT__i))) 
                                                      << 
                                                      (0x0000001fU 
                                                       & ((IData)(0xffU) 
                                                          - t__DOT__i))));
            t__DOT__i = ((IData)(1U) + t__DOT__i);
        }
        vlSelfRef.t__DOT__q[0U] = (0x0fed0000U | (0xf000ffffU 
                                                  & vlSelfRef.t__DOT__q[0U]));
        t__DOT__i = 0x00000040U;
    }

    VLCU_D static void Vt_select_loop___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt_select_loop___024root___eval_phase__act(Vt_select_loop___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt_select_loop___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vt_select_loop___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vt_select_loop___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vt_select_loop___024root___eval_phase__nba(Vt_select_loop___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vt_select_loop___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    Vt_select_loop___024root___nba_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
            Vt_select_loop___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vt_select_loop___024root___eval(Vt_select_loop___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt_select_loop___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("/root/reference/test_regress/t/t_select_loop.v", 7, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vt_select_loop___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("/root/reference/test_regress/t/t_select_loop.v", 7, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vt_select_loop___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vt_select_loop___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vt_select_loop___024root___ctor_var_reset(Vt_select_loop___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        vlSelf->t__DOT__cyc = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 15508623621340971111ull);
        VL_SCOPED_RAND_RESET_W(256, vlSelf->t__DOT__a, __VscopeHash, 568261790826559771ull);
        VL_SCOPED_RAND_RESET_W(256, vlSelf->t__DOT__q, __VscopeHash, 4790251044348089774ull);
        vlSelf->t__DOT__qq = VL_SCOPED_RAND_RESET_Q(64, __VscopeHash, 876090779370479156ull);
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
        Vt_select_loop___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt_select_loop___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt_select_loop___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt_select_loop___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vt_select_loop___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt_select_loop___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt_select_loop__cuda::info() {
    typedef Vt_select_loop__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"__VstlFirstIteration", "__VstlFirstIteration", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlFirstIteration},
        {"__VstlPhaseResult", "__VstlPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"t__DOT__cyc", "t.cyc", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__cyc},
        {"t__DOT__a", "t.a", VLCU_MEMBER_STATE, 4, 0, 8, 256, 255, 0, T::OFF_t__DOT__a},
        {"t__DOT__q", "t.q", VLCU_MEMBER_STATE, 4, 0, 8, 256, 255, 0, T::OFF_t__DOT__q},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"t__DOT__qq", "t.qq", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF_t__DOT__qq},
        {"__VstlTriggered", "__VstlTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VstlTriggered},
        {"__VactTriggered", "__VactTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VactTriggered},
        {"__VnbaTriggered", "__VnbaTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VnbaTriggered},
    };
    static const char* const messages[] = {"*-* All Finished *-*\n", nullptr};
    static const char* const plusargs[] = {nullptr};
    static const VlcuTraceDecl traceDecls[] = {
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0}};  // (terminator, not counted)
    static const uint32_t traceOrder[] = {0};  // codes in dump order (+ terminator)
    static const uint32_t saveOrder[] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "t", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 1, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 14, 0, 0xafe40ccf7198d8e7ULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt_select_loop__cuda VLCU_TESTBENCHES);
