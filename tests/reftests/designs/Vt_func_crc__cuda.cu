// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt_func_crc"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt_func_crc___024root final {
    CData/*0:0*/ clk;
    CData/*0:0*/ __VstlFirstIteration;
    CData/*0:0*/ __VstlPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    IData/*31:0*/ t__DOT__c;
    IData/*31:0*/ t__DOT__q;
    IData/*31:0*/ t__DOT__cyc;
    IData/*31:0*/ __VactIterCount;
    QData/*63:0*/ t__DOT__d;
    VlUnpacked<QData/*63:0*/, 1> __VstlTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vt_func_crc__cuda final {
    typedef Vt_func_crc___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 48,
        OFF___VstlFirstIteration = 49,
        OFF___VstlPhaseResult = 50,
        OFF___Vtrigprevexpr___TOP__clk__0 = 51,
        OFF___VactPhaseResult = 52,
        OFF___VnbaPhaseResult = 53,
        OFF_t__DOT__c = 32,
        OFF_t__DOT__q = 36,
        OFF_t__DOT__cyc = 40,
        OFF___VactIterCount = 44,
        OFF_t__DOT__d = 0,
        OFF___VstlTriggered = 8,
        OFF___VactTriggered = 16,
        OFF___VnbaTriggered = 24,
        kStateBytes = 56,  // S_persist
        kTouchedReadBytes = 44,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 43,  // S_state_wr
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
        s.t__DOT__c = st.ld<IData>(OFF_t__DOT__c, i);
        s.t__DOT__q = st.ld<IData>(OFF_t__DOT__q, i);
        s.t__DOT__cyc = st.ld<IData>(OFF_t__DOT__cyc, i);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        s.t__DOT__d = st.ld<QData>(OFF_t__DOT__d, i);
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
        st.st<IData>(OFF_t__DOT__c, i, s.t__DOT__c);
        st.st<IData>(OFF_t__DOT__q, i, s.t__DOT__q);
        st.st<IData>(OFF_t__DOT__cyc, i, s.t__DOT__cyc);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
        st.st<QData>(OFF_t__DOT__d, i, s.t__DOT__d);
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

    VLCU_D static void Vt_func_crc___024root___eval_static(Vt_func_crc___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        {
            // Inlined CFunc: _eval_static__TOP
            QData/*63:0*/ __Vinline_0__eval_static__TOP_t__DOT__crc__Vstatic__drev;
            __Vinline_0__eval_static__TOP_t__DOT__crc__Vstatic__drev = 0;
            IData/*31:0*/ __Vinline_0__eval_static__TOP_t__DOT__newcrc__Vstatic__N;
            __Vinline_0__eval_static__TOP_t__DOT__newcrc__Vstatic__N = 0;
            __Vinline_0__eval_static__TOP_t__DOT__crc__Vstatic__drev = 0;
            __Vinline_0__eval_static__TOP_t__DOT__newcrc__Vstatic__N = 0;
        }
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
    }

    VLCU_D static void Vt_func_crc___024root___eval_initial(Vt_func_crc___024root* vlSelf, VlcuContext& vlCtx) {
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

    VLCU_D static void Vt_func_crc___024root___eval_final(Vt_func_crc___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_func_crc___024root___eval_settle(Vt_func_crc___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt_func_crc___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
#endif
                VL_FATAL_MT("/root/reference/test_regress/t/t_func_crc.v", 6, "", "DIDNOTCONVERGE: Settle region did not converge after '--converge-limit' of 10000 tries");
            }
            __VstlIterCount = ((IData)(1U) + __VstlIterCount);
            vlSelfRef.__VstlPhaseResult = Vt_func_crc___024root___eval_phase__stl(vlSelf, vlCtx);
            vlSelfRef.__VstlFirstIteration = 0U;
        } while (vlSelfRef.__VstlPhaseResult);
    }

    VLCU_D static bool Vt_func_crc___024root___trigger_anySet__stl(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt_func_crc___024root___stl_sequent__TOP__0(Vt_func_crc___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        QData/*63:0*/ t__DOT__crc__Vstatic__drev;
        t__DOT__crc__Vstatic__drev = 0;
        IData/*31:0*/ t__DOT__newcrc__Vstatic__N;
        t__DOT__newcrc__Vstatic__N = 0;
        IData/*31:0*/ __Vfunc_t__DOT__crc__0__Vfuncout;
        __Vfunc_t__DOT__crc__0__Vfuncout = 0;
        QData/*63:0*/ __Vfunc_t__DOT__crc__0__di;
        __Vfunc_t__DOT__crc__0__di = 0;
        IData/*31:0*/ __Vfunc_t__DOT__crc__0__ci;
        __Vfunc_t__DOT__crc__0__ci = 0;
        QData/*63:0*/ __Vfunc_t__DOT__reverse__1__Vfuncout;
        __Vfunc_t__DOT__reverse__1__Vfuncout = 0;
        QData/*63:0*/ __Vfunc_t__DOT__reverse__1__di;
        __Vfunc_t__DOT__reverse__1__di = 0;
        IData/*31:0*/ __Vfunc_t__DOT__newcrc__2__Vfuncout;
        __Vfunc_t__DOT__newcrc__2__Vfuncout = 0;
        QData/*63:0*/ __Vfunc_t__DOT__newcrc__2__D;
        __Vfunc_t__DOT__newcrc__2__D = 0;
        IData/*31:0*/ __Vfunc_t__DOT__newcrc__2__C;
        __Vfunc_t__DOT__newcrc__2__C = 0;
        // Body
        __Vfunc_t__DOT__crc__0__ci = vlSelfRef.t__DOT__c;
        __Vfunc_t__DOT__crc__0__di = vlSelfRef.t__DOT__d;
        __Vfunc_t__DOT__reverse__1__di = __Vfunc_t__DOT__crc__0__di;
        __Vfunc_t__DOT__reverse__1__Vfuncout = 0ULL;
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0xfffffffffffffff8ULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | (IData)((IData)(
                                                                  ((4U 
                                                                    & ((IData)(
                                                                               (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x3dU)) 
                                                                       << 2U)) 
                                                                   | ((2U 
                                                                       & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x3eU)) 
                                                                          << 1U)) 
                                                                      | (1U 
                                                                         & (IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x3fU))))))));
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0xffffffffffffffc7ULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | ((QData)((IData)(
                                                                   ((4U 
                                                                     & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x3aU)) 
                                                                        << 2U)) 
                                                                    | ((2U 
                                                                        & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x3bU)) 
                                                                           << 1U)) 
                                                                       | (1U 
                                                                          & (IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x3cU))))))) 
                                                   << 3U));
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0xfffffffffffffe3fULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | ((QData)((IData)(
                                                                   ((4U 
                                                                     & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x37U)) 
                                                                        << 2U)) 
                                                                    | ((2U 
                                                                        & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x38U)) 
                                                                           << 1U)) 
                                                                       | (1U 
                                                                          & (IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x39U))))))) 
                                                   << 6U));
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0xfffffffffffff1ffULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | ((QData)((IData)(
                                                                   ((4U 
                                                                     & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x34U)) 
                                                                        << 2U)) 
                                                                    | ((2U 
                                                                        & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x35U)) 
                                                                           << 1U)) 
                                                                       | (1U 
                                                                          & (IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x36U))))))) 
                                                   << 9U));
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0xffffffffffff8fffULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | ((QData)((IData)(
                                                                   ((4U 
                                                                     & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x31U)) 
                                                                        << 2U)) 
                                                                    | ((2U 
                                                                        & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x32U)) 
                                                                           << 1U)) 
                                                                       | (1U 
                                                                          & (IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x33U))))))) 
                                                   << 0x0000000cU));
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0xfffffffffffc7fffULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | ((QData)((IData)(
                                                                   ((4U 
                                                                     & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x2eU)) 
                                                                        << 2U)) 
                                                                    | ((2U 
                                                                        & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x2fU)) 
                                                                           << 1U)) 
                                                                       | (1U 
                                                                          & (IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x30U))))))) 
                                                   << 0x0000000fU));
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0xffffffffffe3ffffULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | ((QData)((IData)(
                                                                   ((4U 
                                                                     & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x2bU)) 
                                                                        << 2U)) 
                                                                    | ((2U 
                                                                        & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x2cU)) 
                                                                           << 1U)) 
                                                                       | (1U 
                                                                          & (IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x2dU))))))) 
                                                   << 0x00000012U));
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0xffffffffff1fffffULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | ((QData)((IData)(
                                                                   ((4U 
                                                                     & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x28U)) 
                                                                        << 2U)) 
                                                                    | ((2U 
                                                                        & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x29U)) 
                                                                           << 1U)) 
                                                                       | (1U 
                                                                          & (IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x2aU))))))) 
                                                   << 0x00000015U));
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0xfffffffff8ffffffULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | ((QData)((IData)(
                                                                   ((4U 
                                                                     & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x25U)) 
                                                                        << 2U)) 
                                                                    | ((2U 
                                                                        & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x26U)) 
                                                                           << 1U)) 
                                                                       | (1U 
                                                                          & (IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x27U))))))) 
                                                   << 0x00000018U));
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0xffffffffc7ffffffULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | ((QData)((IData)(
                                                                   ((4U 
                                                                     & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x22U)) 
                                                                        << 2U)) 
                                                                    | ((2U 
                                                                        & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x23U)) 
                                                                           << 1U)) 
                                                                       | (1U 
                                                                          & (IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x24U))))))) 
                                                   << 0x0000001bU));
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0xfffffffe3fffffffULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | ((QData)((IData)(
                                                                   ((4U 
                                                                     & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x1fU)) 
                                                                        << 2U)) 
                                                                    | ((2U 
                                                                        & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x20U)) 
                                                                           << 1U)) 
                                                                       | (1U 
                                                                          & (IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x21U))))))) 
                                                   << 0x0000001eU));
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0xfffffff1ffffffffULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | ((QData)((IData)(
                                                                   ((4U 
                                                                     & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x1cU)) 
                                                                        << 2U)) 
                                                                    | ((2U 
                                                                        & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x1dU)) 
                                                                           << 1U)) 
                                                                       | (1U 
                                                                          & (IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x1eU))))))) 
                                                   << 0x00000021U));
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0xffffff8fffffffffULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | ((QData)((IData)(
                                                                   ((4U 
                                                                     & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x19U)) 
                                                                        << 2U)) 
                                                                    | ((2U 
                                                                        & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x1aU)) 
                                                                           << 1U)) 
                                                                       | (1U 
                                                                          & (IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x1bU))))))) 
                                                   << 0x00000024U));
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0xfffffc7fffffffffULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | ((QData)((IData)(
                                                                   ((4U 
                                                                     & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x16U)) 
                                                                        << 2U)) 
                                                                    | ((2U 
                                                                        & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x17U)) 
                                                                           << 1U)) 
                                                                       | (1U 
                                                                          & (IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x18U))))))) 
                                                   << 0x00000027U));
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0xffffe3ffffffffffULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | ((QData)((IData)(
                                                                   ((4U 
                                                                     & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x13U)) 
                                                                        << 2U)) 
                                                                    | ((2U 
                                                                        & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x14U)) 
                                                                           << 1U)) 
                                                                       | (1U 
                                                                          & (IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x15U))))))) 
                                                   << 0x0000002aU));
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0xffff1fffffffffffULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | ((QData)((IData)(
                                                                   ((4U 
                                                                     & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x10U)) 
                                                                        << 2U)) 
                                                                    | ((2U 
                                                                        & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x11U)) 
                                                                           << 1U)) 
                                                                       | (1U 
                                                                          & (IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x12U))))))) 
                                                   << 0x0000002dU));
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0xfff8ffffffffffffULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | ((QData)((IData)(
                                                                   ((4U 
                                                                     & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x0dU)) 
                                                                        << 2U)) 
                                                                    | ((2U 
                                                                        & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x0eU)) 
                                                                           << 1U)) 
                                                                       | (1U 
                                                                          & (IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x0fU))))))) 
                                                   << 0x00000030U));
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0xffc7ffffffffffffULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | ((QData)((IData)(
                                                                   ((4U 
                                                                     & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x0aU)) 
                                                                        << 2U)) 
                                                                    | ((2U 
                                                                        & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x0bU)) 
                                                                           << 1U)) 
                                                                       | (1U 
                                                                          & (IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x0cU))))))) 
                                                   << 0x00000033U));
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0xfe3fffffffffffffULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | ((QData)((IData)(
                                                                   ((4U 
                                                                     & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 7U)) 
                                                                        << 2U)) 
                                                                    | ((2U 
                                                                        & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 8U)) 
                                                                           << 1U)) 
                                                                       | (1U 
                                                                          & (IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 9U))))))) 
                                                   << 0x00000036U));
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0xf1ffffffffffffffULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | ((QData)((IData)(
                                                                   ((4U 
                                                                     & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 4U)) 
                                                                        << 2U)) 
                                                                    | ((2U 
                                                                        & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 5U)) 
                                                                           << 1U)) 
                                                                       | (1U 
                                                                          & (IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 6U))))))) 
                                                   << 0x00000039U));
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0x8fffffffffffffffULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | ((QData)((IData)(
                                                                   ((4U 
                                                                     & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 1U)) 
                                                                        << 2U)) 
                                                                    | ((2U 
                                                                        & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 2U)) 
                                                                           << 1U)) 
                                                                       | (1U 
                                                                          & (IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 3U))))))) 
                                                   << 0x0000003cU));
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0x7fffffffffffffffULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | ((QData)((IData)(
                                                                   (1U 
                                                                    & (IData)(__Vfunc_t__DOT__reverse__1__di)))) 
                                                   << 0x0000003fU));
        t__DOT__crc__Vstatic__drev = __Vfunc_t__DOT__reverse__1__Vfuncout;
        __Vfunc_t__DOT__newcrc__2__C = __Vfunc_t__DOT__crc__0__ci;
        __Vfunc_t__DOT__newcrc__2__D = t__DOT__crc__Vstatic__drev;
        t__DOT__newcrc__Vstatic__N = 0U;
        t__DOT__newcrc__Vstatic__N = ((0xfffffffeU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (1U & (VL_REDXOR_64(
                                                            (0x083213a12471c281ULL 
                                                             & __Vfunc_t__DOT__newcrc__2__D)) 
                                               ^ VL_REDXOR_32(
                                                              (0x29e8a50aU 
                                                               & __Vfunc_t__DOT__newcrc__2__C)))));
        t__DOT__newcrc__Vstatic__N = ((0xfffffffdU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (2U & ((VL_REDXOR_64(
                                                             (0x2208882911711dc3ULL 
                                                              & __Vfunc_t__DOT__newcrc__2__D)) 
                                                ^ VL_REDXOR_32(
                                                               (0x4c116d21U 
                                                                & __Vfunc_t__DOT__newcrc__2__C))) 
                                               << 1U)));
        t__DOT__newcrc__Vstatic__N = ((0xfffffffbU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (4U & ((VL_REDXOR_64(
                                                             (0xf0d01c2633c5b800ULL 
                                                              & __Vfunc_t__DOT__newcrc__2__D)) 
                                                ^ VL_REDXOR_32(
                                                               (0xc8649c80U 
                                                                & __Vfunc_t__DOT__newcrc__2__C))) 
                                               << 2U)));
        t__DOT__newcrc__Vstatic__N = ((0xfffffff7U 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (8U & ((VL_REDXOR_64(
                                                             (0x4044c05f4a3a892bULL 
                                                              & __Vfunc_t__DOT__newcrc__2__D)) 
                                                ^ VL_REDXOR_32(
                                                               (0x13002b8aU 
                                                                & __Vfunc_t__DOT__newcrc__2__C))) 
                                               << 3U)));
        t__DOT__newcrc__Vstatic__N = ((0xffffffefU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x00000010U 
                                         & ((VL_REDXOR_64(
                                                          (0x027038a79198c443ULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0x411149e4U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 4U)));
        t__DOT__newcrc__Vstatic__N = ((0xffffffdfU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x00000020U 
                                         & ((VL_REDXOR_64(
                                                          (0x0607888024d44672ULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0x091c7982U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 5U)));
        t__DOT__newcrc__Vstatic__N = ((0xffffffbfU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x00000040U 
                                         & ((VL_REDXOR_64(
                                                          (0xe20c80645a218666ULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0x8a012381U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 6U)));
        t__DOT__newcrc__Vstatic__N = ((0xffffff7fU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x00000080U 
                                         & ((VL_REDXOR_64(
                                                          (0x685a664825009a05ULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0x1c1c9090U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 7U)));
        t__DOT__newcrc__Vstatic__N = ((0xfffffeffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x00000100U 
                                         & ((VL_REDXOR_64(
                                                          (0x7815f70041ca0cc4ULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0xa86a0ac0U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 8U)));
        t__DOT__newcrc__Vstatic__N = ((0xfffffdffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x00000200U 
                                         & ((VL_REDXOR_64(
                                                          (0x4e4c0a8013e58080ULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0x6ac030e2U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 9U)));
        t__DOT__newcrc__Vstatic__N = ((0xfffffbffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x00000400U 
                                         & ((VL_REDXOR_64(
                                                          (0x35694c6acd080542ULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0xc721100cU 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 0x0000000aU)));
        t__DOT__newcrc__Vstatic__N = ((0xfffff7ffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x00000800U 
                                         & ((VL_REDXOR_64(
                                                          (0x0b07e7034220f822ULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0x0b21108eU 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 0x0000000bU)));
        t__DOT__newcrc__Vstatic__N = ((0xffffefffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x00001000U 
                                         & ((VL_REDXOR_64(
                                                          (0x6da19808d1a04024ULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0x16c58040U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 0x0000000cU)));
        t__DOT__newcrc__Vstatic__N = ((0xffffdfffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x00002000U 
                                         & ((VL_REDXOR_64(
                                                          (0x96e8a6401656b800ULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0x324bd600U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 0x0000000dU)));
        t__DOT__newcrc__Vstatic__N = ((0xffffbfffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x00004000U 
                                         & ((VL_REDXOR_64(
                                                          (0x05908bc0464a2f0dULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0x9017a816U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 0x0000000eU)));
        t__DOT__newcrc__Vstatic__N = ((0xffff7fffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x00008000U 
                                         & ((VL_REDXOR_64(
                                                          (0xec018c0817fa0884ULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0x69138a20U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 0x0000000fU)));
        t__DOT__newcrc__Vstatic__N = ((0xfffeffffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x00010000U 
                                         & ((VL_REDXOR_64(
                                                          (0x120268917c9a0922ULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0x14c53640U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 0x00000010U)));
        t__DOT__newcrc__Vstatic__N = ((0xfffdffffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x00020000U 
                                         & ((VL_REDXOR_64(
                                                          (0xf46906341d003291ULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0xc8924259U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 0x00000011U)));
        t__DOT__newcrc__Vstatic__N = ((0xfffbffffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x00040000U 
                                         & ((VL_REDXOR_64(
                                                          (0xa914842ce04b0b80ULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0x04590961U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 0x00000012U)));
        t__DOT__newcrc__Vstatic__N = ((0xfff7ffffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x00080000U 
                                         & ((VL_REDXOR_64(
                                                          (0x50121c23240b14dcULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0x664840e5U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 0x00000013U)));
        t__DOT__newcrc__Vstatic__N = ((0xffefffffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x00100000U 
                                         & ((VL_REDXOR_64(
                                                          (0x8441818d9a0412beULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0xb20e8650U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 0x00000014U)));
        t__DOT__newcrc__Vstatic__N = ((0xffdfffffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x00200000U 
                                         & ((VL_REDXOR_64(
                                                          (0x2b211a082617126aULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0x49f0220cU 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 0x00000015U)));
        t__DOT__newcrc__Vstatic__N = ((0xffbfffffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x00400000U 
                                         & ((VL_REDXOR_64(
                                                          (0xd2283406081c8718ULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0x01875e60U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 0x00000016U)));
        t__DOT__newcrc__Vstatic__N = ((0xff7fffffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x00800000U 
                                         & ((VL_REDXOR_64(
                                                          (0x05488d32029595a8ULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0x8ca48d82U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 0x00000017U)));
        t__DOT__newcrc__Vstatic__N = ((0xfeffffffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x01000000U 
                                         & ((VL_REDXOR_64(
                                                          (0x18151098d8e87213ULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0x0aa20c11U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 0x00000018U)));
        t__DOT__newcrc__Vstatic__N = ((0xfdffffffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x02000000U 
                                         & ((VL_REDXOR_64(
                                                          (0x31484b7038428494ULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0x34847b40U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 0x00000019U)));
        t__DOT__newcrc__Vstatic__N = ((0xfbffffffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x04000000U 
                                         & ((VL_REDXOR_64(
                                                          (0xc4d400baa804440dULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0x8f203402U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 0x0000001aU)));
        t__DOT__newcrc__Vstatic__N = ((0xf7ffffffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x08000000U 
                                         & ((VL_REDXOR_64(
                                                          (0x55c81215b8821954ULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0xd88b5c08U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 0x0000001bU)));
        t__DOT__newcrc__Vstatic__N = ((0xefffffffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x10000000U 
                                         & ((VL_REDXOR_64(
                                                          (0x026d0146e8a042f1ULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0xc500e0d9U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 0x0000001cU)));
        t__DOT__newcrc__Vstatic__N = ((0xdfffffffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x20000000U 
                                         & ((VL_REDXOR_64(
                                                          (0x50807b0b52c021fbULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0x98040957U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 0x0000001dU)));
        t__DOT__newcrc__Vstatic__N = ((0xbfffffffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x40000000U 
                                         & ((VL_REDXOR_64(
                                                          (0xcc90901ca078c44dULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0x1350d611U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 0x0000001eU)));
        t__DOT__newcrc__Vstatic__N = ((0x7fffffffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | ((VL_REDXOR_64(
                                                       (0x25de0474ccec1001ULL 
                                                        & __Vfunc_t__DOT__newcrc__2__D)) 
                                          ^ VL_REDXOR_32(
                                                         (0x15236504U 
                                                          & __Vfunc_t__DOT__newcrc__2__C))) 
                                         << 0x0000001fU));
        __Vfunc_t__DOT__newcrc__2__Vfuncout = t__DOT__newcrc__Vstatic__N;
        __Vfunc_t__DOT__crc__0__Vfuncout = __Vfunc_t__DOT__newcrc__2__Vfuncout;
        vlSelfRef.t__DOT__q = __Vfunc_t__DOT__crc__0__Vfuncout;
    }

    VLCU_D static bool Vt_func_crc___024root___eval_phase__stl(Vt_func_crc___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt_func_crc___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
        }
#endif
        __VstlExecute = Vt_func_crc___024root___trigger_anySet__stl(vlSelfRef.__VstlTriggered);
        if (__VstlExecute) {
            {
                // Inlined CFunc: _eval_stl
                if ((1ULL & vlSelfRef.__VstlTriggered[0U])) {
                    Vt_func_crc___024root___stl_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
        }
        return (__VstlExecute);
    }

    VLCU_D static bool Vt_func_crc___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt_func_crc___024root___nba_sequent__TOP__0(Vt_func_crc___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        QData/*63:0*/ t__DOT__crc__Vstatic__drev;
        t__DOT__crc__Vstatic__drev = 0;
        IData/*31:0*/ t__DOT__newcrc__Vstatic__N;
        t__DOT__newcrc__Vstatic__N = 0;
        IData/*31:0*/ __Vfunc_t__DOT__crc__0__Vfuncout;
        __Vfunc_t__DOT__crc__0__Vfuncout = 0;
        QData/*63:0*/ __Vfunc_t__DOT__crc__0__di;
        __Vfunc_t__DOT__crc__0__di = 0;
        IData/*31:0*/ __Vfunc_t__DOT__crc__0__ci;
        __Vfunc_t__DOT__crc__0__ci = 0;
        QData/*63:0*/ __Vfunc_t__DOT__reverse__1__Vfuncout;
        __Vfunc_t__DOT__reverse__1__Vfuncout = 0;
        QData/*63:0*/ __Vfunc_t__DOT__reverse__1__di;
        __Vfunc_t__DOT__reverse__1__di = 0;
        IData/*31:0*/ __Vfunc_t__DOT__newcrc__2__Vfuncout;
        __Vfunc_t__DOT__newcrc__2__Vfuncout = 0;
        QData/*63:0*/ __Vfunc_t__DOT__newcrc__2__D;
        __Vfunc_t__DOT__newcrc__2__D = 0;
        IData/*31:0*/ __Vfunc_t__DOT__newcrc__2__C;
        __Vfunc_t__DOT__newcrc__2__C = 0;
        IData/*31:0*/ __Vdly__t__DOT__cyc;
        __Vdly__t__DOT__cyc = 0;
        // Body
        __Vdly__t__DOT__cyc = vlSelfRef.t__DOT__cyc;
        if ((0U != vlSelfRef.t__DOT__cyc)) {
            __Vdly__t__DOT__cyc = ((IData)(1U) + vlSelfRef.t__DOT__cyc);
            vlSelfRef.t__DOT__c = vlSelfRef.t__DOT__q;
            vlSelfRef.t__DOT__d = ((vlSelfRef.t__DOT__d 
                                    << 1U) | (QData)((IData)(
                                                             (1U 
                                                              & VL_REDXOR_32(
                                                                             (0x0000ffffU 
                                                                              & (IData)(
                                                                                (vlSelfRef.t__DOT__d 
                                                                                >> 0x30U))))))));
            if ((1U == vlSelfRef.t__DOT__cyc)) {
                vlSelfRef.t__DOT__c = 0x12345678U;
                vlSelfRef.t__DOT__d = 0xffffffffffffffffULL;
            }
            if ((2U == vlSelfRef.t__DOT__cyc)) {
                vlSelfRef.t__DOT__d = 0xffffffffffffffffULL;
            }
            if ((3U == vlSelfRef.t__DOT__cyc)) {
                vlSelfRef.t__DOT__d = 0xffffffffffffffffULL;
            }
            if ((4U == vlSelfRef.t__DOT__cyc)) {
                vlSelfRef.t__DOT__d = 0x5018372181a04b1aULL;
            }
            if (VL_UNLIKELY(((9U == vlSelfRef.t__DOT__cyc)))) {
                if (VL_UNLIKELY(((0x38295e96U != vlSelfRef.t__DOT__q)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_func_crc.v", 43, "");
                }
                vlCtx.writef(0U, 0U);
                VL_FINISH_MT("/root/reference/test_regress/t/t_func_crc.v", 45, "");
            }
        }
        vlSelfRef.t__DOT__cyc = __Vdly__t__DOT__cyc;
        __Vfunc_t__DOT__crc__0__ci = vlSelfRef.t__DOT__c;
        __Vfunc_t__DOT__crc__0__di = vlSelfRef.t__DOT__d;
        __Vfunc_t__DOT__reverse__1__di = __Vfunc_t__DOT__crc__0__di;
        __Vfunc_t__DOT__reverse__1__Vfuncout = 0ULL;
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0xfffffffffffffff8ULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | (IData)((IData)(
                                                                  ((4U 
                                                                    & ((IData)(
                                                                               (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x3dU)) 
                                                                       << 2U)) 
                                                                   | ((2U 
                                                                       & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x3eU)) 
                                                                          << 1U)) 
                                                                      | (1U 
                                                                         & (IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x3fU))))))));
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0xffffffffffffffc7ULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | ((QData)((IData)(
                                                                   ((4U 
                                                                     & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x3aU)) 
                                                                        << 2U)) 
                                                                    | ((2U 
                                                                        & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x3bU)) 
                                                                           << 1U)) 
                                                                       | (1U 
                                                                          & (IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x3cU))))))) 
                                                   << 3U));
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0xfffffffffffffe3fULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | ((QData)((IData)(
                                                                   ((4U 
                                                                     & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x37U)) 
                                                                        << 2U)) 
                                                                    | ((2U 
                                                                        & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x38U)) 
                                                                           << 1U)) 
                                                                       | (1U 
                                                                          & (IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x39U))))))) 
                                                   << 6U));
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0xfffffffffffff1ffULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | ((QData)((IData)(
                                                                   ((4U 
                                                                     & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x34U)) 
                                                                        << 2U)) 
                                                                    | ((2U 
                                                                        & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x35U)) 
                                                                           << 1U)) 
                                                                       | (1U 
                                                                          & (IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x36U))))))) 
                                                   << 9U));
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0xffffffffffff8fffULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | ((QData)((IData)(
                                                                   ((4U 
                                                                     & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x31U)) 
                                                                        << 2U)) 
                                                                    | ((2U 
                                                                        & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x32U)) 
                                                                           << 1U)) 
                                                                       | (1U 
                                                                          & (IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x33U))))))) 
                                                   << 0x0000000cU));
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0xfffffffffffc7fffULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | ((QData)((IData)(
                                                                   ((4U 
                                                                     & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x2eU)) 
                                                                        << 2U)) 
                                                                    | ((2U 
                                                                        & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x2fU)) 
                                                                           << 1U)) 
                                                                       | (1U 
                                                                          & (IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x30U))))))) 
                                                   << 0x0000000fU));
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0xffffffffffe3ffffULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | ((QData)((IData)(
                                                                   ((4U 
                                                                     & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x2bU)) 
                                                                        << 2U)) 
                                                                    | ((2U 
                                                                        & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x2cU)) 
                                                                           << 1U)) 
                                                                       | (1U 
                                                                          & (IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x2dU))))))) 
                                                   << 0x00000012U));
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0xffffffffff1fffffULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | ((QData)((IData)(
                                                                   ((4U 
                                                                     & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x28U)) 
                                                                        << 2U)) 
                                                                    | ((2U 
                                                                        & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x29U)) 
                                                                           << 1U)) 
                                                                       | (1U 
                                                                          & (IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x2aU))))))) 
                                                   << 0x00000015U));
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0xfffffffff8ffffffULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | ((QData)((IData)(
                                                                   ((4U 
                                                                     & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x25U)) 
                                                                        << 2U)) 
                                                                    | ((2U 
                                                                        & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x26U)) 
                                                                           << 1U)) 
                                                                       | (1U 
                                                                          & (IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x27U))))))) 
                                                   << 0x00000018U));
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0xffffffffc7ffffffULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | ((QData)((IData)(
                                                                   ((4U 
                                                                     & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x22U)) 
                                                                        << 2U)) 
                                                                    | ((2U 
                                                                        & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x23U)) 
                                                                           << 1U)) 
                                                                       | (1U 
                                                                          & (IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x24U))))))) 
                                                   << 0x0000001bU));
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0xfffffffe3fffffffULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | ((QData)((IData)(
                                                                   ((4U 
                                                                     & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x1fU)) 
                                                                        << 2U)) 
                                                                    | ((2U 
                                                                        & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x20U)) 
                                                                           << 1U)) 
                                                                       | (1U 
                                                                          & (IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x21U))))))) 
                                                   << 0x0000001eU));
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0xfffffff1ffffffffULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | ((QData)((IData)(
                                                                   ((4U 
                                                                     & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x1cU)) 
                                                                        << 2U)) 
                                                                    | ((2U 
                                                                        & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x1dU)) 
                                                                           << 1U)) 
                                                                       | (1U 
                                                                          & (IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x1eU))))))) 
                                                   << 0x00000021U));
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0xffffff8fffffffffULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | ((QData)((IData)(
                                                                   ((4U 
                                                                     & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x19U)) 
                                                                        << 2U)) 
                                                                    | ((2U 
                                                                        & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x1aU)) 
                                                                           << 1U)) 
                                                                       | (1U 
                                                                          & (IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x1bU))))))) 
                                                   << 0x00000024U));
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0xfffffc7fffffffffULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | ((QData)((IData)(
                                                                   ((4U 
                                                                     & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x16U)) 
                                                                        << 2U)) 
                                                                    | ((2U 
                                                                        & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x17U)) 
                                                                           << 1U)) 
                                                                       | (1U 
                                                                          & (IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x18U))))))) 
                                                   << 0x00000027U));
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0xffffe3ffffffffffULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | ((QData)((IData)(
                                                                   ((4U 
                                                                     & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x13U)) 
                                                                        << 2U)) 
                                                                    | ((2U 
                                                                        & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x14U)) 
                                                                           << 1U)) 
                                                                       | (1U 
                                                                          & (IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x15U))))))) 
                                                   << 0x0000002aU));
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0xffff1fffffffffffULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | ((QData)((IData)(
                                                                   ((4U 
                                                                     & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x10U)) 
                                                                        << 2U)) 
                                                                    | ((2U 
                                                                        & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x11U)) 
                                                                           << 1U)) 
                                                                       | (1U 
                                                                          & (IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x12U))))))) 
                                                   << 0x0000002dU));
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0xfff8ffffffffffffULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | ((QData)((IData)(
                                                                   ((4U 
                                                                     & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x0dU)) 
                                                                        << 2U)) 
                                                                    | ((2U 
                                                                        & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x0eU)) 
                                                                           << 1U)) 
                                                                       | (1U 
                                                                          & (IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x0fU))))))) 
                                                   << 0x00000030U));
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0xffc7ffffffffffffULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | ((QData)((IData)(
                                                                   ((4U 
                                                                     & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x0aU)) 
                                                                        << 2U)) 
                                                                    | ((2U 
                                                                        & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x0bU)) 
                                                                           << 1U)) 
                                                                       | (1U 
                                                                          & (IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 0x0cU))))))) 
                                                   << 0x00000033U));
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0xfe3fffffffffffffULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | ((QData)((IData)(
                                                                   ((4U 
                                                                     & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 7U)) 
                                                                        << 2U)) 
                                                                    | ((2U 
                                                                        & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 8U)) 
                                                                           << 1U)) 
                                                                       | (1U 
                                                                          & (IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 9U))))))) 
                                                   << 0x00000036U));
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0xf1ffffffffffffffULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | ((QData)((IData)(
                                                                   ((4U 
                                                                     & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 4U)) 
                                                                        << 2U)) 
                                                                    | ((2U 
                                                                        & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 5U)) 
                                                                           << 1U)) 
                                                                       | (1U 
                                                                          & (IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 6U))))))) 
                                                   << 0x00000039U));
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0x8fffffffffffffffULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | ((QData)((IData)(
                                                                   ((4U 
                                                                     & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 1U)) 
                                                                        << 2U)) 
                                                                    | ((2U 
                                                                        & ((IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 2U)) 
                                                                           << 1U)) 
                                                                       | (1U 
                                                                          & (IData)(
                                                                                (__Vfunc_t__DOT__reverse__1__di 
                                                                                >> 3U))))))) 
                                                   << 0x0000003cU));
        __Vfunc_t__DOT__reverse__1__Vfuncout = ((0x7fffffffffffffffULL 
                                                 & __Vfunc_t__DOT__reverse__1__Vfuncout) 
                                                | ((QData)((IData)(
                                                                   (1U 
                                                                    & (IData)(__Vfunc_t__DOT__reverse__1__di)))) 
                                                   << 0x0000003fU));
        t__DOT__crc__Vstatic__drev = __Vfunc_t__DOT__reverse__1__Vfuncout;
        __Vfunc_t__DOT__newcrc__2__C = __Vfunc_t__DOT__crc__0__ci;
        __Vfunc_t__DOT__newcrc__2__D = t__DOT__crc__Vstatic__drev;
        t__DOT__newcrc__Vstatic__N = 0U;
        t__DOT__newcrc__Vstatic__N = ((0xfffffffeU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (1U & (VL_REDXOR_64(
                                                            (0x083213a12471c281ULL 
                                                             & __Vfunc_t__DOT__newcrc__2__D)) 
                                               ^ VL_REDXOR_32(
                                                              (0x29e8a50aU 
                                                               & __Vfunc_t__DOT__newcrc__2__C)))));
        t__DOT__newcrc__Vstatic__N = ((0xfffffffdU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (2U & ((VL_REDXOR_64(
                                                             (0x2208882911711dc3ULL 
                                                              & __Vfunc_t__DOT__newcrc__2__D)) 
                                                ^ VL_REDXOR_32(
                                                               (0x4c116d21U 
                                                                & __Vfunc_t__DOT__newcrc__2__C))) 
                                               << 1U)));
        t__DOT__newcrc__Vstatic__N = ((0xfffffffbU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (4U & ((VL_REDXOR_64(
                                                             (0xf0d01c2633c5b800ULL 
                                                              & __Vfunc_t__DOT__newcrc__2__D)) 
                                                ^ VL_REDXOR_32(
                                                               (0xc8649c80U 
                                                                & __Vfunc_t__DOT__newcrc__2__C))) 
                                               << 2U)));
        t__DOT__newcrc__Vstatic__N = ((0xfffffff7U 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (8U & ((VL_REDXOR_64(
                                                             (0x4044c05f4a3a892bULL 
                                                              & __Vfunc_t__DOT__newcrc__2__D)) 
                                                ^ VL_REDXOR_32(
                                                               (0x13002b8aU 
                                                                & __Vfunc_t__DOT__newcrc__2__C))) 
                                               << 3U)));
        t__DOT__newcrc__Vstatic__N = ((0xffffffefU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x00000010U 
                                         & ((VL_REDXOR_64(
                                                          (0x027038a79198c443ULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0x411149e4U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 4U)));
        t__DOT__newcrc__Vstatic__N = ((0xffffffdfU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x00000020U 
                                         & ((VL_REDXOR_64(
                                                          (0x0607888024d44672ULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0x091c7982U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 5U)));
        t__DOT__newcrc__Vstatic__N = ((0xffffffbfU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x00000040U 
                                         & ((VL_REDXOR_64(
                                                          (0xe20c80645a218666ULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0x8a012381U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 6U)));
        t__DOT__newcrc__Vstatic__N = ((0xffffff7fU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x00000080U 
                                         & ((VL_REDXOR_64(
                                                          (0x685a664825009a05ULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0x1c1c9090U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 7U)));
        t__DOT__newcrc__Vstatic__N = ((0xfffffeffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x00000100U 
                                         & ((VL_REDXOR_64(
                                                          (0x7815f70041ca0cc4ULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0xa86a0ac0U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 8U)));
        t__DOT__newcrc__Vstatic__N = ((0xfffffdffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x00000200U 
                                         & ((VL_REDXOR_64(
                                                          (0x4e4c0a8013e58080ULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0x6ac030e2U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 9U)));
        t__DOT__newcrc__Vstatic__N = ((0xfffffbffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x00000400U 
                                         & ((VL_REDXOR_64(
                                                          (0x35694c6acd080542ULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0xc721100cU 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 0x0000000aU)));
        t__DOT__newcrc__Vstatic__N = ((0xfffff7ffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x00000800U 
                                         & ((VL_REDXOR_64(
                                                          (0x0b07e7034220f822ULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0x0b21108eU 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 0x0000000bU)));
        t__DOT__newcrc__Vstatic__N = ((0xffffefffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x00001000U 
                                         & ((VL_REDXOR_64(
                                                          (0x6da19808d1a04024ULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0x16c58040U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 0x0000000cU)));
        t__DOT__newcrc__Vstatic__N = ((0xffffdfffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x00002000U 
                                         & ((VL_REDXOR_64(
                                                          (0x96e8a6401656b800ULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0x324bd600U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 0x0000000dU)));
        t__DOT__newcrc__Vstatic__N = ((0xffffbfffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x00004000U 
                                         & ((VL_REDXOR_64(
                                                          (0x05908bc0464a2f0dULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0x9017a816U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 0x0000000eU)));
        t__DOT__newcrc__Vstatic__N = ((0xffff7fffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x00008000U 
                                         & ((VL_REDXOR_64(
                                                          (0xec018c0817fa0884ULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0x69138a20U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 0x0000000fU)));
        t__DOT__newcrc__Vstatic__N = ((0xfffeffffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x00010000U 
                                         & ((VL_REDXOR_64(
                                                          (0x120268917c9a0922ULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0x14c53640U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 0x00000010U)));
        t__DOT__newcrc__Vstatic__N = ((0xfffdffffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x00020000U 
                                         & ((VL_REDXOR_64(
                                                          (0xf46906341d003291ULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0xc8924259U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 0x00000011U)));
        t__DOT__newcrc__Vstatic__N = ((0xfffbffffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x00040000U 
                                         & ((VL_REDXOR_64(
                                                          (0xa914842ce04b0b80ULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0x04590961U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 0x00000012U)));
        t__DOT__newcrc__Vstatic__N = ((0xfff7ffffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x00080000U 
                                         & ((VL_REDXOR_64(
                                                          (0x50121c23240b14dcULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0x664840e5U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 0x00000013U)));
        t__DOT__newcrc__Vstatic__N = ((0xffefffffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x00100000U 
                                         & ((VL_REDXOR_64(
                                                          (0x8441818d9a0412beULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0xb20e8650U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 0x00000014U)));
        t__DOT__newcrc__Vstatic__N = ((0xffdfffffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x00200000U 
                                         & ((VL_REDXOR_64(
                                                          (0x2b211a082617126aULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0x49f0220cU 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 0x00000015U)));
        t__DOT__newcrc__Vstatic__N = ((0xffbfffffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x00400000U 
                                         & ((VL_REDXOR_64(
                                                          (0xd2283406081c8718ULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0x01875e60U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 0x00000016U)));
        t__DOT__newcrc__Vstatic__N = ((0xff7fffffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x00800000U 
                                         & ((VL_REDXOR_64(
                                                          (0x05488d32029595a8ULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0x8ca48d82U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 0x00000017U)));
        t__DOT__newcrc__Vstatic__N = ((0xfeffffffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x01000000U 
                                         & ((VL_REDXOR_64(
                                                          (0x18151098d8e87213ULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0x0aa20c11U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 0x00000018U)));
        t__DOT__newcrc__Vstatic__N = ((0xfdffffffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x02000000U 
                                         & ((VL_REDXOR_64(
                                                          (0x31484b7038428494ULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0x34847b40U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 0x00000019U)));
        t__DOT__newcrc__Vstatic__N = ((0xfbffffffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x04000000U 
                                         & ((VL_REDXOR_64(
                                                          (0xc4d400baa804440dULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0x8f203402U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 0x0000001aU)));
        t__DOT__newcrc__Vstatic__N = ((0xf7ffffffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x08000000U 
                                         & ((VL_REDXOR_64(
                                                          (0x55c81215b8821954ULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0xd88b5c08U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 0x0000001bU)));
        t__DOT__newcrc__Vstatic__N = ((0xefffffffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x10000000U 
                                         & ((VL_REDXOR_64(
                                                          (0x026d0146e8a042f1ULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0xc500e0d9U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 0x0000001cU)));
        t__DOT__newcrc__Vstatic__N = ((0xdfffffffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x20000000U 
                                         & ((VL_REDXOR_64(
                                                          (0x50807b0b52c021fbULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0x98040957U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 0x0000001dU)));
        t__DOT__newcrc__Vstatic__N = ((0xbfffffffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | (0x40000000U 
                                         & ((VL_REDXOR_64(
                                                          (0xcc90901ca078c44dULL 
                                                           & __Vfunc_t__DOT__newcrc__2__D)) 
                                             ^ VL_REDXOR_32(
                                                            (0x1350d611U 
                                                             & __Vfunc_t__DOT__newcrc__2__C))) 
                                            << 0x0000001eU)));
        t__DOT__newcrc__Vstatic__N = ((0x7fffffffU 
                                       & t__DOT__newcrc__Vstatic__N) 
                                      | ((VL_REDXOR_64(
                                                       (0x25de0474ccec1001ULL 
                                                        & __Vfunc_t__DOT__newcrc__2__D)) 
                                          ^ VL_REDXOR_32(
                                                         (0x15236504U 
                                                          & __Vfunc_t__DOT__newcrc__2__C))) 
                                         << 0x0000001fU));
        __Vfunc_t__DOT__newcrc__2__Vfuncout = t__DOT__newcrc__Vstatic__N;
        __Vfunc_t__DOT__crc__0__Vfuncout = __Vfunc_t__DOT__newcrc__2__Vfuncout;
        vlSelfRef.t__DOT__q = __Vfunc_t__DOT__crc__0__Vfuncout;
    }

    VLCU_D static void Vt_func_crc___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt_func_crc___024root___eval_phase__act(Vt_func_crc___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt_func_crc___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vt_func_crc___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vt_func_crc___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vt_func_crc___024root___eval_phase__nba(Vt_func_crc___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vt_func_crc___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    Vt_func_crc___024root___nba_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
            Vt_func_crc___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vt_func_crc___024root___eval(Vt_func_crc___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt_func_crc___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("/root/reference/test_regress/t/t_func_crc.v", 6, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vt_func_crc___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("/root/reference/test_regress/t/t_func_crc.v", 6, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vt_func_crc___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vt_func_crc___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vt_func_crc___024root___ctor_var_reset(Vt_func_crc___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        vlSelf->t__DOT__d = VL_SCOPED_RAND_RESET_Q(64, __VscopeHash, 12713365125299031450ull);
        vlSelf->t__DOT__c = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 7928863572445662412ull);
        vlSelf->t__DOT__q = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 4790251044348089774ull);
        vlSelf->t__DOT__cyc = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 15508623621340971111ull);
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
        Vt_func_crc___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt_func_crc___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt_func_crc___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt_func_crc___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vt_func_crc___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt_func_crc___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt_func_crc__cuda::info() {
    typedef Vt_func_crc__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"__VstlFirstIteration", "__VstlFirstIteration", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlFirstIteration},
        {"__VstlPhaseResult", "__VstlPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"t__DOT__c", "t.c", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__c},
        {"t__DOT__q", "t.q", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__q},
        {"t__DOT__cyc", "t.cyc", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__cyc},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"t__DOT__d", "t.d", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF_t__DOT__d},
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
        saveOrder, 14, 0, 0x3a3eb0dd5c930618ULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt_func_crc__cuda VLCU_TESTBENCHES);
