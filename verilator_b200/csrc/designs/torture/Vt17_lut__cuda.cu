// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt17_lut"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt17_lut___024root final {
    CData/*0:0*/ clk;
    CData/*0:0*/ we;
    CData/*7:0*/ a;
    CData/*7:0*/ d;
    CData/*7:0*/ q;
    SData/*15:0*/ sum;
    CData/*7:0*/ nib;
    CData/*0:0*/ __VstlFirstIteration;
    CData/*0:0*/ __VstlPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*0:0*/ __Vtrigprevexpr___TOP__we__0;
    CData/*7:0*/ __Vtrigprevexpr___TOP__a__0;
    CData/*7:0*/ __Vtrigprevexpr___TOP__d__0;
    CData/*0:0*/ __VicoDidInit;
    CData/*0:0*/ __VicoPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__1;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    IData/*31:0*/ __VactIterCount;
    VlcuLocalArray<CData/*7:0*/, 256> t17_lut__DOT__lut;  // thread-local memory
    VlcuSharedArray<CData/*7:0*/, 16> t17_lut__DOT__small;  // run-time indexed: shared memory [index][lane]
    VlUnpacked<QData/*63:0*/, 1> __VstlTriggered;
    VlUnpacked<QData/*63:0*/, 2> __VicoTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vt17_lut__cuda final {
    typedef Vt17_lut___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 46,
        OFF_we = 47,
        OFF_a = 48,
        OFF_d = 49,
        OFF_q = 50,
        OFF_sum = 44,
        OFF_nib = 51,
        OFF___VstlFirstIteration = 52,
        OFF___VstlPhaseResult = 53,
        OFF___Vtrigprevexpr___TOP__clk__0 = 54,
        OFF___Vtrigprevexpr___TOP__we__0 = 55,
        OFF___Vtrigprevexpr___TOP__a__0 = 56,
        OFF___Vtrigprevexpr___TOP__d__0 = 57,
        OFF___VicoDidInit = 58,
        OFF___VicoPhaseResult = 59,
        OFF___Vtrigprevexpr___TOP__clk__1 = 60,
        OFF___VactPhaseResult = 61,
        OFF___VnbaPhaseResult = 62,
        OFF___VactIterCount = 40,
        OFF_t17_lut__DOT__lut = 63,
        OFF_t17_lut__DOT__small = 319,
        OFF___VstlTriggered = 0,
        OFF___VicoTriggered = 8,
        OFF___VactTriggered = 24,
        OFF___VnbaTriggered = 32,
        kStateBytes = 336,  // S_persist
        kTouchedReadBytes = 60,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 51,  // S_state_wr
    };
    struct Locals final {
        CData/*7:0*/ t17_lut__DOT__lut[256];
    };
    struct Shared final {
        CData/*7:0*/ t17_lut__DOT__small[16][VLCU_BLOCK];
    };
    VLCU_D static void bind(Root& s, Locals& l, Shared& sh, const VlcuState& st, uint64_t i) {
        (void)l; (void)sh; (void)st; (void)i;
        s.vlNamep = "TOP";
        s.t17_lut__DOT__lut.bind(l.t17_lut__DOT__lut);
        s.t17_lut__DOT__small.bind(sh.t17_lut__DOT__small, threadIdx.x);
    }
    VLCU_D static void load(Root& s, const VlcuState& st, uint64_t i) {
        s.clk = st.ld<VlcuMem8>(OFF_clk, i);
        s.we = st.ld<VlcuMem8>(OFF_we, i);
        s.a = st.ld<VlcuMem8>(OFF_a, i);
        s.d = st.ld<VlcuMem8>(OFF_d, i);
        s.q = st.ld<VlcuMem8>(OFF_q, i);
        s.sum = st.ld<VlcuMem16>(OFF_sum, i);
        s.nib = st.ld<VlcuMem8>(OFF_nib, i);
        s.__VstlFirstIteration = st.ld<VlcuMem8>(OFF___VstlFirstIteration, i);
        s.__VstlPhaseResult = st.ld<VlcuMem8>(OFF___VstlPhaseResult, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__Vtrigprevexpr___TOP__we__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__we__0, i);
        s.__Vtrigprevexpr___TOP__a__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__a__0, i);
        s.__Vtrigprevexpr___TOP__d__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__d__0, i);
        s.__VicoDidInit = st.ld<VlcuMem8>(OFF___VicoDidInit, i);
        s.__VicoPhaseResult = st.ld<VlcuMem8>(OFF___VicoPhaseResult, i);
        s.__Vtrigprevexpr___TOP__clk__1 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__1, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        for (uint32_t e = 0; e < 256; ++e) s.t17_lut__DOT__lut[e] = st.ldE<VlcuMem8>(OFF_t17_lut__DOT__lut, e, i);
        for (uint32_t e = 0; e < 16; ++e) s.t17_lut__DOT__small[e] = st.ldE<VlcuMem8>(OFF_t17_lut__DOT__small, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VstlTriggered[e] = st.ldE<QData>(OFF___VstlTriggered, e, i);
        for (uint32_t e = 0; e < 2; ++e) s.__VicoTriggered[e] = st.ldE<QData>(OFF___VicoTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.st<VlcuMem8>(OFF_we, i, s.we);
        st.st<VlcuMem8>(OFF_a, i, s.a);
        st.st<VlcuMem8>(OFF_d, i, s.d);
        st.st<VlcuMem8>(OFF_q, i, s.q);
        st.st<VlcuMem16>(OFF_sum, i, s.sum);
        st.st<VlcuMem8>(OFF_nib, i, s.nib);
        st.st<VlcuMem8>(OFF___VstlFirstIteration, i, s.__VstlFirstIteration);
        st.st<VlcuMem8>(OFF___VstlPhaseResult, i, s.__VstlPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__we__0, i, s.__Vtrigprevexpr___TOP__we__0);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__a__0, i, s.__Vtrigprevexpr___TOP__a__0);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__d__0, i, s.__Vtrigprevexpr___TOP__d__0);
        st.st<VlcuMem8>(OFF___VicoDidInit, i, s.__VicoDidInit);
        st.st<VlcuMem8>(OFF___VicoPhaseResult, i, s.__VicoPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__1, i, s.__Vtrigprevexpr___TOP__clk__1);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
        for (uint32_t e = 0; e < 256; ++e) st.stE<VlcuMem8>(OFF_t17_lut__DOT__lut, e, i, s.t17_lut__DOT__lut[e]);
        for (uint32_t e = 0; e < 16; ++e) st.stE<VlcuMem8>(OFF_t17_lut__DOT__small, e, i, s.t17_lut__DOT__small[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VstlTriggered, e, i, s.__VstlTriggered[e]);
        for (uint32_t e = 0; e < 2; ++e) st.stE<QData>(OFF___VicoTriggered, e, i, s.__VicoTriggered[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VactTriggered, e, i, s.__VactTriggered[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VnbaTriggered, e, i, s.__VnbaTriggered[e]);
    }
    VLCU_D static void load_inputs_row(Root& s, const uint32_t* rowp, uint64_t stride, uint64_t i) {
        s.clk = static_cast<VlcuMem8>(rowp[0 * stride + i] & 0x1U);
        s.we = static_cast<VlcuMem8>(rowp[1 * stride + i] & 0x1U);
        s.a = static_cast<VlcuMem8>(rowp[2 * stride + i] & 0xffU);
        s.d = static_cast<VlcuMem8>(rowp[3 * stride + i] & 0xffU);
    }
    VLCU_D static void store_outputs_row(const Root& s, uint32_t* rowp, uint64_t stride, uint64_t i) {
        rowp[0 * stride + i] = s.q;
        rowp[1 * stride + i] = s.sum;
        rowp[2 * stride + i] = s.nib;
    }
    VLCU_D static uint64_t outputs_sig(const Root& s, uint64_t sig) {
        uint32_t acc = VLCU_SIG_ACC_INIT;
        acc = vlcu_sig_word(acc, s.q);
        acc = vlcu_sig_word(acc, s.sum);
        acc = vlcu_sig_word(acc, s.nib);
        return vlcu_sig_mix(sig, acc);
    }
    enum : uint32_t { kInWords = 4, kOutWords = 3, kInBytes = 4, kOutBytes = 4, kTraceWords = 0 };

    VLCU_D static void Vt17_lut___024root___eval_static(Vt17_lut___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
        vlSelfRef.__Vtrigprevexpr___TOP__we__0 = vlSelfRef.we;
        vlSelfRef.__Vtrigprevexpr___TOP__a__0 = vlSelfRef.a;
        vlSelfRef.__Vtrigprevexpr___TOP__d__0 = vlSelfRef.d;
        vlSelfRef.__Vtrigprevexpr___TOP__clk__1 = vlSelfRef.clk;
    }

    VLCU_D static void Vt17_lut___024root___eval_initial(Vt17_lut___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt17_lut___024root___eval_final(Vt17_lut___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt17_lut___024root___eval_settle(Vt17_lut___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt17_lut___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t17_lut.v", 5, "", "DIDNOTCONVERGE: Settle region did not converge after '--converge-limit' of 10000 tries");
            }
            __VstlIterCount = ((IData)(1U) + __VstlIterCount);
            vlSelfRef.__VstlPhaseResult = Vt17_lut___024root___eval_phase__stl(vlSelf, vlCtx);
            vlSelfRef.__VstlFirstIteration = 0U;
        } while (vlSelfRef.__VstlPhaseResult);
    }

    VLCU_D static bool Vt17_lut___024root___trigger_anySet__stl(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt17_lut___024root___eval_phase__stl(Vt17_lut___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt17_lut___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
        }
#endif
        __VstlExecute = Vt17_lut___024root___trigger_anySet__stl(vlSelfRef.__VstlTriggered);
        if (__VstlExecute) {
            {
                // Inlined CFunc: _eval_stl
                if ((1ULL & vlSelfRef.__VstlTriggered[0U])) {
                    {
                        // Inlined CFunc: _ico_comb__TOP__0
                        vlSelfRef.nib = (vlSelfRef.t17_lut__DOT__small
                                         [(0x0000000fU 
                                           & (IData)(vlSelfRef.a))] 
                                         ^ vlSelfRef.t17_lut__DOT__small
                                         [(0x0000000fU 
                                           & ((IData)(vlSelfRef.d) 
                                              >> 4U))]);
                    }
                }
            }
        }
        return (__VstlExecute);
    }

    VLCU_D static bool Vt17_lut___024root___trigger_anySet__ico(const VlUnpacked<QData/*63:0*/, 2> &in) {
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
        } while ((2U > n));
        return (0U);
    }

    VLCU_D static bool Vt17_lut___024root___eval_phase__ico(Vt17_lut___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VicoExecute;
        // Body
        {
            // Inlined CFunc: _eval_triggers_vec__ico
            vlSelfRef.__VicoTriggered[0U] = (QData)((IData)(
                                                            (((((IData)(vlSelfRef.d) 
                                                                != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__d__0)) 
                                                               << 3U) 
                                                              | (((IData)(vlSelfRef.a) 
                                                                  != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__a__0)) 
                                                                 << 2U)) 
                                                             | ((((IData)(vlSelfRef.we) 
                                                                  != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__we__0)) 
                                                                 << 1U) 
                                                                | ((IData)(vlSelfRef.clk) 
                                                                   != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__clk__0))))));
            vlSelfRef.__Vtrigprevexpr___TOP__clk__0 
                = vlSelfRef.clk;
            vlSelfRef.__Vtrigprevexpr___TOP__we__0 
                = vlSelfRef.we;
            vlSelfRef.__Vtrigprevexpr___TOP__a__0 = vlSelfRef.a;
            vlSelfRef.__Vtrigprevexpr___TOP__d__0 = vlSelfRef.d;
            if (VL_UNLIKELY(((1U & (~ (IData)(vlSelfRef.__VicoDidInit)))))) {
                vlSelfRef.__VicoDidInit = 1U;
                vlSelfRef.__VicoTriggered[0U] = (1ULL 
                                                 | vlSelfRef.__VicoTriggered[0U]);
                vlSelfRef.__VicoTriggered[0U] = (2ULL 
                                                 | vlSelfRef.__VicoTriggered[0U]);
                vlSelfRef.__VicoTriggered[0U] = (4ULL 
                                                 | vlSelfRef.__VicoTriggered[0U]);
                vlSelfRef.__VicoTriggered[0U] = (8ULL 
                                                 | vlSelfRef.__VicoTriggered[0U]);
            }
        }
#ifdef VL_DEBUG
        if (VL_UNLIKELY(vlSymsp->_vm_contextp__->debug())) {
            Vt17_lut___024root___dump_triggers__ico(vlSelfRef.__VicoTriggered, "ico"s);
        }
#endif
        __VicoExecute = Vt17_lut___024root___trigger_anySet__ico(vlSelfRef.__VicoTriggered);
        if (__VicoExecute) {
            {
                // Inlined CFunc: _eval_ico
                if ((0x000000000000000cULL & vlSelfRef.__VicoTriggered[0U])) {
                    {
                        // Inlined CFunc: _ico_comb__TOP__0
                        vlSelfRef.nib = (vlSelfRef.t17_lut__DOT__small
                                         [(0x0000000fU 
                                           & (IData)(vlSelfRef.a))] 
                                         ^ vlSelfRef.t17_lut__DOT__small
                                         [(0x0000000fU 
                                           & ((IData)(vlSelfRef.d) 
                                              >> 4U))]);
                    }
                }
            }
        }
        return (__VicoExecute);
    }

    VLCU_D static bool Vt17_lut___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt17_lut___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt17_lut___024root___eval_phase__act(Vt17_lut___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        {
            // Inlined CFunc: _eval_triggers_vec__act
            vlSelfRef.__VactTriggered[0U] = (QData)((IData)(
                                                            ((IData)(vlSelfRef.clk) 
                                                             & (~ (IData)(vlSelfRef.__Vtrigprevexpr___TOP__clk__1)))));
            vlSelfRef.__Vtrigprevexpr___TOP__clk__1 
                = vlSelfRef.clk;
        }
#ifdef VL_DEBUG
        if (VL_UNLIKELY(vlSymsp->_vm_contextp__->debug())) {
            Vt17_lut___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vt17_lut___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vt17_lut___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vt17_lut___024root___eval_phase__nba(Vt17_lut___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vt17_lut___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    {
                        // Inlined CFunc: _nba_sequent__TOP__0
                        CData/*7:0*/ __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___Vdly__q;
                        __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___Vdly__q = 0;
                        CData/*7:0*/ __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyVal__t17_lut__DOT__lut__v0;
                        __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyVal__t17_lut__DOT__lut__v0 = 0;
                        CData/*7:0*/ __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyDim0__t17_lut__DOT__lut__v0;
                        __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyDim0__t17_lut__DOT__lut__v0 = 0;
                        CData/*0:0*/ __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlySet__t17_lut__DOT__lut__v0;
                        __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlySet__t17_lut__DOT__lut__v0 = 0;
                        CData/*7:0*/ __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyVal__t17_lut__DOT__small__v0;
                        __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyVal__t17_lut__DOT__small__v0 = 0;
                        CData/*3:0*/ __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyDim0__t17_lut__DOT__small__v0;
                        __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyDim0__t17_lut__DOT__small__v0 = 0;
                        CData/*0:0*/ __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlySet__t17_lut__DOT__small__v0;
                        __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlySet__t17_lut__DOT__small__v0 = 0;
                        __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlySet__t17_lut__DOT__lut__v0 = 0U;
                        __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___Vdly__q 
                            = vlSelfRef.q;
                        __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlySet__t17_lut__DOT__small__v0 = 0U;
                        if (vlSelfRef.we) {
                            __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyVal__t17_lut__DOT__lut__v0 
                                = vlSelfRef.d;
                            __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyDim0__t17_lut__DOT__lut__v0 
                                = vlSelfRef.a;
                            __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlySet__t17_lut__DOT__lut__v0 = 1U;
                        }
                        __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___Vdly__q 
                            = (0x000000ffU & ((IData)(vlSelfRef.t17_lut__DOT__lut
                                                      [
                                                      ((IData)(vlSelfRef.a) 
                                                       ^ (IData)(vlSelfRef.d))]) 
                                              + (IData)(vlSelfRef.t17_lut__DOT__small
                                                        [
                                                        (0x0000000fU 
                                                         & (IData)(vlSelfRef.q))])));
                        vlSelfRef.sum = (0x0000ffffU 
                                         & (((IData)(vlSelfRef.sum) 
                                             + vlSelfRef.t17_lut__DOT__lut
                                             [vlSelfRef.d]) 
                                            + vlSelfRef.t17_lut__DOT__lut
                                            [vlSelfRef.q]));
                        if ((1U & (~ (IData)(vlSelfRef.we)))) {
                            __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyVal__t17_lut__DOT__small__v0 
                                = vlSelfRef.a;
                            __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyDim0__t17_lut__DOT__small__v0 
                                = (0x0000000fU & (IData)(vlSelfRef.d));
                            __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlySet__t17_lut__DOT__small__v0 = 1U;
                        }
                        vlSelfRef.q = __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___Vdly__q;
                        if (__Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlySet__t17_lut__DOT__lut__v0) {
                            vlSelfRef.t17_lut__DOT__lut[__Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyDim0__t17_lut__DOT__lut__v0] 
                                = __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyVal__t17_lut__DOT__lut__v0;
                        }
                        if (__Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlySet__t17_lut__DOT__small__v0) {
                            vlSelfRef.t17_lut__DOT__small[__Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyDim0__t17_lut__DOT__small__v0] 
                                = __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyVal__t17_lut__DOT__small__v0;
                        }
                        vlSelfRef.nib = (vlSelfRef.t17_lut__DOT__small
                                         [(0x0000000fU 
                                           & (IData)(vlSelfRef.a))] 
                                         ^ vlSelfRef.t17_lut__DOT__small
                                         [(0x0000000fU 
                                           & ((IData)(vlSelfRef.d) 
                                              >> 4U))]);
                    }
                }
            }
            Vt17_lut___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vt17_lut___024root___eval(Vt17_lut___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        IData/*31:0*/ __VicoIterCount;
        IData/*31:0*/ __VnbaIterCount;
        // Body
        __VicoIterCount = 0U;
        do {
            if (VL_UNLIKELY(((0x00002710U < __VicoIterCount)))) {
#ifdef VL_DEBUG
                Vt17_lut___024root___dump_triggers__ico(vlSelfRef.__VicoTriggered, "ico"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t17_lut.v", 5, "", "DIDNOTCONVERGE: Input combinational region did not converge after '--converge-limit' of 10000 tries");
            }
            __VicoIterCount = ((IData)(1U) + __VicoIterCount);
            vlSelfRef.__VicoPhaseResult = Vt17_lut___024root___eval_phase__ico(vlSelf, vlCtx);
        } while (vlSelfRef.__VicoPhaseResult);
        __VnbaIterCount = 0U;
        do {
            if (VL_UNLIKELY(((0x00002710U < __VnbaIterCount)))) {
#ifdef VL_DEBUG
                Vt17_lut___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t17_lut.v", 5, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vt17_lut___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t17_lut.v", 5, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vt17_lut___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vt17_lut___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vt17_lut___024root___ctor_var_reset(Vt17_lut___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        vlSelf->we = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 10105644630884274164ull);
        vlSelf->a = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 510903276987443985ull);
        vlSelf->d = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 1720370409040345145ull);
        vlSelf->q = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 8861071527689086543ull);
        vlSelf->sum = VL_SCOPED_RAND_RESET_I(16, __VscopeHash, 17823321413984766096ull);
        vlSelf->nib = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 6260273604009582164ull);
        for (int __Vi0 = 0; __Vi0 < 256; ++__Vi0) {
            vlSelf->t17_lut__DOT__lut[__Vi0] = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 3064526521071891906ull);
        }
        for (int __Vi0 = 0; __Vi0 < 16; ++__Vi0) {
            vlSelf->t17_lut__DOT__small[__Vi0] = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 1984862242711935834ull);
        }
        for (int __Vi0 = 0; __Vi0 < 1; ++__Vi0) {
            vlSelf->__VstlTriggered[__Vi0] = 0;
        }
        for (int __Vi0 = 0; __Vi0 < 2; ++__Vi0) {
            vlSelf->__VicoTriggered[__Vi0] = 0;
        }
        vlSelf->__Vtrigprevexpr___TOP__clk__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__we__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__a__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__d__0 = 0;
        vlSelf->__VicoDidInit = 0;
        for (int __Vi0 = 0; __Vi0 < 1; ++__Vi0) {
            vlSelf->__VactTriggered[__Vi0] = 0;
        }
        vlSelf->__Vtrigprevexpr___TOP__clk__1 = 0;
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
        Vt17_lut___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt17_lut___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt17_lut___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt17_lut___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vt17_lut___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt17_lut___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt17_lut__cuda::info() {
    typedef Vt17_lut__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"we", "we", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_we},
        {"a", "a", VLCU_MEMBER_IN, 1, 0, 1, 8, 7, 0, T::OFF_a},
        {"d", "d", VLCU_MEMBER_IN, 1, 0, 1, 8, 7, 0, T::OFF_d},
        {"q", "q", VLCU_MEMBER_OUT, 1, 0, 1, 8, 7, 0, T::OFF_q},
        {"sum", "sum", VLCU_MEMBER_OUT, 2, 0, 1, 16, 15, 0, T::OFF_sum},
        {"nib", "nib", VLCU_MEMBER_OUT, 1, 0, 1, 8, 7, 0, T::OFF_nib},
        {"__VstlFirstIteration", "__VstlFirstIteration", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlFirstIteration},
        {"__VstlPhaseResult", "__VstlPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__Vtrigprevexpr___TOP__we__0", "__Vtrigprevexpr___TOP__we__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__we__0},
        {"__Vtrigprevexpr___TOP__a__0", "__Vtrigprevexpr___TOP__a__0", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF___Vtrigprevexpr___TOP__a__0},
        {"__Vtrigprevexpr___TOP__d__0", "__Vtrigprevexpr___TOP__d__0", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF___Vtrigprevexpr___TOP__d__0},
        {"__VicoDidInit", "__VicoDidInit", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VicoDidInit},
        {"__VicoPhaseResult", "__VicoPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VicoPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__1", "__Vtrigprevexpr___TOP__clk__1", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__1},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"t17_lut__DOT__lut", "t17_lut.lut", VLCU_MEMBER_STATE, 1, 0, 256, 8, 7, 0, T::OFF_t17_lut__DOT__lut},
        {"t17_lut__DOT__small", "t17_lut.small", VLCU_MEMBER_STATE, 1, 0, 16, 8, 7, 0, T::OFF_t17_lut__DOT__small},
        {"__VstlTriggered", "__VstlTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VstlTriggered},
        {"__VicoTriggered", "__VicoTriggered", VLCU_MEMBER_STATE, 8, 0, 2, 64, 63, 0, T::OFF___VicoTriggered},
        {"__VactTriggered", "__VactTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VactTriggered},
        {"__VnbaTriggered", "__VnbaTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VnbaTriggered},
    };
    static const char* const messages[] = {nullptr};
    static const char* const plusargs[] = {nullptr};
    static const VlcuTraceDecl traceDecls[] = {
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0}};  // (terminator, not counted)
    static const uint32_t traceOrder[] = {0};  // codes in dump order (+ terminator)
    static const uint32_t saveOrder[] = {0, 1, 2, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 5, 18, 19, 20, 21, 22, 23, 24, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "t17_lut", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 0, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 25, 0, 0x494ebbadaed0dae4ULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt17_lut__cuda VLCU_TESTBENCHES);
