// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt12_sys"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt12_sys___024root final {
    CData/*0:0*/ clk;
    CData/*7:0*/ a;
    CData/*0:0*/ oe;
    CData/*7:0*/ bus;
    IData/*31:0*/ rnd;
    QData/*63:0*/ tnow;
    SData/*15:0*/ romq;
    CData/*7:0*/ busq;
    CData/*0:0*/ __VstlFirstIteration;
    CData/*0:0*/ __VstlPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*7:0*/ __Vtrigprevexpr___TOP__a__0;
    CData/*0:0*/ __Vtrigprevexpr___TOP__oe__0;
    CData/*0:0*/ __VicoDidInit;
    CData/*0:0*/ __VicoPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__1;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    IData/*31:0*/ __VactIterCount;
    VlcuSharedArray<SData/*15:0*/, 32> t12_sys__DOT__rom;  // run-time indexed: shared memory [index][lane]
    VlUnpacked<QData/*63:0*/, 1> __VstlTriggered;
    VlUnpacked<QData/*63:0*/, 2> __VicoTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vt12_sys__cuda final {
    typedef Vt12_sys___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 122,
        OFF_a = 123,
        OFF_oe = 124,
        OFF_bus = 125,
        OFF_rnd = 48,
        OFF_tnow = 0,
        OFF_romq = 56,
        OFF_busq = 126,
        OFF___VstlFirstIteration = 127,
        OFF___VstlPhaseResult = 128,
        OFF___Vtrigprevexpr___TOP__clk__0 = 129,
        OFF___Vtrigprevexpr___TOP__a__0 = 130,
        OFF___Vtrigprevexpr___TOP__oe__0 = 131,
        OFF___VicoDidInit = 132,
        OFF___VicoPhaseResult = 133,
        OFF___Vtrigprevexpr___TOP__clk__1 = 134,
        OFF___VactPhaseResult = 135,
        OFF___VnbaPhaseResult = 136,
        OFF___VactIterCount = 52,
        OFF_t12_sys__DOT__rom = 58,
        OFF___VstlTriggered = 8,
        OFF___VicoTriggered = 16,
        OFF___VactTriggered = 32,
        OFF___VnbaTriggered = 40,
        kStateBytes = 144,  // S_persist
        kTouchedReadBytes = 50,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 60,  // S_state_wr
    };
    struct Locals final {
    };
    struct Shared final {
        SData/*15:0*/ t12_sys__DOT__rom[32][VLCU_BLOCK];
    };
    VLCU_D static void bind(Root& s, Locals& l, Shared& sh, const VlcuState& st, uint64_t i) {
        (void)l; (void)sh; (void)st; (void)i;
        s.vlNamep = "TOP";
        s.t12_sys__DOT__rom.bind(sh.t12_sys__DOT__rom, threadIdx.x);
    }
    VLCU_D static void load(Root& s, const VlcuState& st, uint64_t i) {
        s.clk = st.ld<VlcuMem8>(OFF_clk, i);
        s.a = st.ld<VlcuMem8>(OFF_a, i);
        s.oe = st.ld<VlcuMem8>(OFF_oe, i);
        s.bus = st.ld<VlcuMem8>(OFF_bus, i);
        s.rnd = st.ld<IData>(OFF_rnd, i);
        s.tnow = st.ld<QData>(OFF_tnow, i);
        s.romq = st.ld<VlcuMem16>(OFF_romq, i);
        s.busq = st.ld<VlcuMem8>(OFF_busq, i);
        s.__VstlFirstIteration = st.ld<VlcuMem8>(OFF___VstlFirstIteration, i);
        s.__VstlPhaseResult = st.ld<VlcuMem8>(OFF___VstlPhaseResult, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__Vtrigprevexpr___TOP__a__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__a__0, i);
        s.__Vtrigprevexpr___TOP__oe__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__oe__0, i);
        s.__VicoDidInit = st.ld<VlcuMem8>(OFF___VicoDidInit, i);
        s.__VicoPhaseResult = st.ld<VlcuMem8>(OFF___VicoPhaseResult, i);
        s.__Vtrigprevexpr___TOP__clk__1 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__1, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        for (uint32_t e = 0; e < 32; ++e) s.t12_sys__DOT__rom[e] = st.ldE<VlcuMem16>(OFF_t12_sys__DOT__rom, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VstlTriggered[e] = st.ldE<QData>(OFF___VstlTriggered, e, i);
        for (uint32_t e = 0; e < 2; ++e) s.__VicoTriggered[e] = st.ldE<QData>(OFF___VicoTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.st<VlcuMem8>(OFF_a, i, s.a);
        st.st<VlcuMem8>(OFF_oe, i, s.oe);
        st.st<VlcuMem8>(OFF_bus, i, s.bus);
        st.st<IData>(OFF_rnd, i, s.rnd);
        st.st<QData>(OFF_tnow, i, s.tnow);
        st.st<VlcuMem16>(OFF_romq, i, s.romq);
        st.st<VlcuMem8>(OFF_busq, i, s.busq);
        st.st<VlcuMem8>(OFF___VstlFirstIteration, i, s.__VstlFirstIteration);
        st.st<VlcuMem8>(OFF___VstlPhaseResult, i, s.__VstlPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__a__0, i, s.__Vtrigprevexpr___TOP__a__0);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__oe__0, i, s.__Vtrigprevexpr___TOP__oe__0);
        st.st<VlcuMem8>(OFF___VicoDidInit, i, s.__VicoDidInit);
        st.st<VlcuMem8>(OFF___VicoPhaseResult, i, s.__VicoPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__1, i, s.__Vtrigprevexpr___TOP__clk__1);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
        for (uint32_t e = 0; e < 32; ++e) st.stE<VlcuMem16>(OFF_t12_sys__DOT__rom, e, i, s.t12_sys__DOT__rom[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VstlTriggered, e, i, s.__VstlTriggered[e]);
        for (uint32_t e = 0; e < 2; ++e) st.stE<QData>(OFF___VicoTriggered, e, i, s.__VicoTriggered[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VactTriggered, e, i, s.__VactTriggered[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VnbaTriggered, e, i, s.__VnbaTriggered[e]);
    }
    VLCU_D static void load_inputs_row(Root& s, const uint32_t* rowp, uint64_t stride, uint64_t i) {
        s.clk = static_cast<VlcuMem8>(rowp[0 * stride + i] & 0x1U);
        s.a = static_cast<VlcuMem8>(rowp[1 * stride + i] & 0xffU);
        s.oe = static_cast<VlcuMem8>(rowp[2 * stride + i] & 0x1U);
        s.bus = static_cast<VlcuMem8>(rowp[3 * stride + i] & 0xffU);
    }
    VLCU_D static void store_outputs_row(const Root& s, uint32_t* rowp, uint64_t stride, uint64_t i) {
        rowp[0 * stride + i] = s.bus;
        rowp[1 * stride + i] = s.rnd;
        rowp[2 * stride + i] = static_cast<uint32_t>(s.tnow);
        rowp[3 * stride + i] = static_cast<uint32_t>(s.tnow >> 32);
        rowp[4 * stride + i] = s.romq;
        rowp[5 * stride + i] = s.busq;
    }
    VLCU_D static uint64_t outputs_sig(const Root& s, uint64_t sig) {
        uint32_t acc = VLCU_SIG_ACC_INIT;
        acc = vlcu_sig_word(acc, s.bus);
        acc = vlcu_sig_word(acc, s.rnd);
        acc = vlcu_sig_word(acc, static_cast<uint32_t>(s.tnow));
        acc = vlcu_sig_word(acc, static_cast<uint32_t>(s.tnow >> 32));
        acc = vlcu_sig_word(acc, s.romq);
        acc = vlcu_sig_word(acc, s.busq);
        return vlcu_sig_mix(sig, acc);
    }
    enum : uint32_t { kInWords = 4, kOutWords = 6, kInBytes = 4, kOutBytes = 16, kTraceWords = 0 };

    VLCU_D static void Vt12_sys___024root___eval_static(Vt12_sys___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
        vlSelfRef.__Vtrigprevexpr___TOP__a__0 = vlSelfRef.a;
        vlSelfRef.__Vtrigprevexpr___TOP__oe__0 = vlSelfRef.oe;
        vlSelfRef.__Vtrigprevexpr___TOP__clk__1 = vlSelfRef.clk;
    }

    VLCU_D static void Vt12_sys___024root___eval_initial(Vt12_sys___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        {
            // Inlined CFunc: _eval_initial__TOP
            vlSelfRef.t12_sys__DOT__rom[0U] = 0x1234U;
            vlSelfRef.t12_sys__DOT__rom[1U] = 0x1c92U;
            vlSelfRef.t12_sys__DOT__rom[2U] = 0x26f0U;
            vlSelfRef.t12_sys__DOT__rom[3U] = 0x314eU;
            vlSelfRef.t12_sys__DOT__rom[4U] = 0x3bacU;
            vlSelfRef.t12_sys__DOT__rom[5U] = 0x460aU;
            vlSelfRef.t12_sys__DOT__rom[6U] = 0x5068U;
            vlSelfRef.t12_sys__DOT__rom[7U] = 0x5ac6U;
            vlSelfRef.t12_sys__DOT__rom[8U] = 0x6524U;
            vlSelfRef.t12_sys__DOT__rom[9U] = 0x6f82U;
            vlSelfRef.t12_sys__DOT__rom[10U] = 0x79e0U;
            vlSelfRef.t12_sys__DOT__rom[11U] = 0x843eU;
            vlSelfRef.t12_sys__DOT__rom[12U] = 0x8e9cU;
            vlSelfRef.t12_sys__DOT__rom[13U] = 0x98faU;
            vlSelfRef.t12_sys__DOT__rom[14U] = 0xa358U;
            vlSelfRef.t12_sys__DOT__rom[15U] = 0xadb6U;
            vlSelfRef.t12_sys__DOT__rom[16U] = 0xb814U;
            vlSelfRef.t12_sys__DOT__rom[17U] = 0xc272U;
            vlSelfRef.t12_sys__DOT__rom[18U] = 0xccd0U;
            vlSelfRef.t12_sys__DOT__rom[19U] = 0xd72eU;
            vlSelfRef.t12_sys__DOT__rom[20U] = 0xe18cU;
            vlSelfRef.t12_sys__DOT__rom[21U] = 0xebeaU;
            vlSelfRef.t12_sys__DOT__rom[22U] = 0xf648U;
            vlSelfRef.t12_sys__DOT__rom[23U] = 0x00a6U;
            vlSelfRef.t12_sys__DOT__rom[24U] = 0x0b04U;
            vlSelfRef.t12_sys__DOT__rom[25U] = 0x1562U;
            vlSelfRef.t12_sys__DOT__rom[26U] = 0x1fc0U;
            vlSelfRef.t12_sys__DOT__rom[27U] = 0x2a1eU;
            vlSelfRef.t12_sys__DOT__rom[28U] = 0x347cU;
            vlSelfRef.t12_sys__DOT__rom[29U] = 0x3edaU;
            vlSelfRef.t12_sys__DOT__rom[30U] = 0x4938U;
            vlSelfRef.t12_sys__DOT__rom[31U] = 0x5396U;
            vlSelfRef.rnd = 0U;
        }
    }

    VLCU_D static void Vt12_sys___024root___eval_final(Vt12_sys___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        {
            // Inlined CFunc: _eval_final__TOP
            vlCtx.writef(0U, 1U, vlcu_msg_arg(vlSelfRef.romq));
        }
    }

    VLCU_D static void Vt12_sys___024root___eval_settle(Vt12_sys___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt12_sys___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t12_sys.v", 4, "", "DIDNOTCONVERGE: Settle region did not converge after '--converge-limit' of 10000 tries");
            }
            __VstlIterCount = ((IData)(1U) + __VstlIterCount);
            vlSelfRef.__VstlPhaseResult = Vt12_sys___024root___eval_phase__stl(vlSelf, vlCtx);
            vlSelfRef.__VstlFirstIteration = 0U;
        } while (vlSelfRef.__VstlPhaseResult);
    }

    VLCU_D static bool Vt12_sys___024root___trigger_anySet__stl(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt12_sys___024root___eval_phase__stl(Vt12_sys___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt12_sys___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
        }
#endif
        __VstlExecute = Vt12_sys___024root___trigger_anySet__stl(vlSelfRef.__VstlTriggered);
        if (__VstlExecute) {
            {
                // Inlined CFunc: _eval_stl
                if ((1ULL & vlSelfRef.__VstlTriggered[0U])) {
                    {
                        // Inlined CFunc: _ico_comb__TOP__0
                        vlSelfRef.bus = ((0x5aU ^ (IData)(vlSelfRef.a)) 
                                         & (- (IData)((IData)(vlSelfRef.oe))));
                    }
                }
            }
        }
        return (__VstlExecute);
    }

    VLCU_D static bool Vt12_sys___024root___trigger_anySet__ico(const VlUnpacked<QData/*63:0*/, 2> &in) {
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

    VLCU_D static bool Vt12_sys___024root___eval_phase__ico(Vt12_sys___024root* vlSelf, VlcuContext& vlCtx) {
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
                                                            ((((IData)(vlSelfRef.oe) 
                                                               != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__oe__0)) 
                                                              << 2U) 
                                                             | ((((IData)(vlSelfRef.a) 
                                                                  != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__a__0)) 
                                                                 << 1U) 
                                                                | ((IData)(vlSelfRef.clk) 
                                                                   != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__clk__0))))));
            vlSelfRef.__Vtrigprevexpr___TOP__clk__0 
                = vlSelfRef.clk;
            vlSelfRef.__Vtrigprevexpr___TOP__a__0 = vlSelfRef.a;
            vlSelfRef.__Vtrigprevexpr___TOP__oe__0 
                = vlSelfRef.oe;
            if (VL_UNLIKELY(((1U & (~ (IData)(vlSelfRef.__VicoDidInit)))))) {
                vlSelfRef.__VicoDidInit = 1U;
                vlSelfRef.__VicoTriggered[0U] = (1ULL 
                                                 | vlSelfRef.__VicoTriggered[0U]);
                vlSelfRef.__VicoTriggered[0U] = (2ULL 
                                                 | vlSelfRef.__VicoTriggered[0U]);
                vlSelfRef.__VicoTriggered[0U] = (4ULL 
                                                 | vlSelfRef.__VicoTriggered[0U]);
            }
        }
#ifdef VL_DEBUG
        if (VL_UNLIKELY(vlSymsp->_vm_contextp__->debug())) {
            Vt12_sys___024root___dump_triggers__ico(vlSelfRef.__VicoTriggered, "ico"s);
        }
#endif
        __VicoExecute = Vt12_sys___024root___trigger_anySet__ico(vlSelfRef.__VicoTriggered);
        if (__VicoExecute) {
            {
                // Inlined CFunc: _eval_ico
                if ((6ULL & vlSelfRef.__VicoTriggered[0U])) {
                    {
                        // Inlined CFunc: _ico_comb__TOP__0
                        vlSelfRef.bus = ((0x5aU ^ (IData)(vlSelfRef.a)) 
                                         & (- (IData)((IData)(vlSelfRef.oe))));
                    }
                }
            }
        }
        return (__VicoExecute);
    }

    VLCU_D static bool Vt12_sys___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt12_sys___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt12_sys___024root___eval_phase__act(Vt12_sys___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt12_sys___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vt12_sys___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vt12_sys___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vt12_sys___024root___eval_phase__nba(Vt12_sys___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vt12_sys___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    {
                        // Inlined CFunc: _nba_sequent__TOP__0
                        if (VL_UNLIKELY((((0xffU == (IData)(vlSelfRef.a)) 
                                          & (IData)(vlSelfRef.oe))))) {
                            vlCtx.writef(1U, 2U, vlcu_msg_arg(VL_TIME_UNITED_Q(1)), vlcu_msg_arg(VL_TIME_UNITED_Q(1)));
                            VL_STOP_MT("../verilator_b200/csrc/designs/torture/t12_sys.v", 26, "");
                        }
                        if (VL_UNLIKELY((((0xfeU == (IData)(vlSelfRef.a)) 
                                          & (~ (IData)(vlSelfRef.oe)))))) {
                            vlCtx.writef(2U, 2U, vlcu_msg_arg(VL_TIME_UNITED_Q(1)), vlcu_msg_arg((IData)(vlSelfRef.a)));
                            VL_STOP_MT("../verilator_b200/csrc/designs/torture/t12_sys.v", 27, "", false);
                        }
                        vlSelfRef.rnd = VL_RANDOM_I();
                        vlSelfRef.tnow = VL_TIME_UNITED_Q(1);
                        vlSelfRef.romq = vlSelfRef.t12_sys__DOT__rom
                            [(0x0000001fU & (IData)(vlSelfRef.a))];
                        vlSelfRef.busq = vlSelfRef.bus;
                    }
                }
            }
            Vt12_sys___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vt12_sys___024root___eval(Vt12_sys___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt12_sys___024root___dump_triggers__ico(vlSelfRef.__VicoTriggered, "ico"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t12_sys.v", 4, "", "DIDNOTCONVERGE: Input combinational region did not converge after '--converge-limit' of 10000 tries");
            }
            __VicoIterCount = ((IData)(1U) + __VicoIterCount);
            vlSelfRef.__VicoPhaseResult = Vt12_sys___024root___eval_phase__ico(vlSelf, vlCtx);
        } while (vlSelfRef.__VicoPhaseResult);
        __VnbaIterCount = 0U;
        do {
            if (VL_UNLIKELY(((0x00002710U < __VnbaIterCount)))) {
#ifdef VL_DEBUG
                Vt12_sys___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t12_sys.v", 4, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vt12_sys___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t12_sys.v", 4, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vt12_sys___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vt12_sys___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vt12_sys___024root___ctor_var_reset(Vt12_sys___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        vlSelf->a = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 510903276987443985ull);
        vlSelf->oe = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 6065859282030310819ull);
        vlSelf->bus = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 16368316895068468989ull);
        vlSelf->rnd = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 17761107180159243168ull);
        vlSelf->tnow = VL_SCOPED_RAND_RESET_Q(64, __VscopeHash, 12686218706349747783ull);
        vlSelf->romq = VL_SCOPED_RAND_RESET_I(16, __VscopeHash, 365927244690536090ull);
        vlSelf->busq = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 10709550563425763757ull);
        for (int __Vi0 = 0; __Vi0 < 32; ++__Vi0) {
            vlSelf->t12_sys__DOT__rom[__Vi0] = VL_SCOPED_RAND_RESET_I(16, __VscopeHash, 14648671155743178627ull);
        }
        for (int __Vi0 = 0; __Vi0 < 1; ++__Vi0) {
            vlSelf->__VstlTriggered[__Vi0] = 0;
        }
        for (int __Vi0 = 0; __Vi0 < 2; ++__Vi0) {
            vlSelf->__VicoTriggered[__Vi0] = 0;
        }
        vlSelf->__Vtrigprevexpr___TOP__clk__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__a__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__oe__0 = 0;
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
        Vt12_sys___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt12_sys___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt12_sys___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt12_sys___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vt12_sys___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt12_sys___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt12_sys__cuda::info() {
    typedef Vt12_sys__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"a", "a", VLCU_MEMBER_IN, 1, 0, 1, 8, 7, 0, T::OFF_a},
        {"oe", "oe", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_oe},
        {"bus", "bus", VLCU_MEMBER_INOUT, 1, 0, 1, 8, 7, 0, T::OFF_bus},
        {"rnd", "rnd", VLCU_MEMBER_OUT, 4, 0, 1, 32, 31, 0, T::OFF_rnd},
        {"tnow", "tnow", VLCU_MEMBER_OUT, 8, 0, 1, 64, 63, 0, T::OFF_tnow},
        {"romq", "romq", VLCU_MEMBER_OUT, 2, 0, 1, 16, 15, 0, T::OFF_romq},
        {"busq", "busq", VLCU_MEMBER_OUT, 1, 0, 1, 8, 7, 0, T::OFF_busq},
        {"__VstlFirstIteration", "__VstlFirstIteration", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlFirstIteration},
        {"__VstlPhaseResult", "__VstlPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__Vtrigprevexpr___TOP__a__0", "__Vtrigprevexpr___TOP__a__0", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF___Vtrigprevexpr___TOP__a__0},
        {"__Vtrigprevexpr___TOP__oe__0", "__Vtrigprevexpr___TOP__oe__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__oe__0},
        {"__VicoDidInit", "__VicoDidInit", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VicoDidInit},
        {"__VicoPhaseResult", "__VicoPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VicoPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__1", "__Vtrigprevexpr___TOP__clk__1", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__1},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"t12_sys__DOT__rom", "t12_sys.rom", VLCU_MEMBER_STATE, 2, 0, 32, 16, 15, 0, T::OFF_t12_sys__DOT__rom},
        {"__VstlTriggered", "__VstlTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VstlTriggered},
        {"__VicoTriggered", "__VicoTriggered", VLCU_MEMBER_STATE, 8, 0, 2, 64, 63, 0, T::OFF___VicoTriggered},
        {"__VactTriggered", "__VactTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VactTriggered},
        {"__VnbaTriggered", "__VnbaTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VnbaTriggered},
    };
    static const char* const messages[] = {"done %d\n", "[%0t] %%Error: t12_sys.v:26: Assertion failed in TOP.t12_sys: a is all ones at %0t\n", "[%0t] %%Fatal: t12_sys.v:27: Assertion failed in TOP.t12_sys: fatal on %h\n", nullptr};
    static const char* const plusargs[] = {nullptr};
    static const VlcuTraceDecl traceDecls[] = {
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0}};  // (terminator, not counted)
    static const uint32_t traceOrder[] = {0};  // codes in dump order (+ terminator)
    static const uint32_t saveOrder[] = {0, 1, 2, 3, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 6, 4, 18, 5, 19, 20, 21, 22, 23, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "t12_sys", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 3, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 24, 0, 0xf1ec08fd9ee39637ULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt12_sys__cuda VLCU_TESTBENCHES);
