// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt02_fsm"
#endif

// Constant pool
static __device__ const VlWide<256>/*8191:0*/ Vt02_fsm__ConstPool__CONST_h4a7ef657_0 = VlWide<256>{{
        0x00000000, 0x00000001, 0x00000002, 0x00000002,
        0x00000003, 0x00000003, 0x00000003, 0x00000003,
        0x00000004, 0x00000004, 0x00000004, 0x00000004,
        0x00000004, 0x00000004, 0x00000004, 0x00000004,
        0x00000005, 0x00000005, 0x00000005, 0x00000005,
        0x00000005, 0x00000005, 0x00000005, 0x00000005,
        0x00000005, 0x00000005, 0x00000005, 0x00000005,
        0x00000005, 0x00000005, 0x00000005, 0x00000005,
        0x00000006, 0x00000006, 0x00000006, 0x00000006,
        0x00000006, 0x00000006, 0x00000006, 0x00000006,
        0x00000006, 0x00000006, 0x00000006, 0x00000006,
        0x00000006, 0x00000006, 0x00000006, 0x00000006,
        0x00000006, 0x00000006, 0x00000006, 0x00000006,
        0x00000006, 0x00000006, 0x00000006, 0x00000006,
        0x00000006, 0x00000006, 0x00000006, 0x00000006,
        0x00000006, 0x00000006, 0x00000006, 0x00000006,
        0x00000007, 0x00000007, 0x00000007, 0x00000007,
        0x00000007, 0x00000007, 0x00000007, 0x00000007,
        0x00000007, 0x00000007, 0x00000007, 0x00000007,
        0x00000007, 0x00000007, 0x00000007, 0x00000007,
        0x00000007, 0x00000007, 0x00000007, 0x00000007,
        0x00000007, 0x00000007, 0x00000007, 0x00000007,
        0x00000007, 0x00000007, 0x00000007, 0x00000007,
        0x00000007, 0x00000007, 0x00000007, 0x00000007,
        0x00000007, 0x00000007, 0x00000007, 0x00000007,
        0x00000007, 0x00000007, 0x00000007, 0x00000007,
        0x00000007, 0x00000007, 0x00000007, 0x00000007,
        0x00000007, 0x00000007, 0x00000007, 0x00000007,
        0x00000007, 0x00000007, 0x00000007, 0x00000007,
        0x00000007, 0x00000007, 0x00000007, 0x00000007,
        0x00000007, 0x00000007, 0x00000007, 0x00000007,
        0x00000007, 0x00000007, 0x00000007, 0x00000007,
        0x00000008, 0x00000008, 0x00000008, 0x00000008,
        0x00000008, 0x00000008, 0x00000008, 0x00000008,
        0x00000008, 0x00000008, 0x00000008, 0x00000008,
        0x00000008, 0x00000008, 0x00000008, 0x00000008,
        0x00000008, 0x00000008, 0x00000008, 0x00000008,
        0x00000008, 0x00000008, 0x00000008, 0x00000008,
        0x00000008, 0x00000008, 0x00000008, 0x00000008,
        0x00000008, 0x00000008, 0x00000008, 0x00000008,
        0x00000008, 0x00000008, 0x00000008, 0x00000008,
        0x00000008, 0x00000008, 0x00000008, 0x00000008,
        0x00000008, 0x00000008, 0x00000008, 0x00000008,
        0x00000008, 0x00000008, 0x00000008, 0x00000008,
        0x00000008, 0x00000008, 0x00000008, 0x00000008,
        0x00000008, 0x00000008, 0x00000008, 0x00000008,
        0x00000008, 0x00000008, 0x00000008, 0x00000008,
        0x00000008, 0x00000008, 0x00000008, 0x00000008,
        0x00000008, 0x00000008, 0x00000008, 0x00000008,
        0x00000008, 0x00000008, 0x00000008, 0x00000008,
        0x00000008, 0x00000008, 0x00000008, 0x00000008,
        0x00000008, 0x00000008, 0x00000008, 0x00000008,
        0x00000008, 0x00000008, 0x00000008, 0x00000008,
        0x00000008, 0x00000008, 0x00000008, 0x00000008,
        0x00000008, 0x00000008, 0x00000008, 0x00000008,
        0x00000008, 0x00000008, 0x00000008, 0x00000008,
        0x00000008, 0x00000008, 0x00000008, 0x00000008,
        0x00000008, 0x00000008, 0x00000008, 0x00000008,
        0x00000008, 0x00000008, 0x00000008, 0x00000008,
        0x00000008, 0x00000008, 0x00000008, 0x00000008,
        0x00000008, 0x00000008, 0x00000008, 0x00000008,
        0x00000008, 0x00000008, 0x00000008, 0x00000008,
        0x00000008, 0x00000008, 0x00000008, 0x00000008,
        0x00000008, 0x00000008, 0x00000008, 0x00000008
}};
static __device__ const VlWide<16>/*511:0*/ Vt02_fsm__ConstPool__CONST_h7bccbeed_0 = VlWide<16>{{
        0x00000063, 0x0000007c, 0x00000077, 0x0000007b,
        0x000000f2, 0x0000006b, 0x0000006f, 0x000000c5,
        0x00000030, 0x00000001, 0x00000067, 0x0000002b,
        0x000000fe, 0x000000d7, 0x000000ab, 0x00000076
}};

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt02_fsm___024root final {
    CData/*0:0*/ clk;
    CData/*0:0*/ rst;
    CData/*7:0*/ din;
    CData/*0:0*/ go;
    CData/*2:0*/ state;
    CData/*7:0*/ acc;
    CData/*7:0*/ lut;
    CData/*3:0*/ prio;
    SData/*15:0*/ count;
    CData/*7:0*/ t02_fsm__DOT__tbl;
    CData/*0:0*/ __VstlFirstIteration;
    CData/*0:0*/ __VstlPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*0:0*/ __Vtrigprevexpr___TOP__rst__0;
    CData/*7:0*/ __Vtrigprevexpr___TOP__din__0;
    CData/*0:0*/ __Vtrigprevexpr___TOP__go__0;
    CData/*0:0*/ __VicoDidInit;
    CData/*0:0*/ __VicoPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__1;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    IData/*31:0*/ __VactIterCount;
    VlUnpacked<QData/*63:0*/, 1> __VstlTriggered;
    VlUnpacked<QData/*63:0*/, 2> __VicoTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vt02_fsm__cuda final {
    typedef Vt02_fsm___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 46,
        OFF_rst = 47,
        OFF_din = 48,
        OFF_go = 49,
        OFF_state = 50,
        OFF_acc = 51,
        OFF_lut = 52,
        OFF_prio = 53,
        OFF_count = 44,
        OFF_t02_fsm__DOT__tbl = 54,
        OFF___VstlFirstIteration = 55,
        OFF___VstlPhaseResult = 56,
        OFF___Vtrigprevexpr___TOP__clk__0 = 57,
        OFF___Vtrigprevexpr___TOP__rst__0 = 58,
        OFF___Vtrigprevexpr___TOP__din__0 = 59,
        OFF___Vtrigprevexpr___TOP__go__0 = 60,
        OFF___VicoDidInit = 61,
        OFF___VicoPhaseResult = 62,
        OFF___Vtrigprevexpr___TOP__clk__1 = 63,
        OFF___VactPhaseResult = 64,
        OFF___VnbaPhaseResult = 65,
        OFF___VactIterCount = 40,
        OFF___VstlTriggered = 0,
        OFF___VicoTriggered = 8,
        OFF___VactTriggered = 24,
        OFF___VnbaTriggered = 32,
        kStateBytes = 72,  // S_persist
        kTouchedReadBytes = 54,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 52,  // S_state_wr
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
        s.rst = st.ld<VlcuMem8>(OFF_rst, i);
        s.din = st.ld<VlcuMem8>(OFF_din, i);
        s.go = st.ld<VlcuMem8>(OFF_go, i);
        s.state = st.ld<VlcuMem8>(OFF_state, i);
        s.acc = st.ld<VlcuMem8>(OFF_acc, i);
        s.lut = st.ld<VlcuMem8>(OFF_lut, i);
        s.prio = st.ld<VlcuMem8>(OFF_prio, i);
        s.count = st.ld<VlcuMem16>(OFF_count, i);
        s.t02_fsm__DOT__tbl = st.ld<VlcuMem8>(OFF_t02_fsm__DOT__tbl, i);
        s.__VstlFirstIteration = st.ld<VlcuMem8>(OFF___VstlFirstIteration, i);
        s.__VstlPhaseResult = st.ld<VlcuMem8>(OFF___VstlPhaseResult, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__Vtrigprevexpr___TOP__rst__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__rst__0, i);
        s.__Vtrigprevexpr___TOP__din__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__din__0, i);
        s.__Vtrigprevexpr___TOP__go__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__go__0, i);
        s.__VicoDidInit = st.ld<VlcuMem8>(OFF___VicoDidInit, i);
        s.__VicoPhaseResult = st.ld<VlcuMem8>(OFF___VicoPhaseResult, i);
        s.__Vtrigprevexpr___TOP__clk__1 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__1, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VstlTriggered[e] = st.ldE<QData>(OFF___VstlTriggered, e, i);
        for (uint32_t e = 0; e < 2; ++e) s.__VicoTriggered[e] = st.ldE<QData>(OFF___VicoTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.st<VlcuMem8>(OFF_rst, i, s.rst);
        st.st<VlcuMem8>(OFF_din, i, s.din);
        st.st<VlcuMem8>(OFF_go, i, s.go);
        st.st<VlcuMem8>(OFF_state, i, s.state);
        st.st<VlcuMem8>(OFF_acc, i, s.acc);
        st.st<VlcuMem8>(OFF_lut, i, s.lut);
        st.st<VlcuMem8>(OFF_prio, i, s.prio);
        st.st<VlcuMem16>(OFF_count, i, s.count);
        st.st<VlcuMem8>(OFF_t02_fsm__DOT__tbl, i, s.t02_fsm__DOT__tbl);
        st.st<VlcuMem8>(OFF___VstlFirstIteration, i, s.__VstlFirstIteration);
        st.st<VlcuMem8>(OFF___VstlPhaseResult, i, s.__VstlPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__rst__0, i, s.__Vtrigprevexpr___TOP__rst__0);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__din__0, i, s.__Vtrigprevexpr___TOP__din__0);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__go__0, i, s.__Vtrigprevexpr___TOP__go__0);
        st.st<VlcuMem8>(OFF___VicoDidInit, i, s.__VicoDidInit);
        st.st<VlcuMem8>(OFF___VicoPhaseResult, i, s.__VicoPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__1, i, s.__Vtrigprevexpr___TOP__clk__1);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VstlTriggered, e, i, s.__VstlTriggered[e]);
        for (uint32_t e = 0; e < 2; ++e) st.stE<QData>(OFF___VicoTriggered, e, i, s.__VicoTriggered[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VactTriggered, e, i, s.__VactTriggered[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VnbaTriggered, e, i, s.__VnbaTriggered[e]);
    }
    VLCU_D static void load_inputs_row(Root& s, const uint32_t* rowp, uint64_t stride, uint64_t i) {
        s.clk = static_cast<VlcuMem8>(rowp[0 * stride + i] & 0x1U);
        s.rst = static_cast<VlcuMem8>(rowp[1 * stride + i] & 0x1U);
        s.din = static_cast<VlcuMem8>(rowp[2 * stride + i] & 0xffU);
        s.go = static_cast<VlcuMem8>(rowp[3 * stride + i] & 0x1U);
    }
    VLCU_D static void store_outputs_row(const Root& s, uint32_t* rowp, uint64_t stride, uint64_t i) {
        rowp[0 * stride + i] = s.state;
        rowp[1 * stride + i] = s.acc;
        rowp[2 * stride + i] = s.lut;
        rowp[3 * stride + i] = s.prio;
        rowp[4 * stride + i] = s.count;
    }
    VLCU_D static uint64_t outputs_sig(const Root& s, uint64_t sig) {
        uint32_t acc = VLCU_SIG_ACC_INIT;
        acc = vlcu_sig_word(acc, s.state);
        acc = vlcu_sig_word(acc, s.acc);
        acc = vlcu_sig_word(acc, s.lut);
        acc = vlcu_sig_word(acc, s.prio);
        acc = vlcu_sig_word(acc, s.count);
        return vlcu_sig_mix(sig, acc);
    }
    enum : uint32_t { kInWords = 4, kOutWords = 5, kInBytes = 4, kOutBytes = 6, kTraceWords = 0 };

    VLCU_D static void Vt02_fsm___024root___eval_static(Vt02_fsm___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
        vlSelfRef.__Vtrigprevexpr___TOP__rst__0 = vlSelfRef.rst;
        vlSelfRef.__Vtrigprevexpr___TOP__din__0 = vlSelfRef.din;
        vlSelfRef.__Vtrigprevexpr___TOP__go__0 = vlSelfRef.go;
        vlSelfRef.__Vtrigprevexpr___TOP__clk__1 = vlSelfRef.clk;
    }

    VLCU_D static void Vt02_fsm___024root___eval_initial(Vt02_fsm___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt02_fsm___024root___eval_final(Vt02_fsm___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt02_fsm___024root___eval_settle(Vt02_fsm___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt02_fsm___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t02_fsm.v", 3, "", "DIDNOTCONVERGE: Settle region did not converge after '--converge-limit' of 10000 tries");
            }
            __VstlIterCount = ((IData)(1U) + __VstlIterCount);
            vlSelfRef.__VstlPhaseResult = Vt02_fsm___024root___eval_phase__stl(vlSelf, vlCtx);
            vlSelfRef.__VstlFirstIteration = 0U;
        } while (vlSelfRef.__VstlPhaseResult);
    }

    VLCU_D static bool Vt02_fsm___024root___trigger_anySet__stl(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt02_fsm___024root___eval_phase__stl(Vt02_fsm___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt02_fsm___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
        }
#endif
        __VstlExecute = Vt02_fsm___024root___trigger_anySet__stl(vlSelfRef.__VstlTriggered);
        if (__VstlExecute) {
            {
                // Inlined CFunc: _eval_stl
                if ((1ULL & vlSelfRef.__VstlTriggered[0U])) {
                    {
                        // Inlined CFunc: _ico_sequent__TOP__0
                        vlSelfRef.prio = (0x0000000fU 
                                          & Vt02_fsm__ConstPool__CONST_h4a7ef657_0
                                          [(0x07ffffffU 
                                            & (IData)(vlSelfRef.din))]);
                        vlSelfRef.t02_fsm__DOT__tbl 
                            = (0x000000ffU & Vt02_fsm__ConstPool__CONST_h7bccbeed_0
                               [(0x0000000fU & (IData)(vlSelfRef.din))]);
                        vlSelfRef.lut = ((IData)(vlSelfRef.acc) 
                                         ^ (IData)(vlSelfRef.t02_fsm__DOT__tbl));
                    }
                }
            }
        }
        return (__VstlExecute);
    }

    VLCU_D static bool Vt02_fsm___024root___trigger_anySet__ico(const VlUnpacked<QData/*63:0*/, 2> &in) {
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

    VLCU_D static bool Vt02_fsm___024root___eval_phase__ico(Vt02_fsm___024root* vlSelf, VlcuContext& vlCtx) {
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
                                                            (((((IData)(vlSelfRef.go) 
                                                                != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__go__0)) 
                                                               << 3U) 
                                                              | (((IData)(vlSelfRef.din) 
                                                                  != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__din__0)) 
                                                                 << 2U)) 
                                                             | ((((IData)(vlSelfRef.rst) 
                                                                  != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__rst__0)) 
                                                                 << 1U) 
                                                                | ((IData)(vlSelfRef.clk) 
                                                                   != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__clk__0))))));
            vlSelfRef.__Vtrigprevexpr___TOP__clk__0 
                = vlSelfRef.clk;
            vlSelfRef.__Vtrigprevexpr___TOP__rst__0 
                = vlSelfRef.rst;
            vlSelfRef.__Vtrigprevexpr___TOP__din__0 
                = vlSelfRef.din;
            vlSelfRef.__Vtrigprevexpr___TOP__go__0 
                = vlSelfRef.go;
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
            Vt02_fsm___024root___dump_triggers__ico(vlSelfRef.__VicoTriggered, "ico"s);
        }
#endif
        __VicoExecute = Vt02_fsm___024root___trigger_anySet__ico(vlSelfRef.__VicoTriggered);
        if (__VicoExecute) {
            {
                // Inlined CFunc: _eval_ico
                if ((4ULL & vlSelfRef.__VicoTriggered[0U])) {
                    {
                        // Inlined CFunc: _ico_sequent__TOP__0
                        vlSelfRef.prio = (0x0000000fU 
                                          & Vt02_fsm__ConstPool__CONST_h4a7ef657_0
                                          [(0x07ffffffU 
                                            & (IData)(vlSelfRef.din))]);
                        vlSelfRef.t02_fsm__DOT__tbl 
                            = (0x000000ffU & Vt02_fsm__ConstPool__CONST_h7bccbeed_0
                               [(0x0000000fU & (IData)(vlSelfRef.din))]);
                        vlSelfRef.lut = ((IData)(vlSelfRef.acc) 
                                         ^ (IData)(vlSelfRef.t02_fsm__DOT__tbl));
                    }
                }
            }
        }
        return (__VicoExecute);
    }

    VLCU_D static bool Vt02_fsm___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt02_fsm___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt02_fsm___024root___eval_phase__act(Vt02_fsm___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt02_fsm___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vt02_fsm___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vt02_fsm___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vt02_fsm___024root___eval_phase__nba(Vt02_fsm___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vt02_fsm___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    {
                        // Inlined CFunc: _nba_sequent__TOP__0
                        CData/*2:0*/ __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___Vdly__state;
                        __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___Vdly__state = 0;
                        CData/*7:0*/ __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___Vdly__acc;
                        __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___Vdly__acc = 0;
                        SData/*15:0*/ __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___Vdly__count;
                        __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___Vdly__count = 0;
                        __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___Vdly__state 
                            = vlSelfRef.state;
                        __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___Vdly__count 
                            = vlSelfRef.count;
                        __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___Vdly__acc 
                            = vlSelfRef.acc;
                        if (vlSelfRef.rst) {
                            __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___Vdly__state = 0U;
                            __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___Vdly__acc = 0U;
                            __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___Vdly__count = 0U;
                        } else {
                            __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___Vdly__count 
                                = (0x0000ffffU & ((IData)(1U) 
                                                  + (IData)(vlSelfRef.count)));
                            if ((4U & (IData)(vlSelfRef.state))) {
                                __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___Vdly__state = 0U;
                            } else if ((2U & (IData)(vlSelfRef.state))) {
                                if ((1U & (IData)(vlSelfRef.state))) {
                                    if (vlSelfRef.go) {
                                        __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___Vdly__state = 2U;
                                    }
                                } else {
                                    __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___Vdly__acc 
                                        = (0x000000ffU 
                                           & ((IData)(vlSelfRef.acc) 
                                              + (IData)(vlSelfRef.t02_fsm__DOT__tbl)));
                                    if ((IData)((((IData)(vlSelfRef.acc) 
                                                  >> 7U) 
                                                 & (~ (IData)(vlSelfRef.go))))) {
                                        __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___Vdly__state = 3U;
                                    } else if ((0x0fU 
                                                == 
                                                (0x0000000fU 
                                                 & (IData)(vlSelfRef.count)))) {
                                        __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___Vdly__state = 4U;
                                    }
                                }
                            } else if ((1U & (IData)(vlSelfRef.state))) {
                                __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___Vdly__acc 
                                    = vlSelfRef.din;
                                __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___Vdly__state = 2U;
                            } else if (vlSelfRef.go) {
                                __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___Vdly__state = 1U;
                            }
                        }
                        vlSelfRef.state = __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___Vdly__state;
                        vlSelfRef.count = __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___Vdly__count;
                        vlSelfRef.acc = __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___Vdly__acc;
                        vlSelfRef.lut = ((IData)(vlSelfRef.acc) 
                                         ^ (IData)(vlSelfRef.t02_fsm__DOT__tbl));
                    }
                }
            }
            Vt02_fsm___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vt02_fsm___024root___eval(Vt02_fsm___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt02_fsm___024root___dump_triggers__ico(vlSelfRef.__VicoTriggered, "ico"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t02_fsm.v", 3, "", "DIDNOTCONVERGE: Input combinational region did not converge after '--converge-limit' of 10000 tries");
            }
            __VicoIterCount = ((IData)(1U) + __VicoIterCount);
            vlSelfRef.__VicoPhaseResult = Vt02_fsm___024root___eval_phase__ico(vlSelf, vlCtx);
        } while (vlSelfRef.__VicoPhaseResult);
        __VnbaIterCount = 0U;
        do {
            if (VL_UNLIKELY(((0x00002710U < __VnbaIterCount)))) {
#ifdef VL_DEBUG
                Vt02_fsm___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t02_fsm.v", 3, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vt02_fsm___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t02_fsm.v", 3, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vt02_fsm___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vt02_fsm___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vt02_fsm___024root___ctor_var_reset(Vt02_fsm___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        vlSelf->rst = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 18209466448985614591ull);
        vlSelf->din = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 15192908731043726583ull);
        vlSelf->go = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 9942418676787815235ull);
        vlSelf->state = VL_SCOPED_RAND_RESET_I(3, __VscopeHash, 9404372463396948974ull);
        vlSelf->acc = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 14651699267163694450ull);
        vlSelf->lut = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 5263195267054218399ull);
        vlSelf->prio = VL_SCOPED_RAND_RESET_I(4, __VscopeHash, 13964444300711714917ull);
        vlSelf->count = VL_SCOPED_RAND_RESET_I(16, __VscopeHash, 10730142128712957955ull);
        vlSelf->t02_fsm__DOT__tbl = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 15257330042475601073ull);
        for (int __Vi0 = 0; __Vi0 < 1; ++__Vi0) {
            vlSelf->__VstlTriggered[__Vi0] = 0;
        }
        for (int __Vi0 = 0; __Vi0 < 2; ++__Vi0) {
            vlSelf->__VicoTriggered[__Vi0] = 0;
        }
        vlSelf->__Vtrigprevexpr___TOP__clk__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__rst__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__din__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__go__0 = 0;
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
        Vt02_fsm___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt02_fsm___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt02_fsm___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt02_fsm___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vt02_fsm___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt02_fsm___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt02_fsm__cuda::info() {
    typedef Vt02_fsm__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"rst", "rst", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_rst},
        {"din", "din", VLCU_MEMBER_IN, 1, 0, 1, 8, 7, 0, T::OFF_din},
        {"go", "go", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_go},
        {"state", "state", VLCU_MEMBER_OUT, 1, 0, 1, 3, 2, 0, T::OFF_state},
        {"acc", "acc", VLCU_MEMBER_OUT, 1, 0, 1, 8, 7, 0, T::OFF_acc},
        {"lut", "lut", VLCU_MEMBER_OUT, 1, 0, 1, 8, 7, 0, T::OFF_lut},
        {"prio", "prio", VLCU_MEMBER_OUT, 1, 0, 1, 4, 3, 0, T::OFF_prio},
        {"count", "count", VLCU_MEMBER_OUT, 2, 0, 1, 16, 15, 0, T::OFF_count},
        {"t02_fsm__DOT__tbl", "t02_fsm.tbl", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF_t02_fsm__DOT__tbl},
        {"__VstlFirstIteration", "__VstlFirstIteration", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlFirstIteration},
        {"__VstlPhaseResult", "__VstlPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__Vtrigprevexpr___TOP__rst__0", "__Vtrigprevexpr___TOP__rst__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__rst__0},
        {"__Vtrigprevexpr___TOP__din__0", "__Vtrigprevexpr___TOP__din__0", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF___Vtrigprevexpr___TOP__din__0},
        {"__Vtrigprevexpr___TOP__go__0", "__Vtrigprevexpr___TOP__go__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__go__0},
        {"__VicoDidInit", "__VicoDidInit", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VicoDidInit},
        {"__VicoPhaseResult", "__VicoPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VicoPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__1", "__Vtrigprevexpr___TOP__clk__1", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__1},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
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
    static const uint32_t saveOrder[] = {0, 1, 2, 3, 4, 5, 6, 7, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 8, 21, 22, 23, 24, 25, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "t02_fsm", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 0, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 26, 0, 0x3b73d8ee909aca3fULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt02_fsm__cuda VLCU_TESTBENCHES);
