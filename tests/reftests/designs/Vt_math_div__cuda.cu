// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt_math_div"
#endif

// Constant pool
static __device__ const VlWide<8>/*255:0*/ Vt_math_div__ConstPool__CONST_h4313a09b_0 = VlWide<8>{{
        0x00f395b8, 0xad20a48e, 0xfd00ddbd, 0xc3e236dd,
        0x8002bd68, 0xa982646c, 0x00d5fe51, 0x0e17c88f
}};
static __device__ const VlWide<8>/*255:0*/ Vt_math_div__ConstPool__CONST_h6bf974be_0 = VlWide<8>{{
        0xc9f395b8, 0x0a48e039, 0xdddbdad2, 0xe236ddfd,
        0xe2bd68c3, 0x982646c8, 0x3d5fe51a, 0x0e17c88f
}};
static __device__ const VlWide<8>/*255:0*/ Vt_math_div__ConstPool__CONST_hc278fede_0 = VlWide<8>{{
        0x147e7f26, 0x7491f3da, 0x42b14853, 0x45e0eb76,
        0x2413d770, 0xde489bab, 0x6c843d35, 0xed388e64
}};

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt_math_div___024root final {
    CData/*0:0*/ clk;
    CData/*0:0*/ __VstlFirstIteration;
    CData/*0:0*/ __VstlPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    VlWide<8>/*255:0*/ t__DOT__a;
    IData/*31:0*/ t__DOT__cyc;
    IData/*31:0*/ __VactIterCount;
    QData/*60:0*/ t__DOT__divisor;
    QData/*60:0*/ t__DOT__qq;
    QData/*60:0*/ t__DOT__rq;
    QData/*60:0*/ t__DOT__qq4;
    QData/*60:0*/ t__DOT__rq4;
    QData/*60:0*/ t__DOT__qqs;
    QData/*60:0*/ t__DOT__rqs;
    VlUnpacked<QData/*63:0*/, 1> __VstlTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vt_math_div__cuda final {
    typedef Vt_math_div___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 120,
        OFF___VstlFirstIteration = 121,
        OFF___VstlPhaseResult = 122,
        OFF___Vtrigprevexpr___TOP__clk__0 = 123,
        OFF___VactPhaseResult = 124,
        OFF___VnbaPhaseResult = 125,
        OFF_t__DOT__a = 80,
        OFF_t__DOT__cyc = 112,
        OFF___VactIterCount = 116,
        OFF_t__DOT__divisor = 0,
        OFF_t__DOT__qq = 8,
        OFF_t__DOT__rq = 16,
        OFF_t__DOT__qq4 = 24,
        OFF_t__DOT__rq4 = 32,
        OFF_t__DOT__qqs = 40,
        OFF_t__DOT__rqs = 48,
        OFF___VstlTriggered = 56,
        OFF___VactTriggered = 64,
        OFF___VnbaTriggered = 72,
        kStateBytes = 128,  // S_persist
        kTouchedReadBytes = 116,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 115,  // S_state_wr
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
        st.ldW(OFF_t__DOT__a, i, s.t__DOT__a);
        s.t__DOT__cyc = st.ld<IData>(OFF_t__DOT__cyc, i);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        s.t__DOT__divisor = st.ld<QData>(OFF_t__DOT__divisor, i);
        s.t__DOT__qq = st.ld<QData>(OFF_t__DOT__qq, i);
        s.t__DOT__rq = st.ld<QData>(OFF_t__DOT__rq, i);
        s.t__DOT__qq4 = st.ld<QData>(OFF_t__DOT__qq4, i);
        s.t__DOT__rq4 = st.ld<QData>(OFF_t__DOT__rq4, i);
        s.t__DOT__qqs = st.ld<QData>(OFF_t__DOT__qqs, i);
        s.t__DOT__rqs = st.ld<QData>(OFF_t__DOT__rqs, i);
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
        st.stW(OFF_t__DOT__a, i, s.t__DOT__a);
        st.st<IData>(OFF_t__DOT__cyc, i, s.t__DOT__cyc);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
        st.st<QData>(OFF_t__DOT__divisor, i, s.t__DOT__divisor);
        st.st<QData>(OFF_t__DOT__qq, i, s.t__DOT__qq);
        st.st<QData>(OFF_t__DOT__rq, i, s.t__DOT__rq);
        st.st<QData>(OFF_t__DOT__qq4, i, s.t__DOT__qq4);
        st.st<QData>(OFF_t__DOT__rq4, i, s.t__DOT__rq4);
        st.st<QData>(OFF_t__DOT__qqs, i, s.t__DOT__qqs);
        st.st<QData>(OFF_t__DOT__rqs, i, s.t__DOT__rqs);
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

    VLCU_D static void Vt_math_div___024root___eval_static(Vt_math_div___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
    }

    VLCU_D static void Vt_math_div___024root___eval_initial(Vt_math_div___024root* vlSelf, VlcuContext& vlCtx) {
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

    VLCU_D static void Vt_math_div___024root___eval_final(Vt_math_div___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_math_div___024root___eval_settle(Vt_math_div___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt_math_div___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
#endif
                VL_FATAL_MT("/root/reference/test_regress/t/t_math_div.v", 7, "", "DIDNOTCONVERGE: Settle region did not converge after '--converge-limit' of 10000 tries");
            }
            __VstlIterCount = ((IData)(1U) + __VstlIterCount);
            vlSelfRef.__VstlPhaseResult = Vt_math_div___024root___eval_phase__stl(vlSelf, vlCtx);
            vlSelfRef.__VstlFirstIteration = 0U;
        } while (vlSelfRef.__VstlPhaseResult);
    }

    VLCU_D static bool Vt_math_div___024root___trigger_anySet__stl(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt_math_div___024root___eval_phase__stl(Vt_math_div___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt_math_div___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
        }
#endif
        __VstlExecute = Vt_math_div___024root___trigger_anySet__stl(vlSelfRef.__VstlTriggered);
        if (__VstlExecute) {
            {
                // Inlined CFunc: _eval_stl
                if ((1ULL & vlSelfRef.__VstlTriggered[0U])) {
                    {
                        // Inlined CFunc: _stl_sequent__TOP__0
                        vlSelfRef.t__DOT__qq4 = (0x07ffffffffffffffULL 
                                                 & ((0x3fffffffc0000000ULL 
                                                     & ((QData)((IData)(vlSelfRef.t__DOT__a[1U])) 
                                                        << 0x0000001eU)) 
                                                    | ((QData)((IData)(vlSelfRef.t__DOT__a[0U])) 
                                                       >> 2U)));
                        vlSelfRef.t__DOT__rq4 = (3ULL 
                                                 & (QData)((IData)(vlSelfRef.t__DOT__a[0U])));
                        vlSelfRef.t__DOT__qq = (0x1fffffffffffffffULL 
                                                & VL_DIV_QQQ(61, 
                                                             (0x1fffffffffffffffULL 
                                                              & (((QData)((IData)(vlSelfRef.t__DOT__a[1U])) 
                                                                  << 0x00000020U) 
                                                                 | (QData)((IData)(vlSelfRef.t__DOT__a[0U])))), vlSelfRef.t__DOT__divisor));
                        vlSelfRef.t__DOT__rq = (0x1fffffffffffffffULL 
                                                & VL_MODDIV_QQQ(61, 
                                                                (0x1fffffffffffffffULL 
                                                                 & (((QData)((IData)(vlSelfRef.t__DOT__a[1U])) 
                                                                     << 0x00000020U) 
                                                                    | (QData)((IData)(vlSelfRef.t__DOT__a[0U])))), vlSelfRef.t__DOT__divisor));
                        vlSelfRef.t__DOT__qqs = (0x1fffffffffffffffULL 
                                                 & VL_DIVS_QQQ(61, 
                                                               (0x1fffffffffffffffULL 
                                                                & (((QData)((IData)(vlSelfRef.t__DOT__a[1U])) 
                                                                    << 0x00000020U) 
                                                                   | (QData)((IData)(vlSelfRef.t__DOT__a[0U])))), vlSelfRef.t__DOT__divisor));
                        vlSelfRef.t__DOT__rqs = (0x1fffffffffffffffULL 
                                                 & VL_MODDIVS_QQQ(61, 
                                                                  (0x1fffffffffffffffULL 
                                                                   & (((QData)((IData)(vlSelfRef.t__DOT__a[1U])) 
                                                                       << 0x00000020U) 
                                                                      | (QData)((IData)(vlSelfRef.t__DOT__a[0U])))), vlSelfRef.t__DOT__divisor));
                    }
                }
            }
        }
        return (__VstlExecute);
    }

    VLCU_D static bool Vt_math_div___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt_math_div___024root___nba_sequent__TOP__0(Vt_math_div___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        IData/*31:0*/ __Vdly__t__DOT__cyc;
        __Vdly__t__DOT__cyc = 0;
        VlWide<8>/*255:0*/ __Vdly__t__DOT__a;
        VL_ZERO_W(256, __Vdly__t__DOT__a);
        // Body
        __Vdly__t__DOT__cyc = vlSelfRef.t__DOT__cyc;
        __Vdly__t__DOT__a[0U] = vlSelfRef.t__DOT__a[0U];
        __Vdly__t__DOT__a[1U] = vlSelfRef.t__DOT__a[1U];
        __Vdly__t__DOT__a[2U] = vlSelfRef.t__DOT__a[2U];
        __Vdly__t__DOT__a[3U] = vlSelfRef.t__DOT__a[3U];
        __Vdly__t__DOT__a[4U] = vlSelfRef.t__DOT__a[4U];
        __Vdly__t__DOT__a[5U] = vlSelfRef.t__DOT__a[5U];
        __Vdly__t__DOT__a[6U] = vlSelfRef.t__DOT__a[6U];
        __Vdly__t__DOT__a[7U] = vlSelfRef.t__DOT__a[7U];
        if ((0U != vlSelfRef.t__DOT__cyc)) {
            __Vdly__t__DOT__cyc = ((IData)(1U) + vlSelfRef.t__DOT__cyc);
            if ((1U == vlSelfRef.t__DOT__cyc)) {
                VL_ASSIGN_W(256, __Vdly__t__DOT__a, Vt_math_div__ConstPool__CONST_hc278fede_0);
                vlSelfRef.t__DOT__divisor = 0x0000000000012371ULL;
                __Vdly__t__DOT__a[1U] = (0xefffffffU 
                                         & __Vdly__t__DOT__a[1U]);
                vlSelfRef.t__DOT__divisor = (0x0fffffffffffffffULL 
                                             & vlSelfRef.t__DOT__divisor);
            }
            if (VL_LTS_III(32, 1U, vlSelfRef.t__DOT__cyc)) {
                if (VL_UNLIKELY(((vlSelfRef.t__DOT__qq4 
                                  != (0x07ffffffffffffffULL 
                                      & (((QData)((IData)(vlSelfRef.t__DOT__a[1U])) 
                                          << 0x0000001eU) 
                                         | ((QData)((IData)(vlSelfRef.t__DOT__a[0U])) 
                                            >> 2U))))))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_div.v", 49, "");
                }
                if (VL_UNLIKELY(((vlSelfRef.t__DOT__rq4 
                                  != (QData)((IData)(
                                                     (3U 
                                                      & vlSelfRef.t__DOT__a[0U]))))))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_div.v", 50, "");
                }
            }
            if ((2U == vlSelfRef.t__DOT__cyc)) {
                if (VL_UNLIKELY(((0x00000403ad81c0daULL 
                                  != vlSelfRef.t__DOT__qq)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_div.v", 57, "");
                }
                VL_ASSIGN_W(256, __Vdly__t__DOT__a, Vt_math_div__ConstPool__CONST_h6bf974be_0);
                vlSelfRef.t__DOT__divisor = 0x0000001238123771ULL;
                if (VL_UNLIKELY(((0x00000000000090ecULL 
                                  != vlSelfRef.t__DOT__rq)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_div.v", 58, "");
                }
                __Vdly__t__DOT__a[1U] = (0xefffffffU 
                                         & __Vdly__t__DOT__a[1U]);
                vlSelfRef.t__DOT__divisor = (0x0fffffffffffffffULL 
                                             & vlSelfRef.t__DOT__divisor);
                if (VL_UNLIKELY(((0x00000403ad81c0daULL 
                                  != vlSelfRef.t__DOT__qqs)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_div.v", 59, "");
                }
                if (VL_UNLIKELY(((0x00000000000090ecULL 
                                  != vlSelfRef.t__DOT__rqs)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_div.v", 60, "");
                }
                if (VL_UNLIKELY(((0x01247cf6851f9fc9ULL 
                                  != vlSelfRef.t__DOT__qq4)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_div.v", 61, "");
                }
                if (VL_UNLIKELY(((2ULL != vlSelfRef.t__DOT__rq4)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_div.v", 62, "");
                }
            }
            if ((3U == vlSelfRef.t__DOT__cyc)) {
                if (VL_UNLIKELY(((0x000000000090832eULL 
                                  != vlSelfRef.t__DOT__qq)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_div.v", 69, "");
                }
                VL_ASSIGN_W(256, __Vdly__t__DOT__a, Vt_math_div__ConstPool__CONST_h4313a09b_0);
                vlSelfRef.t__DOT__divisor = 0x0000000000000f1bULL;
                if (VL_UNLIKELY(((0x0000000334becc6aULL 
                                  != vlSelfRef.t__DOT__rq)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_div.v", 70, "");
                }
                __Vdly__t__DOT__a[1U] = (0x10000000U 
                                         | __Vdly__t__DOT__a[1U]);
                vlSelfRef.t__DOT__divisor = (0x0fffffffffffffffULL 
                                             & vlSelfRef.t__DOT__divisor);
                if (VL_UNLIKELY(((0x000000000090832eULL 
                                  != vlSelfRef.t__DOT__qqs)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_div.v", 71, "");
                }
                if (VL_UNLIKELY(((0x0000000334becc6aULL 
                                  != vlSelfRef.t__DOT__rqs)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_div.v", 72, "");
                }
                if (VL_UNLIKELY(((0x0292380e727ce56eULL 
                                  != vlSelfRef.t__DOT__qq4)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_div.v", 73, "");
                }
                if (VL_UNLIKELY(((0ULL != vlSelfRef.t__DOT__rq4)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_div.v", 74, "");
                }
            }
            if ((4U == vlSelfRef.t__DOT__cyc)) {
                if (VL_UNLIKELY(((0x0001eda37cca1be8ULL 
                                  != vlSelfRef.t__DOT__qq)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_div.v", 79, "");
                }
                __Vdly__t__DOT__a[1U] = (0xefffffffU 
                                         & __Vdly__t__DOT__a[1U]);
                vlSelfRef.t__DOT__divisor = (0x1000000000000000ULL 
                                             | vlSelfRef.t__DOT__divisor);
                if (VL_UNLIKELY(((0x0000000000000c40ULL 
                                  != vlSelfRef.t__DOT__rq)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_div.v", 80, "");
                }
                if (VL_UNLIKELY(((0x1fffcf5187c76510ULL 
                                  != vlSelfRef.t__DOT__qqs)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_div.v", 81, "");
                }
                if (VL_UNLIKELY(((0x1ffffffffffffd08ULL 
                                  != vlSelfRef.t__DOT__rqs)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_div.v", 82, "");
                }
                if (VL_UNLIKELY(((0x07482923803ce56eULL 
                                  != vlSelfRef.t__DOT__qq4)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_div.v", 83, "");
                }
                if (VL_UNLIKELY(((0ULL != vlSelfRef.t__DOT__rq4)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_div.v", 84, "");
                }
            }
            if ((5U == vlSelfRef.t__DOT__cyc)) {
                if (VL_UNLIKELY(((0ULL != vlSelfRef.t__DOT__qq)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_div.v", 89, "");
                }
                __Vdly__t__DOT__a[1U] = (0x10000000U 
                                         | __Vdly__t__DOT__a[1U]);
                vlSelfRef.t__DOT__divisor = (0x1000000000000000ULL 
                                             | vlSelfRef.t__DOT__divisor);
                if (VL_UNLIKELY(((0x0d20a48e00f395b8ULL 
                                  != vlSelfRef.t__DOT__rq)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_div.v", 90, "");
                }
                if (VL_UNLIKELY(((0ULL != vlSelfRef.t__DOT__qqs)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_div.v", 91, "");
                }
                if (VL_UNLIKELY(((0x0d20a48e00f395b8ULL 
                                  != vlSelfRef.t__DOT__rqs)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_div.v", 92, "");
                }
            }
            if ((6U == vlSelfRef.t__DOT__cyc)) {
                if (VL_UNLIKELY(((1ULL != vlSelfRef.t__DOT__qq)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_div.v", 95, "");
                }
                if (VL_UNLIKELY(((0x0d20a48e00f3869dULL 
                                  != vlSelfRef.t__DOT__rq)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_div.v", 96, "");
                }
                if (VL_UNLIKELY(((0ULL != vlSelfRef.t__DOT__qqs)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_div.v", 97, "");
                }
                if (VL_UNLIKELY(((0x1d20a48e00f395b8ULL 
                                  != vlSelfRef.t__DOT__rqs)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_div.v", 98, "");
                }
            }
            if ((9U == vlSelfRef.t__DOT__cyc)) {
                vlSelfRef.t__DOT__divisor = 0ULL;
            }
            if ((0x0000000aU == vlSelfRef.t__DOT__cyc)) {
                if (VL_UNLIKELY(((0ULL != vlSelfRef.t__DOT__qq)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_div.v", 106, "");
                }
                if (VL_UNLIKELY(((0ULL != vlSelfRef.t__DOT__rq)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_div.v", 107, "");
                }
            }
            if (VL_UNLIKELY(((0x00000013U == vlSelfRef.t__DOT__cyc)))) {
                vlCtx.writef(0U, 0U);
                VL_FINISH_MT("/root/reference/test_regress/t/t_math_div.v", 117, "");
            }
        }
        vlSelfRef.t__DOT__cyc = __Vdly__t__DOT__cyc;
        vlSelfRef.t__DOT__a[0U] = __Vdly__t__DOT__a[0U];
        vlSelfRef.t__DOT__a[1U] = __Vdly__t__DOT__a[1U];
        vlSelfRef.t__DOT__a[2U] = __Vdly__t__DOT__a[2U];
        vlSelfRef.t__DOT__a[3U] = __Vdly__t__DOT__a[3U];
        vlSelfRef.t__DOT__a[4U] = __Vdly__t__DOT__a[4U];
        vlSelfRef.t__DOT__a[5U] = __Vdly__t__DOT__a[5U];
        vlSelfRef.t__DOT__a[6U] = __Vdly__t__DOT__a[6U];
        vlSelfRef.t__DOT__a[7U] = __Vdly__t__DOT__a[7U];
        vlSelfRef.t__DOT__qq4 = (0x07ffffffffffffffULL 
                                 & ((0x3fffffffc0000000ULL 
                                     & ((QData)((IData)(vlSelfRef.t__DOT__a[1U])) 
                                        << 0x0000001eU)) 
                                    | ((QData)((IData)(vlSelfRef.t__DOT__a[0U])) 
                                       >> 2U)));
        vlSelfRef.t__DOT__rq4 = (3ULL & (QData)((IData)(vlSelfRef.t__DOT__a[0U])));
        vlSelfRef.t__DOT__qq = (0x1fffffffffffffffULL 
                                & VL_DIV_QQQ(61, (0x1fffffffffffffffULL 
                                                  & (((QData)((IData)(vlSelfRef.t__DOT__a[1U])) 
                                                      << 0x00000020U) 
                                                     | (QData)((IData)(vlSelfRef.t__DOT__a[0U])))), vlSelfRef.t__DOT__divisor));
        vlSelfRef.t__DOT__rq = (0x1fffffffffffffffULL 
                                & VL_MODDIV_QQQ(61, 
                                                (0x1fffffffffffffffULL 
                                                 & (((QData)((IData)(vlSelfRef.t__DOT__a[1U])) 
                                                     << 0x00000020U) 
                                                    | (QData)((IData)(vlSelfRef.t__DOT__a[0U])))), vlSelfRef.t__DOT__divisor));
        vlSelfRef.t__DOT__qqs = (0x1fffffffffffffffULL 
                                 & VL_DIVS_QQQ(61, 
                                               (0x1fffffffffffffffULL 
                                                & (((QData)((IData)(vlSelfRef.t__DOT__a[1U])) 
                                                    << 0x00000020U) 
                                                   | (QData)((IData)(vlSelfRef.t__DOT__a[0U])))), vlSelfRef.t__DOT__divisor));
        vlSelfRef.t__DOT__rqs = (0x1fffffffffffffffULL 
                                 & VL_MODDIVS_QQQ(61, 
                                                  (0x1fffffffffffffffULL 
                                                   & (((QData)((IData)(vlSelfRef.t__DOT__a[1U])) 
                                                       << 0x00000020U) 
                                                      | (QData)((IData)(vlSelfRef.t__DOT__a[0U])))), vlSelfRef.t__DOT__divisor));
    }

    VLCU_D static void Vt_math_div___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt_math_div___024root___eval_phase__act(Vt_math_div___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt_math_div___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vt_math_div___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vt_math_div___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vt_math_div___024root___eval_phase__nba(Vt_math_div___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vt_math_div___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    Vt_math_div___024root___nba_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
            Vt_math_div___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vt_math_div___024root___eval(Vt_math_div___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt_math_div___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("/root/reference/test_regress/t/t_math_div.v", 7, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vt_math_div___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("/root/reference/test_regress/t/t_math_div.v", 7, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vt_math_div___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vt_math_div___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vt_math_div___024root___ctor_var_reset(Vt_math_div___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        VL_SCOPED_RAND_RESET_W(256, vlSelf->t__DOT__a, __VscopeHash, 568261790826559771ull);
        vlSelf->t__DOT__divisor = VL_SCOPED_RAND_RESET_Q(61, __VscopeHash, 15002985049426033138ull);
        vlSelf->t__DOT__qq = VL_SCOPED_RAND_RESET_Q(61, __VscopeHash, 876090779370479156ull);
        vlSelf->t__DOT__rq = VL_SCOPED_RAND_RESET_Q(61, __VscopeHash, 6810423746990386560ull);
        vlSelf->t__DOT__qq4 = VL_SCOPED_RAND_RESET_Q(61, __VscopeHash, 5902011272748700838ull);
        vlSelf->t__DOT__rq4 = VL_SCOPED_RAND_RESET_Q(61, __VscopeHash, 14806090927500684041ull);
        vlSelf->t__DOT__qqs = VL_SCOPED_RAND_RESET_Q(61, __VscopeHash, 5701884365260607164ull);
        vlSelf->t__DOT__rqs = VL_SCOPED_RAND_RESET_Q(61, __VscopeHash, 15809558147883397135ull);
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
        Vt_math_div___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_div___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_div___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_div___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_div___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_div___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt_math_div__cuda::info() {
    typedef Vt_math_div__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"__VstlFirstIteration", "__VstlFirstIteration", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlFirstIteration},
        {"__VstlPhaseResult", "__VstlPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"t__DOT__a", "t.a", VLCU_MEMBER_STATE, 4, 0, 8, 256, 255, 0, T::OFF_t__DOT__a},
        {"t__DOT__cyc", "t.cyc", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__cyc},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"t__DOT__divisor", "t.divisor", VLCU_MEMBER_STATE, 8, 0, 1, 61, 60, 0, T::OFF_t__DOT__divisor},
        {"t__DOT__qq", "t.qq", VLCU_MEMBER_STATE, 8, 0, 1, 61, 60, 0, T::OFF_t__DOT__qq},
        {"t__DOT__rq", "t.rq", VLCU_MEMBER_STATE, 8, 0, 1, 61, 60, 0, T::OFF_t__DOT__rq},
        {"t__DOT__qq4", "t.qq4", VLCU_MEMBER_STATE, 8, 0, 1, 61, 60, 0, T::OFF_t__DOT__qq4},
        {"t__DOT__rq4", "t.rq4", VLCU_MEMBER_STATE, 8, 0, 1, 61, 60, 0, T::OFF_t__DOT__rq4},
        {"t__DOT__qqs", "t.qqs", VLCU_MEMBER_STATE, 8, 0, 1, 61, 60, 0, T::OFF_t__DOT__qqs},
        {"t__DOT__rqs", "t.rqs", VLCU_MEMBER_STATE, 8, 0, 1, 61, 60, 0, T::OFF_t__DOT__rqs},
        {"__VstlTriggered", "__VstlTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VstlTriggered},
        {"__VactTriggered", "__VactTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VactTriggered},
        {"__VnbaTriggered", "__VnbaTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VnbaTriggered},
    };
    static const char* const messages[] = {"*-* All Finished *-*\n", nullptr};
    static const char* const plusargs[] = {nullptr};
    static const VlcuTraceDecl traceDecls[] = {
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0}};  // (terminator, not counted)
    static const uint32_t traceOrder[] = {0};  // codes in dump order (+ terminator)
    static const uint32_t saveOrder[] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "t", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 1, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 19, 0, 0x841edae0554c77afULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt_math_div__cuda VLCU_TESTBENCHES);
