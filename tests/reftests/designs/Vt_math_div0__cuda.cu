// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

// 56 bytes of register-resident state: latency-bound, compile for full occupancy
#ifndef VLCU_MIN_BLOCKS
#define VLCU_MIN_BLOCKS 32
#endif

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt_math_div0"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt_math_div0___024root final {
    CData/*3:0*/ y;
    IData/*31:0*/ d2;
    IData/*31:0*/ m2;
    QData/*63:0*/ d3;
    QData/*63:0*/ m3;
    CData/*0:0*/ __VstlFirstIteration;
    CData/*0:0*/ __VstlPhaseResult;
    IData/*31:0*/ t__DOT__b;
    QData/*63:0*/ t__DOT__b3;
    VlUnpacked<QData/*63:0*/, 1> __VstlTriggered;
    const char* vlNamep;
};

struct Vt_math_div0__cuda final {
    typedef Vt_math_div0___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_y = 44,
        OFF_d2 = 32,
        OFF_m2 = 36,
        OFF_d3 = 0,
        OFF_m3 = 8,
        OFF___VstlFirstIteration = 45,
        OFF___VstlPhaseResult = 46,
        OFF_t__DOT__b = 40,
        OFF_t__DOT__b3 = 16,
        OFF___VstlTriggered = 24,
        kStateBytes = 48,  // S_persist
        kTouchedReadBytes = 0,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 0,  // S_state_wr
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
        s.y = st.ld<VlcuMem8>(OFF_y, i);
        s.d2 = st.ld<IData>(OFF_d2, i);
        s.m2 = st.ld<IData>(OFF_m2, i);
        s.d3 = st.ld<QData>(OFF_d3, i);
        s.m3 = st.ld<QData>(OFF_m3, i);
        s.__VstlFirstIteration = st.ld<VlcuMem8>(OFF___VstlFirstIteration, i);
        s.__VstlPhaseResult = st.ld<VlcuMem8>(OFF___VstlPhaseResult, i);
        s.t__DOT__b = st.ld<IData>(OFF_t__DOT__b, i);
        s.t__DOT__b3 = st.ld<QData>(OFF_t__DOT__b3, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VstlTriggered[e] = st.ldE<QData>(OFF___VstlTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_y, i, s.y);
        st.st<IData>(OFF_d2, i, s.d2);
        st.st<IData>(OFF_m2, i, s.m2);
        st.st<QData>(OFF_d3, i, s.d3);
        st.st<QData>(OFF_m3, i, s.m3);
        st.st<VlcuMem8>(OFF___VstlFirstIteration, i, s.__VstlFirstIteration);
        st.st<VlcuMem8>(OFF___VstlPhaseResult, i, s.__VstlPhaseResult);
        st.st<IData>(OFF_t__DOT__b, i, s.t__DOT__b);
        st.st<QData>(OFF_t__DOT__b3, i, s.t__DOT__b3);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VstlTriggered, e, i, s.__VstlTriggered[e]);
    }
    VLCU_D static void load_inputs_row(Root& s, const uint32_t* rowp, uint64_t stride, uint64_t i) {
    }
    VLCU_D static void store_outputs_row(const Root& s, uint32_t* rowp, uint64_t stride, uint64_t i) {
        rowp[0 * stride + i] = s.y;
        rowp[1 * stride + i] = s.d2;
        rowp[2 * stride + i] = s.m2;
        rowp[3 * stride + i] = static_cast<uint32_t>(s.d3);
        rowp[4 * stride + i] = static_cast<uint32_t>(s.d3 >> 32);
        rowp[5 * stride + i] = static_cast<uint32_t>(s.m3);
        rowp[6 * stride + i] = static_cast<uint32_t>(s.m3 >> 32);
    }
    VLCU_D static uint64_t outputs_sig(const Root& s, uint64_t sig) {
        uint32_t acc = VLCU_SIG_ACC_INIT;
        acc = vlcu_sig_word(acc, s.y);
        acc = vlcu_sig_word(acc, s.d2);
        acc = vlcu_sig_word(acc, s.m2);
        acc = vlcu_sig_word(acc, static_cast<uint32_t>(s.d3));
        acc = vlcu_sig_word(acc, static_cast<uint32_t>(s.d3 >> 32));
        acc = vlcu_sig_word(acc, static_cast<uint32_t>(s.m3));
        acc = vlcu_sig_word(acc, static_cast<uint32_t>(s.m3 >> 32));
        return vlcu_sig_mix(sig, acc);
    }
    enum : uint32_t { kInWords = 0, kOutWords = 7, kInBytes = 0, kOutBytes = 25, kTraceWords = 0 };

    VLCU_D static void Vt_math_div0___024root___eval_static(Vt_math_div0___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_math_div0___024root___eval_initial(Vt_math_div0___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        {
            // Inlined CFunc: _eval_initial__TOP
            vlSelfRef.y = 0U;
            vlSelfRef.t__DOT__b = 0xffffffffU;
            vlSelfRef.t__DOT__b3 = 0xffffffffffffffffULL;
            vlCtx.writef(0U, 0U);
            VL_FINISH_MT("/root/reference/test_regress/t/t_math_div0.v", 35, "");
        }
    }

    VLCU_D static void Vt_math_div0___024root___eval_final(Vt_math_div0___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_math_div0___024root___eval_settle(Vt_math_div0___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt_math_div0___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
#endif
                VL_FATAL_MT("/root/reference/test_regress/t/t_math_div0.v", 6, "", "DIDNOTCONVERGE: Settle region did not converge after '--converge-limit' of 10000 tries");
            }
            __VstlIterCount = ((IData)(1U) + __VstlIterCount);
            vlSelfRef.__VstlPhaseResult = Vt_math_div0___024root___eval_phase__stl(vlSelf, vlCtx);
            vlSelfRef.__VstlFirstIteration = 0U;
        } while (vlSelfRef.__VstlPhaseResult);
    }

    VLCU_D static bool Vt_math_div0___024root___trigger_anySet__stl(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt_math_div0___024root___eval_phase__stl(Vt_math_div0___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt_math_div0___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
        }
#endif
        __VstlExecute = Vt_math_div0___024root___trigger_anySet__stl(vlSelfRef.__VstlTriggered);
        if (__VstlExecute) {
            {
                // Inlined CFunc: _eval_stl
                if ((1ULL & vlSelfRef.__VstlTriggered[0U])) {
                    {
                        // Inlined CFunc: _stl_sequent__TOP__0
                        vlSelfRef.d2 = VL_DIVS_III(32, (IData)(0x80000000U), vlSelfRef.t__DOT__b);
                        vlSelfRef.m2 = VL_MODDIVS_III(32, (IData)(0x80000000U), vlSelfRef.t__DOT__b);
                        vlSelfRef.d3 = VL_DIVS_QQQ(64, 0x8000000000000000ULL, vlSelfRef.t__DOT__b3);
                        vlSelfRef.m3 = VL_MODDIVS_QQQ(64, 0x8000000000000000ULL, vlSelfRef.t__DOT__b3);
                    }
                }
            }
        }
        return (__VstlExecute);
    }

    VLCU_D static void Vt_math_div0___024root___eval(Vt_math_div0___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_math_div0___024root___ctor_var_reset(Vt_math_div0___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlSelf->y = 0U;
        ;
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->d2 = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 12637791950585568159ull);
        vlSelf->m2 = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 4502822743263013197ull);
        vlSelf->d3 = VL_SCOPED_RAND_RESET_Q(64, __VscopeHash, 4552505006284688023ull);
        vlSelf->m3 = VL_SCOPED_RAND_RESET_Q(64, __VscopeHash, 15123813510401005367ull);
        vlSelf->t__DOT__b = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 11816130287683444364ull);
        vlSelf->t__DOT__b3 = VL_SCOPED_RAND_RESET_Q(64, __VscopeHash, 16000718971521264823ull);
        for (int __Vi0 = 0; __Vi0 < 1; ++__Vi0) {
            vlSelf->__VstlTriggered[__Vi0] = 0;
        }
    }

    // tfp->dump(): every traced value stored at its trace code.  Full functions, then the
    // constants, as VerilatedTrace::dump runs them the first time (verilated_trace_imp.h:348-358)
    VLCU_D static void trace_sample(Root* vlSelf, VlcuContext& vlCtx, uint32_t* valsp) {
        (void)vlSelf; (void)vlCtx; (void)valsp;
    }

    VLCU_D static void ctor_var_reset(Root* vlSelf, VlcuContext& vlCtx) {
        Vt_math_div0___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_div0___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_div0___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_div0___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_div0___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_div0___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt_math_div0__cuda::info() {
    typedef Vt_math_div0__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"y", "y", VLCU_MEMBER_OUT, 1, 0, 1, 4, 3, 0, T::OFF_y},
        {"d2", "d2", VLCU_MEMBER_OUT, 4, 0, 1, 32, 31, 0, T::OFF_d2},
        {"m2", "m2", VLCU_MEMBER_OUT, 4, 0, 1, 32, 31, 0, T::OFF_m2},
        {"d3", "d3", VLCU_MEMBER_OUT, 8, 0, 1, 64, 63, 0, T::OFF_d3},
        {"m3", "m3", VLCU_MEMBER_OUT, 8, 0, 1, 64, 63, 0, T::OFF_m3},
        {"__VstlFirstIteration", "__VstlFirstIteration", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlFirstIteration},
        {"__VstlPhaseResult", "__VstlPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlPhaseResult},
        {"t__DOT__b", "t.b", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__b},
        {"t__DOT__b3", "t.b3", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF_t__DOT__b3},
        {"__VstlTriggered", "__VstlTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VstlTriggered},
    };
    static const char* const messages[] = {"*-* All Finished *-*\n", nullptr};
    static const char* const plusargs[] = {nullptr};
    static const VlcuTraceDecl traceDecls[] = {
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0}};  // (terminator, not counted)
    static const uint32_t traceOrder[] = {0};  // codes in dump order (+ terminator)
    static const uint32_t saveOrder[] = {0, 5, 6, 1, 2, 7, 3, 4, 8, 9, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "t", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 1, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 10, 0, 0x97f7bbf8a8d23cd8ULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt_math_div0__cuda VLCU_TESTBENCHES);
