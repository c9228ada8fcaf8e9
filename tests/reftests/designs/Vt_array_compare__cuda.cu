// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

// 12 bytes of register-resident state: latency-bound, compile for full occupancy
#ifndef VLCU_MIN_BLOCKS
#define VLCU_MIN_BLOCKS 32
#endif

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt_array_compare"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt_array_compare___024root final {
    VlUnpacked<CData/*3:0*/, 3> t__DOT__array_1;
    VlUnpacked<CData/*3:0*/, 3> t__DOT__array_2;
    VlUnpacked<CData/*3:0*/, 3> t__DOT__array_3;
    const char* vlNamep;
};

struct Vt_array_compare__cuda final {
    typedef Vt_array_compare___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_t__DOT__array_1 = 0,
        OFF_t__DOT__array_2 = 3,
        OFF_t__DOT__array_3 = 6,
        kStateBytes = 16,  // S_persist
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
        for (uint32_t e = 0; e < 3; ++e) s.t__DOT__array_1[e] = st.ldE<VlcuMem8>(OFF_t__DOT__array_1, e, i);
        for (uint32_t e = 0; e < 3; ++e) s.t__DOT__array_2[e] = st.ldE<VlcuMem8>(OFF_t__DOT__array_2, e, i);
        for (uint32_t e = 0; e < 3; ++e) s.t__DOT__array_3[e] = st.ldE<VlcuMem8>(OFF_t__DOT__array_3, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        for (uint32_t e = 0; e < 3; ++e) st.stE<VlcuMem8>(OFF_t__DOT__array_1, e, i, s.t__DOT__array_1[e]);
        for (uint32_t e = 0; e < 3; ++e) st.stE<VlcuMem8>(OFF_t__DOT__array_2, e, i, s.t__DOT__array_2[e]);
        for (uint32_t e = 0; e < 3; ++e) st.stE<VlcuMem8>(OFF_t__DOT__array_3, e, i, s.t__DOT__array_3[e]);
    }
    VLCU_D static void load_inputs_row(Root& s, const uint32_t* rowp, uint64_t stride, uint64_t i) {
    }
    VLCU_D static void store_outputs_row(const Root& s, uint32_t* rowp, uint64_t stride, uint64_t i) {
    }
    VLCU_D static uint64_t outputs_sig(const Root& s, uint64_t sig) {
        uint32_t acc = VLCU_SIG_ACC_INIT;
        return vlcu_sig_mix(sig, acc);
    }
    enum : uint32_t { kInWords = 0, kOutWords = 0, kInBytes = 0, kOutBytes = 0, kTraceWords = 0 };

    VLCU_D static void Vt_array_compare___024root___eval_static(Vt_array_compare___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_array_compare___024root___eval_initial(Vt_array_compare___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        {
            // Inlined CFunc: _eval_initial__TOP
            CData/*0:0*/ __Vinline_0__eval_initial__TOP_t__DOT__array_1_ne_array_2;
            __Vinline_0__eval_initial__TOP_t__DOT__array_1_ne_array_2 = 0;
            CData/*0:0*/ __Vinline_0__eval_initial__TOP_t__DOT__array_1_eq_array_2;
            __Vinline_0__eval_initial__TOP_t__DOT__array_1_eq_array_2 = 0;
            CData/*0:0*/ __Vinline_0__eval_initial__TOP_t__DOT__array_1_ne_array_3;
            __Vinline_0__eval_initial__TOP_t__DOT__array_1_ne_array_3 = 0;
            CData/*0:0*/ __Vinline_0__eval_initial__TOP_t__DOT__array_1_eq_array_3;
            __Vinline_0__eval_initial__TOP_t__DOT__array_1_eq_array_3 = 0;
            vlSelfRef.t__DOT__array_1[0U] = 8U;
            vlSelfRef.t__DOT__array_1[1U] = 8U;
            vlSelfRef.t__DOT__array_1[2U] = 8U;
            vlSelfRef.t__DOT__array_2[0U] = 8U;
            vlSelfRef.t__DOT__array_2[1U] = 8U;
            vlSelfRef.t__DOT__array_2[2U] = 8U;
            vlSelfRef.t__DOT__array_3[0U] = 8U;
            vlSelfRef.t__DOT__array_3[1U] = 4U;
            vlSelfRef.t__DOT__array_3[2U] = 4U;
            __Vinline_0__eval_initial__TOP_t__DOT__array_1_ne_array_2 
                = (((vlSelfRef.t__DOT__array_1[2U] 
                     != vlSelfRef.t__DOT__array_2[2U]) 
                    || (vlSelfRef.t__DOT__array_1[1U] 
                        != vlSelfRef.t__DOT__array_2[1U])) 
                   || (vlSelfRef.t__DOT__array_1[0U] 
                       != vlSelfRef.t__DOT__array_2[0U]));
            __Vinline_0__eval_initial__TOP_t__DOT__array_1_eq_array_2 
                = (((vlSelfRef.t__DOT__array_1[2U] 
                     == vlSelfRef.t__DOT__array_2[2U]) 
                    && (vlSelfRef.t__DOT__array_1[1U] 
                        == vlSelfRef.t__DOT__array_2[1U])) 
                   && (vlSelfRef.t__DOT__array_1[0U] 
                       == vlSelfRef.t__DOT__array_2[0U]));
            __Vinline_0__eval_initial__TOP_t__DOT__array_1_ne_array_3 
                = (((vlSelfRef.t__DOT__array_1[2U] 
                     != vlSelfRef.t__DOT__array_3[2U]) 
                    || (vlSelfRef.t__DOT__array_1[1U] 
                        != vlSelfRef.t__DOT__array_3[1U])) 
                   || (vlSelfRef.t__DOT__array_1[0U] 
                       != vlSelfRef.t__DOT__array_3[0U]));
            __Vinline_0__eval_initial__TOP_t__DOT__array_1_eq_array_3 
                = (((vlSelfRef.t__DOT__array_1[2U] 
                     == vlSelfRef.t__DOT__array_3[2U]) 
                    && (vlSelfRef.t__DOT__array_1[1U] 
                        == vlSelfRef.t__DOT__array_3[1U])) 
                   && (vlSelfRef.t__DOT__array_1[0U] 
                       == vlSelfRef.t__DOT__array_3[0U]));
            if (VL_UNLIKELY((__Vinline_0__eval_initial__TOP_t__DOT__array_1_ne_array_2))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_array_compare.v", 50, "");
            }
            if (VL_UNLIKELY(((1U & (~ (IData)(__Vinline_0__eval_initial__TOP_t__DOT__array_1_eq_array_2)))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_array_compare.v", 51, "");
            }
            if (VL_UNLIKELY(((1U & (~ (IData)(__Vinline_0__eval_initial__TOP_t__DOT__array_1_ne_array_3)))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_array_compare.v", 52, "");
            }
            if (VL_UNLIKELY((__Vinline_0__eval_initial__TOP_t__DOT__array_1_eq_array_3))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_array_compare.v", 53, "");
            }
            vlCtx.writef(0U, 0U);
            VL_FINISH_MT("/root/reference/test_regress/t/t_array_compare.v", 56, "");
        }
    }

    VLCU_D static void Vt_array_compare___024root___eval_final(Vt_array_compare___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_array_compare___024root___eval_settle(Vt_array_compare___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_array_compare___024root___eval(Vt_array_compare___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_array_compare___024root___ctor_var_reset(Vt_array_compare___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        for (int __Vi0 = 0; __Vi0 < 3; ++__Vi0) {
            vlSelf->t__DOT__array_1[__Vi0] = VL_SCOPED_RAND_RESET_I(4, __VscopeHash, 4195296930053658977ull);
        }
        for (int __Vi0 = 0; __Vi0 < 3; ++__Vi0) {
            vlSelf->t__DOT__array_2[__Vi0] = VL_SCOPED_RAND_RESET_I(4, __VscopeHash, 3287700222295240399ull);
        }
        for (int __Vi0 = 0; __Vi0 < 3; ++__Vi0) {
            vlSelf->t__DOT__array_3[__Vi0] = VL_SCOPED_RAND_RESET_I(4, __VscopeHash, 15094748091711627531ull);
        }
    }

    // tfp->dump(): every traced value stored at its trace code.  Full functions, then the
    // constants, as VerilatedTrace::dump runs them the first time (verilated_trace_imp.h:348-358)
    VLCU_D static void trace_sample(Root* vlSelf, VlcuContext& vlCtx, uint32_t* valsp) {
        (void)vlSelf; (void)vlCtx; (void)valsp;
    }

    VLCU_D static void ctor_var_reset(Root* vlSelf, VlcuContext& vlCtx) {
        Vt_array_compare___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt_array_compare___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt_array_compare___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt_array_compare___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vt_array_compare___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt_array_compare___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt_array_compare__cuda::info() {
    typedef Vt_array_compare__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"t__DOT__array_1", "t.array_1", VLCU_MEMBER_STATE, 1, 0, 3, 4, 3, 0, T::OFF_t__DOT__array_1},
        {"t__DOT__array_2", "t.array_2", VLCU_MEMBER_STATE, 1, 0, 3, 4, 3, 0, T::OFF_t__DOT__array_2},
        {"t__DOT__array_3", "t.array_3", VLCU_MEMBER_STATE, 1, 0, 3, 4, 3, 0, T::OFF_t__DOT__array_3},
    };
    static const char* const messages[] = {"*-* All Finished *-*\n", nullptr};
    static const char* const plusargs[] = {nullptr};
    static const VlcuTraceDecl traceDecls[] = {
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0}};  // (terminator, not counted)
    static const uint32_t traceOrder[] = {0};  // codes in dump order (+ terminator)
    static const uint32_t saveOrder[] = {0, 1, 2, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "t", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 1, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 3, 0, 0x7d9a61755df4f783ULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt_array_compare__cuda VLCU_TESTBENCHES);
