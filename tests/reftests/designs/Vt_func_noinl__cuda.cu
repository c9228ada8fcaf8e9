// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

// 64 bytes of register-resident state: latency-bound, compile for full occupancy
#ifndef VLCU_MIN_BLOCKS
#define VLCU_MIN_BLOCKS 32
#endif

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt_func_noinl"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt_func_noinl___024root final {
    CData/*0:0*/ clk;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    IData/*31:0*/ t__DOT__cyc;
    IData/*31:0*/ t__DOT__outp;
    IData/*31:0*/ t__DOT__test__DOT__inp_d;
    IData/*31:0*/ __VactIterCount;
    QData/*63:0*/ t__DOT__crc;
    QData/*63:0*/ t__DOT__sum;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vt_func_noinl__cuda final {
    typedef Vt_func_noinl___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 48,
        OFF___Vtrigprevexpr___TOP__clk__0 = 49,
        OFF___VactPhaseResult = 50,
        OFF___VnbaPhaseResult = 51,
        OFF_t__DOT__cyc = 32,
        OFF_t__DOT__outp = 36,
        OFF_t__DOT__test__DOT__inp_d = 40,
        OFF___VactIterCount = 44,
        OFF_t__DOT__crc = 0,
        OFF_t__DOT__sum = 8,
        OFF___VactTriggered = 16,
        OFF___VnbaTriggered = 24,
        kStateBytes = 56,  // S_persist
        kTouchedReadBytes = 52,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 51,  // S_state_wr
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
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        s.t__DOT__cyc = st.ld<IData>(OFF_t__DOT__cyc, i);
        s.t__DOT__outp = st.ld<IData>(OFF_t__DOT__outp, i);
        s.t__DOT__test__DOT__inp_d = st.ld<IData>(OFF_t__DOT__test__DOT__inp_d, i);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        s.t__DOT__crc = st.ld<QData>(OFF_t__DOT__crc, i);
        s.t__DOT__sum = st.ld<QData>(OFF_t__DOT__sum, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.st<IData>(OFF_t__DOT__cyc, i, s.t__DOT__cyc);
        st.st<IData>(OFF_t__DOT__outp, i, s.t__DOT__outp);
        st.st<IData>(OFF_t__DOT__test__DOT__inp_d, i, s.t__DOT__test__DOT__inp_d);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
        st.st<QData>(OFF_t__DOT__crc, i, s.t__DOT__crc);
        st.st<QData>(OFF_t__DOT__sum, i, s.t__DOT__sum);
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

    VLCU_D static void Vt_func_noinl___024root___eval_static(Vt_func_noinl___024root* vlSelf, VlcuContext& vlCtx) {
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

    VLCU_D static void Vt_func_noinl___024root___eval_initial(Vt_func_noinl___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_func_noinl___024root___eval_final(Vt_func_noinl___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_func_noinl___024root___eval_settle(Vt_func_noinl___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static bool Vt_func_noinl___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt_func_noinl___024root___nba_sequent__TOP__0(Vt_func_noinl___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        IData/*31:0*/ __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout;
        __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout = 0;
        IData/*31:0*/ __Vfunc_t__DOT__test__DOT__no_inline_function__0__var1;
        __Vfunc_t__DOT__test__DOT__no_inline_function__0__var1 = 0;
        IData/*31:0*/ __Vfunc_t__DOT__test__DOT__no_inline_function__0__var2;
        __Vfunc_t__DOT__test__DOT__no_inline_function__0__var2 = 0;
        QData/*62:0*/ __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1;
        __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1 = 0;
        QData/*62:0*/ __Vfunc_t__DOT__test__DOT__no_inline_function__0__product2;
        __Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 = 0;
        IData/*31:0*/ __Vdly__t__DOT__cyc;
        __Vdly__t__DOT__cyc = 0;
        QData/*63:0*/ __Vdly__t__DOT__crc;
        __Vdly__t__DOT__crc = 0;
        QData/*63:0*/ __Vdly__t__DOT__sum;
        __Vdly__t__DOT__sum = 0;
        // Body
        __Vdly__t__DOT__cyc = vlSelfRef.t__DOT__cyc;
        __Vdly__t__DOT__crc = vlSelfRef.t__DOT__crc;
        __Vdly__t__DOT__sum = vlSelfRef.t__DOT__sum;
        __Vdly__t__DOT__cyc = ((IData)(1U) + vlSelfRef.t__DOT__cyc);
        __Vdly__t__DOT__crc = ((vlSelfRef.t__DOT__crc 
                                << 1U) | (QData)((IData)(
                                                         (1U 
                                                          & VL_REDXOR_64(
                                                                         (0x8000000000000005ULL 
                                                                          & vlSelfRef.t__DOT__crc))))));
        __Vdly__t__DOT__sum = ((QData)((IData)(vlSelfRef.t__DOT__outp)) 
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
                    VL_STOP_MT("/root/reference/test_regress/t/t_func_noinl.v", 57, "");
                }
                if (VL_UNLIKELY(((0xa7f0a34f9cf56ccbULL 
                                  != vlSelfRef.t__DOT__sum)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_func_noinl.v", 58, "");
                }
                vlCtx.writef(1U, 0U);
                VL_FINISH_MT("/root/reference/test_regress/t/t_func_noinl.v", 60, "");
            }
        }
        vlSelfRef.t__DOT__sum = __Vdly__t__DOT__sum;
        if (VL_GTS_III(32, 5U, vlSelfRef.t__DOT__cyc)) {
            vlSelfRef.t__DOT__outp = 0U;
        } else {
            __Vfunc_t__DOT__test__DOT__no_inline_function__0__var2 
                = vlSelfRef.t__DOT__test__DOT__inp_d;
            __Vfunc_t__DOT__test__DOT__no_inline_function__0__var1 
                = (IData)(vlSelfRef.t__DOT__crc);
            __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1 = 0;
            __Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 = 0ULL;
            if ((1U & __Vfunc_t__DOT__test__DOT__no_inline_function__0__var2)) {
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1 
                    = (QData)((IData)(__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1));
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                    = (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                       ^ __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1);
            }
            if ((2U & __Vfunc_t__DOT__test__DOT__no_inline_function__0__var2)) {
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1 
                    = (0x7fffffffffffffffULL & ((QData)((IData)(__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1)) 
                                                << 1U));
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                    = (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                       ^ __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1);
            }
            if ((4U & __Vfunc_t__DOT__test__DOT__no_inline_function__0__var2)) {
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1 
                    = (0x7fffffffffffffffULL & ((QData)((IData)(__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1)) 
                                                << 2U));
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                    = (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                       ^ __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1);
            }
            if ((8U & __Vfunc_t__DOT__test__DOT__no_inline_function__0__var2)) {
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1 
                    = (0x7fffffffffffffffULL & ((QData)((IData)(__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1)) 
                                                << 3U));
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                    = (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                       ^ __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1);
            }
            if ((0x00000010U & __Vfunc_t__DOT__test__DOT__no_inline_function__0__var2)) {
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1 
                    = (0x7fffffffffffffffULL & ((QData)((IData)(__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1)) 
                                                << 4U));
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                    = (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                       ^ __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1);
            }
            if ((0x00000020U & __Vfunc_t__DOT__test__DOT__no_inline_function__0__var2)) {
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1 
                    = (0x7fffffffffffffffULL & ((QData)((IData)(__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1)) 
                                                << 5U));
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                    = (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                       ^ __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1);
            }
            if ((0x00000040U & __Vfunc_t__DOT__test__DOT__no_inline_function__0__var2)) {
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1 
                    = (0x7fffffffffffffffULL & ((QData)((IData)(__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1)) 
                                                << 6U));
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                    = (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                       ^ __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1);
            }
            if ((0x00000080U & __Vfunc_t__DOT__test__DOT__no_inline_function__0__var2)) {
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1 
                    = (0x7fffffffffffffffULL & ((QData)((IData)(__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1)) 
                                                << 7U));
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                    = (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                       ^ __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1);
            }
            if ((0x00000100U & __Vfunc_t__DOT__test__DOT__no_inline_function__0__var2)) {
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1 
                    = (0x7fffffffffffffffULL & ((QData)((IData)(__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1)) 
                                                << 8U));
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                    = (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                       ^ __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1);
            }
            if ((0x00000200U & __Vfunc_t__DOT__test__DOT__no_inline_function__0__var2)) {
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1 
                    = (0x7fffffffffffffffULL & ((QData)((IData)(__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1)) 
                                                << 9U));
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                    = (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                       ^ __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1);
            }
            if ((0x00000400U & __Vfunc_t__DOT__test__DOT__no_inline_function__0__var2)) {
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1 
                    = (0x7fffffffffffffffULL & ((QData)((IData)(__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1)) 
                                                << 0x0000000aU));
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                    = (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                       ^ __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1);
            }
            if ((0x00000800U & __Vfunc_t__DOT__test__DOT__no_inline_function__0__var2)) {
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1 
                    = (0x7fffffffffffffffULL & ((QData)((IData)(__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1)) 
                                                << 0x0000000bU));
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                    = (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                       ^ __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1);
            }
            if ((0x00001000U & __Vfunc_t__DOT__test__DOT__no_inline_function__0__var2)) {
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1 
                    = (0x7fffffffffffffffULL & ((QData)((IData)(__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1)) 
                                                << 0x0000000cU));
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                    = (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                       ^ __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1);
            }
            if ((0x00002000U & __Vfunc_t__DOT__test__DOT__no_inline_function__0__var2)) {
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1 
                    = (0x7fffffffffffffffULL & ((QData)((IData)(__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1)) 
                                                << 0x0000000dU));
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                    = (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                       ^ __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1);
            }
            if ((0x00004000U & __Vfunc_t__DOT__test__DOT__no_inline_function__0__var2)) {
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1 
                    = (0x7fffffffffffffffULL & ((QData)((IData)(__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1)) 
                                                << 0x0000000eU));
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                    = (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                       ^ __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1);
            }
            if ((0x00008000U & __Vfunc_t__DOT__test__DOT__no_inline_function__0__var2)) {
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1 
                    = (0x7fffffffffffffffULL & ((QData)((IData)(__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1)) 
                                                << 0x0000000fU));
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                    = (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                       ^ __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1);
            }
            if ((0x00010000U & __Vfunc_t__DOT__test__DOT__no_inline_function__0__var2)) {
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1 
                    = (0x7fffffffffffffffULL & ((QData)((IData)(__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1)) 
                                                << 0x00000010U));
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                    = (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                       ^ __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1);
            }
            if ((0x00020000U & __Vfunc_t__DOT__test__DOT__no_inline_function__0__var2)) {
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1 
                    = (0x7fffffffffffffffULL & ((QData)((IData)(__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1)) 
                                                << 0x00000011U));
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                    = (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                       ^ __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1);
            }
            if ((0x00040000U & __Vfunc_t__DOT__test__DOT__no_inline_function__0__var2)) {
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1 
                    = (0x7fffffffffffffffULL & ((QData)((IData)(__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1)) 
                                                << 0x00000012U));
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                    = (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                       ^ __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1);
            }
            if ((0x00080000U & __Vfunc_t__DOT__test__DOT__no_inline_function__0__var2)) {
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1 
                    = (0x7fffffffffffffffULL & ((QData)((IData)(__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1)) 
                                                << 0x00000013U));
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                    = (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                       ^ __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1);
            }
            if ((0x00100000U & __Vfunc_t__DOT__test__DOT__no_inline_function__0__var2)) {
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1 
                    = (0x7fffffffffffffffULL & ((QData)((IData)(__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1)) 
                                                << 0x00000014U));
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                    = (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                       ^ __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1);
            }
            if ((0x00200000U & __Vfunc_t__DOT__test__DOT__no_inline_function__0__var2)) {
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1 
                    = (0x7fffffffffffffffULL & ((QData)((IData)(__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1)) 
                                                << 0x00000015U));
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                    = (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                       ^ __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1);
            }
            if ((0x00400000U & __Vfunc_t__DOT__test__DOT__no_inline_function__0__var2)) {
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1 
                    = (0x7fffffffffffffffULL & ((QData)((IData)(__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1)) 
                                                << 0x00000016U));
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                    = (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                       ^ __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1);
            }
            if ((0x00800000U & __Vfunc_t__DOT__test__DOT__no_inline_function__0__var2)) {
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1 
                    = (0x7fffffffffffffffULL & ((QData)((IData)(__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1)) 
                                                << 0x00000017U));
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                    = (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                       ^ __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1);
            }
            if ((0x01000000U & __Vfunc_t__DOT__test__DOT__no_inline_function__0__var2)) {
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1 
                    = (0x7fffffffffffffffULL & ((QData)((IData)(__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1)) 
                                                << 0x00000018U));
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                    = (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                       ^ __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1);
            }
            if ((0x02000000U & __Vfunc_t__DOT__test__DOT__no_inline_function__0__var2)) {
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1 
                    = (0x7fffffffffffffffULL & ((QData)((IData)(__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1)) 
                                                << 0x00000019U));
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                    = (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                       ^ __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1);
            }
            if ((0x04000000U & __Vfunc_t__DOT__test__DOT__no_inline_function__0__var2)) {
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1 
                    = (0x7fffffffffffffffULL & ((QData)((IData)(__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1)) 
                                                << 0x0000001aU));
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                    = (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                       ^ __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1);
            }
            if ((0x08000000U & __Vfunc_t__DOT__test__DOT__no_inline_function__0__var2)) {
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1 
                    = (0x7fffffffffffffffULL & ((QData)((IData)(__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1)) 
                                                << 0x0000001bU));
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                    = (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                       ^ __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1);
            }
            if ((0x10000000U & __Vfunc_t__DOT__test__DOT__no_inline_function__0__var2)) {
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1 
                    = (0x7fffffffffffffffULL & ((QData)((IData)(__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1)) 
                                                << 0x0000001cU));
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                    = (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                       ^ __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1);
            }
            if ((0x20000000U & __Vfunc_t__DOT__test__DOT__no_inline_function__0__var2)) {
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1 
                    = (0x7fffffffffffffffULL & ((QData)((IData)(__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1)) 
                                                << 0x0000001dU));
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                    = (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                       ^ __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1);
            }
            if ((0x40000000U & __Vfunc_t__DOT__test__DOT__no_inline_function__0__var2)) {
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1 
                    = (0x7fffffffffffffffULL & ((QData)((IData)(__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1)) 
                                                << 0x0000001eU));
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                    = (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                       ^ __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1);
            }
            if ((__Vfunc_t__DOT__test__DOT__no_inline_function__0__var2 
                 >> 0x1fU)) {
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1 
                    = (0x7fffffffffffffffULL & ((QData)((IData)(__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1)) 
                                                << 0x0000001fU));
                __Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                    = (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                       ^ __Vfunc_t__DOT__test__DOT__no_inline_function__0__product1);
            }
            __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout = 0U;
            __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                = ((0xfffffffdU & __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout) 
                   | (2U & (((IData)(__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2) 
                             ^ __Vfunc_t__DOT__test__DOT__no_inline_function__0__var1) 
                            << 1U)));
            __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                = ((0xfffffffbU & __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout) 
                   | (4U & ((((__Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                               >> 1U) ^ (IData)((__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                                                 >> 1U))) 
                             ^ (__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1 
                                >> 1U)) << 2U)));
            __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                = ((0xfffffff7U & __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout) 
                   | (8U & ((((__Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                               >> 2U) ^ (IData)((__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                                                 >> 2U))) 
                             ^ (__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1 
                                >> 2U)) << 3U)));
            __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                = ((0xffffffefU & __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout) 
                   | (0x00000010U & ((((__Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                                        >> 3U) ^ (IData)(
                                                         (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                                                          >> 3U))) 
                                      ^ (__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1 
                                         >> 3U)) << 4U)));
            __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                = ((0xffffffdfU & __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout) 
                   | (0x00000020U & ((((__Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                                        >> 4U) ^ (IData)(
                                                         (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                                                          >> 4U))) 
                                      ^ (__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1 
                                         >> 4U)) << 5U)));
            __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                = ((0xffffffbfU & __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout) 
                   | (0x00000040U & ((((__Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                                        >> 5U) ^ (IData)(
                                                         (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                                                          >> 5U))) 
                                      ^ (__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1 
                                         >> 5U)) << 6U)));
            __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                = ((0xffffff7fU & __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout) 
                   | (0x00000080U & ((((__Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                                        >> 6U) ^ (IData)(
                                                         (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                                                          >> 6U))) 
                                      ^ (__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1 
                                         >> 6U)) << 7U)));
            __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                = ((0xfffffeffU & __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout) 
                   | (0x00000100U & ((((__Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                                        >> 7U) ^ (IData)(
                                                         (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                                                          >> 7U))) 
                                      ^ (__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1 
                                         >> 7U)) << 8U)));
            __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                = ((0xfffffdffU & __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout) 
                   | (0x00000200U & ((((__Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                                        >> 8U) ^ (IData)(
                                                         (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                                                          >> 8U))) 
                                      ^ (__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1 
                                         >> 8U)) << 9U)));
            __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                = ((0xfffffbffU & __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout) 
                   | (0x00000400U & ((((__Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                                        >> 9U) ^ (IData)(
                                                         (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                                                          >> 9U))) 
                                      ^ (__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1 
                                         >> 9U)) << 0x0000000aU)));
            __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                = ((0xfffff7ffU & __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout) 
                   | (0x00000800U & ((((__Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                                        >> 0x0aU) ^ (IData)(
                                                            (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                                                             >> 0x0aU))) 
                                      ^ (__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1 
                                         >> 0x0aU)) 
                                     << 0x0000000bU)));
            __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                = ((0xffffefffU & __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout) 
                   | (0x00001000U & ((((__Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                                        >> 0x0bU) ^ (IData)(
                                                            (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                                                             >> 0x0bU))) 
                                      ^ (__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1 
                                         >> 0x0bU)) 
                                     << 0x0000000cU)));
            __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                = ((0xffffdfffU & __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout) 
                   | (0x00002000U & ((((__Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                                        >> 0x0cU) ^ (IData)(
                                                            (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                                                             >> 0x0cU))) 
                                      ^ (__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1 
                                         >> 0x0cU)) 
                                     << 0x0000000dU)));
            __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                = ((0xffffbfffU & __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout) 
                   | (0x00004000U & ((((__Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                                        >> 0x0dU) ^ (IData)(
                                                            (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                                                             >> 0x0dU))) 
                                      ^ (__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1 
                                         >> 0x0dU)) 
                                     << 0x0000000eU)));
            __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                = ((0xffff7fffU & __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout) 
                   | (0x00008000U & ((((__Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                                        >> 0x0eU) ^ (IData)(
                                                            (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                                                             >> 0x0eU))) 
                                      ^ (__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1 
                                         >> 0x0eU)) 
                                     << 0x0000000fU)));
            __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                = ((0xfffeffffU & __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout) 
                   | (0x00010000U & ((((__Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                                        >> 0x0fU) ^ (IData)(
                                                            (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                                                             >> 0x0fU))) 
                                      ^ (__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1 
                                         >> 0x0fU)) 
                                     << 0x00000010U)));
            __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                = ((0xfffdffffU & __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout) 
                   | (0x00020000U & ((((__Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                                        >> 0x10U) ^ (IData)(
                                                            (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                                                             >> 0x10U))) 
                                      ^ (__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1 
                                         >> 0x10U)) 
                                     << 0x00000011U)));
            __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                = ((0xfffbffffU & __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout) 
                   | (0x00040000U & ((((__Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                                        >> 0x11U) ^ (IData)(
                                                            (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                                                             >> 0x11U))) 
                                      ^ (__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1 
                                         >> 0x11U)) 
                                     << 0x00000012U)));
            __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                = ((0xfff7ffffU & __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout) 
                   | (0x00080000U & ((((__Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                                        >> 0x12U) ^ (IData)(
                                                            (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                                                             >> 0x12U))) 
                                      ^ (__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1 
                                         >> 0x12U)) 
                                     << 0x00000013U)));
            __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                = ((0xffefffffU & __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout) 
                   | (0x00100000U & ((((__Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                                        >> 0x13U) ^ (IData)(
                                                            (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                                                             >> 0x13U))) 
                                      ^ (__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1 
                                         >> 0x13U)) 
                                     << 0x00000014U)));
            __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                = ((0xffdfffffU & __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout) 
                   | (0x00200000U & ((((__Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                                        >> 0x14U) ^ (IData)(
                                                            (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                                                             >> 0x14U))) 
                                      ^ (__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1 
                                         >> 0x14U)) 
                                     << 0x00000015U)));
            __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                = ((0xffbfffffU & __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout) 
                   | (0x00400000U & ((((__Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                                        >> 0x15U) ^ (IData)(
                                                            (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                                                             >> 0x15U))) 
                                      ^ (__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1 
                                         >> 0x15U)) 
                                     << 0x00000016U)));
            __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                = ((0xff7fffffU & __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout) 
                   | (0x00800000U & ((((__Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                                        >> 0x16U) ^ (IData)(
                                                            (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                                                             >> 0x16U))) 
                                      ^ (__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1 
                                         >> 0x16U)) 
                                     << 0x00000017U)));
            __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                = ((0xfeffffffU & __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout) 
                   | (0x01000000U & ((((__Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                                        >> 0x17U) ^ (IData)(
                                                            (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                                                             >> 0x17U))) 
                                      ^ (__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1 
                                         >> 0x17U)) 
                                     << 0x00000018U)));
            __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                = ((0xfdffffffU & __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout) 
                   | (0x02000000U & ((((__Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                                        >> 0x18U) ^ (IData)(
                                                            (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                                                             >> 0x18U))) 
                                      ^ (__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1 
                                         >> 0x18U)) 
                                     << 0x00000019U)));
            __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                = ((0xfbffffffU & __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout) 
                   | (0x04000000U & ((((__Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                                        >> 0x19U) ^ (IData)(
                                                            (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                                                             >> 0x19U))) 
                                      ^ (__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1 
                                         >> 0x19U)) 
                                     << 0x0000001aU)));
            __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                = ((0xf7ffffffU & __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout) 
                   | (0x08000000U & ((((__Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                                        >> 0x1aU) ^ (IData)(
                                                            (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                                                             >> 0x1aU))) 
                                      ^ (__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1 
                                         >> 0x1aU)) 
                                     << 0x0000001bU)));
            __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                = ((0xefffffffU & __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout) 
                   | (0x10000000U & ((((__Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                                        >> 0x1bU) ^ (IData)(
                                                            (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                                                             >> 0x1bU))) 
                                      ^ (__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1 
                                         >> 0x1bU)) 
                                     << 0x0000001cU)));
            __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                = ((0xdfffffffU & __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout) 
                   | (0x20000000U & ((((__Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                                        >> 0x1cU) ^ (IData)(
                                                            (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                                                             >> 0x1cU))) 
                                      ^ (__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1 
                                         >> 0x1cU)) 
                                     << 0x0000001dU)));
            __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                = ((0xbfffffffU & __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout) 
                   | (0x40000000U & ((((__Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                                        >> 0x1dU) ^ (IData)(
                                                            (__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                                                             >> 0x1dU))) 
                                      ^ (__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1 
                                         >> 0x1dU)) 
                                     << 0x0000001eU)));
            __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                = ((0x7fffffffU & __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout) 
                   | ((((__Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout 
                         >> 0x1eU) ^ (IData)((__Vfunc_t__DOT__test__DOT__no_inline_function__0__product2 
                                              >> 0x1eU))) 
                       ^ (__Vfunc_t__DOT__test__DOT__no_inline_function__0__var1 
                          >> 0x1eU)) << 0x0000001fU));
            vlSelfRef.t__DOT__outp = __Vfunc_t__DOT__test__DOT__no_inline_function__0__Vfuncout;
        }
        if (VL_LTES_III(32, 5U, vlSelfRef.t__DOT__cyc)) {
            vlSelfRef.t__DOT__test__DOT__inp_d = (IData)(vlSelfRef.t__DOT__crc);
        }
        vlSelfRef.t__DOT__cyc = __Vdly__t__DOT__cyc;
        vlSelfRef.t__DOT__crc = __Vdly__t__DOT__crc;
    }

    VLCU_D static void Vt_func_noinl___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt_func_noinl___024root___eval_phase__act(Vt_func_noinl___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt_func_noinl___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vt_func_noinl___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vt_func_noinl___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vt_func_noinl___024root___eval_phase__nba(Vt_func_noinl___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vt_func_noinl___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    Vt_func_noinl___024root___nba_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
            Vt_func_noinl___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vt_func_noinl___024root___eval(Vt_func_noinl___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt_func_noinl___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("/root/reference/test_regress/t/t_func_noinl.v", 7, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vt_func_noinl___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("/root/reference/test_regress/t/t_func_noinl.v", 7, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vt_func_noinl___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vt_func_noinl___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vt_func_noinl___024root___ctor_var_reset(Vt_func_noinl___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        vlSelf->t__DOT__crc = VL_SCOPED_RAND_RESET_Q(64, __VscopeHash, 17680824324711171631ull);
        vlSelf->t__DOT__sum = VL_SCOPED_RAND_RESET_Q(64, __VscopeHash, 1514072735762163239ull);
        vlSelf->t__DOT__outp = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 10649294371443871603ull);
        vlSelf->t__DOT__test__DOT__inp_d = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 16471205566143858363ull);
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
        Vt_func_noinl___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt_func_noinl___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt_func_noinl___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt_func_noinl___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vt_func_noinl___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt_func_noinl___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt_func_noinl__cuda::info() {
    typedef Vt_func_noinl__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"t__DOT__cyc", "t.cyc", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__cyc},
        {"t__DOT__outp", "t.outp", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__outp},
        {"t__DOT__test__DOT__inp_d", "t.test.inp_d", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__test__DOT__inp_d},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"t__DOT__crc", "t.crc", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF_t__DOT__crc},
        {"t__DOT__sum", "t.sum", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF_t__DOT__sum},
        {"__VactTriggered", "__VactTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VactTriggered},
        {"__VnbaTriggered", "__VnbaTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VnbaTriggered},
    };
    static const char* const messages[] = {"[%0t] cyc==%0d crc=%x sum=%x\n", "*-* All Finished *-*\n", nullptr};
    static const char* const plusargs[] = {nullptr};
    static const VlcuTraceDecl traceDecls[] = {
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0}};  // (terminator, not counted)
    static const uint32_t traceOrder[] = {0};  // codes in dump order (+ terminator)
    static const uint32_t saveOrder[] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "t", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 2, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 12, 0, 0x5bbfd3b03f86b54ULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt_func_noinl__cuda VLCU_TESTBENCHES);
