// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt_always_split_rst"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt_always_split_rst___024root final {
    CData/*0:0*/ clk;
    CData/*0:0*/ t__DOT__rstn;
    CData/*3:0*/ t__DOT__ff_out;
    CData/*3:0*/ t__DOT__fg_out;
    CData/*3:0*/ t__DOT__fh_out;
    CData/*3:0*/ t__DOT__test__DOT__ff_10;
    CData/*3:0*/ t__DOT__test__DOT__ff_11;
    CData/*3:0*/ t__DOT__test__DOT__ff_12;
    CData/*3:0*/ t__DOT__test__DOT__ff_13;
    CData/*3:0*/ t__DOT__test__DOT__fg_10;
    CData/*3:0*/ t__DOT__test__DOT__fg_11;
    CData/*3:0*/ t__DOT__test__DOT__fg_12;
    CData/*3:0*/ t__DOT__test__DOT__fg_13;
    CData/*3:0*/ t__DOT__test__DOT__fh_10;
    CData/*3:0*/ t__DOT__test__DOT__fh_11;
    CData/*3:0*/ t__DOT__test__DOT__fh_12;
    CData/*3:0*/ t__DOT__test__DOT__fh_13;
    CData/*0:0*/ __VstlFirstIteration;
    CData/*0:0*/ __VstlPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    IData/*31:0*/ t__DOT__cyc;
    IData/*31:0*/ __VactIterCount;
    QData/*63:0*/ t__DOT__crc;
    QData/*63:0*/ t__DOT__sum;
    VlUnpacked<QData/*63:0*/, 1> __VstlTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vt_always_split_rst__cuda final {
    typedef Vt_always_split_rst___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 48,
        OFF_t__DOT__rstn = 49,
        OFF_t__DOT__ff_out = 50,
        OFF_t__DOT__fg_out = 51,
        OFF_t__DOT__fh_out = 52,
        OFF_t__DOT__test__DOT__ff_10 = 53,
        OFF_t__DOT__test__DOT__ff_11 = 54,
        OFF_t__DOT__test__DOT__ff_12 = 55,
        OFF_t__DOT__test__DOT__ff_13 = 56,
        OFF_t__DOT__test__DOT__fg_10 = 57,
        OFF_t__DOT__test__DOT__fg_11 = 58,
        OFF_t__DOT__test__DOT__fg_12 = 59,
        OFF_t__DOT__test__DOT__fg_13 = 60,
        OFF_t__DOT__test__DOT__fh_10 = 61,
        OFF_t__DOT__test__DOT__fh_11 = 62,
        OFF_t__DOT__test__DOT__fh_12 = 63,
        OFF_t__DOT__test__DOT__fh_13 = 64,
        OFF___VstlFirstIteration = 65,
        OFF___VstlPhaseResult = 66,
        OFF___Vtrigprevexpr___TOP__clk__0 = 67,
        OFF___VactPhaseResult = 68,
        OFF___VnbaPhaseResult = 69,
        OFF_t__DOT__cyc = 40,
        OFF___VactIterCount = 44,
        OFF_t__DOT__crc = 0,
        OFF_t__DOT__sum = 8,
        OFF___VstlTriggered = 16,
        OFF___VactTriggered = 24,
        OFF___VnbaTriggered = 32,
        kStateBytes = 72,  // S_persist
        kTouchedReadBytes = 60,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 59,  // S_state_wr
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
        s.t__DOT__rstn = st.ld<VlcuMem8>(OFF_t__DOT__rstn, i);
        s.t__DOT__ff_out = st.ld<VlcuMem8>(OFF_t__DOT__ff_out, i);
        s.t__DOT__fg_out = st.ld<VlcuMem8>(OFF_t__DOT__fg_out, i);
        s.t__DOT__fh_out = st.ld<VlcuMem8>(OFF_t__DOT__fh_out, i);
        s.t__DOT__test__DOT__ff_10 = st.ld<VlcuMem8>(OFF_t__DOT__test__DOT__ff_10, i);
        s.t__DOT__test__DOT__ff_11 = st.ld<VlcuMem8>(OFF_t__DOT__test__DOT__ff_11, i);
        s.t__DOT__test__DOT__ff_12 = st.ld<VlcuMem8>(OFF_t__DOT__test__DOT__ff_12, i);
        s.t__DOT__test__DOT__ff_13 = st.ld<VlcuMem8>(OFF_t__DOT__test__DOT__ff_13, i);
        s.t__DOT__test__DOT__fg_10 = st.ld<VlcuMem8>(OFF_t__DOT__test__DOT__fg_10, i);
        s.t__DOT__test__DOT__fg_11 = st.ld<VlcuMem8>(OFF_t__DOT__test__DOT__fg_11, i);
        s.t__DOT__test__DOT__fg_12 = st.ld<VlcuMem8>(OFF_t__DOT__test__DOT__fg_12, i);
        s.t__DOT__test__DOT__fg_13 = st.ld<VlcuMem8>(OFF_t__DOT__test__DOT__fg_13, i);
        s.t__DOT__test__DOT__fh_10 = st.ld<VlcuMem8>(OFF_t__DOT__test__DOT__fh_10, i);
        s.t__DOT__test__DOT__fh_11 = st.ld<VlcuMem8>(OFF_t__DOT__test__DOT__fh_11, i);
        s.t__DOT__test__DOT__fh_12 = st.ld<VlcuMem8>(OFF_t__DOT__test__DOT__fh_12, i);
        s.t__DOT__test__DOT__fh_13 = st.ld<VlcuMem8>(OFF_t__DOT__test__DOT__fh_13, i);
        s.__VstlFirstIteration = st.ld<VlcuMem8>(OFF___VstlFirstIteration, i);
        s.__VstlPhaseResult = st.ld<VlcuMem8>(OFF___VstlPhaseResult, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        s.t__DOT__cyc = st.ld<IData>(OFF_t__DOT__cyc, i);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        s.t__DOT__crc = st.ld<QData>(OFF_t__DOT__crc, i);
        s.t__DOT__sum = st.ld<QData>(OFF_t__DOT__sum, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VstlTriggered[e] = st.ldE<QData>(OFF___VstlTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.st<VlcuMem8>(OFF_t__DOT__rstn, i, s.t__DOT__rstn);
        st.st<VlcuMem8>(OFF_t__DOT__ff_out, i, s.t__DOT__ff_out);
        st.st<VlcuMem8>(OFF_t__DOT__fg_out, i, s.t__DOT__fg_out);
        st.st<VlcuMem8>(OFF_t__DOT__fh_out, i, s.t__DOT__fh_out);
        st.st<VlcuMem8>(OFF_t__DOT__test__DOT__ff_10, i, s.t__DOT__test__DOT__ff_10);
        st.st<VlcuMem8>(OFF_t__DOT__test__DOT__ff_11, i, s.t__DOT__test__DOT__ff_11);
        st.st<VlcuMem8>(OFF_t__DOT__test__DOT__ff_12, i, s.t__DOT__test__DOT__ff_12);
        st.st<VlcuMem8>(OFF_t__DOT__test__DOT__ff_13, i, s.t__DOT__test__DOT__ff_13);
        st.st<VlcuMem8>(OFF_t__DOT__test__DOT__fg_10, i, s.t__DOT__test__DOT__fg_10);
        st.st<VlcuMem8>(OFF_t__DOT__test__DOT__fg_11, i, s.t__DOT__test__DOT__fg_11);
        st.st<VlcuMem8>(OFF_t__DOT__test__DOT__fg_12, i, s.t__DOT__test__DOT__fg_12);
        st.st<VlcuMem8>(OFF_t__DOT__test__DOT__fg_13, i, s.t__DOT__test__DOT__fg_13);
        st.st<VlcuMem8>(OFF_t__DOT__test__DOT__fh_10, i, s.t__DOT__test__DOT__fh_10);
        st.st<VlcuMem8>(OFF_t__DOT__test__DOT__fh_11, i, s.t__DOT__test__DOT__fh_11);
        st.st<VlcuMem8>(OFF_t__DOT__test__DOT__fh_12, i, s.t__DOT__test__DOT__fh_12);
        st.st<VlcuMem8>(OFF_t__DOT__test__DOT__fh_13, i, s.t__DOT__test__DOT__fh_13);
        st.st<VlcuMem8>(OFF___VstlFirstIteration, i, s.__VstlFirstIteration);
        st.st<VlcuMem8>(OFF___VstlPhaseResult, i, s.__VstlPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.st<IData>(OFF_t__DOT__cyc, i, s.t__DOT__cyc);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
        st.st<QData>(OFF_t__DOT__crc, i, s.t__DOT__crc);
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

    VLCU_D static void Vt_always_split_rst___024root___eval_static(Vt_always_split_rst___024root* vlSelf, VlcuContext& vlCtx) {
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

    VLCU_D static void Vt_always_split_rst___024root___eval_initial(Vt_always_split_rst___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_always_split_rst___024root___eval_final(Vt_always_split_rst___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_always_split_rst___024root___eval_settle(Vt_always_split_rst___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt_always_split_rst___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
#endif
                VL_FATAL_MT("/root/reference/test_regress/t/t_always_split_rst.v", 8, "", "DIDNOTCONVERGE: Settle region did not converge after '--converge-limit' of 10000 tries");
            }
            __VstlIterCount = ((IData)(1U) + __VstlIterCount);
            vlSelfRef.__VstlPhaseResult = Vt_always_split_rst___024root___eval_phase__stl(vlSelf, vlCtx);
            vlSelfRef.__VstlFirstIteration = 0U;
        } while (vlSelfRef.__VstlPhaseResult);
    }

    VLCU_D static bool Vt_always_split_rst___024root___trigger_anySet__stl(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt_always_split_rst___024root___eval_phase__stl(Vt_always_split_rst___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt_always_split_rst___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
        }
#endif
        __VstlExecute = Vt_always_split_rst___024root___trigger_anySet__stl(vlSelfRef.__VstlTriggered);
        if (__VstlExecute) {
            {
                // Inlined CFunc: _eval_stl
                if ((1ULL & vlSelfRef.__VstlTriggered[0U])) {
                    {
                        // Inlined CFunc: _stl_sequent__TOP__0
                        vlSelfRef.t__DOT__rstn = (1U 
                                                  & (~ 
                                                     (VL_GTS_III(32, 0x00000014U, vlSelfRef.t__DOT__cyc) 
                                                      | (0U 
                                                         == 
                                                         (0x0000000fU 
                                                          & (IData)(
                                                                    (vlSelfRef.t__DOT__crc 
                                                                     >> 8U)))))));
                    }
                }
            }
        }
        return (__VstlExecute);
    }

    VLCU_D static bool Vt_always_split_rst___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt_always_split_rst___024root___nba_sequent__TOP__0(Vt_always_split_rst___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        IData/*31:0*/ __Vdly__t__DOT__cyc;
        __Vdly__t__DOT__cyc = 0;
        QData/*63:0*/ __Vdly__t__DOT__crc;
        __Vdly__t__DOT__crc = 0;
        QData/*63:0*/ __Vdly__t__DOT__sum;
        __Vdly__t__DOT__sum = 0;
        // Body
        __Vdly__t__DOT__sum = vlSelfRef.t__DOT__sum;
        __Vdly__t__DOT__cyc = vlSelfRef.t__DOT__cyc;
        __Vdly__t__DOT__crc = vlSelfRef.t__DOT__crc;
        __Vdly__t__DOT__cyc = ((IData)(1U) + vlSelfRef.t__DOT__cyc);
        __Vdly__t__DOT__crc = ((vlSelfRef.t__DOT__crc 
                                << 1U) | (QData)((IData)(
                                                         (1U 
                                                          & VL_REDXOR_64(
                                                                         (0x8000000000000005ULL 
                                                                          & vlSelfRef.t__DOT__crc))))));
        __Vdly__t__DOT__sum = ((QData)((IData)((((IData)(vlSelfRef.t__DOT__ff_out) 
                                                 << 8U) 
                                                | (((IData)(vlSelfRef.t__DOT__fg_out) 
                                                    << 4U) 
                                                   | (IData)(vlSelfRef.t__DOT__fh_out))))) 
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
        } else if (VL_UNLIKELY(((0x00000063U == vlSelfRef.t__DOT__cyc)))) {
            vlCtx.writef(0U, 4U, vlcu_msg_arg(VL_TIME_UNITED_Q(1)), vlcu_msg_arg(vlSelfRef.t__DOT__cyc), vlcu_msg_arg(vlSelfRef.t__DOT__crc), vlcu_msg_arg(vlSelfRef.t__DOT__sum));
            if (VL_UNLIKELY(((0xc77bb9b3784ea091ULL 
                              != vlSelfRef.t__DOT__crc)))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_always_split_rst.v", 59, "");
            }
            if (VL_UNLIKELY(((0x77979747fd86e9fdULL 
                              != vlSelfRef.t__DOT__sum)))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_always_split_rst.v", 62, "");
            }
            vlCtx.writef(1U, 0U);
            VL_FINISH_MT("/root/reference/test_regress/t/t_always_split_rst.v", 64, "");
        }
        vlSelfRef.t__DOT__sum = __Vdly__t__DOT__sum;
        vlSelfRef.t__DOT__cyc = __Vdly__t__DOT__cyc;
        if (vlSelfRef.t__DOT__rstn) {
            vlSelfRef.t__DOT__ff_out = vlSelfRef.t__DOT__test__DOT__ff_13;
            if ((1U & (IData)((vlSelfRef.t__DOT__crc 
                               >> 4U)))) {
                vlSelfRef.t__DOT__fh_out = vlSelfRef.t__DOT__test__DOT__fh_13;
                vlSelfRef.t__DOT__test__DOT__fh_13 
                    = vlSelfRef.t__DOT__test__DOT__fh_12;
                vlSelfRef.t__DOT__test__DOT__fh_12 
                    = vlSelfRef.t__DOT__test__DOT__fh_11;
                vlSelfRef.t__DOT__test__DOT__fh_11 
                    = vlSelfRef.t__DOT__test__DOT__fh_10;
                vlSelfRef.t__DOT__test__DOT__fh_10 
                    = (0x0000000fU & (IData)(vlSelfRef.t__DOT__crc));
            }
            vlSelfRef.t__DOT__test__DOT__ff_13 = vlSelfRef.t__DOT__test__DOT__ff_12;
            vlSelfRef.t__DOT__test__DOT__ff_12 = vlSelfRef.t__DOT__test__DOT__ff_11;
            vlSelfRef.t__DOT__test__DOT__ff_11 = vlSelfRef.t__DOT__test__DOT__ff_10;
            vlSelfRef.t__DOT__test__DOT__ff_10 = (0x0000000fU 
                                                  & (IData)(vlSelfRef.t__DOT__crc));
        } else {
            vlSelfRef.t__DOT__ff_out = 0U;
            vlSelfRef.t__DOT__fh_out = 0U;
            vlSelfRef.t__DOT__test__DOT__ff_13 = 0U;
            vlSelfRef.t__DOT__test__DOT__fh_13 = 0U;
            vlSelfRef.t__DOT__test__DOT__ff_12 = 0U;
            vlSelfRef.t__DOT__test__DOT__fh_12 = 0U;
            vlSelfRef.t__DOT__test__DOT__ff_11 = 0U;
            vlSelfRef.t__DOT__test__DOT__fh_11 = 0U;
            vlSelfRef.t__DOT__test__DOT__ff_10 = 0U;
            vlSelfRef.t__DOT__test__DOT__fh_10 = 0U;
        }
        if ((1U & (IData)((vlSelfRef.t__DOT__crc >> 4U)))) {
            if (vlSelfRef.t__DOT__rstn) {
                vlSelfRef.t__DOT__fg_out = vlSelfRef.t__DOT__test__DOT__fg_13;
                vlSelfRef.t__DOT__test__DOT__fg_13 
                    = vlSelfRef.t__DOT__test__DOT__fg_12;
                vlSelfRef.t__DOT__test__DOT__fg_12 
                    = vlSelfRef.t__DOT__test__DOT__fg_11;
                vlSelfRef.t__DOT__test__DOT__fg_11 
                    = vlSelfRef.t__DOT__test__DOT__fg_10;
                vlSelfRef.t__DOT__test__DOT__fg_10 
                    = (0x0000000fU & (IData)(vlSelfRef.t__DOT__crc));
            } else {
                vlSelfRef.t__DOT__fg_out = 0U;
                vlSelfRef.t__DOT__test__DOT__fg_13 = 0U;
                vlSelfRef.t__DOT__test__DOT__fg_12 = 0U;
                vlSelfRef.t__DOT__test__DOT__fg_11 = 0U;
                vlSelfRef.t__DOT__test__DOT__fg_10 = 0U;
            }
        }
        vlSelfRef.t__DOT__crc = __Vdly__t__DOT__crc;
        vlSelfRef.t__DOT__rstn = (1U & (~ (VL_GTS_III(32, 0x00000014U, vlSelfRef.t__DOT__cyc) 
                                           | (0U == 
                                              (0x0000000fU 
                                               & (IData)(
                                                         (vlSelfRef.t__DOT__crc 
                                                          >> 8U)))))));
    }

    VLCU_D static void Vt_always_split_rst___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt_always_split_rst___024root___eval_phase__act(Vt_always_split_rst___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt_always_split_rst___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vt_always_split_rst___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vt_always_split_rst___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vt_always_split_rst___024root___eval_phase__nba(Vt_always_split_rst___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vt_always_split_rst___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    Vt_always_split_rst___024root___nba_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
            Vt_always_split_rst___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vt_always_split_rst___024root___eval(Vt_always_split_rst___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt_always_split_rst___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("/root/reference/test_regress/t/t_always_split_rst.v", 8, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vt_always_split_rst___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("/root/reference/test_regress/t/t_always_split_rst.v", 8, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vt_always_split_rst___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vt_always_split_rst___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vt_always_split_rst___024root___ctor_var_reset(Vt_always_split_rst___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        vlSelf->t__DOT__crc = VL_SCOPED_RAND_RESET_Q(64, __VscopeHash, 17680824324711171631ull);
        vlSelf->t__DOT__sum = VL_SCOPED_RAND_RESET_Q(64, __VscopeHash, 1514072735762163239ull);
        vlSelf->t__DOT__rstn = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 10275593677058802194ull);
        vlSelf->t__DOT__ff_out = VL_SCOPED_RAND_RESET_I(4, __VscopeHash, 2082641103291980491ull);
        vlSelf->t__DOT__fg_out = VL_SCOPED_RAND_RESET_I(4, __VscopeHash, 1490781451479804920ull);
        vlSelf->t__DOT__fh_out = VL_SCOPED_RAND_RESET_I(4, __VscopeHash, 16453111798234557149ull);
        vlSelf->t__DOT__test__DOT__ff_10 = VL_SCOPED_RAND_RESET_I(4, __VscopeHash, 7389424546721167408ull);
        vlSelf->t__DOT__test__DOT__ff_11 = VL_SCOPED_RAND_RESET_I(4, __VscopeHash, 9851794693045046695ull);
        vlSelf->t__DOT__test__DOT__ff_12 = VL_SCOPED_RAND_RESET_I(4, __VscopeHash, 9672324093102902156ull);
        vlSelf->t__DOT__test__DOT__ff_13 = VL_SCOPED_RAND_RESET_I(4, __VscopeHash, 8425633211346165962ull);
        vlSelf->t__DOT__test__DOT__fg_10 = VL_SCOPED_RAND_RESET_I(4, __VscopeHash, 5802111607209510640ull);
        vlSelf->t__DOT__test__DOT__fg_11 = VL_SCOPED_RAND_RESET_I(4, __VscopeHash, 12112215280735268044ull);
        vlSelf->t__DOT__test__DOT__fg_12 = VL_SCOPED_RAND_RESET_I(4, __VscopeHash, 15909715058816513674ull);
        vlSelf->t__DOT__test__DOT__fg_13 = VL_SCOPED_RAND_RESET_I(4, __VscopeHash, 15427206207405072951ull);
        vlSelf->t__DOT__test__DOT__fh_10 = VL_SCOPED_RAND_RESET_I(4, __VscopeHash, 8499143754463022536ull);
        vlSelf->t__DOT__test__DOT__fh_11 = VL_SCOPED_RAND_RESET_I(4, __VscopeHash, 9037982868887562593ull);
        vlSelf->t__DOT__test__DOT__fh_12 = VL_SCOPED_RAND_RESET_I(4, __VscopeHash, 9087371259535734119ull);
        vlSelf->t__DOT__test__DOT__fh_13 = VL_SCOPED_RAND_RESET_I(4, __VscopeHash, 7166401848182774910ull);
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
        Vt_always_split_rst___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt_always_split_rst___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt_always_split_rst___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt_always_split_rst___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vt_always_split_rst___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt_always_split_rst___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt_always_split_rst__cuda::info() {
    typedef Vt_always_split_rst__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"t__DOT__rstn", "t.rstn", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__rstn},
        {"t__DOT__ff_out", "t.ff_out", VLCU_MEMBER_STATE, 1, 0, 1, 4, 3, 0, T::OFF_t__DOT__ff_out},
        {"t__DOT__fg_out", "t.fg_out", VLCU_MEMBER_STATE, 1, 0, 1, 4, 3, 0, T::OFF_t__DOT__fg_out},
        {"t__DOT__fh_out", "t.fh_out", VLCU_MEMBER_STATE, 1, 0, 1, 4, 3, 0, T::OFF_t__DOT__fh_out},
        {"t__DOT__test__DOT__ff_10", "t.test.ff_10", VLCU_MEMBER_STATE, 1, 0, 1, 4, 3, 0, T::OFF_t__DOT__test__DOT__ff_10},
        {"t__DOT__test__DOT__ff_11", "t.test.ff_11", VLCU_MEMBER_STATE, 1, 0, 1, 4, 3, 0, T::OFF_t__DOT__test__DOT__ff_11},
        {"t__DOT__test__DOT__ff_12", "t.test.ff_12", VLCU_MEMBER_STATE, 1, 0, 1, 4, 3, 0, T::OFF_t__DOT__test__DOT__ff_12},
        {"t__DOT__test__DOT__ff_13", "t.test.ff_13", VLCU_MEMBER_STATE, 1, 0, 1, 4, 3, 0, T::OFF_t__DOT__test__DOT__ff_13},
        {"t__DOT__test__DOT__fg_10", "t.test.fg_10", VLCU_MEMBER_STATE, 1, 0, 1, 4, 3, 0, T::OFF_t__DOT__test__DOT__fg_10},
        {"t__DOT__test__DOT__fg_11", "t.test.fg_11", VLCU_MEMBER_STATE, 1, 0, 1, 4, 3, 0, T::OFF_t__DOT__test__DOT__fg_11},
        {"t__DOT__test__DOT__fg_12", "t.test.fg_12", VLCU_MEMBER_STATE, 1, 0, 1, 4, 3, 0, T::OFF_t__DOT__test__DOT__fg_12},
        {"t__DOT__test__DOT__fg_13", "t.test.fg_13", VLCU_MEMBER_STATE, 1, 0, 1, 4, 3, 0, T::OFF_t__DOT__test__DOT__fg_13},
        {"t__DOT__test__DOT__fh_10", "t.test.fh_10", VLCU_MEMBER_STATE, 1, 0, 1, 4, 3, 0, T::OFF_t__DOT__test__DOT__fh_10},
        {"t__DOT__test__DOT__fh_11", "t.test.fh_11", VLCU_MEMBER_STATE, 1, 0, 1, 4, 3, 0, T::OFF_t__DOT__test__DOT__fh_11},
        {"t__DOT__test__DOT__fh_12", "t.test.fh_12", VLCU_MEMBER_STATE, 1, 0, 1, 4, 3, 0, T::OFF_t__DOT__test__DOT__fh_12},
        {"t__DOT__test__DOT__fh_13", "t.test.fh_13", VLCU_MEMBER_STATE, 1, 0, 1, 4, 3, 0, T::OFF_t__DOT__test__DOT__fh_13},
        {"__VstlFirstIteration", "__VstlFirstIteration", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlFirstIteration},
        {"__VstlPhaseResult", "__VstlPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"t__DOT__cyc", "t.cyc", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__cyc},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"t__DOT__crc", "t.crc", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF_t__DOT__crc},
        {"t__DOT__sum", "t.sum", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF_t__DOT__sum},
        {"__VstlTriggered", "__VstlTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VstlTriggered},
        {"__VactTriggered", "__VactTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VactTriggered},
        {"__VnbaTriggered", "__VnbaTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VnbaTriggered},
    };
    static const char* const messages[] = {"[%0t] cyc==%0d crc=%x sum=%x\n", "*-* All Finished *-*\n", nullptr};
    static const char* const plusargs[] = {nullptr};
    static const VlcuTraceDecl traceDecls[] = {
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0}};  // (terminator, not counted)
    static const uint32_t traceOrder[] = {0};  // codes in dump order (+ terminator)
    static const uint32_t saveOrder[] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "t", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 2, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 29, 0, 0x72ea31a112838193ULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt_always_split_rst__cuda VLCU_TESTBENCHES);
