// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt_clk_gen"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt_clk_gen___024root final {
    CData/*0:0*/ clk;
    CData/*0:0*/ t__DOT__gendlyclk_r;
    CData/*0:0*/ t__DOT__genblkclk;
    CData/*0:0*/ __Vdly__t__DOT__gendlyclk_r;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*0:0*/ __Vtrigprevexpr___TOP__t__DOT__gendlyclk_r__0;
    CData/*0:0*/ __Vtrigprevexpr___TOP__t__DOT__genblkclk__0;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    IData/*31:0*/ t__DOT__cyc;
    IData/*31:0*/ t__DOT__gendlydata_r;
    IData/*31:0*/ t__DOT__dlydata_gr;
    IData/*31:0*/ t__DOT__genblkdata;
    IData/*31:0*/ t__DOT__blkdata_gr;
    IData/*31:0*/ t__DOT__initwire;
    IData/*31:0*/ __Vdly__t__DOT__gendlydata_r;
    IData/*31:0*/ __Vdly__t__DOT__cyc;
    IData/*31:0*/ __VactIterCount;
    VlUnpacked<QData/*63:0*/, 3> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 2> __VnbaTriggered;
    const char* vlNamep;
};

struct Vt_clk_gen__cuda final {
    typedef Vt_clk_gen___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 76,
        OFF_t__DOT__gendlyclk_r = 77,
        OFF_t__DOT__genblkclk = 78,
        OFF___Vdly__t__DOT__gendlyclk_r = 79,
        OFF___Vtrigprevexpr___TOP__clk__0 = 80,
        OFF___Vtrigprevexpr___TOP__t__DOT__gendlyclk_r__0 = 81,
        OFF___Vtrigprevexpr___TOP__t__DOT__genblkclk__0 = 82,
        OFF___VactPhaseResult = 83,
        OFF___VnbaPhaseResult = 84,
        OFF_t__DOT__cyc = 40,
        OFF_t__DOT__gendlydata_r = 44,
        OFF_t__DOT__dlydata_gr = 48,
        OFF_t__DOT__genblkdata = 52,
        OFF_t__DOT__blkdata_gr = 56,
        OFF_t__DOT__initwire = 60,
        OFF___Vdly__t__DOT__gendlydata_r = 64,
        OFF___Vdly__t__DOT__cyc = 68,
        OFF___VactIterCount = 72,
        OFF___VactTriggered = 0,
        OFF___VnbaTriggered = 24,
        kStateBytes = 88,  // S_persist
        kTouchedReadBytes = 85,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 72,  // S_state_wr
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
        s.t__DOT__gendlyclk_r = st.ld<VlcuMem8>(OFF_t__DOT__gendlyclk_r, i);
        s.t__DOT__genblkclk = st.ld<VlcuMem8>(OFF_t__DOT__genblkclk, i);
        s.__Vdly__t__DOT__gendlyclk_r = st.ld<VlcuMem8>(OFF___Vdly__t__DOT__gendlyclk_r, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__Vtrigprevexpr___TOP__t__DOT__gendlyclk_r__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__t__DOT__gendlyclk_r__0, i);
        s.__Vtrigprevexpr___TOP__t__DOT__genblkclk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__t__DOT__genblkclk__0, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        s.t__DOT__cyc = st.ld<IData>(OFF_t__DOT__cyc, i);
        s.t__DOT__gendlydata_r = st.ld<IData>(OFF_t__DOT__gendlydata_r, i);
        s.t__DOT__dlydata_gr = st.ld<IData>(OFF_t__DOT__dlydata_gr, i);
        s.t__DOT__genblkdata = st.ld<IData>(OFF_t__DOT__genblkdata, i);
        s.t__DOT__blkdata_gr = st.ld<IData>(OFF_t__DOT__blkdata_gr, i);
        s.t__DOT__initwire = st.ld<IData>(OFF_t__DOT__initwire, i);
        s.__Vdly__t__DOT__gendlydata_r = st.ld<IData>(OFF___Vdly__t__DOT__gendlydata_r, i);
        s.__Vdly__t__DOT__cyc = st.ld<IData>(OFF___Vdly__t__DOT__cyc, i);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        for (uint32_t e = 0; e < 3; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 2; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.st<VlcuMem8>(OFF_t__DOT__gendlyclk_r, i, s.t__DOT__gendlyclk_r);
        st.st<VlcuMem8>(OFF_t__DOT__genblkclk, i, s.t__DOT__genblkclk);
        st.st<VlcuMem8>(OFF___Vdly__t__DOT__gendlyclk_r, i, s.__Vdly__t__DOT__gendlyclk_r);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__t__DOT__gendlyclk_r__0, i, s.__Vtrigprevexpr___TOP__t__DOT__gendlyclk_r__0);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__t__DOT__genblkclk__0, i, s.__Vtrigprevexpr___TOP__t__DOT__genblkclk__0);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.st<IData>(OFF_t__DOT__cyc, i, s.t__DOT__cyc);
        st.st<IData>(OFF_t__DOT__gendlydata_r, i, s.t__DOT__gendlydata_r);
        st.st<IData>(OFF_t__DOT__dlydata_gr, i, s.t__DOT__dlydata_gr);
        st.st<IData>(OFF_t__DOT__genblkdata, i, s.t__DOT__genblkdata);
        st.st<IData>(OFF_t__DOT__blkdata_gr, i, s.t__DOT__blkdata_gr);
        st.st<IData>(OFF_t__DOT__initwire, i, s.t__DOT__initwire);
        st.st<IData>(OFF___Vdly__t__DOT__gendlydata_r, i, s.__Vdly__t__DOT__gendlydata_r);
        st.st<IData>(OFF___Vdly__t__DOT__cyc, i, s.__Vdly__t__DOT__cyc);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
        for (uint32_t e = 0; e < 3; ++e) st.stE<QData>(OFF___VactTriggered, e, i, s.__VactTriggered[e]);
        for (uint32_t e = 0; e < 2; ++e) st.stE<QData>(OFF___VnbaTriggered, e, i, s.__VnbaTriggered[e]);
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

    VLCU_D static void Vt_clk_gen___024root___eval_static(Vt_clk_gen___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
        vlSelfRef.__Vtrigprevexpr___TOP__t__DOT__gendlyclk_r__0 
            = vlSelfRef.t__DOT__gendlyclk_r;
        vlSelfRef.__Vtrigprevexpr___TOP__t__DOT__genblkclk__0 
            = vlSelfRef.t__DOT__genblkclk;
    }

    VLCU_D static void Vt_clk_gen___024root___eval_initial(Vt_clk_gen___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        {
            // Inlined CFunc: _eval_initial__TOP
            IData/*31:0*/ __Vinline_0__eval_initial__TOP_t__DOT__i;
            __Vinline_0__eval_initial__TOP_t__DOT__i = 0;
            vlSelfRef.t__DOT__cyc = 1U;
            __Vinline_0__eval_initial__TOP_t__DOT__i = 0U;
            while (VL_GTS_III(32, 0x00002710U, __Vinline_0__eval_initial__TOP_t__DOT__i)) {
                vlSelfRef.t__DOT__initwire = 0x00002200U;
                __Vinline_0__eval_initial__TOP_t__DOT__i 
                    = ((IData)(1U) + __Vinline_0__eval_initial__TOP_t__DOT__i);
            }
        }
    }

    VLCU_D static void Vt_clk_gen___024root___eval_final(Vt_clk_gen___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_clk_gen___024root___eval_settle(Vt_clk_gen___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_clk_gen___024root___eval_triggers_ext__act(Vt_clk_gen___024root* vlSelf, VlcuContext& vlCtx, const VlUnpacked<QData/*63:0*/, 2> &latched) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        IData/*31:0*/ __VactTrigPreLoopCounter;
        // Body
        __VactTrigPreLoopCounter = 0U;
        do {
            vlSelfRef.__VactTriggered[((IData)(2U) 
                                       + __VactTrigPreLoopCounter)] 
                = (vlSelfRef.__VactTriggered[__VactTrigPreLoopCounter] 
                   & (~ latched[__VactTrigPreLoopCounter]));
            __VactTrigPreLoopCounter = ((IData)(1U) 
                                        + __VactTrigPreLoopCounter);
        } while ((1U > __VactTrigPreLoopCounter));
    }

    VLCU_D static bool Vt_clk_gen___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 2> &in) {
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

    VLCU_D static void Vt_clk_gen___024root___trigger_orInto__act_ext_vec(VlUnpacked<QData/*63:0*/, 2> &out, const VlUnpacked<QData/*63:0*/, 3> &in) {
        typedef void VlcuFuncRet;
        // Locals
        IData/*31:0*/ n;
        // Body
        n = 0U;
        do {
            out[n] = (out[n] | in[n]);
            n = ((IData)(1U) + n);
        } while ((1U >= n));
    }

    VLCU_D static bool Vt_clk_gen___024root___eval_phase__act(Vt_clk_gen___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VactExecute;
        // Body
        {
            // Inlined CFunc: _eval_triggers_vec__act
            vlSelfRef.__VactTriggered[0U] = (QData)((IData)(
                                                            ((IData)(vlSelfRef.clk) 
                                                             & (~ (IData)(vlSelfRef.__Vtrigprevexpr___TOP__clk__0)))));
            vlSelfRef.__VactTriggered[1U] = (QData)((IData)(
                                                            ((((IData)(vlSelfRef.t__DOT__genblkclk) 
                                                               & (~ (IData)(vlSelfRef.__Vtrigprevexpr___TOP__t__DOT__genblkclk__0))) 
                                                              << 1U) 
                                                             | ((IData)(vlSelfRef.t__DOT__gendlyclk_r) 
                                                                & (~ (IData)(vlSelfRef.__Vtrigprevexpr___TOP__t__DOT__gendlyclk_r__0))))));
            vlSelfRef.__Vtrigprevexpr___TOP__clk__0 
                = vlSelfRef.clk;
            vlSelfRef.__Vtrigprevexpr___TOP__t__DOT__gendlyclk_r__0 
                = vlSelfRef.t__DOT__gendlyclk_r;
            vlSelfRef.__Vtrigprevexpr___TOP__t__DOT__genblkclk__0 
                = vlSelfRef.t__DOT__genblkclk;
        }
        Vt_clk_gen___024root___eval_triggers_ext__act(vlSelf, vlCtx, vlSelfRef.__VnbaTriggered);
#ifdef VL_DEBUG
        if (VL_UNLIKELY(vlSymsp->_vm_contextp__->debug())) {
            Vt_clk_gen___024root___dump_triggers__act_ext(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vt_clk_gen___024root___trigger_orInto__act_ext_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        __VactExecute = Vt_clk_gen___024root___trigger_anySet__act_ext(vlSelfRef.__VactTriggered);
        if (__VactExecute) {
            {
                // Inlined CFunc: _eval_act
                if ((1ULL & vlSelfRef.__VactTriggered[2U])) {
                    {
                        // Inlined CFunc: _act_sequent__TOP__0
                        vlSelfRef.__Vdly__t__DOT__gendlydata_r 
                            = vlSelfRef.t__DOT__gendlydata_r;
                        vlSelfRef.__Vdly__t__DOT__cyc 
                            = vlSelfRef.t__DOT__cyc;
                        vlSelfRef.__Vdly__t__DOT__gendlyclk_r 
                            = vlSelfRef.t__DOT__gendlyclk_r;
                    }
                }
                if ((1ULL & vlSelfRef.__VactTriggered[0U])) {
                    {
                        // Inlined CFunc: _act_sequent__TOP__1
                        vlSelfRef.__Vdly__t__DOT__gendlydata_r = 0x00110000U;
                        vlSelfRef.__Vdly__t__DOT__gendlyclk_r = 0U;
                        vlSelfRef.t__DOT__genblkclk = 0U;
                        vlSelfRef.t__DOT__genblkdata = 0U;
                        if ((0U != vlSelfRef.t__DOT__cyc)) {
                            vlSelfRef.__Vdly__t__DOT__cyc 
                                = ((IData)(1U) + vlSelfRef.t__DOT__cyc);
                            if (VL_UNLIKELY(((2U == vlSelfRef.t__DOT__cyc)))) {
                                vlSelfRef.t__DOT__genblkclk = 1U;
                                vlSelfRef.t__DOT__genblkdata = 0x00000aceU;
                                vlCtx.writef(0U, 1U, vlcu_msg_arg(VL_TIME_UNITED_Q(1)));
                                vlSelfRef.__Vdly__t__DOT__gendlyclk_r = 1U;
                                vlSelfRef.__Vdly__t__DOT__gendlydata_r = 0x00540000U;
                            }
                            if (VL_UNLIKELY(((3U == vlSelfRef.t__DOT__cyc)))) {
                                vlSelfRef.t__DOT__genblkdata = 0x00000dceU;
                                if (VL_UNLIKELY(((0x87542211U 
                                                  != 
                                                  (0x00000011U 
                                                   | (vlSelfRef.t__DOT__gendlydata_r 
                                                      | (vlSelfRef.t__DOT__dlydata_gr 
                                                         | (vlSelfRef.t__DOT__blkdata_gr 
                                                            | vlSelfRef.t__DOT__initwire)))))))) {
                                    VL_STOP_MT("/root/reference/test_regress/t/t_clk_gen.v", 55, "");
                                }
                                vlCtx.writef(1U, 0U);
                                VL_FINISH_MT("/root/reference/test_regress/t/t_clk_gen.v", 57, "");
                                vlSelfRef.__Vdly__t__DOT__gendlydata_r = 0x00ff0000U;
                            }
                        }
                    }
                }
            }
        }
        return (__VactExecute);
    }

    VLCU_D static bool Vt_clk_gen___024root___trigger_anySet__act_ext(const VlUnpacked<QData/*63:0*/, 3> &in) {
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
        } while ((3U > n));
        return (0U);
    }

    VLCU_D static void Vt_clk_gen___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 2> &out) {
        typedef void VlcuFuncRet;
        // Locals
        IData/*31:0*/ n;
        // Body
        n = 0U;
        do {
            out[n] = 0ULL;
            n = ((IData)(1U) + n);
        } while ((2U > n));
    }

    VLCU_D static bool Vt_clk_gen___024root___eval_phase__nba(Vt_clk_gen___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vt_clk_gen___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    {
                        // Inlined CFunc: _nba_sequent__TOP__0
                        vlSelfRef.t__DOT__gendlyclk_r 
                            = vlSelfRef.__Vdly__t__DOT__gendlyclk_r;
                        vlSelfRef.t__DOT__cyc = vlSelfRef.__Vdly__t__DOT__cyc;
                    }
                }
                if ((1ULL & vlSelfRef.__VnbaTriggered[1U])) {
                    {
                        // Inlined CFunc: _nba_sequent__TOP__1
                        if (VL_UNLIKELY(((0ULL < VL_TIME_UNITED_Q(1))))) {
                            vlCtx.writef(2U, 3U, vlcu_msg_arg(VL_TIME_UNITED_Q(1)), vlcu_msg_arg(vlSelfRef.t__DOT__gendlydata_r), vlcu_msg_arg(vlSelfRef.t__DOT__genblkdata));
                            vlSelfRef.t__DOT__dlydata_gr = 0x80000000U;
                            if (VL_UNLIKELY(((0x00540000U 
                                              != vlSelfRef.t__DOT__gendlydata_r)))) {
                                VL_STOP_MT("/root/reference/test_regress/t/t_clk_gen.v", 70, "");
                            }
                            if (VL_UNLIKELY(((0x00000aceU 
                                              != vlSelfRef.t__DOT__genblkdata)))) {
                                VL_STOP_MT("/root/reference/test_regress/t/t_clk_gen.v", 71, "");
                            }
                        }
                    }
                }
                if ((2ULL & vlSelfRef.__VnbaTriggered[1U])) {
                    {
                        // Inlined CFunc: _nba_sequent__TOP__2
                        if (VL_UNLIKELY(((0ULL < VL_TIME_UNITED_Q(1))))) {
                            vlCtx.writef(3U, 3U, vlcu_msg_arg(VL_TIME_UNITED_Q(1)), vlcu_msg_arg(vlSelfRef.t__DOT__gendlydata_r), vlcu_msg_arg(vlSelfRef.t__DOT__genblkdata));
                            vlSelfRef.t__DOT__blkdata_gr = 0x07000000U;
                            if (VL_UNLIKELY(((0x00000aceU 
                                              != vlSelfRef.t__DOT__genblkdata)))) {
                                VL_STOP_MT("/root/reference/test_regress/t/t_clk_gen.v", 84, "");
                            }
                        }
                    }
                }
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    {
                        // Inlined CFunc: _nba_sequent__TOP__3
                        vlSelfRef.t__DOT__gendlydata_r 
                            = vlSelfRef.__Vdly__t__DOT__gendlydata_r;
                    }
                }
            }
            Vt_clk_gen___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vt_clk_gen___024root___eval(Vt_clk_gen___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt_clk_gen___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("/root/reference/test_regress/t/t_clk_gen.v", 7, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vt_clk_gen___024root___dump_triggers__act_ext(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("/root/reference/test_regress/t/t_clk_gen.v", 7, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vt_clk_gen___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vt_clk_gen___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vt_clk_gen___024root___ctor_var_reset(Vt_clk_gen___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        vlSelf->t__DOT__cyc = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 15508623621340971111ull);
        vlSelf->t__DOT__gendlyclk_r = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 17084736132951200016ull);
        vlSelf->t__DOT__gendlydata_r = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 17354153824228876240ull);
        vlSelf->t__DOT__dlydata_gr = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 6676639359898947850ull);
        vlSelf->t__DOT__genblkclk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 4154712316518099688ull);
        vlSelf->t__DOT__genblkdata = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 11216324000925227277ull);
        vlSelf->t__DOT__blkdata_gr = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 12523824713056300033ull);
        vlSelf->t__DOT__initwire = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 7470068425268488702ull);
        vlSelf->__Vdly__t__DOT__gendlydata_r = 0;
        vlSelf->__Vdly__t__DOT__gendlyclk_r = 0;
        vlSelf->__Vdly__t__DOT__cyc = 0;
        for (int __Vi0 = 0; __Vi0 < 3; ++__Vi0) {
            vlSelf->__VactTriggered[__Vi0] = 0;
        }
        vlSelf->__Vtrigprevexpr___TOP__clk__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__t__DOT__gendlyclk_r__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__t__DOT__genblkclk__0 = 0;
        for (int __Vi0 = 0; __Vi0 < 2; ++__Vi0) {
            vlSelf->__VnbaTriggered[__Vi0] = 0;
        }
    }

    // tfp->dump(): every traced value stored at its trace code.  Full functions, then the
    // constants, as VerilatedTrace::dump runs them the first time (verilated_trace_imp.h:348-358)
    VLCU_D static void trace_sample(Root* vlSelf, VlcuContext& vlCtx, uint32_t* valsp) {
        (void)vlSelf; (void)vlCtx; (void)valsp;
    }

    VLCU_D static void ctor_var_reset(Root* vlSelf, VlcuContext& vlCtx) {
        Vt_clk_gen___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt_clk_gen___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt_clk_gen___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt_clk_gen___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vt_clk_gen___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt_clk_gen___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt_clk_gen__cuda::info() {
    typedef Vt_clk_gen__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"t__DOT__gendlyclk_r", "t.gendlyclk_r", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__gendlyclk_r},
        {"t__DOT__genblkclk", "t.genblkclk", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__genblkclk},
        {"__Vdly__t__DOT__gendlyclk_r", "__Vdly__t.gendlyclk_r", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vdly__t__DOT__gendlyclk_r},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__Vtrigprevexpr___TOP__t__DOT__gendlyclk_r__0", "__Vtrigprevexpr___TOP__t.gendlyclk_r__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__t__DOT__gendlyclk_r__0},
        {"__Vtrigprevexpr___TOP__t__DOT__genblkclk__0", "__Vtrigprevexpr___TOP__t.genblkclk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__t__DOT__genblkclk__0},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"t__DOT__cyc", "t.cyc", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__cyc},
        {"t__DOT__gendlydata_r", "t.gendlydata_r", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__gendlydata_r},
        {"t__DOT__dlydata_gr", "t.dlydata_gr", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__dlydata_gr},
        {"t__DOT__genblkdata", "t.genblkdata", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__genblkdata},
        {"t__DOT__blkdata_gr", "t.blkdata_gr", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__blkdata_gr},
        {"t__DOT__initwire", "t.initwire", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__initwire},
        {"__Vdly__t__DOT__gendlydata_r", "__Vdly__t.gendlydata_r", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___Vdly__t__DOT__gendlydata_r},
        {"__Vdly__t__DOT__cyc", "__Vdly__t.cyc", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___Vdly__t__DOT__cyc},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"__VactTriggered", "__VactTriggered", VLCU_MEMBER_STATE, 8, 0, 3, 64, 63, 0, T::OFF___VactTriggered},
        {"__VnbaTriggered", "__VnbaTriggered", VLCU_MEMBER_STATE, 8, 0, 2, 64, 63, 0, T::OFF___VnbaTriggered},
    };
    static const char* const messages[] = {"[%0t] Send pulse\n", "*-* All Finished *-*\n", "[%0t] Got gendlyclk_r, d=%x b=%x\n", "[%0t] Got genblkclk, d=%x b=%x\n", nullptr};
    static const char* const plusargs[] = {nullptr};
    static const VlcuTraceDecl traceDecls[] = {
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0}};  // (terminator, not counted)
    static const uint32_t traceOrder[] = {0};  // codes in dump order (+ terminator)
    static const uint32_t saveOrder[] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "t", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 4, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 20, 0, 0x62ad85a9ec1d2fe7ULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt_clk_gen__cuda VLCU_TESTBENCHES);
