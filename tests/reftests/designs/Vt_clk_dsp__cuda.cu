// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt_clk_dsp"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt_clk_dsp___024root final {
    CData/*0:0*/ clk;
    CData/*7:0*/ t__DOT__cyc;
    CData/*7:0*/ t__DOT__padd;
    CData/*0:0*/ t__DOT__dsp_ph1;
    CData/*0:0*/ t__DOT__dsp_ph2;
    CData/*0:0*/ t__DOT__dsp_reset;
    CData/*7:0*/ t__DOT__out;
    CData/*0:0*/ t__DOT__dspchip__DOT__out1_r;
    CData/*7:0*/ t__DOT__dspchip__DOT__out2_r;
    CData/*7:0*/ t__DOT__dspchip__DOT__t_dsppla__DOT__latched_r;
    CData/*7:0*/ __Vdly__t__DOT__cyc;
    CData/*0:0*/ __VstlFirstIteration;
    CData/*0:0*/ __VstlPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*0:0*/ __Vtrigprevexpr___TOP__t__DOT__dsp_ph1__0;
    CData/*0:0*/ __Vtrigprevexpr___TOP__t__DOT__dsp_reset__0;
    CData/*0:0*/ __Vtrigprevexpr___TOP__t__DOT__dsp_ph2__0;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    IData/*31:0*/ __VactIterCount;
    VlUnpacked<QData/*63:0*/, 1> __VstlTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vt_clk_dsp__cuda final {
    typedef Vt_clk_dsp___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 28,
        OFF_t__DOT__cyc = 29,
        OFF_t__DOT__padd = 30,
        OFF_t__DOT__dsp_ph1 = 31,
        OFF_t__DOT__dsp_ph2 = 32,
        OFF_t__DOT__dsp_reset = 33,
        OFF_t__DOT__out = 34,
        OFF_t__DOT__dspchip__DOT__out1_r = 35,
        OFF_t__DOT__dspchip__DOT__out2_r = 36,
        OFF_t__DOT__dspchip__DOT__t_dsppla__DOT__latched_r = 37,
        OFF___Vdly__t__DOT__cyc = 38,
        OFF___VstlFirstIteration = 39,
        OFF___VstlPhaseResult = 40,
        OFF___Vtrigprevexpr___TOP__clk__0 = 41,
        OFF___Vtrigprevexpr___TOP__t__DOT__dsp_ph1__0 = 42,
        OFF___Vtrigprevexpr___TOP__t__DOT__dsp_reset__0 = 43,
        OFF___Vtrigprevexpr___TOP__t__DOT__dsp_ph2__0 = 44,
        OFF___VactPhaseResult = 45,
        OFF___VnbaPhaseResult = 46,
        OFF___VactIterCount = 24,
        OFF___VstlTriggered = 0,
        OFF___VactTriggered = 8,
        OFF___VnbaTriggered = 16,
        kStateBytes = 48,  // S_persist
        kTouchedReadBytes = 37,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 36,  // S_state_wr
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
        s.t__DOT__cyc = st.ld<VlcuMem8>(OFF_t__DOT__cyc, i);
        s.t__DOT__padd = st.ld<VlcuMem8>(OFF_t__DOT__padd, i);
        s.t__DOT__dsp_ph1 = st.ld<VlcuMem8>(OFF_t__DOT__dsp_ph1, i);
        s.t__DOT__dsp_ph2 = st.ld<VlcuMem8>(OFF_t__DOT__dsp_ph2, i);
        s.t__DOT__dsp_reset = st.ld<VlcuMem8>(OFF_t__DOT__dsp_reset, i);
        s.t__DOT__out = st.ld<VlcuMem8>(OFF_t__DOT__out, i);
        s.t__DOT__dspchip__DOT__out1_r = st.ld<VlcuMem8>(OFF_t__DOT__dspchip__DOT__out1_r, i);
        s.t__DOT__dspchip__DOT__out2_r = st.ld<VlcuMem8>(OFF_t__DOT__dspchip__DOT__out2_r, i);
        s.t__DOT__dspchip__DOT__t_dsppla__DOT__latched_r = st.ld<VlcuMem8>(OFF_t__DOT__dspchip__DOT__t_dsppla__DOT__latched_r, i);
        s.__Vdly__t__DOT__cyc = st.ld<VlcuMem8>(OFF___Vdly__t__DOT__cyc, i);
        s.__VstlFirstIteration = st.ld<VlcuMem8>(OFF___VstlFirstIteration, i);
        s.__VstlPhaseResult = st.ld<VlcuMem8>(OFF___VstlPhaseResult, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__Vtrigprevexpr___TOP__t__DOT__dsp_ph1__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__t__DOT__dsp_ph1__0, i);
        s.__Vtrigprevexpr___TOP__t__DOT__dsp_reset__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__t__DOT__dsp_reset__0, i);
        s.__Vtrigprevexpr___TOP__t__DOT__dsp_ph2__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__t__DOT__dsp_ph2__0, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VstlTriggered[e] = st.ldE<QData>(OFF___VstlTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.st<VlcuMem8>(OFF_t__DOT__cyc, i, s.t__DOT__cyc);
        st.st<VlcuMem8>(OFF_t__DOT__padd, i, s.t__DOT__padd);
        st.st<VlcuMem8>(OFF_t__DOT__dsp_ph1, i, s.t__DOT__dsp_ph1);
        st.st<VlcuMem8>(OFF_t__DOT__dsp_ph2, i, s.t__DOT__dsp_ph2);
        st.st<VlcuMem8>(OFF_t__DOT__dsp_reset, i, s.t__DOT__dsp_reset);
        st.st<VlcuMem8>(OFF_t__DOT__out, i, s.t__DOT__out);
        st.st<VlcuMem8>(OFF_t__DOT__dspchip__DOT__out1_r, i, s.t__DOT__dspchip__DOT__out1_r);
        st.st<VlcuMem8>(OFF_t__DOT__dspchip__DOT__out2_r, i, s.t__DOT__dspchip__DOT__out2_r);
        st.st<VlcuMem8>(OFF_t__DOT__dspchip__DOT__t_dsppla__DOT__latched_r, i, s.t__DOT__dspchip__DOT__t_dsppla__DOT__latched_r);
        st.st<VlcuMem8>(OFF___Vdly__t__DOT__cyc, i, s.__Vdly__t__DOT__cyc);
        st.st<VlcuMem8>(OFF___VstlFirstIteration, i, s.__VstlFirstIteration);
        st.st<VlcuMem8>(OFF___VstlPhaseResult, i, s.__VstlPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__t__DOT__dsp_ph1__0, i, s.__Vtrigprevexpr___TOP__t__DOT__dsp_ph1__0);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__t__DOT__dsp_reset__0, i, s.__Vtrigprevexpr___TOP__t__DOT__dsp_reset__0);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__t__DOT__dsp_ph2__0, i, s.__Vtrigprevexpr___TOP__t__DOT__dsp_ph2__0);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
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

    VLCU_D static void Vt_clk_dsp___024root___eval_static(Vt_clk_dsp___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
        vlSelfRef.__Vtrigprevexpr___TOP__t__DOT__dsp_ph1__0 
            = vlSelfRef.t__DOT__dsp_ph1;
        vlSelfRef.__Vtrigprevexpr___TOP__t__DOT__dsp_reset__0 
            = vlSelfRef.t__DOT__dsp_reset;
        vlSelfRef.__Vtrigprevexpr___TOP__t__DOT__dsp_ph2__0 
            = vlSelfRef.t__DOT__dsp_ph2;
    }

    VLCU_D static void Vt_clk_dsp___024root___eval_initial(Vt_clk_dsp___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        {
            // Inlined CFunc: _eval_initial__TOP
            vlSelfRef.t__DOT__cyc = 0U;
        }
    }

    VLCU_D static void Vt_clk_dsp___024root___eval_final(Vt_clk_dsp___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_clk_dsp___024root___eval_settle(Vt_clk_dsp___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt_clk_dsp___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
#endif
                VL_FATAL_MT("/root/reference/test_regress/t/t_clk_dsp.v", 7, "", "DIDNOTCONVERGE: Settle region did not converge after '--converge-limit' of 10000 tries");
            }
            __VstlIterCount = ((IData)(1U) + __VstlIterCount);
            vlSelfRef.__VstlPhaseResult = Vt_clk_dsp___024root___eval_phase__stl(vlSelf, vlCtx);
            vlSelfRef.__VstlFirstIteration = 0U;
        } while (vlSelfRef.__VstlPhaseResult);
    }

    VLCU_D static bool Vt_clk_dsp___024root___trigger_anySet__stl(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt_clk_dsp___024root___eval_phase__stl(Vt_clk_dsp___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt_clk_dsp___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
        }
#endif
        __VstlExecute = Vt_clk_dsp___024root___trigger_anySet__stl(vlSelfRef.__VstlTriggered);
        if (__VstlExecute) {
            {
                // Inlined CFunc: _eval_stl
                if ((1ULL & vlSelfRef.__VstlTriggered[0U])) {
                    {
                        // Inlined CFunc: _nba_comb__TOP__0
                        vlSelfRef.t__DOT__out = ((IData)(vlSelfRef.t__DOT__dspchip__DOT__out2_r) 
                                                 & (- (IData)(
                                                              (1U 
                                                               & (~ (IData)(vlSelfRef.t__DOT__dspchip__DOT__out1_r))))));
                    }
                }
            }
        }
        return (__VstlExecute);
    }

    VLCU_D static bool Vt_clk_dsp___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt_clk_dsp___024root___nba_sequent__TOP__4(Vt_clk_dsp___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlCtx.writef(0U, 1U, vlcu_msg_arg(vlSelfRef.t__DOT__cyc));
        if ((0U == (IData)(vlSelfRef.t__DOT__cyc))) {
            vlSelfRef.__Vdly__t__DOT__cyc = 1U;
            vlSelfRef.t__DOT__dsp_reset = 0U;
            vlSelfRef.t__DOT__padd = 0U;
        } else if (VL_UNLIKELY(((0x14U == (IData)(vlSelfRef.t__DOT__cyc))))) {
            vlCtx.writef(1U, 0U);
            VL_FINISH_MT("/root/reference/test_regress/t/t_clk_dsp.v", 43, "");
        } else {
            vlSelfRef.__Vdly__t__DOT__cyc = (0x000000ffU 
                                             & ((IData)(1U) 
                                                + (IData)(vlSelfRef.t__DOT__cyc)));
            vlSelfRef.t__DOT__dsp_ph1 = (0U == (3U 
                                                & (IData)(vlSelfRef.t__DOT__cyc)));
            vlSelfRef.t__DOT__dsp_ph2 = (2U == (3U 
                                                & (IData)(vlSelfRef.t__DOT__cyc)));
            vlSelfRef.t__DOT__dsp_reset = (1U == (IData)(vlSelfRef.t__DOT__cyc));
            vlSelfRef.t__DOT__padd = vlSelfRef.t__DOT__cyc;
            if (VL_UNLIKELY(((0x00000080U & (IData)(vlSelfRef.t__DOT__cyc))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_clk_dsp.v", 53, "");
            } else if (VL_UNLIKELY(((0x00000040U & (IData)(vlSelfRef.t__DOT__cyc))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_clk_dsp.v", 53, "");
            } else if (VL_UNLIKELY(((0x00000020U & (IData)(vlSelfRef.t__DOT__cyc))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_clk_dsp.v", 53, "");
            } else if ((0x00000010U & (IData)(vlSelfRef.t__DOT__cyc))) {
                if (VL_UNLIKELY(((8U & (IData)(vlSelfRef.t__DOT__cyc))))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_clk_dsp.v", 53, "");
                } else if (VL_UNLIKELY(((4U & (IData)(vlSelfRef.t__DOT__cyc))))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_clk_dsp.v", 53, "");
                } else if ((2U & (IData)(vlSelfRef.t__DOT__cyc))) {
                    if ((1U & (IData)(vlSelfRef.t__DOT__cyc))) {
                        if (VL_UNLIKELY(((0x10U != (IData)(vlSelfRef.t__DOT__out))))) {
                            VL_STOP_MT("/root/reference/test_regress/t/t_clk_dsp.v", 72, "");
                        }
                    } else if (VL_UNLIKELY(((0x0cU 
                                             != (IData)(vlSelfRef.t__DOT__out))))) {
                        VL_STOP_MT("/root/reference/test_regress/t/t_clk_dsp.v", 71, "");
                    }
                } else if ((1U & (IData)(vlSelfRef.t__DOT__cyc))) {
                    if (VL_UNLIKELY(((0x0cU != (IData)(vlSelfRef.t__DOT__out))))) {
                        VL_STOP_MT("/root/reference/test_regress/t/t_clk_dsp.v", 70, "");
                    }
                } else if (VL_UNLIKELY(((0U != (IData)(vlSelfRef.t__DOT__out))))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_clk_dsp.v", 69, "");
                }
            } else if ((8U & (IData)(vlSelfRef.t__DOT__cyc))) {
                if ((4U & (IData)(vlSelfRef.t__DOT__cyc))) {
                    if ((2U & (IData)(vlSelfRef.t__DOT__cyc))) {
                        if ((1U & (IData)(vlSelfRef.t__DOT__cyc))) {
                            if (VL_UNLIKELY(((0U != (IData)(vlSelfRef.t__DOT__out))))) {
                                VL_STOP_MT("/root/reference/test_regress/t/t_clk_dsp.v", 68, "");
                            }
                        } else if (VL_UNLIKELY(((0U 
                                                 != (IData)(vlSelfRef.t__DOT__out))))) {
                            VL_STOP_MT("/root/reference/test_regress/t/t_clk_dsp.v", 67, "");
                        }
                    } else if ((1U & (IData)(vlSelfRef.t__DOT__cyc))) {
                        if (VL_UNLIKELY(((0U != (IData)(vlSelfRef.t__DOT__out))))) {
                            VL_STOP_MT("/root/reference/test_regress/t/t_clk_dsp.v", 66, "");
                        }
                    } else if (VL_UNLIKELY(((8U != (IData)(vlSelfRef.t__DOT__out))))) {
                        VL_STOP_MT("/root/reference/test_regress/t/t_clk_dsp.v", 65, "");
                    }
                } else if ((2U & (IData)(vlSelfRef.t__DOT__cyc))) {
                    if ((1U & (IData)(vlSelfRef.t__DOT__cyc))) {
                        if (VL_UNLIKELY(((8U != (IData)(vlSelfRef.t__DOT__out))))) {
                            VL_STOP_MT("/root/reference/test_regress/t/t_clk_dsp.v", 64, "");
                        }
                    } else if (VL_UNLIKELY(((4U != (IData)(vlSelfRef.t__DOT__out))))) {
                        VL_STOP_MT("/root/reference/test_regress/t/t_clk_dsp.v", 63, "");
                    }
                } else if ((1U & (IData)(vlSelfRef.t__DOT__cyc))) {
                    if (VL_UNLIKELY(((4U != (IData)(vlSelfRef.t__DOT__out))))) {
                        VL_STOP_MT("/root/reference/test_regress/t/t_clk_dsp.v", 62, "");
                    }
                }
            } else if ((1U & (~ ((IData)(vlSelfRef.t__DOT__cyc) 
                                 >> 2U)))) {
                if ((1U & (~ ((IData)(vlSelfRef.t__DOT__cyc) 
                              >> 1U)))) {
                    if (VL_UNLIKELY(((1U & (~ (IData)(vlSelfRef.t__DOT__cyc)))))) {
                        VL_STOP_MT("/root/reference/test_regress/t/t_clk_dsp.v", 53, "");
                    }
                }
            }
        }
        vlSelfRef.t__DOT__cyc = vlSelfRef.__Vdly__t__DOT__cyc;
    }

    VLCU_D static void Vt_clk_dsp___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt_clk_dsp___024root___eval_phase__act(Vt_clk_dsp___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        {
            // Inlined CFunc: _eval_triggers_vec__act
            vlSelfRef.__VactTriggered[0U] = (QData)((IData)(
                                                            (((((IData)(vlSelfRef.t__DOT__dsp_ph2) 
                                                                & (~ (IData)(vlSelfRef.__Vtrigprevexpr___TOP__t__DOT__dsp_ph2__0))) 
                                                               << 3U) 
                                                              | (((IData)(vlSelfRef.t__DOT__dsp_reset) 
                                                                  & (~ (IData)(vlSelfRef.__Vtrigprevexpr___TOP__t__DOT__dsp_reset__0))) 
                                                                 << 2U)) 
                                                             | ((((IData)(vlSelfRef.t__DOT__dsp_ph1) 
                                                                  & (~ (IData)(vlSelfRef.__Vtrigprevexpr___TOP__t__DOT__dsp_ph1__0))) 
                                                                 << 1U) 
                                                                | ((IData)(vlSelfRef.clk) 
                                                                   & (~ (IData)(vlSelfRef.__Vtrigprevexpr___TOP__clk__0)))))));
            vlSelfRef.__Vtrigprevexpr___TOP__clk__0 
                = vlSelfRef.clk;
            vlSelfRef.__Vtrigprevexpr___TOP__t__DOT__dsp_ph1__0 
                = vlSelfRef.t__DOT__dsp_ph1;
            vlSelfRef.__Vtrigprevexpr___TOP__t__DOT__dsp_reset__0 
                = vlSelfRef.t__DOT__dsp_reset;
            vlSelfRef.__Vtrigprevexpr___TOP__t__DOT__dsp_ph2__0 
                = vlSelfRef.t__DOT__dsp_ph2;
        }
#ifdef VL_DEBUG
        if (VL_UNLIKELY(vlSymsp->_vm_contextp__->debug())) {
            Vt_clk_dsp___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vt_clk_dsp___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vt_clk_dsp___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vt_clk_dsp___024root___eval_phase__nba(Vt_clk_dsp___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vt_clk_dsp___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    {
                        // Inlined CFunc: _nba_sequent__TOP__0
                        vlSelfRef.__Vdly__t__DOT__cyc 
                            = vlSelfRef.t__DOT__cyc;
                    }
                }
                if ((6ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    {
                        // Inlined CFunc: _nba_sequent__TOP__1
                        vlSelfRef.t__DOT__dspchip__DOT__out1_r 
                            = (1U & ((~ (IData)(vlSelfRef.t__DOT__dsp_reset)) 
                                     & (~ (IData)(vlSelfRef.t__DOT__dspchip__DOT__out1_r))));
                    }
                }
                if ((0x000000000000000cULL & vlSelfRef.__VnbaTriggered[0U])) {
                    {
                        // Inlined CFunc: _nba_sequent__TOP__2
                        vlSelfRef.t__DOT__dspchip__DOT__out2_r 
                            = ((IData)(vlSelfRef.t__DOT__dsp_reset)
                                ? 0U : (IData)(vlSelfRef.t__DOT__dspchip__DOT__t_dsppla__DOT__latched_r));
                    }
                }
                if ((6ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    {
                        // Inlined CFunc: _nba_sequent__TOP__3
                        vlSelfRef.t__DOT__dspchip__DOT__t_dsppla__DOT__latched_r 
                            = ((IData)(vlSelfRef.t__DOT__dsp_reset)
                                ? 0U : (IData)(vlSelfRef.t__DOT__padd));
                    }
                }
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    Vt_clk_dsp___024root___nba_sequent__TOP__4(vlSelf, vlCtx);
                }
                if ((0x000000000000000eULL & vlSelfRef.__VnbaTriggered[0U])) {
                    {
                        // Inlined CFunc: _nba_comb__TOP__0
                        vlSelfRef.t__DOT__out = ((IData)(vlSelfRef.t__DOT__dspchip__DOT__out2_r) 
                                                 & (- (IData)(
                                                              (1U 
                                                               & (~ (IData)(vlSelfRef.t__DOT__dspchip__DOT__out1_r))))));
                    }
                }
            }
            Vt_clk_dsp___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vt_clk_dsp___024root___eval(Vt_clk_dsp___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt_clk_dsp___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("/root/reference/test_regress/t/t_clk_dsp.v", 7, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vt_clk_dsp___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("/root/reference/test_regress/t/t_clk_dsp.v", 7, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vt_clk_dsp___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vt_clk_dsp___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vt_clk_dsp___024root___ctor_var_reset(Vt_clk_dsp___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        vlSelf->t__DOT__cyc = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 15508623621340971111ull);
        vlSelf->t__DOT__padd = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 7892429414711426158ull);
        vlSelf->t__DOT__dsp_ph1 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 6443813525634039454ull);
        vlSelf->t__DOT__dsp_ph2 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 9983485472985298851ull);
        vlSelf->t__DOT__dsp_reset = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 435851799945749985ull);
        vlSelf->t__DOT__out = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 9695058467195139273ull);
        vlSelf->t__DOT__dspchip__DOT__out1_r = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 5463916196579462299ull);
        vlSelf->t__DOT__dspchip__DOT__out2_r = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 3670570264759302537ull);
        vlSelf->t__DOT__dspchip__DOT__t_dsppla__DOT__latched_r = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 11052591224198317333ull);
        vlSelf->__Vdly__t__DOT__cyc = 0;
        for (int __Vi0 = 0; __Vi0 < 1; ++__Vi0) {
            vlSelf->__VstlTriggered[__Vi0] = 0;
        }
        for (int __Vi0 = 0; __Vi0 < 1; ++__Vi0) {
            vlSelf->__VactTriggered[__Vi0] = 0;
        }
        vlSelf->__Vtrigprevexpr___TOP__clk__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__t__DOT__dsp_ph1__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__t__DOT__dsp_reset__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__t__DOT__dsp_ph2__0 = 0;
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
        Vt_clk_dsp___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt_clk_dsp___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt_clk_dsp___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt_clk_dsp___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vt_clk_dsp___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt_clk_dsp___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt_clk_dsp__cuda::info() {
    typedef Vt_clk_dsp__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"t__DOT__cyc", "t.cyc", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF_t__DOT__cyc},
        {"t__DOT__padd", "t.padd", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF_t__DOT__padd},
        {"t__DOT__dsp_ph1", "t.dsp_ph1", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__dsp_ph1},
        {"t__DOT__dsp_ph2", "t.dsp_ph2", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__dsp_ph2},
        {"t__DOT__dsp_reset", "t.dsp_reset", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__dsp_reset},
        {"t__DOT__out", "t.out", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF_t__DOT__out},
        {"t__DOT__dspchip__DOT__out1_r", "t.dspchip.out1_r", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__dspchip__DOT__out1_r},
        {"t__DOT__dspchip__DOT__out2_r", "t.dspchip.out2_r", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF_t__DOT__dspchip__DOT__out2_r},
        {"t__DOT__dspchip__DOT__t_dsppla__DOT__latched_r", "t.dspchip.t_dsppla.latched_r", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF_t__DOT__dspchip__DOT__t_dsppla__DOT__latched_r},
        {"__Vdly__t__DOT__cyc", "__Vdly__t.cyc", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF___Vdly__t__DOT__cyc},
        {"__VstlFirstIteration", "__VstlFirstIteration", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlFirstIteration},
        {"__VstlPhaseResult", "__VstlPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__Vtrigprevexpr___TOP__t__DOT__dsp_ph1__0", "__Vtrigprevexpr___TOP__t.dsp_ph1__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__t__DOT__dsp_ph1__0},
        {"__Vtrigprevexpr___TOP__t__DOT__dsp_reset__0", "__Vtrigprevexpr___TOP__t.dsp_reset__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__t__DOT__dsp_reset__0},
        {"__Vtrigprevexpr___TOP__t__DOT__dsp_ph2__0", "__Vtrigprevexpr___TOP__t.dsp_ph2__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__t__DOT__dsp_ph2__0},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"__VstlTriggered", "__VstlTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VstlTriggered},
        {"__VactTriggered", "__VactTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VactTriggered},
        {"__VnbaTriggered", "__VnbaTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VnbaTriggered},
    };
    static const char* const messages[] = {"cyc %d\n", "*-* All Finished *-*\n", nullptr};
    static const char* const plusargs[] = {nullptr};
    static const VlcuTraceDecl traceDecls[] = {
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0}};  // (terminator, not counted)
    static const uint32_t traceOrder[] = {0};  // codes in dump order (+ terminator)
    static const uint32_t saveOrder[] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "t", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 2, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 23, 0, 0x911b6693bd329ef9ULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt_clk_dsp__cuda VLCU_TESTBENCHES);
