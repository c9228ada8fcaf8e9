// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt_clk_powerdn"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt_clk_powerdn___024root final {
    CData/*0:0*/ clk;
    CData/*0:0*/ t__DOT__reset_l;
    CData/*0:0*/ t__DOT__clkgate_e2r;
    CData/*0:0*/ t__DOT__clkgate_e1r_l;
    CData/*0:0*/ t__DOT__clkgate_e1f;
    CData/*7:0*/ t__DOT__cyc;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*0:0*/ __Vtrigprevexpr___TOP__t__DOT__reset_l__0;
    CData/*0:0*/ __Vtrigprevexpr_h6a4fec9a__1;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    IData/*31:0*/ t__DOT__countgated;
    IData/*31:0*/ t__DOT__count;
    IData/*31:0*/ __Vdly__t__DOT__count;
    IData/*31:0*/ __Vdly__t__DOT__countgated;
    IData/*31:0*/ __VactIterCount;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vt_clk_powerdn__cuda final {
    typedef Vt_clk_powerdn___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 36,
        OFF_t__DOT__reset_l = 37,
        OFF_t__DOT__clkgate_e2r = 38,
        OFF_t__DOT__clkgate_e1r_l = 39,
        OFF_t__DOT__clkgate_e1f = 40,
        OFF_t__DOT__cyc = 41,
        OFF___Vtrigprevexpr___TOP__clk__0 = 42,
        OFF___Vtrigprevexpr___TOP__t__DOT__reset_l__0 = 43,
        OFF___Vtrigprevexpr_h6a4fec9a__1 = 44,
        OFF___VactPhaseResult = 45,
        OFF___VnbaPhaseResult = 46,
        OFF_t__DOT__countgated = 16,
        OFF_t__DOT__count = 20,
        OFF___Vdly__t__DOT__count = 24,
        OFF___Vdly__t__DOT__countgated = 28,
        OFF___VactIterCount = 32,
        OFF___VactTriggered = 0,
        OFF___VnbaTriggered = 8,
        kStateBytes = 48,  // S_persist
        kTouchedReadBytes = 47,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 46,  // S_state_wr
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
        s.t__DOT__reset_l = st.ld<VlcuMem8>(OFF_t__DOT__reset_l, i);
        s.t__DOT__clkgate_e2r = st.ld<VlcuMem8>(OFF_t__DOT__clkgate_e2r, i);
        s.t__DOT__clkgate_e1r_l = st.ld<VlcuMem8>(OFF_t__DOT__clkgate_e1r_l, i);
        s.t__DOT__clkgate_e1f = st.ld<VlcuMem8>(OFF_t__DOT__clkgate_e1f, i);
        s.t__DOT__cyc = st.ld<VlcuMem8>(OFF_t__DOT__cyc, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__Vtrigprevexpr___TOP__t__DOT__reset_l__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__t__DOT__reset_l__0, i);
        s.__Vtrigprevexpr_h6a4fec9a__1 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr_h6a4fec9a__1, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        s.t__DOT__countgated = st.ld<IData>(OFF_t__DOT__countgated, i);
        s.t__DOT__count = st.ld<IData>(OFF_t__DOT__count, i);
        s.__Vdly__t__DOT__count = st.ld<IData>(OFF___Vdly__t__DOT__count, i);
        s.__Vdly__t__DOT__countgated = st.ld<IData>(OFF___Vdly__t__DOT__countgated, i);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.st<VlcuMem8>(OFF_t__DOT__reset_l, i, s.t__DOT__reset_l);
        st.st<VlcuMem8>(OFF_t__DOT__clkgate_e2r, i, s.t__DOT__clkgate_e2r);
        st.st<VlcuMem8>(OFF_t__DOT__clkgate_e1r_l, i, s.t__DOT__clkgate_e1r_l);
        st.st<VlcuMem8>(OFF_t__DOT__clkgate_e1f, i, s.t__DOT__clkgate_e1f);
        st.st<VlcuMem8>(OFF_t__DOT__cyc, i, s.t__DOT__cyc);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__t__DOT__reset_l__0, i, s.__Vtrigprevexpr___TOP__t__DOT__reset_l__0);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr_h6a4fec9a__1, i, s.__Vtrigprevexpr_h6a4fec9a__1);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.st<IData>(OFF_t__DOT__countgated, i, s.t__DOT__countgated);
        st.st<IData>(OFF_t__DOT__count, i, s.t__DOT__count);
        st.st<IData>(OFF___Vdly__t__DOT__count, i, s.__Vdly__t__DOT__count);
        st.st<IData>(OFF___Vdly__t__DOT__countgated, i, s.__Vdly__t__DOT__countgated);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
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

    VLCU_D static void Vt_clk_powerdn___024root___eval_static(Vt_clk_powerdn___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
        vlSelfRef.__Vtrigprevexpr___TOP__t__DOT__reset_l__0 
            = vlSelfRef.t__DOT__reset_l;
        vlSelfRef.__Vtrigprevexpr_h6a4fec9a__1 = ((IData)(vlSelfRef.clk) 
                                                  & (IData)(vlSelfRef.t__DOT__clkgate_e1f));
    }

    VLCU_D static void Vt_clk_powerdn___024root___eval_initial(Vt_clk_powerdn___024root* vlSelf, VlcuContext& vlCtx) {
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

    VLCU_D static void Vt_clk_powerdn___024root___eval_final(Vt_clk_powerdn___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_clk_powerdn___024root___eval_settle(Vt_clk_powerdn___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static bool Vt_clk_powerdn___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt_clk_powerdn___024root___nba_sequent__TOP__3(Vt_clk_powerdn___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*7:0*/ __Vdly__t__DOT__cyc;
        __Vdly__t__DOT__cyc = 0;
        // Body
        __Vdly__t__DOT__cyc = (0x000000ffU & ((IData)(1U) 
                                              + (IData)(vlSelfRef.t__DOT__cyc)));
        if (((((((((0U == (IData)(vlSelfRef.t__DOT__cyc)) 
                   | (1U == (IData)(vlSelfRef.t__DOT__cyc))) 
                  | (2U == (IData)(vlSelfRef.t__DOT__cyc))) 
                 | (3U == (IData)(vlSelfRef.t__DOT__cyc))) 
                | (4U == (IData)(vlSelfRef.t__DOT__cyc))) 
               | (5U == (IData)(vlSelfRef.t__DOT__cyc))) 
              | (9U == (IData)(vlSelfRef.t__DOT__cyc))) 
             | (0x0bU == (IData)(vlSelfRef.t__DOT__cyc)))) {
            if ((0U == (IData)(vlSelfRef.t__DOT__cyc))) {
                vlSelfRef.t__DOT__reset_l = 1U;
                vlSelfRef.t__DOT__clkgate_e2r = 1U;
            } else if ((1U == (IData)(vlSelfRef.t__DOT__cyc))) {
                vlSelfRef.t__DOT__reset_l = 1U;
            } else if ((2U != (IData)(vlSelfRef.t__DOT__cyc))) {
                if ((3U == (IData)(vlSelfRef.t__DOT__cyc))) {
                    vlSelfRef.t__DOT__reset_l = 0U;
                } else if ((4U != (IData)(vlSelfRef.t__DOT__cyc))) {
                    if ((5U == (IData)(vlSelfRef.t__DOT__cyc))) {
                        vlSelfRef.t__DOT__reset_l = 1U;
                    } else {
                        vlSelfRef.t__DOT__clkgate_e2r 
                            = (9U != (IData)(vlSelfRef.t__DOT__cyc));
                    }
                }
            }
        } else if (VL_UNLIKELY(((0x14U == (IData)(vlSelfRef.t__DOT__cyc))))) {
            vlCtx.writef(0U, 0U);
            VL_FINISH_MT("/root/reference/test_regress/t/t_clk_powerdn.v", 93, "");
        }
        if (VL_UNLIKELY(((0x00000080U & (IData)(vlSelfRef.t__DOT__cyc))))) {
            VL_STOP_MT("/root/reference/test_regress/t/t_clk_powerdn.v", 119, "");
        } else if (VL_UNLIKELY(((0x00000040U & (IData)(vlSelfRef.t__DOT__cyc))))) {
            VL_STOP_MT("/root/reference/test_regress/t/t_clk_powerdn.v", 119, "");
        } else if (VL_UNLIKELY(((0x00000020U & (IData)(vlSelfRef.t__DOT__cyc))))) {
            VL_STOP_MT("/root/reference/test_regress/t/t_clk_powerdn.v", 119, "");
        } else if ((0x00000010U & (IData)(vlSelfRef.t__DOT__cyc))) {
            if (VL_UNLIKELY(((8U & (IData)(vlSelfRef.t__DOT__cyc))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_clk_powerdn.v", 119, "");
            } else if ((4U & (IData)(vlSelfRef.t__DOT__cyc))) {
                if (VL_UNLIKELY(((2U & (IData)(vlSelfRef.t__DOT__cyc))))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_clk_powerdn.v", 119, "");
                } else if (VL_UNLIKELY(((1U & (IData)(vlSelfRef.t__DOT__cyc))))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_clk_powerdn.v", 119, "");
                } else if (VL_UNLIKELY((((0x0000100eU 
                                          != vlSelfRef.t__DOT__count) 
                                         | (0x0000100cU 
                                            != vlSelfRef.t__DOT__countgated))))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_clk_powerdn.v", 118, "");
                }
            } else if ((2U & (IData)(vlSelfRef.t__DOT__cyc))) {
                if ((1U & (IData)(vlSelfRef.t__DOT__cyc))) {
                    if (VL_UNLIKELY((((0x0000100dU 
                                       != vlSelfRef.t__DOT__count) 
                                      | (0x0000100bU 
                                         != vlSelfRef.t__DOT__countgated))))) {
                        VL_STOP_MT("/root/reference/test_regress/t/t_clk_powerdn.v", 117, "");
                    }
                } else if (VL_UNLIKELY((((0x0000100cU 
                                          != vlSelfRef.t__DOT__count) 
                                         | (0x0000100aU 
                                            != vlSelfRef.t__DOT__countgated))))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_clk_powerdn.v", 116, "");
                }
            } else if ((1U & (IData)(vlSelfRef.t__DOT__cyc))) {
                if (VL_UNLIKELY((((0x0000100bU != vlSelfRef.t__DOT__count) 
                                  | (0x00001009U != vlSelfRef.t__DOT__countgated))))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_clk_powerdn.v", 115, "");
                }
            } else if (VL_UNLIKELY((((0x0000100aU != vlSelfRef.t__DOT__count) 
                                     | (0x00001008U 
                                        != vlSelfRef.t__DOT__countgated))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_clk_powerdn.v", 114, "");
            }
        } else if ((8U & (IData)(vlSelfRef.t__DOT__cyc))) {
            if ((4U & (IData)(vlSelfRef.t__DOT__cyc))) {
                if ((2U & (IData)(vlSelfRef.t__DOT__cyc))) {
                    if ((1U & (IData)(vlSelfRef.t__DOT__cyc))) {
                        if (VL_UNLIKELY((((0x00001009U 
                                           != vlSelfRef.t__DOT__count) 
                                          | (0x00001007U 
                                             != vlSelfRef.t__DOT__countgated))))) {
                            VL_STOP_MT("/root/reference/test_regress/t/t_clk_powerdn.v", 113, "");
                        }
                    } else if (VL_UNLIKELY((((0x00001008U 
                                              != vlSelfRef.t__DOT__count) 
                                             | (0x00001006U 
                                                != vlSelfRef.t__DOT__countgated))))) {
                        VL_STOP_MT("/root/reference/test_regress/t/t_clk_powerdn.v", 112, "");
                    }
                } else if ((1U & (IData)(vlSelfRef.t__DOT__cyc))) {
                    if (VL_UNLIKELY((((0x00001007U 
                                       != vlSelfRef.t__DOT__count) 
                                      | (0x00001005U 
                                         != vlSelfRef.t__DOT__countgated))))) {
                        VL_STOP_MT("/root/reference/test_regress/t/t_clk_powerdn.v", 111, "");
                    }
                } else if (VL_UNLIKELY((((0x00001006U 
                                          != vlSelfRef.t__DOT__count) 
                                         | (0x00001005U 
                                            != vlSelfRef.t__DOT__countgated))))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_clk_powerdn.v", 110, "");
                }
            } else if ((2U & (IData)(vlSelfRef.t__DOT__cyc))) {
                if ((1U & (IData)(vlSelfRef.t__DOT__cyc))) {
                    if (VL_UNLIKELY((((0x00001005U 
                                       != vlSelfRef.t__DOT__count) 
                                      | (0x00001005U 
                                         != vlSelfRef.t__DOT__countgated))))) {
                        VL_STOP_MT("/root/reference/test_regress/t/t_clk_powerdn.v", 109, "");
                    }
                } else if (VL_UNLIKELY((((0x00001004U 
                                          != vlSelfRef.t__DOT__count) 
                                         | (0x00001004U 
                                            != vlSelfRef.t__DOT__countgated))))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_clk_powerdn.v", 108, "");
                }
            } else if ((1U & (IData)(vlSelfRef.t__DOT__cyc))) {
                if (VL_UNLIKELY((((0x00001003U != vlSelfRef.t__DOT__count) 
                                  | (0x00001003U != vlSelfRef.t__DOT__countgated))))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_clk_powerdn.v", 107, "");
                }
            } else if (VL_UNLIKELY((((0x00001002U != vlSelfRef.t__DOT__count) 
                                     | (0x00001002U 
                                        != vlSelfRef.t__DOT__countgated))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_clk_powerdn.v", 106, "");
            }
        } else if ((4U & (IData)(vlSelfRef.t__DOT__cyc))) {
            if ((2U & (IData)(vlSelfRef.t__DOT__cyc))) {
                if ((1U & (IData)(vlSelfRef.t__DOT__cyc))) {
                    if (VL_UNLIKELY((((0x00001001U 
                                       != vlSelfRef.t__DOT__count) 
                                      | (0x00001001U 
                                         != vlSelfRef.t__DOT__countgated))))) {
                        VL_STOP_MT("/root/reference/test_regress/t/t_clk_powerdn.v", 105, "");
                    }
                } else if (VL_UNLIKELY((((0x00001000U 
                                          != vlSelfRef.t__DOT__count) 
                                         | (0x00001000U 
                                            != vlSelfRef.t__DOT__countgated))))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_clk_powerdn.v", 104, "");
                }
            } else if ((1U & (IData)(vlSelfRef.t__DOT__cyc))) {
                if (VL_UNLIKELY((((0x00001000U != vlSelfRef.t__DOT__count) 
                                  | (0x00001000U != vlSelfRef.t__DOT__countgated))))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_clk_powerdn.v", 103, "");
                }
            } else if (VL_UNLIKELY((((0x00001000U != vlSelfRef.t__DOT__count) 
                                     | (0x00001000U 
                                        != vlSelfRef.t__DOT__countgated))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_clk_powerdn.v", 102, "");
            }
        }
        vlSelfRef.t__DOT__cyc = __Vdly__t__DOT__cyc;
    }

    VLCU_D static void Vt_clk_powerdn___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt_clk_powerdn___024root___eval_phase__act(Vt_clk_powerdn___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VactExecute;
        // Body
        {
            // Inlined CFunc: _eval_triggers_vec__act
            CData/*0:0*/ __Vinline_0__eval_triggers_vec__act___Vtrigprevexpr_h6a4fec9a__0;
            __Vinline_0__eval_triggers_vec__act___Vtrigprevexpr_h6a4fec9a__0 = 0;
            __Vinline_0__eval_triggers_vec__act___Vtrigprevexpr_h6a4fec9a__0 
                = ((IData)(vlSelfRef.clk) & (IData)(vlSelfRef.t__DOT__clkgate_e1f));
            vlSelfRef.__VactTriggered[0U] = (QData)((IData)(
                                                            ((((__Vinline_0__eval_triggers_vec__act___Vtrigprevexpr_h6a4fec9a__0 
                                                                & (~ (IData)(vlSelfRef.__Vtrigprevexpr_h6a4fec9a__1))) 
                                                               << 3U) 
                                                              | (((~ (IData)(vlSelfRef.t__DOT__reset_l)) 
                                                                  & (IData)(vlSelfRef.__Vtrigprevexpr___TOP__t__DOT__reset_l__0)) 
                                                                 << 2U)) 
                                                             | ((((IData)(vlSelfRef.clk) 
                                                                  & (~ (IData)(vlSelfRef.__Vtrigprevexpr___TOP__clk__0))) 
                                                                 << 1U) 
                                                                | ((~ (IData)(vlSelfRef.clk)) 
                                                                   & (IData)(vlSelfRef.__Vtrigprevexpr___TOP__clk__0))))));
            vlSelfRef.__Vtrigprevexpr___TOP__clk__0 
                = vlSelfRef.clk;
            vlSelfRef.__Vtrigprevexpr___TOP__t__DOT__reset_l__0 
                = vlSelfRef.t__DOT__reset_l;
            vlSelfRef.__Vtrigprevexpr_h6a4fec9a__1 
                = __Vinline_0__eval_triggers_vec__act___Vtrigprevexpr_h6a4fec9a__0;
        }
#ifdef VL_DEBUG
        if (VL_UNLIKELY(vlSymsp->_vm_contextp__->debug())) {
            Vt_clk_powerdn___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vt_clk_powerdn___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        __VactExecute = Vt_clk_powerdn___024root___trigger_anySet__act(vlSelfRef.__VactTriggered);
        if (__VactExecute) {
            {
                // Inlined CFunc: _eval_act
                if ((1ULL & vlSelfRef.__VactTriggered[0U])) {
                    {
                        // Inlined CFunc: _act_sequent__TOP__0
                        vlSelfRef.t__DOT__clkgate_e1f 
                            = (1U & ((~ (IData)(vlSelfRef.t__DOT__clkgate_e1r_l)) 
                                     | (~ (IData)(vlSelfRef.t__DOT__reset_l))));
                    }
                }
            }
        }
        return (__VactExecute);
    }

    VLCU_D static void Vt_clk_powerdn___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vt_clk_powerdn___024root___eval_phase__nba(Vt_clk_powerdn___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vt_clk_powerdn___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((2ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    {
                        // Inlined CFunc: _nba_sequent__TOP__0
                        CData/*7:0*/ __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___Vdly__t__DOT__cyc;
                        __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___Vdly__t__DOT__cyc = 0;
                        __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___Vdly__t__DOT__cyc 
                            = vlSelfRef.t__DOT__cyc;
                    }
                }
                if ((6ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    {
                        // Inlined CFunc: _nba_sequent__TOP__1
                        vlSelfRef.__Vdly__t__DOT__count 
                            = vlSelfRef.t__DOT__count;
                        vlSelfRef.__Vdly__t__DOT__count 
                            = ((IData)(vlSelfRef.t__DOT__reset_l)
                                ? ((IData)(1U) + vlSelfRef.t__DOT__count)
                                : 0x00001000U);
                        vlSelfRef.t__DOT__clkgate_e1r_l 
                            = ((IData)(vlSelfRef.t__DOT__reset_l) 
                               && (1U & (~ (IData)(vlSelfRef.t__DOT__clkgate_e2r))));
                    }
                }
                if ((0x000000000000000cULL & vlSelfRef.__VnbaTriggered[0U])) {
                    {
                        // Inlined CFunc: _nba_sequent__TOP__2
                        vlSelfRef.__Vdly__t__DOT__countgated 
                            = vlSelfRef.t__DOT__countgated;
                        vlSelfRef.__Vdly__t__DOT__countgated 
                            = ((IData)(vlSelfRef.t__DOT__reset_l)
                                ? ((IData)(1U) + vlSelfRef.t__DOT__countgated)
                                : 0x00001000U);
                    }
                }
                if ((2ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    Vt_clk_powerdn___024root___nba_sequent__TOP__3(vlSelf, vlCtx);
                }
                if ((6ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    {
                        // Inlined CFunc: _nba_sequent__TOP__4
                        vlSelfRef.t__DOT__count = vlSelfRef.__Vdly__t__DOT__count;
                    }
                }
                if ((0x000000000000000cULL & vlSelfRef.__VnbaTriggered[0U])) {
                    {
                        // Inlined CFunc: _nba_sequent__TOP__5
                        vlSelfRef.t__DOT__countgated 
                            = vlSelfRef.__Vdly__t__DOT__countgated;
                    }
                }
            }
            Vt_clk_powerdn___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vt_clk_powerdn___024root___eval(Vt_clk_powerdn___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt_clk_powerdn___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("/root/reference/test_regress/t/t_clk_powerdn.v", 7, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vt_clk_powerdn___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("/root/reference/test_regress/t/t_clk_powerdn.v", 7, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vt_clk_powerdn___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vt_clk_powerdn___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vt_clk_powerdn___024root___ctor_var_reset(Vt_clk_powerdn___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        vlSelf->t__DOT__reset_l = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 14476596066648728246ull);
        vlSelf->t__DOT__clkgate_e2r = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 2063779974173110091ull);
        vlSelf->t__DOT__clkgate_e1r_l = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 5106523499339345510ull);
        vlSelf->t__DOT__clkgate_e1f = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 4302420924902900838ull);
        vlSelf->t__DOT__countgated = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 3417870225628500299ull);
        vlSelf->t__DOT__count = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 7599402141995636686ull);
        vlSelf->t__DOT__cyc = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 15508623621340971111ull);
        vlSelf->__Vdly__t__DOT__count = 0;
        vlSelf->__Vdly__t__DOT__countgated = 0;
        for (int __Vi0 = 0; __Vi0 < 1; ++__Vi0) {
            vlSelf->__VactTriggered[__Vi0] = 0;
        }
        vlSelf->__Vtrigprevexpr___TOP__clk__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__t__DOT__reset_l__0 = 0;
        vlSelf->__Vtrigprevexpr_h6a4fec9a__1 = 0;
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
        Vt_clk_powerdn___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt_clk_powerdn___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt_clk_powerdn___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt_clk_powerdn___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vt_clk_powerdn___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt_clk_powerdn___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt_clk_powerdn__cuda::info() {
    typedef Vt_clk_powerdn__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"t__DOT__reset_l", "t.reset_l", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__reset_l},
        {"t__DOT__clkgate_e2r", "t.clkgate_e2r", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__clkgate_e2r},
        {"t__DOT__clkgate_e1r_l", "t.clkgate_e1r_l", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__clkgate_e1r_l},
        {"t__DOT__clkgate_e1f", "t.clkgate_e1f", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__clkgate_e1f},
        {"t__DOT__cyc", "t.cyc", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF_t__DOT__cyc},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__Vtrigprevexpr___TOP__t__DOT__reset_l__0", "__Vtrigprevexpr___TOP__t.reset_l__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__t__DOT__reset_l__0},
        {"__Vtrigprevexpr_h6a4fec9a__1", "__Vtrigprevexpr_h6a4fec9a__1", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr_h6a4fec9a__1},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"t__DOT__countgated", "t.countgated", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__countgated},
        {"t__DOT__count", "t.count", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__count},
        {"__Vdly__t__DOT__count", "__Vdly__t.count", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___Vdly__t__DOT__count},
        {"__Vdly__t__DOT__countgated", "__Vdly__t.countgated", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___Vdly__t__DOT__countgated},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"__VactTriggered", "__VactTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VactTriggered},
        {"__VnbaTriggered", "__VnbaTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VnbaTriggered},
    };
    static const char* const messages[] = {"*-* All Finished *-*\n", nullptr};
    static const char* const plusargs[] = {nullptr};
    static const VlcuTraceDecl traceDecls[] = {
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0}};  // (terminator, not counted)
    static const uint32_t traceOrder[] = {0};  // codes in dump order (+ terminator)
    static const uint32_t saveOrder[] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "t", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 1, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 18, 0, 0x28ea696f4f5177a9ULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt_clk_powerdn__cuda VLCU_TESTBENCHES);
