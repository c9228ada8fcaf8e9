// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt_math_vgen"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt_math_vgen___024root final {
    CData/*0:0*/ clk;
    CData/*0:0*/ t__DOT__check;
    CData/*0:0*/ t__DOT__W1243;
    CData/*0:0*/ t__DOT__W0344;
    CData/*7:0*/ t__DOT__W0593;
    CData/*2:0*/ t__DOT__DW0515;
    CData/*2:0*/ t__DOT__W0772;
    CData/*5:0*/ t__DOT__W123_is_3f;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    IData/*31:0*/ t__DOT__cyc;
    VlWide<4>/*118:0*/ t__DOT__W0420;
    VlWide<4>/*115:0*/ t__DOT__W0421;
    IData/*31:0*/ t__DOT__W0161;
    VlWide<4>/*119:0*/ t__DOT__W0592;
    VlWide<4>/*127:0*/ t__DOT__WB1063;
    VlWide<3>/*89:0*/ t__DOT__WB1066;
    VlWide<4>/*111:0*/ t__DOT__WB1068;
    VlWide<3>/*68:0*/ t__DOT__W0792;
    VlWide<3>/*68:0*/ t__DOT__W0793;
    VlWide<3>/*83:0*/ t__DOT__W0024;
    VlWide<3>/*83:0*/ t__DOT__W0025;
    IData/*31:0*/ __VactIterCount;
    QData/*58:0*/ t__DOT__W0101;
    QData/*62:0*/ t__DOT__W0217;
    QData/*53:0*/ t__DOT__W0219;
    QData/*34:0*/ t__DOT__WB1064;
    QData/*62:0*/ t__DOT__WB1065;
    QData/*34:0*/ t__DOT__WB1067;
    QData/*62:0*/ t__DOT__W0753;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vt_math_vgen__cuda final {
    typedef Vt_math_vgen___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 224,
        OFF_t__DOT__check = 225,
        OFF_t__DOT__W1243 = 226,
        OFF_t__DOT__W0344 = 227,
        OFF_t__DOT__W0593 = 228,
        OFF_t__DOT__DW0515 = 229,
        OFF_t__DOT__W0772 = 230,
        OFF_t__DOT__W123_is_3f = 231,
        OFF___Vtrigprevexpr___TOP__clk__0 = 232,
        OFF___VactPhaseResult = 233,
        OFF___VnbaPhaseResult = 234,
        OFF_t__DOT__cyc = 72,
        OFF_t__DOT__W0420 = 76,
        OFF_t__DOT__W0421 = 92,
        OFF_t__DOT__W0161 = 108,
        OFF_t__DOT__W0592 = 112,
        OFF_t__DOT__WB1063 = 128,
        OFF_t__DOT__WB1066 = 144,
        OFF_t__DOT__WB1068 = 156,
        OFF_t__DOT__W0792 = 172,
        OFF_t__DOT__W0793 = 184,
        OFF_t__DOT__W0024 = 196,
        OFF_t__DOT__W0025 = 208,
        OFF___VactIterCount = 220,
        OFF_t__DOT__W0101 = 0,
        OFF_t__DOT__W0217 = 8,
        OFF_t__DOT__W0219 = 16,
        OFF_t__DOT__WB1064 = 24,
        OFF_t__DOT__WB1065 = 32,
        OFF_t__DOT__WB1067 = 40,
        OFF_t__DOT__W0753 = 48,
        OFF___VactTriggered = 56,
        OFF___VnbaTriggered = 64,
        kStateBytes = 240,  // S_persist
        kTouchedReadBytes = 235,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 234,  // S_state_wr
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
        s.t__DOT__check = st.ld<VlcuMem8>(OFF_t__DOT__check, i);
        s.t__DOT__W1243 = st.ld<VlcuMem8>(OFF_t__DOT__W1243, i);
        s.t__DOT__W0344 = st.ld<VlcuMem8>(OFF_t__DOT__W0344, i);
        s.t__DOT__W0593 = st.ld<VlcuMem8>(OFF_t__DOT__W0593, i);
        s.t__DOT__DW0515 = st.ld<VlcuMem8>(OFF_t__DOT__DW0515, i);
        s.t__DOT__W0772 = st.ld<VlcuMem8>(OFF_t__DOT__W0772, i);
        s.t__DOT__W123_is_3f = st.ld<VlcuMem8>(OFF_t__DOT__W123_is_3f, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        s.t__DOT__cyc = st.ld<IData>(OFF_t__DOT__cyc, i);
        st.ldW(OFF_t__DOT__W0420, i, s.t__DOT__W0420);
        st.ldW(OFF_t__DOT__W0421, i, s.t__DOT__W0421);
        s.t__DOT__W0161 = st.ld<IData>(OFF_t__DOT__W0161, i);
        st.ldW(OFF_t__DOT__W0592, i, s.t__DOT__W0592);
        st.ldW(OFF_t__DOT__WB1063, i, s.t__DOT__WB1063);
        st.ldW(OFF_t__DOT__WB1066, i, s.t__DOT__WB1066);
        st.ldW(OFF_t__DOT__WB1068, i, s.t__DOT__WB1068);
        st.ldW(OFF_t__DOT__W0792, i, s.t__DOT__W0792);
        st.ldW(OFF_t__DOT__W0793, i, s.t__DOT__W0793);
        st.ldW(OFF_t__DOT__W0024, i, s.t__DOT__W0024);
        st.ldW(OFF_t__DOT__W0025, i, s.t__DOT__W0025);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        s.t__DOT__W0101 = st.ld<QData>(OFF_t__DOT__W0101, i);
        s.t__DOT__W0217 = st.ld<QData>(OFF_t__DOT__W0217, i);
        s.t__DOT__W0219 = st.ld<QData>(OFF_t__DOT__W0219, i);
        s.t__DOT__WB1064 = st.ld<QData>(OFF_t__DOT__WB1064, i);
        s.t__DOT__WB1065 = st.ld<QData>(OFF_t__DOT__WB1065, i);
        s.t__DOT__WB1067 = st.ld<QData>(OFF_t__DOT__WB1067, i);
        s.t__DOT__W0753 = st.ld<QData>(OFF_t__DOT__W0753, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.st<VlcuMem8>(OFF_t__DOT__check, i, s.t__DOT__check);
        st.st<VlcuMem8>(OFF_t__DOT__W1243, i, s.t__DOT__W1243);
        st.st<VlcuMem8>(OFF_t__DOT__W0344, i, s.t__DOT__W0344);
        st.st<VlcuMem8>(OFF_t__DOT__W0593, i, s.t__DOT__W0593);
        st.st<VlcuMem8>(OFF_t__DOT__DW0515, i, s.t__DOT__DW0515);
        st.st<VlcuMem8>(OFF_t__DOT__W0772, i, s.t__DOT__W0772);
        st.st<VlcuMem8>(OFF_t__DOT__W123_is_3f, i, s.t__DOT__W123_is_3f);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.st<IData>(OFF_t__DOT__cyc, i, s.t__DOT__cyc);
        st.stW(OFF_t__DOT__W0420, i, s.t__DOT__W0420);
        st.stW(OFF_t__DOT__W0421, i, s.t__DOT__W0421);
        st.st<IData>(OFF_t__DOT__W0161, i, s.t__DOT__W0161);
        st.stW(OFF_t__DOT__W0592, i, s.t__DOT__W0592);
        st.stW(OFF_t__DOT__WB1063, i, s.t__DOT__WB1063);
        st.stW(OFF_t__DOT__WB1066, i, s.t__DOT__WB1066);
        st.stW(OFF_t__DOT__WB1068, i, s.t__DOT__WB1068);
        st.stW(OFF_t__DOT__W0792, i, s.t__DOT__W0792);
        st.stW(OFF_t__DOT__W0793, i, s.t__DOT__W0793);
        st.stW(OFF_t__DOT__W0024, i, s.t__DOT__W0024);
        st.stW(OFF_t__DOT__W0025, i, s.t__DOT__W0025);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
        st.st<QData>(OFF_t__DOT__W0101, i, s.t__DOT__W0101);
        st.st<QData>(OFF_t__DOT__W0217, i, s.t__DOT__W0217);
        st.st<QData>(OFF_t__DOT__W0219, i, s.t__DOT__W0219);
        st.st<QData>(OFF_t__DOT__WB1064, i, s.t__DOT__WB1064);
        st.st<QData>(OFF_t__DOT__WB1065, i, s.t__DOT__WB1065);
        st.st<QData>(OFF_t__DOT__WB1067, i, s.t__DOT__WB1067);
        st.st<QData>(OFF_t__DOT__W0753, i, s.t__DOT__W0753);
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

    VLCU_D static void Vt_math_vgen___024root___eval_static(Vt_math_vgen___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
    }

    VLCU_D static void Vt_math_vgen___024root___eval_initial(Vt_math_vgen___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        {
            // Inlined CFunc: _eval_initial__TOP
            vlSelfRef.t__DOT__cyc = 1U;
            vlSelfRef.t__DOT__check = 0U;
        }
    }

    VLCU_D static void Vt_math_vgen___024root___eval_final(Vt_math_vgen___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_math_vgen___024root___eval_settle(Vt_math_vgen___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static bool Vt_math_vgen___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt_math_vgen___024root___nba_sequent__TOP__0(Vt_math_vgen___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        IData/*31:0*/ __Vdly__t__DOT__cyc;
        __Vdly__t__DOT__cyc = 0;
        VlWide<3>/*95:0*/ __Vtemp_4;
        VlWide<4>/*127:0*/ __Vtemp_9;
        VlWide<4>/*127:0*/ __Vtemp_10;
        VlWide<3>/*95:0*/ __Vtemp_17;
        VlWide<3>/*95:0*/ __Vtemp_18;
        IData/*31:0*/ __VExpandSel_LoShift_2;
        CData/*0:0*/ __VExpandSel_Aligned_2;
        IData/*31:0*/ __VExpandSel_HiShift_2;
        IData/*31:0*/ __VExpandSel_HiMask_2;
        IData/*31:0*/ __VExpandSel_LoShift_3;
        CData/*0:0*/ __VExpandSel_Aligned_3;
        IData/*31:0*/ __VExpandSel_HiShift_3;
        IData/*31:0*/ __VExpandSel_HiMask_3;
        // Body
        __Vdly__t__DOT__cyc = vlSelfRef.t__DOT__cyc;
        if ((1U & (~ ((IData)(vlSelfRef.t__DOT__DW0515) 
                      >> 1U)))) {
            if (VL_UNLIKELY((vlSelfRef.t__DOT__check))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_vgen.v", 240, "");
            }
        }
        if ((0xffffffffU != VL_SHIFTRS_III(32,32,6, 
                                           (~ VL_SHIFTL_III(32,32,6, (IData)(0x088d1bcbU), (IData)(vlSelfRef.t__DOT__W123_is_3f))), 0x3fU))) {
            if (VL_UNLIKELY((vlSelfRef.t__DOT__check))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_vgen.v", 273, "");
            }
        }
        if ((0U != (((0x51247e0cU ^ vlSelfRef.t__DOT__W0792[0U]) 
                     | (0x93515695U ^ vlSelfRef.t__DOT__W0792[1U])) 
                    | (0x00000016U ^ vlSelfRef.t__DOT__W0792[2U])))) {
            if (VL_UNLIKELY((vlSelfRef.t__DOT__check))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_vgen.v", 231, "");
            }
        }
        if ((1U & (~ ((IData)(vlSelfRef.t__DOT__W0772) 
                      >> VL_GTS_IQQ(63, 0x0876c66a7e29fabfULL, vlSelfRef.t__DOT__W0753))))) {
            if (VL_UNLIKELY((vlSelfRef.t__DOT__check))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_vgen.v", 251, "");
            }
        }
        if ((1U & (~ ((IData)(vlSelfRef.t__DOT__W0772) 
                      >> 1U)))) {
            if (VL_UNLIKELY((vlSelfRef.t__DOT__check))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_vgen.v", 252, "");
            }
        }
        VL_MODDIVS_WWW(84, __Vtemp_4, vlSelfRef.t__DOT__W0024, vlSelfRef.t__DOT__W0025);
        if ((0U != (((0xe1fe9094U ^ __Vtemp_4[0U]) 
                     | __Vtemp_4[1U]) | (0x000fffffU 
                                         & __Vtemp_4[2U])))) {
            if (VL_UNLIKELY((vlSelfRef.t__DOT__check))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_vgen.v", 292, "");
            }
        }
        if ((2U == vlSelfRef.t__DOT__cyc)) {
            if ((0x00000000ffffffffULL != vlSelfRef.t__DOT__W0101)) {
                if (VL_UNLIKELY((vlSelfRef.t__DOT__check))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_vgen.v", 31, "");
                }
            }
            if ((1U > (- (IData)(vlSelfRef.t__DOT__W1243)))) {
                if (VL_UNLIKELY((vlSelfRef.t__DOT__check))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_vgen.v", 46, "");
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_vgen.v", 47, "");
                }
            }
            if (vlSelfRef.t__DOT__W0344) {
                if (VL_UNLIKELY((vlSelfRef.t__DOT__check))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_vgen.v", 61, "");
                }
            }
            if ((1U & (~ (vlSelfRef.t__DOT__W0420[(1U 
                                                   & ((- 
                                                       ((vlSelfRef.t__DOT__W0421[3U] 
                                                         << 0x00000012U) 
                                                        | (vlSelfRef.t__DOT__W0421[3U] 
                                                           >> 0x0000000eU))) 
                                                      >> 5U))] 
                          >> (0x0000001fU & (- ((vlSelfRef.t__DOT__W0421[3U] 
                                                 << 0x00000012U) 
                                                | (vlSelfRef.t__DOT__W0421[3U] 
                                                   >> 0x0000000eU)))))))) {
                if (VL_UNLIKELY((vlSelfRef.t__DOT__check))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_vgen.v", 80, "");
                }
            }
            __Vtemp_9[0U] = ((vlSelfRef.t__DOT__W0592[1U] 
                              << 0x00000017U) | (vlSelfRef.t__DOT__W0592[0U] 
                                                 >> 9U));
            __Vtemp_9[1U] = ((vlSelfRef.t__DOT__W0592[2U] 
                              << 0x00000017U) | (vlSelfRef.t__DOT__W0592[1U] 
                                                 >> 9U));
            __Vtemp_9[2U] = ((vlSelfRef.t__DOT__W0592[3U] 
                              << 0x00000017U) | (vlSelfRef.t__DOT__W0592[2U] 
                                                 >> 9U));
            __Vtemp_9[3U] = (0x00007fffU & (vlSelfRef.t__DOT__W0592[3U] 
                                            >> 9U));
            VL_SHIFTR_WWI(111,111,8, __Vtemp_10, __Vtemp_9, (IData)(vlSelfRef.t__DOT__W0593));
            if ((0U != (((__Vtemp_10[0U] | __Vtemp_10[1U]) 
                         | __Vtemp_10[2U]) | __Vtemp_10[3U]))) {
                if (VL_UNLIKELY((vlSelfRef.t__DOT__check))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_vgen.v", 116, "");
                }
            }
            if ((0xffffffffU != vlSelfRef.t__DOT__W0161)) {
                if (VL_UNLIKELY((vlSelfRef.t__DOT__check))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_vgen.v", 99, "");
                }
            }
            if ((0U != (0x0000007fU & ((- vlSelfRef.t__DOT__W0161) 
                                       & ((IData)((vlSelfRef.t__DOT__W0217 
                                                   >> 0x1fU)) 
                                          & (IData)(
                                                    (vlSelfRef.t__DOT__W0219 
                                                     >> 0x2fU))))))) {
                if (VL_UNLIKELY((vlSelfRef.t__DOT__check))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_vgen.v", 100, "");
                }
            }
        }
        if ((9U == vlSelfRef.t__DOT__cyc)) {
            if ((0U != (((vlSelfRef.t__DOT__WB1068[0U] 
                          | vlSelfRef.t__DOT__WB1068[1U]) 
                         | vlSelfRef.t__DOT__WB1068[2U]) 
                        | vlSelfRef.t__DOT__WB1068[3U]))) {
                if (VL_UNLIKELY((vlSelfRef.t__DOT__check))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_vgen.v", 163, "");
                }
            }
            __VExpandSel_LoShift_2 = (7U & (IData)(
                                                   (vlSelfRef.t__DOT__WB1064 
                                                    >> 
                                                    (0x0000000fU 
                                                     & (IData)(
                                                               (vlSelfRef.t__DOT__WB1065 
                                                                >> 4U))))));
            __VExpandSel_Aligned_2 = (0U == __VExpandSel_LoShift_2);
            if (__VExpandSel_Aligned_2) {
                __VExpandSel_HiShift_2 = 0U;
                __VExpandSel_HiMask_2 = 0U;
            } else {
                __VExpandSel_HiShift_2 = ((IData)(0x00000020U) 
                                          - __VExpandSel_LoShift_2);
                __VExpandSel_HiMask_2 = 0xffffffffU;
            }
            if ((0U != ((((((vlSelfRef.t__DOT__WB1063[1U] 
                             << __VExpandSel_HiShift_2) 
                            & __VExpandSel_HiMask_2) 
                           | (vlSelfRef.t__DOT__WB1063[0U] 
                              >> __VExpandSel_LoShift_2)) 
                          | (((vlSelfRef.t__DOT__WB1063[2U] 
                               << __VExpandSel_HiShift_2) 
                              & __VExpandSel_HiMask_2) 
                             | (vlSelfRef.t__DOT__WB1063[1U] 
                                >> __VExpandSel_LoShift_2))) 
                         | (((vlSelfRef.t__DOT__WB1063[3U] 
                              << __VExpandSel_HiShift_2) 
                             & __VExpandSel_HiMask_2) 
                            | (vlSelfRef.t__DOT__WB1063[2U] 
                               >> __VExpandSel_LoShift_2))) 
                        | (0x0000ffffU & (vlSelfRef.t__DOT__WB1063[3U] 
                                          >> __VExpandSel_LoShift_2))))) {
                if (VL_UNLIKELY((vlSelfRef.t__DOT__check))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_vgen.v", 164, "");
                }
            }
        }
        vlSelfRef.t__DOT__DW0515 = 7U;
        vlSelfRef.t__DOT__W123_is_3f = 0x3fU;
        __Vtemp_17[0U] = 0xf15f509eU;
        __Vtemp_17[1U] = 0x0e989f3eU;
        __Vtemp_17[2U] = 0x0000001fU;
        VL_MULS_WWW(69, __Vtemp_18, __Vtemp_17, vlSelfRef.t__DOT__W0793);
        vlSelfRef.t__DOT__W0792[0U] = __Vtemp_18[0U];
        vlSelfRef.t__DOT__W0792[1U] = __Vtemp_18[1U];
        vlSelfRef.t__DOT__W0792[2U] = (0x0000001fU 
                                       & __Vtemp_18[2U]);
        vlSelfRef.t__DOT__W0772 = 7U;
        vlSelfRef.t__DOT__W0753 = 0x004e20004ed93e26ULL;
        vlSelfRef.t__DOT__W0024[0U] = 0xe1fe9094U;
        vlSelfRef.t__DOT__W0024[1U] = 0U;
        vlSelfRef.t__DOT__W0024[2U] = 0U;
        vlSelfRef.t__DOT__W0025[0U] = 0x308b3d7cU;
        vlSelfRef.t__DOT__W0025[1U] = 0xfffffffeU;
        vlSelfRef.t__DOT__W0025[2U] = 0x0000f66aU;
        if ((1U == vlSelfRef.t__DOT__cyc)) {
            vlSelfRef.t__DOT__W0101 = 0x00000000ffffffffULL;
            vlSelfRef.t__DOT__W1243 = 1U;
            vlSelfRef.t__DOT__W0344 = 0U;
            vlSelfRef.t__DOT__W0421[0U] = 0U;
            vlSelfRef.t__DOT__W0421[1U] = 0U;
            vlSelfRef.t__DOT__W0421[2U] = 0U;
            vlSelfRef.t__DOT__W0421[3U] = 0U;
            vlSelfRef.t__DOT__W0420[0U] = 0xffffffffU;
            vlSelfRef.t__DOT__W0420[1U] = 0xffffffffU;
            vlSelfRef.t__DOT__W0420[2U] = 0xffffffffU;
            vlSelfRef.t__DOT__W0420[3U] = 0x007fffffU;
            vlSelfRef.t__DOT__W0161 = 0xffffffffU;
            vlSelfRef.t__DOT__W0217 = 0ULL;
            vlSelfRef.t__DOT__W0219 = 0ULL;
            vlSelfRef.t__DOT__W0593 = 0x70U;
            vlSelfRef.t__DOT__W0592[0U] = 0U;
            vlSelfRef.t__DOT__W0592[1U] = 0U;
            vlSelfRef.t__DOT__W0592[2U] = 0U;
            vlSelfRef.t__DOT__W0592[3U] = 0U;
        }
        if ((0U != vlSelfRef.t__DOT__cyc)) {
            __Vdly__t__DOT__cyc = ((IData)(1U) + vlSelfRef.t__DOT__cyc);
            if ((0x00000012U == vlSelfRef.t__DOT__cyc)) {
                vlSelfRef.t__DOT__check = 1U;
            }
            if (VL_UNLIKELY(((0x00000014U == vlSelfRef.t__DOT__cyc)))) {
                vlCtx.writef(0U, 0U);
                VL_FINISH_MT("/root/reference/test_regress/t/t_math_vgen.v", 305, "");
            }
        }
        if ((1U == vlSelfRef.t__DOT__cyc)) {
            vlSelfRef.t__DOT__WB1067 = 0x00000007ffffffffULL;
            vlSelfRef.t__DOT__WB1066[0U] = 1U;
            vlSelfRef.t__DOT__WB1066[1U] = 0U;
            vlSelfRef.t__DOT__WB1066[2U] = 0U;
        }
        if ((5U == vlSelfRef.t__DOT__cyc)) {
            __VExpandSel_LoShift_3 = (7U & (IData)(
                                                   (vlSelfRef.t__DOT__WB1064 
                                                    >> 
                                                    (0x0000000fU 
                                                     & (IData)(
                                                               (vlSelfRef.t__DOT__WB1065 
                                                                >> 4U))))));
            __VExpandSel_Aligned_3 = (0U == __VExpandSel_LoShift_3);
            if (__VExpandSel_Aligned_3) {
                __VExpandSel_HiShift_3 = 0U;
                __VExpandSel_HiMask_3 = 0U;
            } else {
                __VExpandSel_HiShift_3 = ((IData)(0x00000020U) 
                                          - __VExpandSel_LoShift_3);
                __VExpandSel_HiMask_3 = 0xffffffffU;
            }
            vlSelfRef.t__DOT__WB1068[0U] = (((vlSelfRef.t__DOT__WB1063[1U] 
                                              << __VExpandSel_HiShift_3) 
                                             & __VExpandSel_HiMask_3) 
                                            | (vlSelfRef.t__DOT__WB1063[0U] 
                                               >> __VExpandSel_LoShift_3));
            vlSelfRef.t__DOT__WB1068[1U] = (((vlSelfRef.t__DOT__WB1063[2U] 
                                              << __VExpandSel_HiShift_3) 
                                             & __VExpandSel_HiMask_3) 
                                            | (vlSelfRef.t__DOT__WB1063[1U] 
                                               >> __VExpandSel_LoShift_3));
            vlSelfRef.t__DOT__WB1068[2U] = (((vlSelfRef.t__DOT__WB1063[3U] 
                                              << __VExpandSel_HiShift_3) 
                                             & __VExpandSel_HiMask_3) 
                                            | (vlSelfRef.t__DOT__WB1063[2U] 
                                               >> __VExpandSel_LoShift_3));
            vlSelfRef.t__DOT__WB1068[3U] = (0x0000ffffU 
                                            & (vlSelfRef.t__DOT__WB1063[3U] 
                                               >> __VExpandSel_LoShift_3));
        }
        if ((2U == vlSelfRef.t__DOT__cyc)) {
            vlSelfRef.t__DOT__WB1065 = (0x7fffffffffffffffULL 
                                        & (((QData)((IData)(vlSelfRef.t__DOT__WB1066[2U])) 
                                            << 0x00000025U) 
                                           | (((QData)((IData)(vlSelfRef.t__DOT__WB1066[1U])) 
                                               << 5U) 
                                              | ((QData)((IData)(vlSelfRef.t__DOT__WB1066[0U])) 
                                                 >> 0x0000001bU))));
        }
        vlSelfRef.t__DOT__W0793[0U] = 0x4eb1a91aU;
        vlSelfRef.t__DOT__W0793[1U] = 0xffffffffU;
        vlSelfRef.t__DOT__W0793[2U] = 0x0000001fU;
        if ((4U == vlSelfRef.t__DOT__cyc)) {
            vlSelfRef.t__DOT__WB1063[0U] = 1U;
            vlSelfRef.t__DOT__WB1063[1U] = 0U;
            vlSelfRef.t__DOT__WB1063[2U] = 0U;
            vlSelfRef.t__DOT__WB1063[3U] = 0U;
        }
        if ((3U == vlSelfRef.t__DOT__cyc)) {
            vlSelfRef.t__DOT__WB1064 = vlSelfRef.t__DOT__WB1067;
        }
        vlSelfRef.t__DOT__cyc = __Vdly__t__DOT__cyc;
    }

    VLCU_D static void Vt_math_vgen___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt_math_vgen___024root___eval_phase__act(Vt_math_vgen___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt_math_vgen___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vt_math_vgen___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vt_math_vgen___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vt_math_vgen___024root___eval_phase__nba(Vt_math_vgen___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vt_math_vgen___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    Vt_math_vgen___024root___nba_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
            Vt_math_vgen___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vt_math_vgen___024root___eval(Vt_math_vgen___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt_math_vgen___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("/root/reference/test_regress/t/t_math_vgen.v", 7, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vt_math_vgen___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("/root/reference/test_regress/t/t_math_vgen.v", 7, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vt_math_vgen___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vt_math_vgen___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vt_math_vgen___024root___ctor_var_reset(Vt_math_vgen___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        vlSelf->t__DOT__cyc = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 15508623621340971111ull);
        vlSelf->t__DOT__check = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 3884252797450265655ull);
        vlSelf->t__DOT__W0101 = VL_SCOPED_RAND_RESET_Q(59, __VscopeHash, 9011926962591223105ull);
        vlSelf->t__DOT__W1243 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 808898171089397119ull);
        vlSelf->t__DOT__W0344 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 10810341623752518954ull);
        VL_SCOPED_RAND_RESET_W(119, vlSelf->t__DOT__W0420, __VscopeHash, 15381742542538016590ull);
        VL_SCOPED_RAND_RESET_W(116, vlSelf->t__DOT__W0421, __VscopeHash, 11142363016469817538ull);
        vlSelf->t__DOT__W0161 = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 7507768286456264343ull);
        vlSelf->t__DOT__W0217 = VL_SCOPED_RAND_RESET_Q(63, __VscopeHash, 16095459557959691770ull);
        vlSelf->t__DOT__W0219 = VL_SCOPED_RAND_RESET_Q(54, __VscopeHash, 10883416889689074456ull);
        VL_SCOPED_RAND_RESET_W(120, vlSelf->t__DOT__W0592, __VscopeHash, 11602595709250438554ull);
        vlSelf->t__DOT__W0593 = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 6692047709031579340ull);
        VL_SCOPED_RAND_RESET_W(128, vlSelf->t__DOT__WB1063, __VscopeHash, 3180661227264021301ull);
        vlSelf->t__DOT__WB1064 = VL_SCOPED_RAND_RESET_Q(35, __VscopeHash, 12763728245063856103ull);
        vlSelf->t__DOT__WB1065 = VL_SCOPED_RAND_RESET_Q(63, __VscopeHash, 13212149570744578732ull);
        VL_SCOPED_RAND_RESET_W(90, vlSelf->t__DOT__WB1066, __VscopeHash, 5101399657448204183ull);
        vlSelf->t__DOT__WB1067 = VL_SCOPED_RAND_RESET_Q(35, __VscopeHash, 7914842415927453015ull);
        VL_SCOPED_RAND_RESET_W(112, vlSelf->t__DOT__WB1068, __VscopeHash, 7567983453049059226ull);
        VL_SCOPED_RAND_RESET_W(69, vlSelf->t__DOT__W0792, __VscopeHash, 2272491696769227557ull);
        VL_SCOPED_RAND_RESET_W(69, vlSelf->t__DOT__W0793, __VscopeHash, 6528100985757451332ull);
        vlSelf->t__DOT__DW0515 = VL_SCOPED_RAND_RESET_I(3, __VscopeHash, 5034227539424009929ull);
        vlSelf->t__DOT__W0753 = VL_SCOPED_RAND_RESET_Q(63, __VscopeHash, 15075452037044007879ull);
        vlSelf->t__DOT__W0772 = VL_SCOPED_RAND_RESET_I(3, __VscopeHash, 7184854979887603741ull);
        vlSelf->t__DOT__W123_is_3f = VL_SCOPED_RAND_RESET_I(6, __VscopeHash, 2799312502924863718ull);
        VL_SCOPED_RAND_RESET_W(84, vlSelf->t__DOT__W0024, __VscopeHash, 5617741843454706871ull);
        VL_SCOPED_RAND_RESET_W(84, vlSelf->t__DOT__W0025, __VscopeHash, 9873351132442951315ull);
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
        Vt_math_vgen___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_vgen___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_vgen___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_vgen___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_vgen___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_vgen___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt_math_vgen__cuda::info() {
    typedef Vt_math_vgen__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"t__DOT__check", "t.check", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__check},
        {"t__DOT__W1243", "t.W1243", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__W1243},
        {"t__DOT__W0344", "t.W0344", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__W0344},
        {"t__DOT__W0593", "t.W0593", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF_t__DOT__W0593},
        {"t__DOT__DW0515", "t.DW0515", VLCU_MEMBER_STATE, 1, 0, 1, 3, 2, 0, T::OFF_t__DOT__DW0515},
        {"t__DOT__W0772", "t.W0772", VLCU_MEMBER_STATE, 1, 0, 1, 3, 2, 0, T::OFF_t__DOT__W0772},
        {"t__DOT__W123_is_3f", "t.W123_is_3f", VLCU_MEMBER_STATE, 1, 0, 1, 6, 5, 0, T::OFF_t__DOT__W123_is_3f},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"t__DOT__cyc", "t.cyc", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__cyc},
        {"t__DOT__W0420", "t.W0420", VLCU_MEMBER_STATE, 4, 0, 4, 119, 118, 0, T::OFF_t__DOT__W0420},
        {"t__DOT__W0421", "t.W0421", VLCU_MEMBER_STATE, 4, 0, 4, 116, 115, 0, T::OFF_t__DOT__W0421},
        {"t__DOT__W0161", "t.W0161", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__W0161},
        {"t__DOT__W0592", "t.W0592", VLCU_MEMBER_STATE, 4, 0, 4, 120, 119, 0, T::OFF_t__DOT__W0592},
        {"t__DOT__WB1063", "t.WB1063", VLCU_MEMBER_STATE, 4, 0, 4, 128, 127, 0, T::OFF_t__DOT__WB1063},
        {"t__DOT__WB1066", "t.WB1066", VLCU_MEMBER_STATE, 4, 0, 3, 90, 89, 0, T::OFF_t__DOT__WB1066},
        {"t__DOT__WB1068", "t.WB1068", VLCU_MEMBER_STATE, 4, 0, 4, 112, 111, 0, T::OFF_t__DOT__WB1068},
        {"t__DOT__W0792", "t.W0792", VLCU_MEMBER_STATE, 4, 0, 3, 69, 68, 0, T::OFF_t__DOT__W0792},
        {"t__DOT__W0793", "t.W0793", VLCU_MEMBER_STATE, 4, 0, 3, 69, 68, 0, T::OFF_t__DOT__W0793},
        {"t__DOT__W0024", "t.W0024", VLCU_MEMBER_STATE, 4, 0, 3, 84, 83, 0, T::OFF_t__DOT__W0024},
        {"t__DOT__W0025", "t.W0025", VLCU_MEMBER_STATE, 4, 0, 3, 84, 83, 0, T::OFF_t__DOT__W0025},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"t__DOT__W0101", "t.W0101", VLCU_MEMBER_STATE, 8, 0, 1, 59, 58, 0, T::OFF_t__DOT__W0101},
        {"t__DOT__W0217", "t.W0217", VLCU_MEMBER_STATE, 8, 0, 1, 63, 62, 0, T::OFF_t__DOT__W0217},
        {"t__DOT__W0219", "t.W0219", VLCU_MEMBER_STATE, 8, 0, 1, 54, 53, 0, T::OFF_t__DOT__W0219},
        {"t__DOT__WB1064", "t.WB1064", VLCU_MEMBER_STATE, 8, 0, 1, 35, 34, 0, T::OFF_t__DOT__WB1064},
        {"t__DOT__WB1065", "t.WB1065", VLCU_MEMBER_STATE, 8, 0, 1, 63, 62, 0, T::OFF_t__DOT__WB1065},
        {"t__DOT__WB1067", "t.WB1067", VLCU_MEMBER_STATE, 8, 0, 1, 35, 34, 0, T::OFF_t__DOT__WB1067},
        {"t__DOT__W0753", "t.W0753", VLCU_MEMBER_STATE, 8, 0, 1, 63, 62, 0, T::OFF_t__DOT__W0753},
        {"__VactTriggered", "__VactTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VactTriggered},
        {"__VnbaTriggered", "__VnbaTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VnbaTriggered},
    };
    static const char* const messages[] = {"*-* All Finished *-*\n", nullptr};
    static const char* const plusargs[] = {nullptr};
    static const VlcuTraceDecl traceDecls[] = {
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0}};  // (terminator, not counted)
    static const uint32_t traceOrder[] = {0};  // codes in dump order (+ terminator)
    static const uint32_t saveOrder[] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "t", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 1, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 33, 0, 0xf0c2b11053185e24ULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt_math_vgen__cuda VLCU_TESTBENCHES);
