// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt_const_hi"
#endif

// Constant pool
static __device__ const VlWide<33>/*1055:0*/ Vt_const_hi__ConstPool__CONST_h11ca9c09_0 = VlWide<33>{{
        0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff,
        0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff,
        0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff,
        0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff,
        0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff,
        0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff,
        0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff,
        0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff,
        0x03ffffff
}};
static __device__ const VlUnpacked<VlWide<33>/*1049:0*/, 4> Vt_const_hi__ConstPool__TABLE_h852c65ce_0 = {{
    VlWide<33>{{
            0x00000000, 0x00000000, 0x00000000, 0x00000000,
            0x00000000, 0x00000000, 0x00000000, 0x00000000,
            0x00000000, 0x00000000, 0x00000000, 0x00000000,
            0x00000000, 0x00000000, 0x00000000, 0x00000000,
            0x00000000, 0x00000000, 0x00000000, 0x00000000,
            0x00000000, 0x00000000, 0x00000000, 0x00000000,
            0x00000000, 0x00000000, 0x00000000, 0x00000000,
            0x00000000, 0x00000000, 0x00000000, 0x00000000,
            0x00000000
    }},
    VlWide<33>{{
            0x00000000, 0x00000000, 0x00000000, 0x00000000,
            0x00000000, 0x00000000, 0x00000000, 0x00000000,
            0x00000000, 0x00000000, 0x00000000, 0x00000000,
            0x00000000, 0x00000000, 0x00000000, 0x00000000,
            0x00000000, 0x00000000, 0x00000000, 0x00000000,
            0x00000000, 0x00000000, 0x00000000, 0x00000000,
            0x00000000, 0x00000000, 0x00000000, 0x00000000,
            0x00000000, 0x00000000, 0x00000000, 0x03fffc00,
            0x03fffc00
    }},
    VlWide<33>{{
            0xffff0000, 0xffff0000, 0x00000000, 0x00000000,
            0x00000000, 0x00000000, 0x00000000, 0x00000000,
            0x00000000, 0x00000000, 0x00000000, 0x00000000,
            0x00000000, 0x00000000, 0x00000000, 0x00000000,
            0x00000000, 0x00000000, 0x00000000, 0x00000000,
            0x00000000, 0x00000000, 0x00000000, 0x00000000,
            0x00000000, 0x00000000, 0x00000000, 0x00000000,
            0x00000000, 0x00000000, 0x00000000, 0x00000000,
            0x00000000
    }},
    VlWide<33>{{
            0xffff0000, 0xffff0000, 0x00000000, 0x00000000,
            0x00000000, 0x00000000, 0x00000000, 0x00000000,
            0x00000000, 0x00000000, 0x00000000, 0x00000000,
            0x00000000, 0x00000000, 0x00000000, 0x00000000,
            0x00000000, 0x00000000, 0x00000000, 0x00000000,
            0x00000000, 0x00000000, 0x00000000, 0x00000000,
            0x00000000, 0x00000000, 0x00000000, 0x00000000,
            0x00000000, 0x00000000, 0x00000000, 0x00000000,
            0x00000000
    }}
}};

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt_const_hi___024root final {
    CData/*0:0*/ clk;
    CData/*1:0*/ t__DOT__reg_i;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    IData/*31:0*/ t__DOT__cyc;
    VlWide<33>/*1049:0*/ t__DOT__pad0;
    VlWide<33>/*1049:0*/ t__DOT__spad1;
    IData/*31:0*/ __VactIterCount;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vt_const_hi__cuda final {
    typedef Vt_const_hi___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 288,
        OFF_t__DOT__reg_i = 289,
        OFF___Vtrigprevexpr___TOP__clk__0 = 290,
        OFF___VactPhaseResult = 291,
        OFF___VnbaPhaseResult = 292,
        OFF_t__DOT__cyc = 16,
        OFF_t__DOT__pad0 = 20,
        OFF_t__DOT__spad1 = 152,
        OFF___VactIterCount = 284,
        OFF___VactTriggered = 0,
        OFF___VnbaTriggered = 8,
        kStateBytes = 296,  // S_persist
        kTouchedReadBytes = 293,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 292,  // S_state_wr
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
        s.t__DOT__reg_i = st.ld<VlcuMem8>(OFF_t__DOT__reg_i, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        s.t__DOT__cyc = st.ld<IData>(OFF_t__DOT__cyc, i);
        st.ldW(OFF_t__DOT__pad0, i, s.t__DOT__pad0);
        st.ldW(OFF_t__DOT__spad1, i, s.t__DOT__spad1);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.st<VlcuMem8>(OFF_t__DOT__reg_i, i, s.t__DOT__reg_i);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.st<IData>(OFF_t__DOT__cyc, i, s.t__DOT__cyc);
        st.stW(OFF_t__DOT__pad0, i, s.t__DOT__pad0);
        st.stW(OFF_t__DOT__spad1, i, s.t__DOT__spad1);
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

    VLCU_D static void Vt_const_hi___024root___eval_static(Vt_const_hi___024root* vlSelf, VlcuContext& vlCtx) {
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

    VLCU_D static void Vt_const_hi___024root___eval_initial(Vt_const_hi___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_const_hi___024root___eval_final(Vt_const_hi___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_const_hi___024root___eval_settle(Vt_const_hi___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static bool Vt_const_hi___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt_const_hi___024root___nba_sequent__TOP__0(Vt_const_hi___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        IData/*31:0*/ __Vdly__t__DOT__cyc;
        __Vdly__t__DOT__cyc = 0;
        CData/*1:0*/ __Vdly__t__DOT__reg_i;
        __Vdly__t__DOT__reg_i = 0;
        VlWide<33>/*1049:0*/ __Vdly__t__DOT__pad0;
        VL_ZERO_W(1050, __Vdly__t__DOT__pad0);
        VlWide<33>/*1049:0*/ __Vdly__t__DOT__spad1;
        VL_ZERO_W(1050, __Vdly__t__DOT__spad1);
        // Body
        __Vdly__t__DOT__cyc = vlSelfRef.t__DOT__cyc;
        __Vdly__t__DOT__reg_i = vlSelfRef.t__DOT__reg_i;
        __Vdly__t__DOT__pad0[0U] = vlSelfRef.t__DOT__pad0[0U];
        __Vdly__t__DOT__pad0[1U] = vlSelfRef.t__DOT__pad0[1U];
        __Vdly__t__DOT__pad0[2U] = vlSelfRef.t__DOT__pad0[2U];
        __Vdly__t__DOT__pad0[3U] = vlSelfRef.t__DOT__pad0[3U];
        __Vdly__t__DOT__pad0[4U] = vlSelfRef.t__DOT__pad0[4U];
        __Vdly__t__DOT__pad0[5U] = vlSelfRef.t__DOT__pad0[5U];
        __Vdly__t__DOT__pad0[6U] = vlSelfRef.t__DOT__pad0[6U];
        __Vdly__t__DOT__pad0[7U] = vlSelfRef.t__DOT__pad0[7U];
        __Vdly__t__DOT__pad0[8U] = vlSelfRef.t__DOT__pad0[8U];
        __Vdly__t__DOT__pad0[9U] = vlSelfRef.t__DOT__pad0[9U];
        __Vdly__t__DOT__pad0[10U] = vlSelfRef.t__DOT__pad0[10U];
        __Vdly__t__DOT__pad0[11U] = vlSelfRef.t__DOT__pad0[11U];
        __Vdly__t__DOT__pad0[12U] = vlSelfRef.t__DOT__pad0[12U];
        __Vdly__t__DOT__pad0[13U] = vlSelfRef.t__DOT__pad0[13U];
        __Vdly__t__DOT__pad0[14U] = vlSelfRef.t__DOT__pad0[14U];
        __Vdly__t__DOT__pad0[15U] = vlSelfRef.t__DOT__pad0[15U];
        __Vdly__t__DOT__pad0[16U] = vlSelfRef.t__DOT__pad0[16U];
        __Vdly__t__DOT__pad0[17U] = vlSelfRef.t__DOT__pad0[17U];
        __Vdly__t__DOT__pad0[18U] = vlSelfRef.t__DOT__pad0[18U];
        __Vdly__t__DOT__pad0[19U] = vlSelfRef.t__DOT__pad0[19U];
        __Vdly__t__DOT__pad0[20U] = vlSelfRef.t__DOT__pad0[20U];
        __Vdly__t__DOT__pad0[21U] = vlSelfRef.t__DOT__pad0[21U];
        __Vdly__t__DOT__pad0[22U] = vlSelfRef.t__DOT__pad0[22U];
        __Vdly__t__DOT__pad0[23U] = vlSelfRef.t__DOT__pad0[23U];
        __Vdly__t__DOT__pad0[24U] = vlSelfRef.t__DOT__pad0[24U];
        __Vdly__t__DOT__pad0[25U] = vlSelfRef.t__DOT__pad0[25U];
        __Vdly__t__DOT__pad0[26U] = vlSelfRef.t__DOT__pad0[26U];
        __Vdly__t__DOT__pad0[27U] = vlSelfRef.t__DOT__pad0[27U];
        __Vdly__t__DOT__pad0[28U] = vlSelfRef.t__DOT__pad0[28U];
        __Vdly__t__DOT__pad0[29U] = vlSelfRef.t__DOT__pad0[29U];
        __Vdly__t__DOT__pad0[30U] = vlSelfRef.t__DOT__pad0[30U];
        __Vdly__t__DOT__pad0[31U] = vlSelfRef.t__DOT__pad0[31U];
        __Vdly__t__DOT__pad0[32U] = vlSelfRef.t__DOT__pad0[32U];
        __Vdly__t__DOT__spad1[0U] = vlSelfRef.t__DOT__spad1[0U];
        __Vdly__t__DOT__spad1[1U] = vlSelfRef.t__DOT__spad1[1U];
        __Vdly__t__DOT__spad1[2U] = vlSelfRef.t__DOT__spad1[2U];
        __Vdly__t__DOT__spad1[3U] = vlSelfRef.t__DOT__spad1[3U];
        __Vdly__t__DOT__spad1[4U] = vlSelfRef.t__DOT__spad1[4U];
        __Vdly__t__DOT__spad1[5U] = vlSelfRef.t__DOT__spad1[5U];
        __Vdly__t__DOT__spad1[6U] = vlSelfRef.t__DOT__spad1[6U];
        __Vdly__t__DOT__spad1[7U] = vlSelfRef.t__DOT__spad1[7U];
        __Vdly__t__DOT__spad1[8U] = vlSelfRef.t__DOT__spad1[8U];
        __Vdly__t__DOT__spad1[9U] = vlSelfRef.t__DOT__spad1[9U];
        __Vdly__t__DOT__spad1[10U] = vlSelfRef.t__DOT__spad1[10U];
        __Vdly__t__DOT__spad1[11U] = vlSelfRef.t__DOT__spad1[11U];
        __Vdly__t__DOT__spad1[12U] = vlSelfRef.t__DOT__spad1[12U];
        __Vdly__t__DOT__spad1[13U] = vlSelfRef.t__DOT__spad1[13U];
        __Vdly__t__DOT__spad1[14U] = vlSelfRef.t__DOT__spad1[14U];
        __Vdly__t__DOT__spad1[15U] = vlSelfRef.t__DOT__spad1[15U];
        __Vdly__t__DOT__spad1[16U] = vlSelfRef.t__DOT__spad1[16U];
        __Vdly__t__DOT__spad1[17U] = vlSelfRef.t__DOT__spad1[17U];
        __Vdly__t__DOT__spad1[18U] = vlSelfRef.t__DOT__spad1[18U];
        __Vdly__t__DOT__spad1[19U] = vlSelfRef.t__DOT__spad1[19U];
        __Vdly__t__DOT__spad1[20U] = vlSelfRef.t__DOT__spad1[20U];
        __Vdly__t__DOT__spad1[21U] = vlSelfRef.t__DOT__spad1[21U];
        __Vdly__t__DOT__spad1[22U] = vlSelfRef.t__DOT__spad1[22U];
        __Vdly__t__DOT__spad1[23U] = vlSelfRef.t__DOT__spad1[23U];
        __Vdly__t__DOT__spad1[24U] = vlSelfRef.t__DOT__spad1[24U];
        __Vdly__t__DOT__spad1[25U] = vlSelfRef.t__DOT__spad1[25U];
        __Vdly__t__DOT__spad1[26U] = vlSelfRef.t__DOT__spad1[26U];
        __Vdly__t__DOT__spad1[27U] = vlSelfRef.t__DOT__spad1[27U];
        __Vdly__t__DOT__spad1[28U] = vlSelfRef.t__DOT__spad1[28U];
        __Vdly__t__DOT__spad1[29U] = vlSelfRef.t__DOT__spad1[29U];
        __Vdly__t__DOT__spad1[30U] = vlSelfRef.t__DOT__spad1[30U];
        __Vdly__t__DOT__spad1[31U] = vlSelfRef.t__DOT__spad1[31U];
        __Vdly__t__DOT__spad1[32U] = vlSelfRef.t__DOT__spad1[32U];
        __Vdly__t__DOT__cyc = ((IData)(1U) + vlSelfRef.t__DOT__cyc);
        if ((0U == vlSelfRef.t__DOT__cyc)) {
            __Vdly__t__DOT__reg_i = 0U;
            VL_ASSIGN_W(1050, __Vdly__t__DOT__pad0, Vt_const_hi__ConstPool__CONST_h11ca9c09_0);
            VL_ASSIGN_W(1050, __Vdly__t__DOT__spad1, Vt_const_hi__ConstPool__CONST_h11ca9c09_0);
        } else if ((1U == vlSelfRef.t__DOT__cyc)) {
            __Vdly__t__DOT__reg_i = 1U;
        } else if ((2U == vlSelfRef.t__DOT__cyc)) {
            if (VL_UNLIKELY(((0U != (((((((((((((((
                                                   (((((((((((((((((Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                                    [vlSelfRef.t__DOT__reg_i][0U] 
                                                                    | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                                    [vlSelfRef.t__DOT__reg_i][1U]) 
                                                                   | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                                   [vlSelfRef.t__DOT__reg_i][2U]) 
                                                                  | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                                  [vlSelfRef.t__DOT__reg_i][3U]) 
                                                                 | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                                 [vlSelfRef.t__DOT__reg_i][4U]) 
                                                                | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                                [vlSelfRef.t__DOT__reg_i][5U]) 
                                                               | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                               [vlSelfRef.t__DOT__reg_i][6U]) 
                                                              | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                              [vlSelfRef.t__DOT__reg_i][7U]) 
                                                             | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                             [vlSelfRef.t__DOT__reg_i][8U]) 
                                                            | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                            [vlSelfRef.t__DOT__reg_i][9U]) 
                                                           | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                           [vlSelfRef.t__DOT__reg_i][10U]) 
                                                          | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                          [vlSelfRef.t__DOT__reg_i][11U]) 
                                                         | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                         [vlSelfRef.t__DOT__reg_i][12U]) 
                                                        | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                        [vlSelfRef.t__DOT__reg_i][13U]) 
                                                       | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                       [vlSelfRef.t__DOT__reg_i][14U]) 
                                                      | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                      [vlSelfRef.t__DOT__reg_i][15U]) 
                                                     | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                     [vlSelfRef.t__DOT__reg_i][16U]) 
                                                    | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                    [vlSelfRef.t__DOT__reg_i][17U]) 
                                                   | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                   [vlSelfRef.t__DOT__reg_i][18U]) 
                                                  | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                  [vlSelfRef.t__DOT__reg_i][19U]) 
                                                 | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                 [vlSelfRef.t__DOT__reg_i][20U]) 
                                                | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                [vlSelfRef.t__DOT__reg_i][21U]) 
                                               | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                               [vlSelfRef.t__DOT__reg_i][22U]) 
                                              | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                              [vlSelfRef.t__DOT__reg_i][23U]) 
                                             | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                             [vlSelfRef.t__DOT__reg_i][24U]) 
                                            | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                            [vlSelfRef.t__DOT__reg_i][25U]) 
                                           | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                           [vlSelfRef.t__DOT__reg_i][26U]) 
                                          | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                          [vlSelfRef.t__DOT__reg_i][27U]) 
                                         | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                         [vlSelfRef.t__DOT__reg_i][28U]) 
                                        | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                        [vlSelfRef.t__DOT__reg_i][29U]) 
                                       | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                       [vlSelfRef.t__DOT__reg_i][30U]) 
                                      | (0x03fffc00U 
                                         ^ Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                         [vlSelfRef.t__DOT__reg_i][31U])) 
                                     | (0x03fffc00U 
                                        ^ Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                        [vlSelfRef.t__DOT__reg_i][32U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_const_hi.v", 39, "");
            }
            __Vdly__t__DOT__reg_i = 2U;
        } else if (VL_UNLIKELY(((0x00000063U == vlSelfRef.t__DOT__cyc)))) {
            if (VL_UNLIKELY(((0U != (((((((((((((((
                                                   ((((((((((((((((((0xffff0000U 
                                                                     ^ Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                                     [vlSelfRef.t__DOT__reg_i][0U]) 
                                                                    | (0xffff0000U 
                                                                       ^ Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                                       [vlSelfRef.t__DOT__reg_i][1U])) 
                                                                   | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                                   [vlSelfRef.t__DOT__reg_i][2U]) 
                                                                  | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                                  [vlSelfRef.t__DOT__reg_i][3U]) 
                                                                 | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                                 [vlSelfRef.t__DOT__reg_i][4U]) 
                                                                | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                                [vlSelfRef.t__DOT__reg_i][5U]) 
                                                               | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                               [vlSelfRef.t__DOT__reg_i][6U]) 
                                                              | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                              [vlSelfRef.t__DOT__reg_i][7U]) 
                                                             | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                             [vlSelfRef.t__DOT__reg_i][8U]) 
                                                            | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                            [vlSelfRef.t__DOT__reg_i][9U]) 
                                                           | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                           [vlSelfRef.t__DOT__reg_i][10U]) 
                                                          | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                          [vlSelfRef.t__DOT__reg_i][11U]) 
                                                         | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                         [vlSelfRef.t__DOT__reg_i][12U]) 
                                                        | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                        [vlSelfRef.t__DOT__reg_i][13U]) 
                                                       | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                       [vlSelfRef.t__DOT__reg_i][14U]) 
                                                      | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                      [vlSelfRef.t__DOT__reg_i][15U]) 
                                                     | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                     [vlSelfRef.t__DOT__reg_i][16U]) 
                                                    | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                    [vlSelfRef.t__DOT__reg_i][17U]) 
                                                   | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                   [vlSelfRef.t__DOT__reg_i][18U]) 
                                                  | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                  [vlSelfRef.t__DOT__reg_i][19U]) 
                                                 | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                 [vlSelfRef.t__DOT__reg_i][20U]) 
                                                | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                                [vlSelfRef.t__DOT__reg_i][21U]) 
                                               | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                               [vlSelfRef.t__DOT__reg_i][22U]) 
                                              | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                              [vlSelfRef.t__DOT__reg_i][23U]) 
                                             | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                             [vlSelfRef.t__DOT__reg_i][24U]) 
                                            | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                            [vlSelfRef.t__DOT__reg_i][25U]) 
                                           | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                           [vlSelfRef.t__DOT__reg_i][26U]) 
                                          | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                          [vlSelfRef.t__DOT__reg_i][27U]) 
                                         | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                         [vlSelfRef.t__DOT__reg_i][28U]) 
                                        | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                        [vlSelfRef.t__DOT__reg_i][29U]) 
                                       | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                       [vlSelfRef.t__DOT__reg_i][30U]) 
                                      | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                      [vlSelfRef.t__DOT__reg_i][31U]) 
                                     | Vt_const_hi__ConstPool__TABLE_h852c65ce_0
                                     [vlSelfRef.t__DOT__reg_i][32U]))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_const_hi.v", 43, "");
            }
            if (VL_UNLIKELY(((0U != (((((((((((((((
                                                   ((((((((((((((((((~ vlSelfRef.t__DOT__pad0[0U]) 
                                                                    | (~ vlSelfRef.t__DOT__pad0[1U])) 
                                                                   | (~ vlSelfRef.t__DOT__pad0[2U])) 
                                                                  | (~ vlSelfRef.t__DOT__pad0[3U])) 
                                                                 | (~ vlSelfRef.t__DOT__pad0[4U])) 
                                                                | (~ vlSelfRef.t__DOT__pad0[5U])) 
                                                               | (~ vlSelfRef.t__DOT__pad0[6U])) 
                                                              | (~ vlSelfRef.t__DOT__pad0[7U])) 
                                                             | (~ vlSelfRef.t__DOT__pad0[8U])) 
                                                            | (~ vlSelfRef.t__DOT__pad0[9U])) 
                                                           | (~ vlSelfRef.t__DOT__pad0[10U])) 
                                                          | (~ vlSelfRef.t__DOT__pad0[11U])) 
                                                         | (~ vlSelfRef.t__DOT__pad0[12U])) 
                                                        | (~ vlSelfRef.t__DOT__pad0[13U])) 
                                                       | (~ vlSelfRef.t__DOT__pad0[14U])) 
                                                      | (~ vlSelfRef.t__DOT__pad0[15U])) 
                                                     | (~ vlSelfRef.t__DOT__pad0[16U])) 
                                                    | (~ vlSelfRef.t__DOT__pad0[17U])) 
                                                   | (~ vlSelfRef.t__DOT__pad0[18U])) 
                                                  | (~ vlSelfRef.t__DOT__pad0[19U])) 
                                                 | (~ vlSelfRef.t__DOT__pad0[20U])) 
                                                | (~ vlSelfRef.t__DOT__pad0[21U])) 
                                               | (~ vlSelfRef.t__DOT__pad0[22U])) 
                                              | (~ vlSelfRef.t__DOT__pad0[23U])) 
                                             | (~ vlSelfRef.t__DOT__pad0[24U])) 
                                            | (~ vlSelfRef.t__DOT__pad0[25U])) 
                                           | (~ vlSelfRef.t__DOT__pad0[26U])) 
                                          | (~ vlSelfRef.t__DOT__pad0[27U])) 
                                         | (~ vlSelfRef.t__DOT__pad0[28U])) 
                                        | (~ vlSelfRef.t__DOT__pad0[29U])) 
                                       | (~ vlSelfRef.t__DOT__pad0[30U])) 
                                      | (~ vlSelfRef.t__DOT__pad0[31U])) 
                                     | (0x03ffffffU 
                                        ^ vlSelfRef.t__DOT__pad0[32U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_const_hi.v", 44, "");
            }
            if (VL_UNLIKELY(((0U != (((((((((((((((
                                                   ((((((((((((((((((~ vlSelfRef.t__DOT__spad1[0U]) 
                                                                    | (~ vlSelfRef.t__DOT__spad1[1U])) 
                                                                   | (~ vlSelfRef.t__DOT__spad1[2U])) 
                                                                  | (~ vlSelfRef.t__DOT__spad1[3U])) 
                                                                 | (~ vlSelfRef.t__DOT__spad1[4U])) 
                                                                | (~ vlSelfRef.t__DOT__spad1[5U])) 
                                                               | (~ vlSelfRef.t__DOT__spad1[6U])) 
                                                              | (~ vlSelfRef.t__DOT__spad1[7U])) 
                                                             | (~ vlSelfRef.t__DOT__spad1[8U])) 
                                                            | (~ vlSelfRef.t__DOT__spad1[9U])) 
                                                           | (~ vlSelfRef.t__DOT__spad1[10U])) 
                                                          | (~ vlSelfRef.t__DOT__spad1[11U])) 
                                                         | (~ vlSelfRef.t__DOT__spad1[12U])) 
                                                        | (~ vlSelfRef.t__DOT__spad1[13U])) 
                                                       | (~ vlSelfRef.t__DOT__spad1[14U])) 
                                                      | (~ vlSelfRef.t__DOT__spad1[15U])) 
                                                     | (~ vlSelfRef.t__DOT__spad1[16U])) 
                                                    | (~ vlSelfRef.t__DOT__spad1[17U])) 
                                                   | (~ vlSelfRef.t__DOT__spad1[18U])) 
                                                  | (~ vlSelfRef.t__DOT__spad1[19U])) 
                                                 | (~ vlSelfRef.t__DOT__spad1[20U])) 
                                                | (~ vlSelfRef.t__DOT__spad1[21U])) 
                                               | (~ vlSelfRef.t__DOT__spad1[22U])) 
                                              | (~ vlSelfRef.t__DOT__spad1[23U])) 
                                             | (~ vlSelfRef.t__DOT__spad1[24U])) 
                                            | (~ vlSelfRef.t__DOT__spad1[25U])) 
                                           | (~ vlSelfRef.t__DOT__spad1[26U])) 
                                          | (~ vlSelfRef.t__DOT__spad1[27U])) 
                                         | (~ vlSelfRef.t__DOT__spad1[28U])) 
                                        | (~ vlSelfRef.t__DOT__spad1[29U])) 
                                       | (~ vlSelfRef.t__DOT__spad1[30U])) 
                                      | (~ vlSelfRef.t__DOT__spad1[31U])) 
                                     | (0x03ffffffU 
                                        ^ vlSelfRef.t__DOT__spad1[32U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_const_hi.v", 45, "");
            }
            vlCtx.writef(0U, 0U);
            VL_FINISH_MT("/root/reference/test_regress/t/t_const_hi.v", 47, "");
        }
        vlSelfRef.t__DOT__cyc = __Vdly__t__DOT__cyc;
        vlSelfRef.t__DOT__reg_i = __Vdly__t__DOT__reg_i;
        vlSelfRef.t__DOT__pad0[0U] = __Vdly__t__DOT__pad0[0U];
        vlSelfRef.t__DOT__pad0[1U] = __Vdly__t__DOT__pad0[1U];
        vlSelfRef.t__DOT__pad0[2U] = __Vdly__t__DOT__pad0[2U];
        vlSelfRef.t__DOT__pad0[3U] = __Vdly__t__DOT__pad0[3U];
        vlSelfRef.t__DOT__pad0[4U] = __Vdly__t__DOT__pad0[4U];
        vlSelfRef.t__DOT__pad0[5U] = __Vdly__t__DOT__pad0[5U];
        vlSelfRef.t__DOT__pad0[6U] = __Vdly__t__DOT__pad0[6U];
        vlSelfRef.t__DOT__pad0[7U] = __Vdly__t__DOT__pad0[7U];
        vlSelfRef.t__DOT__pad0[8U] = __Vdly__t__DOT__pad0[8U];
        vlSelfRef.t__DOT__pad0[9U] = __Vdly__t__DOT__pad0[9U];
        vlSelfRef.t__DOT__pad0[10U] = __Vdly__t__DOT__pad0[10U];
        vlSelfRef.t__DOT__pad0[11U] = __Vdly__t__DOT__pad0[11U];
        vlSelfRef.t__DOT__pad0[12U] = __Vdly__t__DOT__pad0[12U];
        vlSelfRef.t__DOT__pad0[13U] = __Vdly__t__DOT__pad0[13U];
        vlSelfRef.t__DOT__pad0[14U] = __Vdly__t__DOT__pad0[14U];
        vlSelfRef.t__DOT__pad0[15U] = __Vdly__t__DOT__pad0[15U];
        vlSelfRef.t__DOT__pad0[16U] = __Vdly__t__DOT__pad0[16U];
        vlSelfRef.t__DOT__pad0[17U] = __Vdly__t__DOT__pad0[17U];
        vlSelfRef.t__DOT__pad0[18U] = __Vdly__t__DOT__pad0[18U];
        vlSelfRef.t__DOT__pad0[19U] = __Vdly__t__DOT__pad0[19U];
        vlSelfRef.t__DOT__pad0[20U] = __Vdly__t__DOT__pad0[20U];
        vlSelfRef.t__DOT__pad0[21U] = __Vdly__t__DOT__pad0[21U];
        vlSelfRef.t__DOT__pad0[22U] = __Vdly__t__DOT__pad0[22U];
        vlSelfRef.t__DOT__pad0[23U] = __Vdly__t__DOT__pad0[23U];
        vlSelfRef.t__DOT__pad0[24U] = __Vdly__t__DOT__pad0[24U];
        vlSelfRef.t__DOT__pad0[25U] = __Vdly__t__DOT__pad0[25U];
        vlSelfRef.t__DOT__pad0[26U] = __Vdly__t__DOT__pad0[26U];
        vlSelfRef.t__DOT__pad0[27U] = __Vdly__t__DOT__pad0[27U];
        vlSelfRef.t__DOT__pad0[28U] = __Vdly__t__DOT__pad0[28U];
        vlSelfRef.t__DOT__pad0[29U] = __Vdly__t__DOT__pad0[29U];
        vlSelfRef.t__DOT__pad0[30U] = __Vdly__t__DOT__pad0[30U];
        vlSelfRef.t__DOT__pad0[31U] = __Vdly__t__DOT__pad0[31U];
        vlSelfRef.t__DOT__pad0[32U] = __Vdly__t__DOT__pad0[32U];
        vlSelfRef.t__DOT__spad1[0U] = __Vdly__t__DOT__spad1[0U];
        vlSelfRef.t__DOT__spad1[1U] = __Vdly__t__DOT__spad1[1U];
        vlSelfRef.t__DOT__spad1[2U] = __Vdly__t__DOT__spad1[2U];
        vlSelfRef.t__DOT__spad1[3U] = __Vdly__t__DOT__spad1[3U];
        vlSelfRef.t__DOT__spad1[4U] = __Vdly__t__DOT__spad1[4U];
        vlSelfRef.t__DOT__spad1[5U] = __Vdly__t__DOT__spad1[5U];
        vlSelfRef.t__DOT__spad1[6U] = __Vdly__t__DOT__spad1[6U];
        vlSelfRef.t__DOT__spad1[7U] = __Vdly__t__DOT__spad1[7U];
        vlSelfRef.t__DOT__spad1[8U] = __Vdly__t__DOT__spad1[8U];
        vlSelfRef.t__DOT__spad1[9U] = __Vdly__t__DOT__spad1[9U];
        vlSelfRef.t__DOT__spad1[10U] = __Vdly__t__DOT__spad1[10U];
        vlSelfRef.t__DOT__spad1[11U] = __Vdly__t__DOT__spad1[11U];
        vlSelfRef.t__DOT__spad1[12U] = __Vdly__t__DOT__spad1[12U];
        vlSelfRef.t__DOT__spad1[13U] = __Vdly__t__DOT__spad1[13U];
        vlSelfRef.t__DOT__spad1[14U] = __Vdly__t__DOT__spad1[14U];
        vlSelfRef.t__DOT__spad1[15U] = __Vdly__t__DOT__spad1[15U];
        vlSelfRef.t__DOT__spad1[16U] = __Vdly__t__DOT__spad1[16U];
        vlSelfRef.t__DOT__spad1[17U] = __Vdly__t__DOT__spad1[17U];
        vlSelfRef.t__DOT__spad1[18U] = __Vdly__t__DOT__spad1[18U];
        vlSelfRef.t__DOT__spad1[19U] = __Vdly__t__DOT__spad1[19U];
        vlSelfRef.t__DOT__spad1[20U] = __Vdly__t__DOT__spad1[20U];
        vlSelfRef.t__DOT__spad1[21U] = __Vdly__t__DOT__spad1[21U];
        vlSelfRef.t__DOT__spad1[22U] = __Vdly__t__DOT__spad1[22U];
        vlSelfRef.t__DOT__spad1[23U] = __Vdly__t__DOT__spad1[23U];
        vlSelfRef.t__DOT__spad1[24U] = __Vdly__t__DOT__spad1[24U];
        vlSelfRef.t__DOT__spad1[25U] = __Vdly__t__DOT__spad1[25U];
        vlSelfRef.t__DOT__spad1[26U] = __Vdly__t__DOT__spad1[26U];
        vlSelfRef.t__DOT__spad1[27U] = __Vdly__t__DOT__spad1[27U];
        vlSelfRef.t__DOT__spad1[28U] = __Vdly__t__DOT__spad1[28U];
        vlSelfRef.t__DOT__spad1[29U] = __Vdly__t__DOT__spad1[29U];
        vlSelfRef.t__DOT__spad1[30U] = __Vdly__t__DOT__spad1[30U];
        vlSelfRef.t__DOT__spad1[31U] = __Vdly__t__DOT__spad1[31U];
        vlSelfRef.t__DOT__spad1[32U] = __Vdly__t__DOT__spad1[32U];
    }

    VLCU_D static void Vt_const_hi___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt_const_hi___024root___eval_phase__act(Vt_const_hi___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt_const_hi___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vt_const_hi___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vt_const_hi___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vt_const_hi___024root___eval_phase__nba(Vt_const_hi___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vt_const_hi___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    Vt_const_hi___024root___nba_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
            Vt_const_hi___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vt_const_hi___024root___eval(Vt_const_hi___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt_const_hi___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("/root/reference/test_regress/t/t_const_hi.v", 7, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vt_const_hi___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("/root/reference/test_regress/t/t_const_hi.v", 7, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vt_const_hi___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vt_const_hi___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vt_const_hi___024root___ctor_var_reset(Vt_const_hi___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        vlSelf->t__DOT__reg_i = VL_SCOPED_RAND_RESET_I(2, __VscopeHash, 509144586799950995ull);
        VL_SCOPED_RAND_RESET_W(1050, vlSelf->t__DOT__pad0, __VscopeHash, 14828470054082254ull);
        VL_SCOPED_RAND_RESET_W(1050, vlSelf->t__DOT__spad1, __VscopeHash, 14937427165637450479ull);
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
        Vt_const_hi___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt_const_hi___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt_const_hi___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt_const_hi___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vt_const_hi___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt_const_hi___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt_const_hi__cuda::info() {
    typedef Vt_const_hi__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"t__DOT__reg_i", "t.reg_i", VLCU_MEMBER_STATE, 1, 0, 1, 2, 1, 0, T::OFF_t__DOT__reg_i},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"t__DOT__cyc", "t.cyc", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__cyc},
        {"t__DOT__pad0", "t.pad0", VLCU_MEMBER_STATE, 4, 0, 33, 1050, 1049, 0, T::OFF_t__DOT__pad0},
        {"t__DOT__spad1", "t.spad1", VLCU_MEMBER_STATE, 4, 0, 33, 1050, 1049, 0, T::OFF_t__DOT__spad1},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"__VactTriggered", "__VactTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VactTriggered},
        {"__VnbaTriggered", "__VnbaTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VnbaTriggered},
    };
    static const char* const messages[] = {"*-* All Finished *-*\n", nullptr};
    static const char* const plusargs[] = {nullptr};
    static const VlcuTraceDecl traceDecls[] = {
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0}};  // (terminator, not counted)
    static const uint32_t traceOrder[] = {0};  // codes in dump order (+ terminator)
    static const uint32_t saveOrder[] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "t", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 1, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 11, 0, 0x6ddc4b72c7b00336ULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt_const_hi__cuda VLCU_TESTBENCHES);
