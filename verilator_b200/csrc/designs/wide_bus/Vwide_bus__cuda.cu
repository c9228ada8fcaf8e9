// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vwide_bus"
#endif

// Constant pool
static __device__ const VlWide<8>/*255:0*/ Vwide_bus__ConstPool__CONST_h4e9f510d_0 = VlWide<8>{{
        0x00000001, 0x00000000, 0x00000000, 0x00000000,
        0x00000000, 0x00000000, 0x00000000, 0x00000000
}};
static __device__ const VlWide<8>/*255:0*/ Vwide_bus__ConstPool__CONST_h9e67c271_0 = VlWide<8>{{
        0x00000000, 0x00000000, 0x00000000, 0x00000000,
        0x00000000, 0x00000000, 0x00000000, 0x00000000
}};

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vwide_bus___024root final {
    CData/*0:0*/ clk;
    CData/*0:0*/ rst_n;
    CData/*0:0*/ in_valid;
    CData/*2:0*/ in_op;
    CData/*7:0*/ in_shift;
    VlWide<8>/*255:0*/ in_data;
    CData/*0:0*/ out_valid;
    CData/*0:0*/ out_lt;
    CData/*0:0*/ out_lts;
    SData/*8:0*/ out_ones;
    VlWide<3>/*69:0*/ out_slice;
    VlWide<4>/*127:0*/ out_hash;
    VlWide<8>/*255:0*/ out_data;
    CData/*7:0*/ wide_bus__DOT__sh;
    CData/*2:0*/ wide_bus__DOT__op;
    CData/*0:0*/ wide_bus__DOT__vld;
    CData/*0:0*/ __VstlFirstIteration;
    CData/*0:0*/ __VstlPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    VlWide<8>/*255:0*/ wide_bus__DOT__acc;
    VlWide<8>/*255:0*/ wide_bus__DOT__stage;
    VlWide<4>/*127:0*/ wide_bus__DOT__hash;
    VlWide<8>/*255:0*/ wide_bus__DOT__shl;
    VlWide<4>/*127:0*/ wide_bus__DOT__prod;
    VlWide<8>/*255:0*/ wide_bus__DOT__res;
    IData/*31:0*/ __VactIterCount;
    VlUnpacked<QData/*63:0*/, 1> __VstlTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    VlUnpacked<CData/*0:0*/, 2> __Vm_traceActivity;
    const char* vlNamep;
};

struct Vwide_bus__cuda final {
    typedef Vwide_bus___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 282,
        OFF_rst_n = 283,
        OFF_in_valid = 284,
        OFF_in_op = 285,
        OFF_in_shift = 286,
        OFF_in_data = 24,
        OFF_out_valid = 287,
        OFF_out_lt = 288,
        OFF_out_lts = 289,
        OFF_out_ones = 280,
        OFF_out_slice = 56,
        OFF_out_hash = 68,
        OFF_out_data = 84,
        OFF_wide_bus__DOT__sh = 290,
        OFF_wide_bus__DOT__op = 291,
        OFF_wide_bus__DOT__vld = 292,
        OFF___VstlFirstIteration = 293,
        OFF___VstlPhaseResult = 294,
        OFF___Vtrigprevexpr___TOP__clk__0 = 295,
        OFF___VactPhaseResult = 296,
        OFF___VnbaPhaseResult = 297,
        OFF_wide_bus__DOT__acc = 116,
        OFF_wide_bus__DOT__stage = 148,
        OFF_wide_bus__DOT__hash = 180,
        OFF_wide_bus__DOT__shl = 196,
        OFF_wide_bus__DOT__prod = 228,
        OFF_wide_bus__DOT__res = 244,
        OFF___VactIterCount = 276,
        OFF___VstlTriggered = 0,
        OFF___VactTriggered = 8,
        OFF___VnbaTriggered = 16,
        OFF___Vm_traceActivity = 298,
        kStateBytes = 304,  // S_persist
        kTouchedReadBytes = 235,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 252,  // S_state_wr
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
        s.rst_n = st.ld<VlcuMem8>(OFF_rst_n, i);
        s.in_valid = st.ld<VlcuMem8>(OFF_in_valid, i);
        s.in_op = st.ld<VlcuMem8>(OFF_in_op, i);
        s.in_shift = st.ld<VlcuMem8>(OFF_in_shift, i);
        st.ldW(OFF_in_data, i, s.in_data);
        s.out_valid = st.ld<VlcuMem8>(OFF_out_valid, i);
        s.out_lt = st.ld<VlcuMem8>(OFF_out_lt, i);
        s.out_lts = st.ld<VlcuMem8>(OFF_out_lts, i);
        s.out_ones = st.ld<VlcuMem16>(OFF_out_ones, i);
        st.ldW(OFF_out_slice, i, s.out_slice);
        st.ldW(OFF_out_hash, i, s.out_hash);
        st.ldW(OFF_out_data, i, s.out_data);
        s.wide_bus__DOT__sh = st.ld<VlcuMem8>(OFF_wide_bus__DOT__sh, i);
        s.wide_bus__DOT__op = st.ld<VlcuMem8>(OFF_wide_bus__DOT__op, i);
        s.wide_bus__DOT__vld = st.ld<VlcuMem8>(OFF_wide_bus__DOT__vld, i);
        s.__VstlFirstIteration = st.ld<VlcuMem8>(OFF___VstlFirstIteration, i);
        s.__VstlPhaseResult = st.ld<VlcuMem8>(OFF___VstlPhaseResult, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        st.ldW(OFF_wide_bus__DOT__acc, i, s.wide_bus__DOT__acc);
        st.ldW(OFF_wide_bus__DOT__stage, i, s.wide_bus__DOT__stage);
        st.ldW(OFF_wide_bus__DOT__hash, i, s.wide_bus__DOT__hash);
        st.ldW(OFF_wide_bus__DOT__shl, i, s.wide_bus__DOT__shl);
        st.ldW(OFF_wide_bus__DOT__prod, i, s.wide_bus__DOT__prod);
        st.ldW(OFF_wide_bus__DOT__res, i, s.wide_bus__DOT__res);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VstlTriggered[e] = st.ldE<QData>(OFF___VstlTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
        for (uint32_t e = 0; e < 2; ++e) s.__Vm_traceActivity[e] = st.ldE<VlcuMem8>(OFF___Vm_traceActivity, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.st<VlcuMem8>(OFF_rst_n, i, s.rst_n);
        st.st<VlcuMem8>(OFF_in_valid, i, s.in_valid);
        st.st<VlcuMem8>(OFF_in_op, i, s.in_op);
        st.st<VlcuMem8>(OFF_in_shift, i, s.in_shift);
        st.stW(OFF_in_data, i, s.in_data);
        st.st<VlcuMem8>(OFF_out_valid, i, s.out_valid);
        st.st<VlcuMem8>(OFF_out_lt, i, s.out_lt);
        st.st<VlcuMem8>(OFF_out_lts, i, s.out_lts);
        st.st<VlcuMem16>(OFF_out_ones, i, s.out_ones);
        st.stW(OFF_out_slice, i, s.out_slice);
        st.stW(OFF_out_hash, i, s.out_hash);
        st.stW(OFF_out_data, i, s.out_data);
        st.st<VlcuMem8>(OFF_wide_bus__DOT__sh, i, s.wide_bus__DOT__sh);
        st.st<VlcuMem8>(OFF_wide_bus__DOT__op, i, s.wide_bus__DOT__op);
        st.st<VlcuMem8>(OFF_wide_bus__DOT__vld, i, s.wide_bus__DOT__vld);
        st.st<VlcuMem8>(OFF___VstlFirstIteration, i, s.__VstlFirstIteration);
        st.st<VlcuMem8>(OFF___VstlPhaseResult, i, s.__VstlPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.stW(OFF_wide_bus__DOT__acc, i, s.wide_bus__DOT__acc);
        st.stW(OFF_wide_bus__DOT__stage, i, s.wide_bus__DOT__stage);
        st.stW(OFF_wide_bus__DOT__hash, i, s.wide_bus__DOT__hash);
        st.stW(OFF_wide_bus__DOT__shl, i, s.wide_bus__DOT__shl);
        st.stW(OFF_wide_bus__DOT__prod, i, s.wide_bus__DOT__prod);
        st.stW(OFF_wide_bus__DOT__res, i, s.wide_bus__DOT__res);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VstlTriggered, e, i, s.__VstlTriggered[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VactTriggered, e, i, s.__VactTriggered[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VnbaTriggered, e, i, s.__VnbaTriggered[e]);
        for (uint32_t e = 0; e < 2; ++e) st.stE<VlcuMem8>(OFF___Vm_traceActivity, e, i, s.__Vm_traceActivity[e]);
    }
    VLCU_D static void load_inputs_row(Root& s, const uint32_t* rowp, uint64_t stride, uint64_t i) {
        s.clk = static_cast<VlcuMem8>(rowp[0 * stride + i] & 0x1U);
        s.rst_n = static_cast<VlcuMem8>(rowp[1 * stride + i] & 0x1U);
        s.in_valid = static_cast<VlcuMem8>(rowp[2 * stride + i] & 0x1U);
        s.in_op = static_cast<VlcuMem8>(rowp[3 * stride + i] & 0x7U);
        s.in_shift = static_cast<VlcuMem8>(rowp[4 * stride + i] & 0xffU);
        s.in_data[0] = rowp[5 * stride + i] & 0xffffffffU;
        s.in_data[1] = rowp[6 * stride + i] & 0xffffffffU;
        s.in_data[2] = rowp[7 * stride + i] & 0xffffffffU;
        s.in_data[3] = rowp[8 * stride + i] & 0xffffffffU;
        s.in_data[4] = rowp[9 * stride + i] & 0xffffffffU;
        s.in_data[5] = rowp[10 * stride + i] & 0xffffffffU;
        s.in_data[6] = rowp[11 * stride + i] & 0xffffffffU;
        s.in_data[7] = rowp[12 * stride + i] & 0xffffffffU;
    }
    VLCU_D static void store_outputs_row(const Root& s, uint32_t* rowp, uint64_t stride, uint64_t i) {
        rowp[0 * stride + i] = s.out_valid;
        rowp[1 * stride + i] = s.out_lt;
        rowp[2 * stride + i] = s.out_lts;
        rowp[3 * stride + i] = s.out_ones;
        rowp[4 * stride + i] = s.out_slice[0];
        rowp[5 * stride + i] = s.out_slice[1];
        rowp[6 * stride + i] = s.out_slice[2];
        rowp[7 * stride + i] = s.out_hash[0];
        rowp[8 * stride + i] = s.out_hash[1];
        rowp[9 * stride + i] = s.out_hash[2];
        rowp[10 * stride + i] = s.out_hash[3];
        rowp[11 * stride + i] = s.out_data[0];
        rowp[12 * stride + i] = s.out_data[1];
        rowp[13 * stride + i] = s.out_data[2];
        rowp[14 * stride + i] = s.out_data[3];
        rowp[15 * stride + i] = s.out_data[4];
        rowp[16 * stride + i] = s.out_data[5];
        rowp[17 * stride + i] = s.out_data[6];
        rowp[18 * stride + i] = s.out_data[7];
    }
    VLCU_D static uint64_t outputs_sig(const Root& s, uint64_t sig) {
        uint32_t acc = VLCU_SIG_ACC_INIT;
        acc = vlcu_sig_word(acc, s.out_valid);
        acc = vlcu_sig_word(acc, s.out_lt);
        acc = vlcu_sig_word(acc, s.out_lts);
        acc = vlcu_sig_word(acc, s.out_ones);
        acc = vlcu_sig_word(acc, s.out_slice[0]);
        acc = vlcu_sig_word(acc, s.out_slice[1]);
        acc = vlcu_sig_word(acc, s.out_slice[2]);
        acc = vlcu_sig_word(acc, s.out_hash[0]);
        acc = vlcu_sig_word(acc, s.out_hash[1]);
        acc = vlcu_sig_word(acc, s.out_hash[2]);
        acc = vlcu_sig_word(acc, s.out_hash[3]);
        acc = vlcu_sig_word(acc, s.out_data[0]);
        acc = vlcu_sig_word(acc, s.out_data[1]);
        acc = vlcu_sig_word(acc, s.out_data[2]);
        acc = vlcu_sig_word(acc, s.out_data[3]);
        acc = vlcu_sig_word(acc, s.out_data[4]);
        acc = vlcu_sig_word(acc, s.out_data[5]);
        acc = vlcu_sig_word(acc, s.out_data[6]);
        acc = vlcu_sig_word(acc, s.out_data[7]);
        return vlcu_sig_mix(sig, acc);
    }
    enum : uint32_t { kInWords = 13, kOutWords = 19, kInBytes = 37, kOutBytes = 65, kTraceWords = 102 };

    VLCU_D static void Vwide_bus___024root___eval_static(Vwide_bus___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
    }

    VLCU_D static void Vwide_bus___024root___eval_initial(Vwide_bus___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vwide_bus___024root___eval_final(Vwide_bus___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vwide_bus___024root___eval_settle(Vwide_bus___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vwide_bus___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/wide_bus/wide_bus.v", 6, "", "DIDNOTCONVERGE: Settle region did not converge after '--converge-limit' of 10000 tries");
            }
            __VstlIterCount = ((IData)(1U) + __VstlIterCount);
            vlSelfRef.__VstlPhaseResult = Vwide_bus___024root___eval_phase__stl(vlSelf, vlCtx);
            vlSelfRef.__VstlFirstIteration = 0U;
        } while (vlSelfRef.__VstlPhaseResult);
    }

    VLCU_D static bool Vwide_bus___024root___trigger_anySet__stl(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vwide_bus___024root___stl_sequent__TOP__0(Vwide_bus___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        VlWide<4>/*127:0*/ __Vtemp_1;
        VlWide<4>/*127:0*/ __Vtemp_2;
        IData/*31:0*/ __VExpandSel_WordIdx_1;
        IData/*31:0*/ __VExpandSel_LoShift_1;
        CData/*0:0*/ __VExpandSel_Aligned_1;
        IData/*31:0*/ __VExpandSel_HiShift_1;
        IData/*31:0*/ __VExpandSel_HiMask_1;
        // Body
        vlSelfRef.out_valid = vlSelfRef.wide_bus__DOT__vld;
        vlSelfRef.out_ones = (0x000001ffU & VL_COUNTONES_W(8, vlSelfRef.wide_bus__DOT__acc));
        vlSelfRef.out_data[0U] = vlSelfRef.wide_bus__DOT__acc[0U];
        vlSelfRef.out_data[1U] = vlSelfRef.wide_bus__DOT__acc[1U];
        vlSelfRef.out_data[2U] = vlSelfRef.wide_bus__DOT__acc[2U];
        vlSelfRef.out_data[3U] = vlSelfRef.wide_bus__DOT__acc[3U];
        vlSelfRef.out_data[4U] = vlSelfRef.wide_bus__DOT__acc[4U];
        vlSelfRef.out_data[5U] = vlSelfRef.wide_bus__DOT__acc[5U];
        vlSelfRef.out_data[6U] = vlSelfRef.wide_bus__DOT__acc[6U];
        vlSelfRef.out_data[7U] = vlSelfRef.wide_bus__DOT__acc[7U];
        vlSelfRef.out_hash[0U] = vlSelfRef.wide_bus__DOT__hash[0U];
        vlSelfRef.out_hash[1U] = vlSelfRef.wide_bus__DOT__hash[1U];
        vlSelfRef.out_hash[2U] = vlSelfRef.wide_bus__DOT__hash[2U];
        vlSelfRef.out_hash[3U] = vlSelfRef.wide_bus__DOT__hash[3U];
        vlSelfRef.out_lt = VL_LT_W(8, vlSelfRef.wide_bus__DOT__acc, vlSelfRef.wide_bus__DOT__stage);
        vlSelfRef.out_lts = VL_LTS_IWW(256, vlSelfRef.wide_bus__DOT__acc, vlSelfRef.wide_bus__DOT__stage);
        VL_SHIFTL_WWI(256,256,8, vlSelfRef.wide_bus__DOT__shl, vlSelfRef.wide_bus__DOT__stage, (IData)(vlSelfRef.wide_bus__DOT__sh));
        __Vtemp_1[0U] = (IData)((((QData)((IData)(vlSelfRef.wide_bus__DOT__acc[3U])) 
                                  << 0x00000020U) | (QData)((IData)(vlSelfRef.wide_bus__DOT__acc[2U]))));
        __Vtemp_1[1U] = (IData)(((((QData)((IData)(vlSelfRef.wide_bus__DOT__acc[3U])) 
                                   << 0x00000020U) 
                                  | (QData)((IData)(vlSelfRef.wide_bus__DOT__acc[2U]))) 
                                 >> 0x00000020U));
        __Vtemp_1[2U] = 0U;
        __Vtemp_1[3U] = 0U;
        __Vtemp_2[0U] = (IData)((((QData)((IData)(vlSelfRef.wide_bus__DOT__stage[1U])) 
                                  << 0x00000020U) | (QData)((IData)(vlSelfRef.wide_bus__DOT__stage[0U]))));
        __Vtemp_2[1U] = (IData)(((((QData)((IData)(vlSelfRef.wide_bus__DOT__stage[1U])) 
                                   << 0x00000020U) 
                                  | (QData)((IData)(vlSelfRef.wide_bus__DOT__stage[0U]))) 
                                 >> 0x00000020U));
        __Vtemp_2[2U] = 0U;
        __Vtemp_2[3U] = 0U;
        VL_MUL_W(4, vlSelfRef.wide_bus__DOT__prod, __Vtemp_1, __Vtemp_2);
        if ((0U == (IData)(vlSelfRef.wide_bus__DOT__op))) {
            VL_ADD_W(8, vlSelfRef.wide_bus__DOT__res, vlSelfRef.wide_bus__DOT__acc, vlSelfRef.wide_bus__DOT__stage);
        } else if ((1U == (IData)(vlSelfRef.wide_bus__DOT__op))) {
            VL_SUB_W(8, vlSelfRef.wide_bus__DOT__res, vlSelfRef.wide_bus__DOT__acc, vlSelfRef.wide_bus__DOT__stage);
        } else if ((2U == (IData)(vlSelfRef.wide_bus__DOT__op))) {
            vlSelfRef.wide_bus__DOT__res[0U] = vlSelfRef.wide_bus__DOT__shl[0U];
            vlSelfRef.wide_bus__DOT__res[1U] = vlSelfRef.wide_bus__DOT__shl[1U];
            vlSelfRef.wide_bus__DOT__res[2U] = vlSelfRef.wide_bus__DOT__shl[2U];
            vlSelfRef.wide_bus__DOT__res[3U] = vlSelfRef.wide_bus__DOT__shl[3U];
            vlSelfRef.wide_bus__DOT__res[4U] = vlSelfRef.wide_bus__DOT__shl[4U];
            vlSelfRef.wide_bus__DOT__res[5U] = vlSelfRef.wide_bus__DOT__shl[5U];
            vlSelfRef.wide_bus__DOT__res[6U] = vlSelfRef.wide_bus__DOT__shl[6U];
            vlSelfRef.wide_bus__DOT__res[7U] = vlSelfRef.wide_bus__DOT__shl[7U];
        } else if ((3U == (IData)(vlSelfRef.wide_bus__DOT__op))) {
            VL_SHIFTR_WWI(256,256,8, vlSelfRef.wide_bus__DOT__res, vlSelfRef.wide_bus__DOT__stage, (IData)(vlSelfRef.wide_bus__DOT__sh));
        } else if ((4U == (IData)(vlSelfRef.wide_bus__DOT__op))) {
            VL_SHIFTRS_WWI(256,256,8, vlSelfRef.wide_bus__DOT__res, vlSelfRef.wide_bus__DOT__stage, (IData)(vlSelfRef.wide_bus__DOT__sh));
        } else if ((5U == (IData)(vlSelfRef.wide_bus__DOT__op))) {
            vlSelfRef.wide_bus__DOT__res[0U] = (vlSelfRef.wide_bus__DOT__acc[4U] 
                                                ^ (~ vlSelfRef.wide_bus__DOT__prod[0U]));
            vlSelfRef.wide_bus__DOT__res[1U] = (vlSelfRef.wide_bus__DOT__acc[5U] 
                                                ^ (~ vlSelfRef.wide_bus__DOT__prod[1U]));
            vlSelfRef.wide_bus__DOT__res[2U] = (vlSelfRef.wide_bus__DOT__acc[6U] 
                                                ^ (~ vlSelfRef.wide_bus__DOT__prod[2U]));
            vlSelfRef.wide_bus__DOT__res[3U] = (vlSelfRef.wide_bus__DOT__acc[7U] 
                                                ^ (~ vlSelfRef.wide_bus__DOT__prod[3U]));
            vlSelfRef.wide_bus__DOT__res[4U] = (vlSelfRef.wide_bus__DOT__acc[0U] 
                                                ^ vlSelfRef.wide_bus__DOT__prod[0U]);
            vlSelfRef.wide_bus__DOT__res[5U] = (vlSelfRef.wide_bus__DOT__acc[1U] 
                                                ^ vlSelfRef.wide_bus__DOT__prod[1U]);
            vlSelfRef.wide_bus__DOT__res[6U] = (vlSelfRef.wide_bus__DOT__acc[2U] 
                                                ^ vlSelfRef.wide_bus__DOT__prod[2U]);
            vlSelfRef.wide_bus__DOT__res[7U] = (vlSelfRef.wide_bus__DOT__acc[3U] 
                                                ^ vlSelfRef.wide_bus__DOT__prod[3U]);
        } else if ((6U == (IData)(vlSelfRef.wide_bus__DOT__op))) {
            vlSelfRef.wide_bus__DOT__res[0U] = ((vlSelfRef.wide_bus__DOT__acc[0U] 
                                                 & vlSelfRef.wide_bus__DOT__stage[0U]) 
                                                | ((~ vlSelfRef.wide_bus__DOT__acc[0U]) 
                                                   & vlSelfRef.wide_bus__DOT__shl[0U]));
            vlSelfRef.wide_bus__DOT__res[1U] = ((vlSelfRef.wide_bus__DOT__acc[1U] 
                                                 & vlSelfRef.wide_bus__DOT__stage[1U]) 
                                                | ((~ vlSelfRef.wide_bus__DOT__acc[1U]) 
                                                   & vlSelfRef.wide_bus__DOT__shl[1U]));
            vlSelfRef.wide_bus__DOT__res[2U] = ((vlSelfRef.wide_bus__DOT__acc[2U] 
                                                 & vlSelfRef.wide_bus__DOT__stage[2U]) 
                                                | ((~ vlSelfRef.wide_bus__DOT__acc[2U]) 
                                                   & vlSelfRef.wide_bus__DOT__shl[2U]));
            vlSelfRef.wide_bus__DOT__res[3U] = ((vlSelfRef.wide_bus__DOT__acc[3U] 
                                                 & vlSelfRef.wide_bus__DOT__stage[3U]) 
                                                | ((~ vlSelfRef.wide_bus__DOT__acc[3U]) 
                                                   & vlSelfRef.wide_bus__DOT__shl[3U]));
            vlSelfRef.wide_bus__DOT__res[4U] = ((vlSelfRef.wide_bus__DOT__acc[4U] 
                                                 & vlSelfRef.wide_bus__DOT__stage[4U]) 
                                                | ((~ vlSelfRef.wide_bus__DOT__acc[4U]) 
                                                   & vlSelfRef.wide_bus__DOT__shl[4U]));
            vlSelfRef.wide_bus__DOT__res[5U] = ((vlSelfRef.wide_bus__DOT__acc[5U] 
                                                 & vlSelfRef.wide_bus__DOT__stage[5U]) 
                                                | ((~ vlSelfRef.wide_bus__DOT__acc[5U]) 
                                                   & vlSelfRef.wide_bus__DOT__shl[5U]));
            vlSelfRef.wide_bus__DOT__res[6U] = ((vlSelfRef.wide_bus__DOT__acc[6U] 
                                                 & vlSelfRef.wide_bus__DOT__stage[6U]) 
                                                | ((~ vlSelfRef.wide_bus__DOT__acc[6U]) 
                                                   & vlSelfRef.wide_bus__DOT__shl[6U]));
            vlSelfRef.wide_bus__DOT__res[7U] = ((vlSelfRef.wide_bus__DOT__acc[7U] 
                                                 & vlSelfRef.wide_bus__DOT__stage[7U]) 
                                                | ((~ vlSelfRef.wide_bus__DOT__acc[7U]) 
                                                   & vlSelfRef.wide_bus__DOT__shl[7U]));
        } else {
            vlSelfRef.wide_bus__DOT__res[0U] = ((vlSelfRef.wide_bus__DOT__acc[0U] 
                                                 << 1U) 
                                                | (1U 
                                                   & VL_REDXOR_32(
                                                                  (((((((vlSelfRef.wide_bus__DOT__acc[0U] 
                                                                         ^ vlSelfRef.wide_bus__DOT__acc[1U]) 
                                                                        ^ vlSelfRef.wide_bus__DOT__acc[2U]) 
                                                                       ^ vlSelfRef.wide_bus__DOT__acc[3U]) 
                                                                      ^ vlSelfRef.wide_bus__DOT__acc[4U]) 
                                                                     ^ vlSelfRef.wide_bus__DOT__acc[5U]) 
                                                                    ^ vlSelfRef.wide_bus__DOT__acc[6U]) 
                                                                   ^ vlSelfRef.wide_bus__DOT__acc[7U]))));
            vlSelfRef.wide_bus__DOT__res[1U] = ((vlSelfRef.wide_bus__DOT__acc[0U] 
                                                 >> 0x0000001fU) 
                                                | (vlSelfRef.wide_bus__DOT__acc[1U] 
                                                   << 1U));
            vlSelfRef.wide_bus__DOT__res[2U] = ((vlSelfRef.wide_bus__DOT__acc[1U] 
                                                 >> 0x0000001fU) 
                                                | (vlSelfRef.wide_bus__DOT__acc[2U] 
                                                   << 1U));
            vlSelfRef.wide_bus__DOT__res[3U] = ((vlSelfRef.wide_bus__DOT__acc[2U] 
                                                 >> 0x0000001fU) 
                                                | (vlSelfRef.wide_bus__DOT__acc[3U] 
                                                   << 1U));
            vlSelfRef.wide_bus__DOT__res[4U] = ((vlSelfRef.wide_bus__DOT__acc[3U] 
                                                 >> 0x0000001fU) 
                                                | (vlSelfRef.wide_bus__DOT__acc[4U] 
                                                   << 1U));
            vlSelfRef.wide_bus__DOT__res[5U] = ((vlSelfRef.wide_bus__DOT__acc[4U] 
                                                 >> 0x0000001fU) 
                                                | (vlSelfRef.wide_bus__DOT__acc[5U] 
                                                   << 1U));
            vlSelfRef.wide_bus__DOT__res[6U] = ((vlSelfRef.wide_bus__DOT__acc[5U] 
                                                 >> 0x0000001fU) 
                                                | (vlSelfRef.wide_bus__DOT__acc[6U] 
                                                   << 1U));
            vlSelfRef.wide_bus__DOT__res[7U] = ((vlSelfRef.wide_bus__DOT__acc[6U] 
                                                 >> 0x0000001fU) 
                                                | (vlSelfRef.wide_bus__DOT__acc[7U] 
                                                   << 1U));
        }
        __VExpandSel_WordIdx_1 = (3U & ((IData)(vlSelfRef.wide_bus__DOT__sh) 
                                        >> 5U));
        __VExpandSel_LoShift_1 = (0x0000001fU & (IData)(vlSelfRef.wide_bus__DOT__sh));
        __VExpandSel_Aligned_1 = (0U == __VExpandSel_LoShift_1);
        if (__VExpandSel_Aligned_1) {
            __VExpandSel_HiShift_1 = 0U;
            __VExpandSel_HiMask_1 = 0U;
        } else {
            __VExpandSel_HiShift_1 = ((IData)(0x00000020U) 
                                      - __VExpandSel_LoShift_1);
            __VExpandSel_HiMask_1 = 0xffffffffU;
        }
        vlSelfRef.out_slice[0U] = (((vlSelfRef.wide_bus__DOT__res
                                     [((IData)(1U) 
                                       + __VExpandSel_WordIdx_1)] 
                                     << __VExpandSel_HiShift_1) 
                                    & __VExpandSel_HiMask_1) 
                                   | (vlSelfRef.wide_bus__DOT__res
                                      [__VExpandSel_WordIdx_1] 
                                      >> __VExpandSel_LoShift_1));
        vlSelfRef.out_slice[1U] = (((vlSelfRef.wide_bus__DOT__res
                                     [((IData)(2U) 
                                       + __VExpandSel_WordIdx_1)] 
                                     << __VExpandSel_HiShift_1) 
                                    & __VExpandSel_HiMask_1) 
                                   | (vlSelfRef.wide_bus__DOT__res
                                      [((IData)(1U) 
                                        + __VExpandSel_WordIdx_1)] 
                                      >> __VExpandSel_LoShift_1));
        vlSelfRef.out_slice[2U] = (0x0000003fU & ((
                                                   (((5U 
                                                      <= __VExpandSel_WordIdx_1)
                                                      ? 0U
                                                      : vlSelfRef.wide_bus__DOT__res
                                                     [
                                                     ((IData)(3U) 
                                                      + __VExpandSel_WordIdx_1)]) 
                                                    << __VExpandSel_HiShift_1) 
                                                   & __VExpandSel_HiMask_1) 
                                                  | (vlSelfRef.wide_bus__DOT__res
                                                     [
                                                     ((IData)(2U) 
                                                      + __VExpandSel_WordIdx_1)] 
                                                     >> __VExpandSel_LoShift_1)));
    }

    VLCU_D static bool Vwide_bus___024root___eval_phase__stl(Vwide_bus___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vwide_bus___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
        }
#endif
        __VstlExecute = Vwide_bus___024root___trigger_anySet__stl(vlSelfRef.__VstlTriggered);
        if (__VstlExecute) {
            {
                // Inlined CFunc: _eval_stl
                if ((1ULL & vlSelfRef.__VstlTriggered[0U])) {
                    Vwide_bus___024root___stl_sequent__TOP__0(vlSelf, vlCtx);
                    {
                        // Inlined CFunc: __Vm_traceActivitySetAll
                        vlSelfRef.__Vm_traceActivity[0U] = 1U;
                        vlSelfRef.__Vm_traceActivity[1U] = 1U;
                    }
                }
            }
        }
        return (__VstlExecute);
    }

    VLCU_D static bool Vwide_bus___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vwide_bus___024root___nba_sequent__TOP__0(Vwide_bus___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        VlWide<4>/*127:0*/ __Vtemp_1;
        VlWide<4>/*127:0*/ __Vtemp_8;
        VlWide<4>/*127:0*/ __Vtemp_9;
        IData/*31:0*/ __VExpandSel_WordIdx_1;
        IData/*31:0*/ __VExpandSel_LoShift_1;
        CData/*0:0*/ __VExpandSel_Aligned_1;
        IData/*31:0*/ __VExpandSel_HiShift_1;
        IData/*31:0*/ __VExpandSel_HiMask_1;
        // Body
        if (vlSelfRef.rst_n) {
            vlSelfRef.wide_bus__DOT__op = vlSelfRef.in_op;
            vlSelfRef.wide_bus__DOT__sh = vlSelfRef.in_shift;
            vlSelfRef.wide_bus__DOT__stage[0U] = vlSelfRef.in_data[0U];
            vlSelfRef.wide_bus__DOT__stage[1U] = vlSelfRef.in_data[1U];
            vlSelfRef.wide_bus__DOT__stage[2U] = vlSelfRef.in_data[2U];
            vlSelfRef.wide_bus__DOT__stage[3U] = vlSelfRef.in_data[3U];
            vlSelfRef.wide_bus__DOT__stage[4U] = vlSelfRef.in_data[4U];
            vlSelfRef.wide_bus__DOT__stage[5U] = vlSelfRef.in_data[5U];
            vlSelfRef.wide_bus__DOT__stage[6U] = vlSelfRef.in_data[6U];
            vlSelfRef.wide_bus__DOT__stage[7U] = vlSelfRef.in_data[7U];
            if (vlSelfRef.wide_bus__DOT__vld) {
                __Vtemp_1[1U] = ((((vlSelfRef.wide_bus__DOT__hash[0U] 
                                    >> 0x0000001fU) 
                                   | (vlSelfRef.wide_bus__DOT__hash[1U] 
                                      << 1U)) ^ vlSelfRef.wide_bus__DOT__prod[1U]) 
                                 ^ vlSelfRef.out_slice[1U]);
                __Vtemp_1[2U] = ((((vlSelfRef.wide_bus__DOT__hash[1U] 
                                    >> 0x0000001fU) 
                                   | (vlSelfRef.wide_bus__DOT__hash[2U] 
                                      << 1U)) ^ vlSelfRef.wide_bus__DOT__prod[2U]) 
                                 ^ vlSelfRef.out_slice[2U]);
                __Vtemp_1[3U] = (((vlSelfRef.wide_bus__DOT__hash[2U] 
                                   >> 0x0000001fU) 
                                  | (vlSelfRef.wide_bus__DOT__hash[3U] 
                                     << 1U)) ^ vlSelfRef.wide_bus__DOT__prod[3U]);
                vlSelfRef.wide_bus__DOT__hash[0U] = 
                    ((((vlSelfRef.wide_bus__DOT__hash[0U] 
                        << 1U) | (vlSelfRef.wide_bus__DOT__hash[3U] 
                                  >> 0x0000001fU)) 
                      ^ vlSelfRef.wide_bus__DOT__prod[0U]) 
                     ^ vlSelfRef.out_slice[0U]);
                vlSelfRef.wide_bus__DOT__hash[1U] = __Vtemp_1[1U];
                vlSelfRef.wide_bus__DOT__hash[2U] = __Vtemp_1[2U];
                vlSelfRef.wide_bus__DOT__hash[3U] = __Vtemp_1[3U];
                vlSelfRef.wide_bus__DOT__acc[0U] = vlSelfRef.wide_bus__DOT__res[0U];
                vlSelfRef.wide_bus__DOT__acc[1U] = vlSelfRef.wide_bus__DOT__res[1U];
                vlSelfRef.wide_bus__DOT__acc[2U] = vlSelfRef.wide_bus__DOT__res[2U];
                vlSelfRef.wide_bus__DOT__acc[3U] = vlSelfRef.wide_bus__DOT__res[3U];
                vlSelfRef.wide_bus__DOT__acc[4U] = vlSelfRef.wide_bus__DOT__res[4U];
                vlSelfRef.wide_bus__DOT__acc[5U] = vlSelfRef.wide_bus__DOT__res[5U];
                vlSelfRef.wide_bus__DOT__acc[6U] = vlSelfRef.wide_bus__DOT__res[6U];
                vlSelfRef.wide_bus__DOT__acc[7U] = vlSelfRef.wide_bus__DOT__res[7U];
            }
        } else {
            vlSelfRef.wide_bus__DOT__op = 0U;
            vlSelfRef.wide_bus__DOT__sh = 0U;
            VL_ASSIGN_W(256, vlSelfRef.wide_bus__DOT__stage, Vwide_bus__ConstPool__CONST_h9e67c271_0);
            vlSelfRef.wide_bus__DOT__hash[0U] = 0U;
            vlSelfRef.wide_bus__DOT__hash[1U] = 0U;
            vlSelfRef.wide_bus__DOT__hash[2U] = 0U;
            vlSelfRef.wide_bus__DOT__hash[3U] = 0U;
            VL_ASSIGN_W(256, vlSelfRef.wide_bus__DOT__acc, Vwide_bus__ConstPool__CONST_h4e9f510d_0);
        }
        VL_SHIFTL_WWI(256,256,8, vlSelfRef.wide_bus__DOT__shl, vlSelfRef.wide_bus__DOT__stage, (IData)(vlSelfRef.wide_bus__DOT__sh));
        vlSelfRef.out_hash[0U] = vlSelfRef.wide_bus__DOT__hash[0U];
        vlSelfRef.out_hash[1U] = vlSelfRef.wide_bus__DOT__hash[1U];
        vlSelfRef.out_hash[2U] = vlSelfRef.wide_bus__DOT__hash[2U];
        vlSelfRef.out_hash[3U] = vlSelfRef.wide_bus__DOT__hash[3U];
        vlSelfRef.wide_bus__DOT__vld = ((IData)(vlSelfRef.rst_n) 
                                        && (IData)(vlSelfRef.in_valid));
        vlSelfRef.out_ones = (0x000001ffU & VL_COUNTONES_W(8, vlSelfRef.wide_bus__DOT__acc));
        vlSelfRef.out_data[0U] = vlSelfRef.wide_bus__DOT__acc[0U];
        vlSelfRef.out_data[1U] = vlSelfRef.wide_bus__DOT__acc[1U];
        vlSelfRef.out_data[2U] = vlSelfRef.wide_bus__DOT__acc[2U];
        vlSelfRef.out_data[3U] = vlSelfRef.wide_bus__DOT__acc[3U];
        vlSelfRef.out_data[4U] = vlSelfRef.wide_bus__DOT__acc[4U];
        vlSelfRef.out_data[5U] = vlSelfRef.wide_bus__DOT__acc[5U];
        vlSelfRef.out_data[6U] = vlSelfRef.wide_bus__DOT__acc[6U];
        vlSelfRef.out_data[7U] = vlSelfRef.wide_bus__DOT__acc[7U];
        vlSelfRef.out_lt = VL_LT_W(8, vlSelfRef.wide_bus__DOT__acc, vlSelfRef.wide_bus__DOT__stage);
        vlSelfRef.out_lts = VL_LTS_IWW(256, vlSelfRef.wide_bus__DOT__acc, vlSelfRef.wide_bus__DOT__stage);
        __Vtemp_8[0U] = (IData)((((QData)((IData)(vlSelfRef.wide_bus__DOT__acc[3U])) 
                                  << 0x00000020U) | (QData)((IData)(vlSelfRef.wide_bus__DOT__acc[2U]))));
        __Vtemp_8[1U] = (IData)(((((QData)((IData)(vlSelfRef.wide_bus__DOT__acc[3U])) 
                                   << 0x00000020U) 
                                  | (QData)((IData)(vlSelfRef.wide_bus__DOT__acc[2U]))) 
                                 >> 0x00000020U));
        __Vtemp_8[2U] = 0U;
        __Vtemp_8[3U] = 0U;
        __Vtemp_9[0U] = (IData)((((QData)((IData)(vlSelfRef.wide_bus__DOT__stage[1U])) 
                                  << 0x00000020U) | (QData)((IData)(vlSelfRef.wide_bus__DOT__stage[0U]))));
        __Vtemp_9[1U] = (IData)(((((QData)((IData)(vlSelfRef.wide_bus__DOT__stage[1U])) 
                                   << 0x00000020U) 
                                  | (QData)((IData)(vlSelfRef.wide_bus__DOT__stage[0U]))) 
                                 >> 0x00000020U));
        __Vtemp_9[2U] = 0U;
        __Vtemp_9[3U] = 0U;
        VL_MUL_W(4, vlSelfRef.wide_bus__DOT__prod, __Vtemp_8, __Vtemp_9);
        if ((0U == (IData)(vlSelfRef.wide_bus__DOT__op))) {
            VL_ADD_W(8, vlSelfRef.wide_bus__DOT__res, vlSelfRef.wide_bus__DOT__acc, vlSelfRef.wide_bus__DOT__stage);
        } else if ((1U == (IData)(vlSelfRef.wide_bus__DOT__op))) {
            VL_SUB_W(8, vlSelfRef.wide_bus__DOT__res, vlSelfRef.wide_bus__DOT__acc, vlSelfRef.wide_bus__DOT__stage);
        } else if ((2U == (IData)(vlSelfRef.wide_bus__DOT__op))) {
            vlSelfRef.wide_bus__DOT__res[0U] = vlSelfRef.wide_bus__DOT__shl[0U];
            vlSelfRef.wide_bus__DOT__res[1U] = vlSelfRef.wide_bus__DOT__shl[1U];
            vlSelfRef.wide_bus__DOT__res[2U] = vlSelfRef.wide_bus__DOT__shl[2U];
            vlSelfRef.wide_bus__DOT__res[3U] = vlSelfRef.wide_bus__DOT__shl[3U];
            vlSelfRef.wide_bus__DOT__res[4U] = vlSelfRef.wide_bus__DOT__shl[4U];
            vlSelfRef.wide_bus__DOT__res[5U] = vlSelfRef.wide_bus__DOT__shl[5U];
            vlSelfRef.wide_bus__DOT__res[6U] = vlSelfRef.wide_bus__DOT__shl[6U];
            vlSelfRef.wide_bus__DOT__res[7U] = vlSelfRef.wide_bus__DOT__shl[7U];
        } else if ((3U == (IData)(vlSelfRef.wide_bus__DOT__op))) {
            VL_SHIFTR_WWI(256,256,8, vlSelfRef.wide_bus__DOT__res, vlSelfRef.wide_bus__DOT__stage, (IData)(vlSelfRef.wide_bus__DOT__sh));
        } else if ((4U == (IData)(vlSelfRef.wide_bus__DOT__op))) {
            VL_SHIFTRS_WWI(256,256,8, vlSelfRef.wide_bus__DOT__res, vlSelfRef.wide_bus__DOT__stage, (IData)(vlSelfRef.wide_bus__DOT__sh));
        } else if ((5U == (IData)(vlSelfRef.wide_bus__DOT__op))) {
            vlSelfRef.wide_bus__DOT__res[0U] = (vlSelfRef.wide_bus__DOT__acc[4U] 
                                                ^ (~ vlSelfRef.wide_bus__DOT__prod[0U]));
            vlSelfRef.wide_bus__DOT__res[1U] = (vlSelfRef.wide_bus__DOT__acc[5U] 
                                                ^ (~ vlSelfRef.wide_bus__DOT__prod[1U]));
            vlSelfRef.wide_bus__DOT__res[2U] = (vlSelfRef.wide_bus__DOT__acc[6U] 
                                                ^ (~ vlSelfRef.wide_bus__DOT__prod[2U]));
            vlSelfRef.wide_bus__DOT__res[3U] = (vlSelfRef.wide_bus__DOT__acc[7U] 
                                                ^ (~ vlSelfRef.wide_bus__DOT__prod[3U]));
            vlSelfRef.wide_bus__DOT__res[4U] = (vlSelfRef.wide_bus__DOT__acc[0U] 
                                                ^ vlSelfRef.wide_bus__DOT__prod[0U]);
            vlSelfRef.wide_bus__DOT__res[5U] = (vlSelfRef.wide_bus__DOT__acc[1U] 
                                                ^ vlSelfRef.wide_bus__DOT__prod[1U]);
            vlSelfRef.wide_bus__DOT__res[6U] = (vlSelfRef.wide_bus__DOT__acc[2U] 
                                                ^ vlSelfRef.wide_bus__DOT__prod[2U]);
            vlSelfRef.wide_bus__DOT__res[7U] = (vlSelfRef.wide_bus__DOT__acc[3U] 
                                                ^ vlSelfRef.wide_bus__DOT__prod[3U]);
        } else if ((6U == (IData)(vlSelfRef.wide_bus__DOT__op))) {
            vlSelfRef.wide_bus__DOT__res[0U] = ((vlSelfRef.wide_bus__DOT__acc[0U] 
                                                 & vlSelfRef.wide_bus__DOT__stage[0U]) 
                                                | ((~ vlSelfRef.wide_bus__DOT__acc[0U]) 
                                                   & vlSelfRef.wide_bus__DOT__shl[0U]));
            vlSelfRef.wide_bus__DOT__res[1U] = ((vlSelfRef.wide_bus__DOT__acc[1U] 
                                                 & vlSelfRef.wide_bus__DOT__stage[1U]) 
                                                | ((~ vlSelfRef.wide_bus__DOT__acc[1U]) 
                                                   & vlSelfRef.wide_bus__DOT__shl[1U]));
            vlSelfRef.wide_bus__DOT__res[2U] = ((vlSelfRef.wide_bus__DOT__acc[2U] 
                                                 & vlSelfRef.wide_bus__DOT__stage[2U]) 
                                                | ((~ vlSelfRef.wide_bus__DOT__acc[2U]) 
                                                   & vlSelfRef.wide_bus__DOT__shl[2U]));
            vlSelfRef.wide_bus__DOT__res[3U] = ((vlSelfRef.wide_bus__DOT__acc[3U] 
                                                 & vlSelfRef.wide_bus__DOT__stage[3U]) 
                                                | ((~ vlSelfRef.wide_bus__DOT__acc[3U]) 
                                                   & vlSelfRef.wide_bus__DOT__shl[3U]));
            vlSelfRef.wide_bus__DOT__res[4U] = ((vlSelfRef.wide_bus__DOT__acc[4U] 
                                                 & vlSelfRef.wide_bus__DOT__stage[4U]) 
                                                | ((~ vlSelfRef.wide_bus__DOT__acc[4U]) 
                                                   & vlSelfRef.wide_bus__DOT__shl[4U]));
            vlSelfRef.wide_bus__DOT__res[5U] = ((vlSelfRef.wide_bus__DOT__acc[5U] 
                                                 & vlSelfRef.wide_bus__DOT__stage[5U]) 
                                                | ((~ vlSelfRef.wide_bus__DOT__acc[5U]) 
                                                   & vlSelfRef.wide_bus__DOT__shl[5U]));
            vlSelfRef.wide_bus__DOT__res[6U] = ((vlSelfRef.wide_bus__DOT__acc[6U] 
                                                 & vlSelfRef.wide_bus__DOT__stage[6U]) 
                                                | ((~ vlSelfRef.wide_bus__DOT__acc[6U]) 
                                                   & vlSelfRef.wide_bus__DOT__shl[6U]));
            vlSelfRef.wide_bus__DOT__res[7U] = ((vlSelfRef.wide_bus__DOT__acc[7U] 
                                                 & vlSelfRef.wide_bus__DOT__stage[7U]) 
                                                | ((~ vlSelfRef.wide_bus__DOT__acc[7U]) 
                                                   & vlSelfRef.wide_bus__DOT__shl[7U]));
        } else {
            vlSelfRef.wide_bus__DOT__res[0U] = ((vlSelfRef.wide_bus__DOT__acc[0U] 
                                                 << 1U) 
                                                | (1U 
                                                   & VL_REDXOR_32(
                                                                  (((((((vlSelfRef.wide_bus__DOT__acc[0U] 
                                                                         ^ vlSelfRef.wide_bus__DOT__acc[1U]) 
                                                                        ^ vlSelfRef.wide_bus__DOT__acc[2U]) 
                                                                       ^ vlSelfRef.wide_bus__DOT__acc[3U]) 
                                                                      ^ vlSelfRef.wide_bus__DOT__acc[4U]) 
                                                                     ^ vlSelfRef.wide_bus__DOT__acc[5U]) 
                                                                    ^ vlSelfRef.wide_bus__DOT__acc[6U]) 
                                                                   ^ vlSelfRef.wide_bus__DOT__acc[7U]))));
            vlSelfRef.wide_bus__DOT__res[1U] = ((vlSelfRef.wide_bus__DOT__acc[0U] 
                                                 >> 0x0000001fU) 
                                                | (vlSelfRef.wide_bus__DOT__acc[1U] 
                                                   << 1U));
            vlSelfRef.wide_bus__DOT__res[2U] = ((vlSelfRef.wide_bus__DOT__acc[1U] 
                                                 >> 0x0000001fU) 
                                                | (vlSelfRef.wide_bus__DOT__acc[2U] 
                                                   << 1U));
            vlSelfRef.wide_bus__DOT__res[3U] = ((vlSelfRef.wide_bus__DOT__acc[2U] 
                                                 >> 0x0000001fU) 
                                                | (vlSelfRef.wide_bus__DOT__acc[3U] 
                                                   << 1U));
            vlSelfRef.wide_bus__DOT__res[4U] = ((vlSelfRef.wide_bus__DOT__acc[3U] 
                                                 >> 0x0000001fU) 
                                                | (vlSelfRef.wide_bus__DOT__acc[4U] 
                                                   << 1U));
            vlSelfRef.wide_bus__DOT__res[5U] = ((vlSelfRef.wide_bus__DOT__acc[4U] 
                                                 >> 0x0000001fU) 
                                                | (vlSelfRef.wide_bus__DOT__acc[5U] 
                                                   << 1U));
            vlSelfRef.wide_bus__DOT__res[6U] = ((vlSelfRef.wide_bus__DOT__acc[5U] 
                                                 >> 0x0000001fU) 
                                                | (vlSelfRef.wide_bus__DOT__acc[6U] 
                                                   << 1U));
            vlSelfRef.wide_bus__DOT__res[7U] = ((vlSelfRef.wide_bus__DOT__acc[6U] 
                                                 >> 0x0000001fU) 
                                                | (vlSelfRef.wide_bus__DOT__acc[7U] 
                                                   << 1U));
        }
        vlSelfRef.out_valid = vlSelfRef.wide_bus__DOT__vld;
        __VExpandSel_WordIdx_1 = (3U & ((IData)(vlSelfRef.wide_bus__DOT__sh) 
                                        >> 5U));
        __VExpandSel_LoShift_1 = (0x0000001fU & (IData)(vlSelfRef.wide_bus__DOT__sh));
        __VExpandSel_Aligned_1 = (0U == __VExpandSel_LoShift_1);
        if (__VExpandSel_Aligned_1) {
            __VExpandSel_HiShift_1 = 0U;
            __VExpandSel_HiMask_1 = 0U;
        } else {
            __VExpandSel_HiShift_1 = ((IData)(0x00000020U) 
                                      - __VExpandSel_LoShift_1);
            __VExpandSel_HiMask_1 = 0xffffffffU;
        }
        vlSelfRef.out_slice[0U] = (((vlSelfRef.wide_bus__DOT__res
                                     [((IData)(1U) 
                                       + __VExpandSel_WordIdx_1)] 
                                     << __VExpandSel_HiShift_1) 
                                    & __VExpandSel_HiMask_1) 
                                   | (vlSelfRef.wide_bus__DOT__res
                                      [__VExpandSel_WordIdx_1] 
                                      >> __VExpandSel_LoShift_1));
        vlSelfRef.out_slice[1U] = (((vlSelfRef.wide_bus__DOT__res
                                     [((IData)(2U) 
                                       + __VExpandSel_WordIdx_1)] 
                                     << __VExpandSel_HiShift_1) 
                                    & __VExpandSel_HiMask_1) 
                                   | (vlSelfRef.wide_bus__DOT__res
                                      [((IData)(1U) 
                                        + __VExpandSel_WordIdx_1)] 
                                      >> __VExpandSel_LoShift_1));
        vlSelfRef.out_slice[2U] = (0x0000003fU & ((
                                                   (((5U 
                                                      <= __VExpandSel_WordIdx_1)
                                                      ? 0U
                                                      : vlSelfRef.wide_bus__DOT__res
                                                     [
                                                     ((IData)(3U) 
                                                      + __VExpandSel_WordIdx_1)]) 
                                                    << __VExpandSel_HiShift_1) 
                                                   & __VExpandSel_HiMask_1) 
                                                  | (vlSelfRef.wide_bus__DOT__res
                                                     [
                                                     ((IData)(2U) 
                                                      + __VExpandSel_WordIdx_1)] 
                                                     >> __VExpandSel_LoShift_1)));
    }

    VLCU_D static void Vwide_bus___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vwide_bus___024root___eval_phase__act(Vwide_bus___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vwide_bus___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vwide_bus___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vwide_bus___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vwide_bus___024root___eval_phase__nba(Vwide_bus___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vwide_bus___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    Vwide_bus___024root___nba_sequent__TOP__0(vlSelf, vlCtx);
                    vlSelfRef.__Vm_traceActivity[1U] = 1U;
                }
            }
            Vwide_bus___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vwide_bus___024root___eval(Vwide_bus___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vwide_bus___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/wide_bus/wide_bus.v", 6, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vwide_bus___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("../verilator_b200/csrc/designs/wide_bus/wide_bus.v", 6, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vwide_bus___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vwide_bus___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vwide_bus___024root___ctor_var_reset(Vwide_bus___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        vlSelf->rst_n = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 1638864771569018232ull);
        vlSelf->in_valid = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 2339549897027650563ull);
        vlSelf->in_op = VL_SCOPED_RAND_RESET_I(3, __VscopeHash, 4275446026273260408ull);
        vlSelf->in_shift = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 7877930319276453391ull);
        VL_SCOPED_RAND_RESET_W(256, vlSelf->in_data, __VscopeHash, 4057622023130387117ull);
        vlSelf->out_valid = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 2886291494070200219ull);
        vlSelf->out_lt = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 4255730895800688037ull);
        vlSelf->out_lts = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 15801040001649358340ull);
        vlSelf->out_ones = VL_SCOPED_RAND_RESET_I(9, __VscopeHash, 6501428637900428922ull);
        VL_SCOPED_RAND_RESET_W(70, vlSelf->out_slice, __VscopeHash, 11106134612885672685ull);
        VL_SCOPED_RAND_RESET_W(128, vlSelf->out_hash, __VscopeHash, 7208313861573857419ull);
        VL_SCOPED_RAND_RESET_W(256, vlSelf->out_data, __VscopeHash, 10144880484820144978ull);
        VL_SCOPED_RAND_RESET_W(256, vlSelf->wide_bus__DOT__acc, __VscopeHash, 14194502259691503642ull);
        VL_SCOPED_RAND_RESET_W(256, vlSelf->wide_bus__DOT__stage, __VscopeHash, 1050150744099148078ull);
        VL_SCOPED_RAND_RESET_W(128, vlSelf->wide_bus__DOT__hash, __VscopeHash, 528772692071959489ull);
        vlSelf->wide_bus__DOT__sh = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 16548142997905002182ull);
        vlSelf->wide_bus__DOT__op = VL_SCOPED_RAND_RESET_I(3, __VscopeHash, 10497530677271438983ull);
        vlSelf->wide_bus__DOT__vld = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 1300867881753208539ull);
        VL_SCOPED_RAND_RESET_W(256, vlSelf->wide_bus__DOT__shl, __VscopeHash, 3463108042687186674ull);
        VL_SCOPED_RAND_RESET_W(128, vlSelf->wide_bus__DOT__prod, __VscopeHash, 12014111489641930510ull);
        VL_SCOPED_RAND_RESET_W(256, vlSelf->wide_bus__DOT__res, __VscopeHash, 16169197656339735969ull);
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
        for (int __Vi0 = 0; __Vi0 < 2; ++__Vi0) {
            vlSelf->__Vm_traceActivity[__Vi0] = 0;
        }
    }

    VLCU_D static void Vwide_bus___024root__trace_full_0_sub_0(Vwide_bus___024root* vlSelf, VlcuContext& vlCtx, uint32_t* valsp) {
        auto& vlSelfRef VL_ATTR_UNUSED = *vlSelf;
        VlWide<8>/*255:0*/ __Vtemp_1;
        VlWide<8>/*255:0*/ __Vtemp_2;
        VlWide<8>/*255:0*/ __Vtemp_3;
        VlWide<8>/*255:0*/ __Vtemp_4;
        VlWide<8>/*255:0*/ __Vtemp_10;
        vlcu_trace_set(valsp + 0, static_cast<IData>(vlSelfRef.wide_bus__DOT__vld));
        vlcu_trace_set_w(valsp + 1, 4, (vlSelfRef.wide_bus__DOT__hash));
        vlcu_trace_set_w(valsp + 5, 8, (vlSelfRef.wide_bus__DOT__acc));
        vlcu_trace_set_w(valsp + 13, 8, (vlSelfRef.wide_bus__DOT__stage));
        vlcu_trace_set(valsp + 21, static_cast<IData>(vlSelfRef.wide_bus__DOT__sh));
        vlcu_trace_set(valsp + 22, static_cast<IData>(vlSelfRef.wide_bus__DOT__op));
        VL_ADD_W(8, __Vtemp_1, vlSelfRef.wide_bus__DOT__acc, vlSelfRef.wide_bus__DOT__stage);
        vlcu_trace_set_w(valsp + 23, 8, (__Vtemp_1));
        VL_SUB_W(8, __Vtemp_2, vlSelfRef.wide_bus__DOT__acc, vlSelfRef.wide_bus__DOT__stage);
        vlcu_trace_set_w(valsp + 31, 8, (__Vtemp_2));
        vlcu_trace_set_w(valsp + 39, 8, (vlSelfRef.wide_bus__DOT__shl));
        VL_SHIFTR_WWI(256,256,8, __Vtemp_3, vlSelfRef.wide_bus__DOT__stage, (IData)(vlSelfRef.wide_bus__DOT__sh));
        vlcu_trace_set_w(valsp + 47, 8, (__Vtemp_3));
        VL_SHIFTRS_WWI(256,256,8, __Vtemp_4, vlSelfRef.wide_bus__DOT__stage, (IData)(vlSelfRef.wide_bus__DOT__sh));
        vlcu_trace_set_w(valsp + 55, 8, (__Vtemp_4));
        vlcu_trace_set_w(valsp + 63, 4, (vlSelfRef.wide_bus__DOT__prod));
        __Vtemp_10[0U] = (vlSelfRef.wide_bus__DOT__acc[4U] 
                          ^ (~ vlSelfRef.wide_bus__DOT__prod[0U]));
        __Vtemp_10[1U] = (vlSelfRef.wide_bus__DOT__acc[5U] 
                          ^ (~ vlSelfRef.wide_bus__DOT__prod[1U]));
        __Vtemp_10[2U] = (vlSelfRef.wide_bus__DOT__acc[6U] 
                          ^ (~ vlSelfRef.wide_bus__DOT__prod[2U]));
        __Vtemp_10[3U] = (vlSelfRef.wide_bus__DOT__acc[7U] 
                          ^ (~ vlSelfRef.wide_bus__DOT__prod[3U]));
        __Vtemp_10[4U] = (vlSelfRef.wide_bus__DOT__acc[0U] 
                          ^ vlSelfRef.wide_bus__DOT__prod[0U]);
        __Vtemp_10[5U] = (vlSelfRef.wide_bus__DOT__acc[1U] 
                          ^ vlSelfRef.wide_bus__DOT__prod[1U]);
        __Vtemp_10[6U] = (vlSelfRef.wide_bus__DOT__acc[2U] 
                          ^ vlSelfRef.wide_bus__DOT__prod[2U]);
        __Vtemp_10[7U] = (vlSelfRef.wide_bus__DOT__acc[3U] 
                          ^ vlSelfRef.wide_bus__DOT__prod[3U]);
        vlcu_trace_set_w(valsp + 67, 8, (__Vtemp_10));
        vlcu_trace_set_w(valsp + 75, 8, (vlSelfRef.wide_bus__DOT__res));
        vlcu_trace_set(valsp + 83, static_cast<IData>(vlSelfRef.clk));
        vlcu_trace_set(valsp + 84, static_cast<IData>(vlSelfRef.rst_n));
        vlcu_trace_set(valsp + 85, static_cast<IData>(vlSelfRef.in_valid));
        vlcu_trace_set(valsp + 86, static_cast<IData>(vlSelfRef.in_op));
        vlcu_trace_set(valsp + 87, static_cast<IData>(vlSelfRef.in_shift));
        vlcu_trace_set_w(valsp + 88, 8, (vlSelfRef.in_data));
        vlcu_trace_set(valsp + 96, static_cast<IData>(vlSelfRef.out_lt));
        vlcu_trace_set(valsp + 97, static_cast<IData>(vlSelfRef.out_lts));
        vlcu_trace_set(valsp + 98, static_cast<IData>(vlSelfRef.out_ones));
        vlcu_trace_set_w(valsp + 99, 3, (vlSelfRef.out_slice));
    }

    // tfp->dump(): every traced value stored at its trace code.  Full functions, then the
    // constants, as VerilatedTrace::dump runs them the first time (verilated_trace_imp.h:348-358)
    VLCU_D static void trace_sample(Root* vlSelf, VlcuContext& vlCtx, uint32_t* valsp) {
        (void)vlSelf; (void)vlCtx; (void)valsp;
        Vwide_bus___024root__trace_full_0_sub_0(vlSelf, vlCtx, valsp);
    }

    VLCU_D static void ctor_var_reset(Root* vlSelf, VlcuContext& vlCtx) {
        Vwide_bus___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vwide_bus___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vwide_bus___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vwide_bus___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vwide_bus___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vwide_bus___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vwide_bus__cuda::info() {
    typedef Vwide_bus__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"rst_n", "rst_n", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_rst_n},
        {"in_valid", "in_valid", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_in_valid},
        {"in_op", "in_op", VLCU_MEMBER_IN, 1, 0, 1, 3, 2, 0, T::OFF_in_op},
        {"in_shift", "in_shift", VLCU_MEMBER_IN, 1, 0, 1, 8, 7, 0, T::OFF_in_shift},
        {"in_data", "in_data", VLCU_MEMBER_IN, 4, 0, 8, 256, 255, 0, T::OFF_in_data},
        {"out_valid", "out_valid", VLCU_MEMBER_OUT, 1, 0, 1, 1, 0, 0, T::OFF_out_valid},
        {"out_lt", "out_lt", VLCU_MEMBER_OUT, 1, 0, 1, 1, 0, 0, T::OFF_out_lt},
        {"out_lts", "out_lts", VLCU_MEMBER_OUT, 1, 0, 1, 1, 0, 0, T::OFF_out_lts},
        {"out_ones", "out_ones", VLCU_MEMBER_OUT, 2, 0, 1, 9, 8, 0, T::OFF_out_ones},
        {"out_slice", "out_slice", VLCU_MEMBER_OUT, 4, 0, 3, 70, 69, 0, T::OFF_out_slice},
        {"out_hash", "out_hash", VLCU_MEMBER_OUT, 4, 0, 4, 128, 127, 0, T::OFF_out_hash},
        {"out_data", "out_data", VLCU_MEMBER_OUT, 4, 0, 8, 256, 255, 0, T::OFF_out_data},
        {"wide_bus__DOT__sh", "wide_bus.sh", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF_wide_bus__DOT__sh},
        {"wide_bus__DOT__op", "wide_bus.op", VLCU_MEMBER_STATE, 1, 0, 1, 3, 2, 0, T::OFF_wide_bus__DOT__op},
        {"wide_bus__DOT__vld", "wide_bus.vld", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_wide_bus__DOT__vld},
        {"__VstlFirstIteration", "__VstlFirstIteration", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlFirstIteration},
        {"__VstlPhaseResult", "__VstlPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"wide_bus__DOT__acc", "wide_bus.acc", VLCU_MEMBER_STATE, 4, 0, 8, 256, 255, 0, T::OFF_wide_bus__DOT__acc},
        {"wide_bus__DOT__stage", "wide_bus.stage", VLCU_MEMBER_STATE, 4, 0, 8, 256, 255, 0, T::OFF_wide_bus__DOT__stage},
        {"wide_bus__DOT__hash", "wide_bus.hash", VLCU_MEMBER_STATE, 4, 0, 4, 128, 127, 0, T::OFF_wide_bus__DOT__hash},
        {"wide_bus__DOT__shl", "wide_bus.shl", VLCU_MEMBER_STATE, 4, 0, 8, 256, 255, 0, T::OFF_wide_bus__DOT__shl},
        {"wide_bus__DOT__prod", "wide_bus.prod", VLCU_MEMBER_STATE, 4, 0, 4, 128, 127, 0, T::OFF_wide_bus__DOT__prod},
        {"wide_bus__DOT__res", "wide_bus.res", VLCU_MEMBER_STATE, 4, 0, 8, 256, 255, 0, T::OFF_wide_bus__DOT__res},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"__VstlTriggered", "__VstlTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VstlTriggered},
        {"__VactTriggered", "__VactTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VactTriggered},
        {"__VnbaTriggered", "__VnbaTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VnbaTriggered},
        {"__Vm_traceActivity", "__Vm_traceActivity", VLCU_MEMBER_STATE, 1, 0, 2, 1, 0, 0, T::OFF___Vm_traceActivity},
    };
    static const char* const messages[] = {nullptr};
    static const char* const plusargs[] = {nullptr};
    static const VlcuTraceDecl traceDecls[] = {
        {VLCU_TRACE_PUSH, 0, "$rootio", 0, VLCU_PREFIX_SCOPE_MODULE, -1, 0, 0},
        {VLCU_TRACE_SIG, 0, "clk", 83, 1, -1, 0, 0},
        {VLCU_TRACE_SIG, 0, "rst_n", 84, 1, -1, 0, 0},
        {VLCU_TRACE_SIG, 0, "in_valid", 85, 1, -1, 0, 0},
        {VLCU_TRACE_SIG, 1, "in_op", 86, 3, -1, 2, 0},
        {VLCU_TRACE_SIG, 1, "in_shift", 87, 8, -1, 7, 0},
        {VLCU_TRACE_SIG, 1, "in_data", 88, 256, -1, 255, 0},
        {VLCU_TRACE_SIG, 0, "out_valid", 0, 1, -1, 0, 0},
        {VLCU_TRACE_SIG, 0, "out_lt", 96, 1, -1, 0, 0},
        {VLCU_TRACE_SIG, 0, "out_lts", 97, 1, -1, 0, 0},
        {VLCU_TRACE_SIG, 1, "out_ones", 98, 9, -1, 8, 0},
        {VLCU_TRACE_SIG, 1, "out_slice", 99, 70, -1, 69, 0},
        {VLCU_TRACE_SIG, 1, "out_hash", 1, 128, -1, 127, 0},
        {VLCU_TRACE_SIG, 1, "out_data", 5, 256, -1, 255, 0},
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0},
        {VLCU_TRACE_PUSH, 0, "wide_bus", 0, VLCU_PREFIX_SCOPE_MODULE, -1, 0, 0},
        {VLCU_TRACE_SIG, 0, "clk", 83, 1, -1, 0, 0},
        {VLCU_TRACE_SIG, 0, "rst_n", 84, 1, -1, 0, 0},
        {VLCU_TRACE_SIG, 0, "in_valid", 85, 1, -1, 0, 0},
        {VLCU_TRACE_SIG, 1, "in_op", 86, 3, -1, 2, 0},
        {VLCU_TRACE_SIG, 1, "in_shift", 87, 8, -1, 7, 0},
        {VLCU_TRACE_SIG, 1, "in_data", 88, 256, -1, 255, 0},
        {VLCU_TRACE_SIG, 0, "out_valid", 0, 1, -1, 0, 0},
        {VLCU_TRACE_SIG, 0, "out_lt", 96, 1, -1, 0, 0},
        {VLCU_TRACE_SIG, 0, "out_lts", 97, 1, -1, 0, 0},
        {VLCU_TRACE_SIG, 1, "out_ones", 98, 9, -1, 8, 0},
        {VLCU_TRACE_SIG, 1, "out_slice", 99, 70, -1, 69, 0},
        {VLCU_TRACE_SIG, 1, "out_hash", 1, 128, -1, 127, 0},
        {VLCU_TRACE_SIG, 1, "out_data", 5, 256, -1, 255, 0},
        {VLCU_TRACE_SIG, 1, "acc", 5, 256, -1, 255, 0},
        {VLCU_TRACE_SIG, 1, "stage", 13, 256, -1, 255, 0},
        {VLCU_TRACE_SIG, 1, "hash", 1, 128, -1, 127, 0},
        {VLCU_TRACE_SIG, 1, "sh", 21, 8, -1, 7, 0},
        {VLCU_TRACE_SIG, 1, "op", 22, 3, -1, 2, 0},
        {VLCU_TRACE_SIG, 0, "vld", 0, 1, -1, 0, 0},
        {VLCU_TRACE_SIG, 1, "sum", 23, 256, -1, 255, 0},
        {VLCU_TRACE_SIG, 1, "dif", 31, 256, -1, 255, 0},
        {VLCU_TRACE_SIG, 1, "shl", 39, 256, -1, 255, 0},
        {VLCU_TRACE_SIG, 1, "shr", 47, 256, -1, 255, 0},
        {VLCU_TRACE_SIG, 1, "sra", 55, 256, -1, 255, 0},
        {VLCU_TRACE_SIG, 1, "prod", 63, 128, -1, 127, 0},
        {VLCU_TRACE_SIG, 1, "mix", 67, 256, -1, 255, 0},
        {VLCU_TRACE_SIG, 1, "res", 75, 256, -1, 255, 0},
        {VLCU_TRACE_SIG, 1, "slice", 99, 70, -1, 69, 0},
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0},
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0}};  // (terminator, not counted)
    static const uint32_t traceOrder[] = {0, 1, 5, 13, 21, 22, 23, 31, 39, 47, 55, 63, 67, 75, 83, 84, 85, 86, 87, 88, 96, 97, 98, 99, 0};  // codes in dump order (+ terminator)
    static const uint32_t saveOrder[] = {0, 1, 2, 3, 4, 6, 7, 8, 13, 14, 15, 16, 17, 18, 19, 20, 9, 5, 10, 11, 12, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "wide_bus", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 0, plusargs, 0,
        traceDecls, 45, T::kTraceWords, traceOrder, 24,
        saveOrder, 32, VLCU_SAVE_ACTIVITY, 0x2819d23c62a96db6ULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vwide_bus__cuda VLCU_TESTBENCHES);
