// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vrnd01"
#endif

// Constant pool
static __device__ const VlWide<8>/*255:0*/ Vrnd01__ConstPool__CONST_h113dda8b_0 = VlWide<8>{{
        0xfa4faf5d, 0xd674136c, 0x31c14190, 0x00000000,
        0x00000000, 0x00000000, 0x00000000, 0x00000000
}};

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vrnd01___024root final {
    CData/*0:0*/ clk;
    VlWide<3>/*95:0*/ i0;
    QData/*47:0*/ i1;
    CData/*7:0*/ i2;
    IData/*30:0*/ i3;
    QData/*39:0*/ i4;
    QData/*32:0*/ i5;
    CData/*0:0*/ o0;
    SData/*15:0*/ o1;
    IData/*31:0*/ o2;
    QData/*39:0*/ o3;
    VlWide<3>/*95:0*/ o4;
    CData/*0:0*/ __VdfgRegularize_h6e95ff9d_0_0;
    CData/*0:0*/ __VdfgRegularize_h6e95ff9d_0_3;
    CData/*0:0*/ __VdfgRegularize_h6e95ff9d_0_4;
    CData/*0:0*/ __VdfgRegularize_h6e95ff9d_0_5;
    CData/*0:0*/ __VstlFirstIteration;
    CData/*0:0*/ __VstlPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*7:0*/ __Vtrigprevexpr___TOP__i2__0;
    CData/*0:0*/ __VicoDidInit;
    CData/*0:0*/ __VicoPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__1;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    IData/*31:0*/ rnd01__DOT__r0;
    VlWide<4>/*99:0*/ rnd01__DOT__r1;
    VlWide<5>/*149:0*/ rnd01__DOT__r3;
    VlWide<5>/*149:0*/ rnd01__DOT__w8;
    VlWide<3>/*64:0*/ rnd01__DOT__w11;
    VlWide<4>/*99:0*/ __VdfgRegularize_h6e95ff9d_0_2;
    VlWide<3>/*95:0*/ __Vtrigprevexpr___TOP__i0__0;
    IData/*30:0*/ __Vtrigprevexpr___TOP__i3__0;
    IData/*31:0*/ __VactIterCount;
    QData/*63:0*/ rnd01__DOT__r2;
    QData/*47:0*/ __VdfgRegularize_h6e95ff9d_0_1;
    QData/*47:0*/ __Vtrigprevexpr___TOP__i1__0;
    QData/*39:0*/ __Vtrigprevexpr___TOP__i4__0;
    QData/*32:0*/ __Vtrigprevexpr___TOP__i5__0;
    VlUnpacked<QData/*63:0*/, 1> __VstlTriggered;
    VlUnpacked<QData/*63:0*/, 2> __VicoTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vrnd01__cuda final {
    typedef Vrnd01___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 254,
        OFF_i0 = 112,
        OFF_i1 = 0,
        OFF_i2 = 255,
        OFF_i3 = 124,
        OFF_i4 = 8,
        OFF_i5 = 16,
        OFF_o0 = 256,
        OFF_o1 = 252,
        OFF_o2 = 128,
        OFF_o3 = 24,
        OFF_o4 = 132,
        OFF___VdfgRegularize_h6e95ff9d_0_0 = 257,
        OFF___VdfgRegularize_h6e95ff9d_0_3 = 258,
        OFF___VdfgRegularize_h6e95ff9d_0_4 = 259,
        OFF___VdfgRegularize_h6e95ff9d_0_5 = 260,
        OFF___VstlFirstIteration = 261,
        OFF___VstlPhaseResult = 262,
        OFF___Vtrigprevexpr___TOP__clk__0 = 263,
        OFF___Vtrigprevexpr___TOP__i2__0 = 264,
        OFF___VicoDidInit = 265,
        OFF___VicoPhaseResult = 266,
        OFF___Vtrigprevexpr___TOP__clk__1 = 267,
        OFF___VactPhaseResult = 268,
        OFF___VnbaPhaseResult = 269,
        OFF_rnd01__DOT__r0 = 144,
        OFF_rnd01__DOT__r1 = 148,
        OFF_rnd01__DOT__r3 = 164,
        OFF_rnd01__DOT__w8 = 184,
        OFF_rnd01__DOT__w11 = 204,
        OFF___VdfgRegularize_h6e95ff9d_0_2 = 216,
        OFF___Vtrigprevexpr___TOP__i0__0 = 232,
        OFF___Vtrigprevexpr___TOP__i3__0 = 244,
        OFF___VactIterCount = 248,
        OFF_rnd01__DOT__r2 = 32,
        OFF___VdfgRegularize_h6e95ff9d_0_1 = 40,
        OFF___Vtrigprevexpr___TOP__i1__0 = 48,
        OFF___Vtrigprevexpr___TOP__i4__0 = 56,
        OFF___Vtrigprevexpr___TOP__i5__0 = 64,
        OFF___VstlTriggered = 72,
        OFF___VicoTriggered = 80,
        OFF___VactTriggered = 96,
        OFF___VnbaTriggered = 104,
        kStateBytes = 272,  // S_persist
        kTouchedReadBytes = 233,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 218,  // S_state_wr
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
        st.ldW(OFF_i0, i, s.i0);
        s.i1 = st.ld<QData>(OFF_i1, i);
        s.i2 = st.ld<VlcuMem8>(OFF_i2, i);
        s.i3 = st.ld<IData>(OFF_i3, i);
        s.i4 = st.ld<QData>(OFF_i4, i);
        s.i5 = st.ld<QData>(OFF_i5, i);
        s.o0 = st.ld<VlcuMem8>(OFF_o0, i);
        s.o1 = st.ld<VlcuMem16>(OFF_o1, i);
        s.o2 = st.ld<IData>(OFF_o2, i);
        s.o3 = st.ld<QData>(OFF_o3, i);
        st.ldW(OFF_o4, i, s.o4);
        s.__VdfgRegularize_h6e95ff9d_0_0 = st.ld<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_0, i);
        s.__VdfgRegularize_h6e95ff9d_0_3 = st.ld<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_3, i);
        s.__VdfgRegularize_h6e95ff9d_0_4 = st.ld<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_4, i);
        s.__VdfgRegularize_h6e95ff9d_0_5 = st.ld<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_5, i);
        s.__VstlFirstIteration = st.ld<VlcuMem8>(OFF___VstlFirstIteration, i);
        s.__VstlPhaseResult = st.ld<VlcuMem8>(OFF___VstlPhaseResult, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__Vtrigprevexpr___TOP__i2__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__i2__0, i);
        s.__VicoDidInit = st.ld<VlcuMem8>(OFF___VicoDidInit, i);
        s.__VicoPhaseResult = st.ld<VlcuMem8>(OFF___VicoPhaseResult, i);
        s.__Vtrigprevexpr___TOP__clk__1 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__1, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        s.rnd01__DOT__r0 = st.ld<IData>(OFF_rnd01__DOT__r0, i);
        st.ldW(OFF_rnd01__DOT__r1, i, s.rnd01__DOT__r1);
        st.ldW(OFF_rnd01__DOT__r3, i, s.rnd01__DOT__r3);
        st.ldW(OFF_rnd01__DOT__w8, i, s.rnd01__DOT__w8);
        st.ldW(OFF_rnd01__DOT__w11, i, s.rnd01__DOT__w11);
        st.ldW(OFF___VdfgRegularize_h6e95ff9d_0_2, i, s.__VdfgRegularize_h6e95ff9d_0_2);
        st.ldW(OFF___Vtrigprevexpr___TOP__i0__0, i, s.__Vtrigprevexpr___TOP__i0__0);
        s.__Vtrigprevexpr___TOP__i3__0 = st.ld<IData>(OFF___Vtrigprevexpr___TOP__i3__0, i);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        s.rnd01__DOT__r2 = st.ld<QData>(OFF_rnd01__DOT__r2, i);
        s.__VdfgRegularize_h6e95ff9d_0_1 = st.ld<QData>(OFF___VdfgRegularize_h6e95ff9d_0_1, i);
        s.__Vtrigprevexpr___TOP__i1__0 = st.ld<QData>(OFF___Vtrigprevexpr___TOP__i1__0, i);
        s.__Vtrigprevexpr___TOP__i4__0 = st.ld<QData>(OFF___Vtrigprevexpr___TOP__i4__0, i);
        s.__Vtrigprevexpr___TOP__i5__0 = st.ld<QData>(OFF___Vtrigprevexpr___TOP__i5__0, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VstlTriggered[e] = st.ldE<QData>(OFF___VstlTriggered, e, i);
        for (uint32_t e = 0; e < 2; ++e) s.__VicoTriggered[e] = st.ldE<QData>(OFF___VicoTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.stW(OFF_i0, i, s.i0);
        st.st<QData>(OFF_i1, i, s.i1);
        st.st<VlcuMem8>(OFF_i2, i, s.i2);
        st.st<IData>(OFF_i3, i, s.i3);
        st.st<QData>(OFF_i4, i, s.i4);
        st.st<QData>(OFF_i5, i, s.i5);
        st.st<VlcuMem8>(OFF_o0, i, s.o0);
        st.st<VlcuMem16>(OFF_o1, i, s.o1);
        st.st<IData>(OFF_o2, i, s.o2);
        st.st<QData>(OFF_o3, i, s.o3);
        st.stW(OFF_o4, i, s.o4);
        st.st<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_0, i, s.__VdfgRegularize_h6e95ff9d_0_0);
        st.st<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_3, i, s.__VdfgRegularize_h6e95ff9d_0_3);
        st.st<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_4, i, s.__VdfgRegularize_h6e95ff9d_0_4);
        st.st<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_5, i, s.__VdfgRegularize_h6e95ff9d_0_5);
        st.st<VlcuMem8>(OFF___VstlFirstIteration, i, s.__VstlFirstIteration);
        st.st<VlcuMem8>(OFF___VstlPhaseResult, i, s.__VstlPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__i2__0, i, s.__Vtrigprevexpr___TOP__i2__0);
        st.st<VlcuMem8>(OFF___VicoDidInit, i, s.__VicoDidInit);
        st.st<VlcuMem8>(OFF___VicoPhaseResult, i, s.__VicoPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__1, i, s.__Vtrigprevexpr___TOP__clk__1);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.st<IData>(OFF_rnd01__DOT__r0, i, s.rnd01__DOT__r0);
        st.stW(OFF_rnd01__DOT__r1, i, s.rnd01__DOT__r1);
        st.stW(OFF_rnd01__DOT__r3, i, s.rnd01__DOT__r3);
        st.stW(OFF_rnd01__DOT__w8, i, s.rnd01__DOT__w8);
        st.stW(OFF_rnd01__DOT__w11, i, s.rnd01__DOT__w11);
        st.stW(OFF___VdfgRegularize_h6e95ff9d_0_2, i, s.__VdfgRegularize_h6e95ff9d_0_2);
        st.stW(OFF___Vtrigprevexpr___TOP__i0__0, i, s.__Vtrigprevexpr___TOP__i0__0);
        st.st<IData>(OFF___Vtrigprevexpr___TOP__i3__0, i, s.__Vtrigprevexpr___TOP__i3__0);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
        st.st<QData>(OFF_rnd01__DOT__r2, i, s.rnd01__DOT__r2);
        st.st<QData>(OFF___VdfgRegularize_h6e95ff9d_0_1, i, s.__VdfgRegularize_h6e95ff9d_0_1);
        st.st<QData>(OFF___Vtrigprevexpr___TOP__i1__0, i, s.__Vtrigprevexpr___TOP__i1__0);
        st.st<QData>(OFF___Vtrigprevexpr___TOP__i4__0, i, s.__Vtrigprevexpr___TOP__i4__0);
        st.st<QData>(OFF___Vtrigprevexpr___TOP__i5__0, i, s.__Vtrigprevexpr___TOP__i5__0);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VstlTriggered, e, i, s.__VstlTriggered[e]);
        for (uint32_t e = 0; e < 2; ++e) st.stE<QData>(OFF___VicoTriggered, e, i, s.__VicoTriggered[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VactTriggered, e, i, s.__VactTriggered[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VnbaTriggered, e, i, s.__VnbaTriggered[e]);
    }
    VLCU_D static void load_inputs_row(Root& s, const uint32_t* rowp, uint64_t stride, uint64_t i) {
        s.clk = static_cast<VlcuMem8>(rowp[0 * stride + i] & 0x1U);
        s.i0[0] = rowp[1 * stride + i] & 0xffffffffU;
        s.i0[1] = rowp[2 * stride + i] & 0xffffffffU;
        s.i0[2] = rowp[3 * stride + i] & 0xffffffffU;
        s.i1 = (static_cast<QData>(rowp[5 * stride + i] & 0xffffU) << 32) | rowp[4 * stride + i];
        s.i2 = static_cast<VlcuMem8>(rowp[6 * stride + i] & 0xffU);
        s.i3 = static_cast<IData>(rowp[7 * stride + i] & 0x7fffffffU);
        s.i4 = (static_cast<QData>(rowp[9 * stride + i] & 0xffU) << 32) | rowp[8 * stride + i];
        s.i5 = (static_cast<QData>(rowp[11 * stride + i] & 0x1U) << 32) | rowp[10 * stride + i];
    }
    VLCU_D static void store_outputs_row(const Root& s, uint32_t* rowp, uint64_t stride, uint64_t i) {
        rowp[0 * stride + i] = s.o0;
        rowp[1 * stride + i] = s.o1;
        rowp[2 * stride + i] = s.o2;
        rowp[3 * stride + i] = static_cast<uint32_t>(s.o3);
        rowp[4 * stride + i] = static_cast<uint32_t>(s.o3 >> 32);
        rowp[5 * stride + i] = s.o4[0];
        rowp[6 * stride + i] = s.o4[1];
        rowp[7 * stride + i] = s.o4[2];
    }
    VLCU_D static uint64_t outputs_sig(const Root& s, uint64_t sig) {
        uint32_t acc = VLCU_SIG_ACC_INIT;
        acc = vlcu_sig_word(acc, s.o0);
        acc = vlcu_sig_word(acc, s.o1);
        acc = vlcu_sig_word(acc, s.o2);
        acc = vlcu_sig_word(acc, static_cast<uint32_t>(s.o3));
        acc = vlcu_sig_word(acc, static_cast<uint32_t>(s.o3 >> 32));
        acc = vlcu_sig_word(acc, s.o4[0]);
        acc = vlcu_sig_word(acc, s.o4[1]);
        acc = vlcu_sig_word(acc, s.o4[2]);
        return vlcu_sig_mix(sig, acc);
    }
    enum : uint32_t { kInWords = 12, kOutWords = 8, kInBytes = 42, kOutBytes = 27, kTraceWords = 0 };

    VLCU_D static void Vrnd01___024root___eval_static(Vrnd01___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
        vlSelfRef.__Vtrigprevexpr___TOP__i0__0[0U] 
            = vlSelfRef.i0[0U];
        vlSelfRef.__Vtrigprevexpr___TOP__i0__0[1U] 
            = vlSelfRef.i0[1U];
        vlSelfRef.__Vtrigprevexpr___TOP__i0__0[2U] 
            = vlSelfRef.i0[2U];
        vlSelfRef.__Vtrigprevexpr___TOP__i1__0 = vlSelfRef.i1;
        vlSelfRef.__Vtrigprevexpr___TOP__i2__0 = vlSelfRef.i2;
        vlSelfRef.__Vtrigprevexpr___TOP__i3__0 = vlSelfRef.i3;
        vlSelfRef.__Vtrigprevexpr___TOP__i4__0 = vlSelfRef.i4;
        vlSelfRef.__Vtrigprevexpr___TOP__i5__0 = vlSelfRef.i5;
        vlSelfRef.__Vtrigprevexpr___TOP__clk__1 = vlSelfRef.clk;
    }

    VLCU_D static void Vrnd01___024root___eval_initial(Vrnd01___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vrnd01___024root___eval_final(Vrnd01___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vrnd01___024root___eval_settle(Vrnd01___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vrnd01___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/rnd01.v", 2, "", "DIDNOTCONVERGE: Settle region did not converge after '--converge-limit' of 10000 tries");
            }
            __VstlIterCount = ((IData)(1U) + __VstlIterCount);
            vlSelfRef.__VstlPhaseResult = Vrnd01___024root___eval_phase__stl(vlSelf, vlCtx);
            vlSelfRef.__VstlFirstIteration = 0U;
        } while (vlSelfRef.__VstlPhaseResult);
    }

    VLCU_D static bool Vrnd01___024root___trigger_anySet__stl(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vrnd01___024root___stl_sequent__TOP__0(Vrnd01___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        VlWide<4>/*127:0*/ __Vtemp_1;
        VlWide<4>/*127:0*/ __Vtemp_3;
        VlWide<4>/*127:0*/ __Vtemp_4;
        VlWide<4>/*127:0*/ __Vtemp_5;
        VlWide<4>/*127:0*/ __Vtemp_8;
        VlWide<4>/*127:0*/ __Vtemp_9;
        VlWide<4>/*127:0*/ __Vtemp_10;
        VlWide<4>/*127:0*/ __Vtemp_11;
        VlWide<4>/*127:0*/ __Vtemp_13;
        VlWide<4>/*127:0*/ __Vtemp_14;
        VlWide<4>/*127:0*/ __Vtemp_15;
        VlWide<4>/*127:0*/ __Vtemp_18;
        VlWide<4>/*127:0*/ __Vtemp_19;
        VlWide<4>/*127:0*/ __Vtemp_20;
        VlWide<4>/*127:0*/ __Vtemp_21;
        VlWide<4>/*127:0*/ __Vtemp_22;
        VlWide<5>/*159:0*/ __Vtemp_24;
        VlWide<5>/*159:0*/ __Vtemp_25;
        VlWide<5>/*159:0*/ __Vtemp_27;
        VlWide<5>/*159:0*/ __Vtemp_28;
        VlWide<8>/*255:0*/ __Vtemp_30;
        VlWide<8>/*255:0*/ __Vtemp_35;
        VlWide<8>/*255:0*/ __Vtemp_36;
        VlWide<8>/*255:0*/ __Vtemp_37;
        VlWide<5>/*159:0*/ __Vtemp_41;
        VlWide<5>/*159:0*/ __Vtemp_43;
        VlWide<5>/*159:0*/ __Vtemp_44;
        VlWide<3>/*95:0*/ __Vtemp_46;
        VlWide<3>/*95:0*/ __Vtemp_47;
        VlWide<3>/*95:0*/ __Vtemp_48;
        VlWide<3>/*95:0*/ __Vtemp_49;
        VlWide<3>/*95:0*/ __Vtemp_50;
        VlWide<3>/*95:0*/ __Vtemp_51;
        VlWide<3>/*95:0*/ __Vtemp_52;
        // Body
        vlSelfRef.o2 = (IData)((0x000000ffffffffffULL 
                                & VL_DIVS_QQQ(40, (0x000000ffffffffffULL 
                                                   & VL_EXTENDS_QI(40,32, vlSelfRef.rnd01__DOT__r0)), vlSelfRef.i4)));
        __Vtemp_1[0U] = 0U;
        __Vtemp_1[1U] = 0U;
        __Vtemp_1[2U] = 0U;
        __Vtemp_1[3U] = 0U;
        __Vtemp_3[0U] = vlSelfRef.i0[0U];
        __Vtemp_3[1U] = vlSelfRef.i0[1U];
        __Vtemp_3[2U] = vlSelfRef.i0[2U];
        __Vtemp_3[3U] = 0U;
        VL_MUL_W(4, __Vtemp_4, vlSelfRef.rnd01__DOT__r1, __Vtemp_3);
        __Vtemp_5[0U] = __Vtemp_4[0U];
        __Vtemp_5[1U] = __Vtemp_4[1U];
        __Vtemp_5[2U] = __Vtemp_4[2U];
        __Vtemp_5[3U] = (0x0000000fU & __Vtemp_4[3U]);
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_3 = 
            VL_LTS_IWW(100, __Vtemp_1, __Vtemp_5);
        __Vtemp_8[0U] = vlSelfRef.i2;
        __Vtemp_8[1U] = 0U;
        __Vtemp_8[2U] = 0U;
        __Vtemp_8[3U] = 0U;
        __Vtemp_9[0U] = vlSelfRef.rnd01__DOT__r0;
        __Vtemp_9[1U] = 0U;
        __Vtemp_9[2U] = 0U;
        __Vtemp_9[3U] = 0U;
        VL_MODDIV_WWW(100, __Vtemp_10, __Vtemp_8, __Vtemp_9);
        __Vtemp_11[0U] = __Vtemp_10[0U];
        __Vtemp_11[1U] = __Vtemp_10[1U];
        __Vtemp_11[2U] = __Vtemp_10[2U];
        __Vtemp_11[3U] = (0x0000000fU & __Vtemp_10[3U]);
        __Vtemp_13[0U] = vlSelfRef.rnd01__DOT__r1[0U];
        __Vtemp_13[1U] = (0x00008000U ^ vlSelfRef.rnd01__DOT__r1[1U]);
        __Vtemp_13[2U] = vlSelfRef.rnd01__DOT__r1[2U];
        __Vtemp_13[3U] = vlSelfRef.rnd01__DOT__r1[3U];
        VL_MUL_W(4, __Vtemp_14, __Vtemp_11, __Vtemp_13);
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[0U] 
            = __Vtemp_14[0U];
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[1U] 
            = __Vtemp_14[1U];
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[2U] 
            = __Vtemp_14[2U];
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[3U] 
            = (0x0000000fU & __Vtemp_14[3U]);
        VL_EXTENDS_WI(128,1, __Vtemp_15, (1U & (~ (
                                                   (vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2
                                                    [
                                                    (1U 
                                                     & ((IData)(vlSelfRef.i2) 
                                                        >> 5U))] 
                                                    >> 
                                                    (0x0000001fU 
                                                     & (IData)(vlSelfRef.i2))) 
                                                   | (0U 
                                                      != 
                                                      (((QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[1U])) 
                                                        << 0x00000020U) 
                                                       | (QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[0U]))))))));
        __Vtemp_18[0U] = (vlSelfRef.rnd01__DOT__r1[0U] 
                          | vlSelfRef.i3);
        __Vtemp_18[1U] = vlSelfRef.rnd01__DOT__r1[1U];
        __Vtemp_18[2U] = vlSelfRef.rnd01__DOT__r1[2U];
        __Vtemp_18[3U] = vlSelfRef.rnd01__DOT__r1[3U];
        __Vtemp_19[0U] = (IData)(vlSelfRef.i1);
        __Vtemp_19[1U] = (IData)((vlSelfRef.i1 >> 0x00000020U));
        __Vtemp_19[2U] = 0U;
        __Vtemp_19[3U] = 0U;
        __Vtemp_20[0U] = 0x33da3873U;
        __Vtemp_20[1U] = 0xf87305faU;
        __Vtemp_20[2U] = 0x96158e5fU;
        __Vtemp_20[3U] = 0x665c8735U;
        VL_SUB_W(4, __Vtemp_21, __Vtemp_19, __Vtemp_20);
        VL_SUB_W(4, __Vtemp_22, __Vtemp_18, __Vtemp_21);
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_4 = 
            VL_GTS_IWW(128, __Vtemp_15, __Vtemp_22);
        VL_EXTENDS_WQ(150,64, __Vtemp_24, (((QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[1U])) 
                                            << 0x00000020U) 
                                           | (QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[0U]))));
        __Vtemp_25[0U] = __Vtemp_24[0U];
        __Vtemp_25[1U] = __Vtemp_24[1U];
        __Vtemp_25[2U] = __Vtemp_24[2U];
        __Vtemp_25[3U] = __Vtemp_24[3U];
        __Vtemp_25[4U] = (0x003fffffU & __Vtemp_24[4U]);
        VL_SHIFTRS_WWI(150,150,8, __Vtemp_27, vlSelfRef.rnd01__DOT__r3, 
                       (0x000000ffU & vlSelfRef.i3));
        __Vtemp_28[0U] = __Vtemp_27[0U];
        __Vtemp_28[1U] = __Vtemp_27[1U];
        __Vtemp_28[2U] = __Vtemp_27[2U];
        __Vtemp_28[3U] = __Vtemp_27[3U];
        __Vtemp_28[4U] = (0x003fffffU & __Vtemp_27[4U]);
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_5 = 
            VL_GTS_III(31, (0x7fffffffU & VL_EXTENDS_II(31,1, 
                                                        VL_LTS_IWW(150, __Vtemp_25, __Vtemp_28))), 
                       (0x7fffffffU & (VL_MULS_III(31, vlSelfRef.i3, 
                                                   (0x7fffffffU 
                                                    & VL_EXTENDS_II(31,8, (IData)(vlSelfRef.i2)))) 
                                       << (vlSelfRef.i4 
                                           > vlSelfRef.i5))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1 = 
            (0x0000ffffffffffffULL & (vlSelfRef.i1 
                                      * (QData)((IData)(
                                                        ((VL_SHIFTR_III(28,28,8, 
                                                                        (vlSelfRef.rnd01__DOT__r0 
                                                                         >> 4U), (IData)(vlSelfRef.i2)) 
                                                          << 1U) 
                                                         | ((((QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[1U])) 
                                                              << 0x00000020U) 
                                                             | (QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[0U]))) 
                                                            <= vlSelfRef.i4))))));
        VL_SHIFTR_WWI(250,250,31, __Vtemp_30, Vrnd01__ConstPool__CONST_h113dda8b_0, vlSelfRef.i3);
        __Vtemp_35[0U] = (0x0000007fU & (vlSelfRef.i3 
                                         >> 2U));
        __Vtemp_35[1U] = 0U;
        __Vtemp_35[2U] = 0U;
        __Vtemp_35[3U] = 0U;
        __Vtemp_35[4U] = 0U;
        __Vtemp_35[5U] = 0U;
        __Vtemp_35[6U] = 0U;
        __Vtemp_35[7U] = 0U;
        __Vtemp_36[0U] = vlSelfRef.__VdfgRegularize_h6e95ff9d_0_3;
        __Vtemp_36[1U] = 0U;
        __Vtemp_36[2U] = 0U;
        __Vtemp_36[3U] = 0U;
        __Vtemp_36[4U] = 0U;
        __Vtemp_36[5U] = 0U;
        __Vtemp_36[6U] = 0U;
        __Vtemp_36[7U] = 0U;
        VL_MUL_W(8, __Vtemp_37, __Vtemp_35, __Vtemp_36);
        vlSelfRef.rnd01__DOT__w8[0U] = ((~ (__Vtemp_30[0U] 
                                            ^ (0x0000000fU 
                                               & ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_5) 
                                                  >> 
                                                  (0x0000000fU 
                                                   & (IData)(vlSelfRef.i4)))))) 
                                        ^ __Vtemp_37[0U]);
        vlSelfRef.rnd01__DOT__w8[1U] = ((~ __Vtemp_30[1U]) 
                                        ^ __Vtemp_37[1U]);
        vlSelfRef.rnd01__DOT__w8[2U] = ((~ __Vtemp_30[2U]) 
                                        ^ __Vtemp_37[2U]);
        vlSelfRef.rnd01__DOT__w8[3U] = ((~ __Vtemp_30[3U]) 
                                        ^ __Vtemp_37[3U]);
        vlSelfRef.rnd01__DOT__w8[4U] = (0x003fffffU 
                                        & ((~ __Vtemp_30[4U]) 
                                           ^ __Vtemp_37[4U]));
        __Vtemp_41[0U] = 0U;
        __Vtemp_41[1U] = 0U;
        __Vtemp_41[2U] = 0U;
        __Vtemp_41[3U] = 0U;
        __Vtemp_41[4U] = 0U;
        vlSelfRef.o1 = (0x0000ffffU & ((0U != (((QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[1U])) 
                                                << 0x00000020U) 
                                               | (QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[0U]))))
                                        ? VL_MODDIV_III(32, vlSelfRef.i3, 
                                                        (0x00ffffffU 
                                                         & (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1)))
                                        : (~ ((VL_COUNTONES_I(
                                                              (1U 
                                                               & (vlSelfRef.i0
                                                                  [
                                                                  (1U 
                                                                   & (vlSelfRef.rnd01__DOT__r3[0U] 
                                                                      >> 5U))] 
                                                                  >> 
                                                                  (0x0000001fU 
                                                                   & vlSelfRef.rnd01__DOT__r3[0U])))) 
                                               + VL_ONEHOT_I(
                                                             (0x000007ffU 
                                                              & (vlSelfRef.i3 
                                                                 >> 0x00000011U)))) 
                                              ^ VL_GTE_W(5, __Vtemp_41, vlSelfRef.rnd01__DOT__r3)))));
        VL_EXTENDS_WI(150,24, __Vtemp_43, (0x00ffffffU 
                                           & (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1)));
        __Vtemp_44[0U] = __Vtemp_43[0U];
        __Vtemp_44[1U] = __Vtemp_43[1U];
        __Vtemp_44[2U] = __Vtemp_43[2U];
        __Vtemp_44[3U] = __Vtemp_43[3U];
        __Vtemp_44[4U] = (0x003fffffU & __Vtemp_43[4U]);
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0 = 
            VL_GTES_IWW(150, vlSelfRef.rnd01__DOT__r3, __Vtemp_44);
        if ((1U & VL_SHIFTL_III(1,1,8, (~ (0U != ((
                                                   ((vlSelfRef.rnd01__DOT__w8[0U] 
                                                     | vlSelfRef.rnd01__DOT__w8[1U]) 
                                                    | vlSelfRef.rnd01__DOT__w8[2U]) 
                                                   | vlSelfRef.rnd01__DOT__w8[3U]) 
                                                  | vlSelfRef.rnd01__DOT__w8[4U]))), 
                                (0x000000ffU & vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[0U])))) {
            __Vtemp_47[0U] = (0x0000007fU & (IData)(vlSelfRef.i2));
            __Vtemp_47[1U] = 0U;
            __Vtemp_47[2U] = 0U;
            __Vtemp_48[0U] = ((0U != (0x00ffffffU & (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1))) 
                              & ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_3) 
                                 >> (1U & (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1))));
            __Vtemp_48[1U] = 0U;
            __Vtemp_48[2U] = 0U;
            VL_ADD_W(3, __Vtemp_46, __Vtemp_47, __Vtemp_48);
        } else {
            __Vtemp_49[0U] = 0x00000094U;
            __Vtemp_49[1U] = 0U;
            __Vtemp_49[2U] = 0U;
            __Vtemp_50[0U] = VL_ONEHOT_Q((0x000000ffffffffffULL 
                                          & VL_DIVS_QQQ(40, 
                                                        (0x000000ffffffffffULL 
                                                         & vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1), 
                                                        (0x000000ffffffffffULL 
                                                         & VL_EXTENDS_QI(40,32, vlSelfRef.rnd01__DOT__r0)))));
            __Vtemp_50[1U] = 0U;
            __Vtemp_50[2U] = 0U;
            VL_ADD_W(3, __Vtemp_46, __Vtemp_49, __Vtemp_50);
        }
        vlSelfRef.rnd01__DOT__w11[0U] = __Vtemp_46[0U];
        vlSelfRef.rnd01__DOT__w11[1U] = __Vtemp_46[1U];
        vlSelfRef.rnd01__DOT__w11[2U] = (1U & __Vtemp_46[2U]);
        vlSelfRef.o0 = (1U & ((~ (vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[0U] 
                                  ^ (IData)(VL_SHIFTL_QQI(64,64,8, vlSelfRef.i5, 
                                                          (0x000000ffU 
                                                           & (IData)(vlSelfRef.i1)))))) 
                              ^ (IData)((1ULL >> (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)))));
        vlSelfRef.o3 = (0x000000ffffffffffULL & (vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1 
                                                 | ((QData)((IData)(
                                                                    VL_ONEHOT_W(5, vlSelfRef.rnd01__DOT__w8))) 
                                                    + (QData)((IData)(
                                                                      VL_COUNTONES_I(
                                                                                ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0) 
                                                                                <= 
                                                                                (1U 
                                                                                & (IData)(
                                                                                (vlSelfRef.i5 
                                                                                >> 
                                                                                (0x0000001fU 
                                                                                & vlSelfRef.rnd01__DOT__r3[0U])))))))))));
        if (vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0) {
            __Vtemp_52[0U] = vlSelfRef.__VdfgRegularize_h6e95ff9d_0_5;
            __Vtemp_52[1U] = 0U;
            __Vtemp_52[2U] = 0U;
            VL_SHIFTL_WWI(96,96,8, __Vtemp_51, __Vtemp_52, 
                          (0x000000ffU & vlSelfRef.rnd01__DOT__w11[0U]));
        } else {
            __Vtemp_51[0U] = (IData)(vlSelfRef.i1);
            __Vtemp_51[1U] = (IData)((vlSelfRef.i1 
                                      >> 0x00000020U));
            __Vtemp_51[2U] = 0U;
        }
        vlSelfRef.o4[0U] = (__Vtemp_51[0U] ^ (IData)(
                                                     (((QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[1U])) 
                                                       << 0x00000020U) 
                                                      | (QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[0U])))));
        vlSelfRef.o4[1U] = (__Vtemp_51[1U] ^ (IData)(
                                                     ((((QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[1U])) 
                                                        << 0x00000020U) 
                                                       | (QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[0U]))) 
                                                      >> 0x00000020U)));
        vlSelfRef.o4[2U] = __Vtemp_51[2U];
    }

    VLCU_D static bool Vrnd01___024root___eval_phase__stl(Vrnd01___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vrnd01___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
        }
#endif
        __VstlExecute = Vrnd01___024root___trigger_anySet__stl(vlSelfRef.__VstlTriggered);
        if (__VstlExecute) {
            {
                // Inlined CFunc: _eval_stl
                if ((1ULL & vlSelfRef.__VstlTriggered[0U])) {
                    Vrnd01___024root___stl_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
        }
        return (__VstlExecute);
    }

    VLCU_D static bool Vrnd01___024root___trigger_anySet__ico(const VlUnpacked<QData/*63:0*/, 2> &in) {
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

    VLCU_D static void Vrnd01___024root___ico_comb__TOP__6(Vrnd01___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        VlWide<3>/*95:0*/ __Vtemp_2;
        VlWide<3>/*95:0*/ __Vtemp_3;
        VlWide<3>/*95:0*/ __Vtemp_4;
        VlWide<3>/*95:0*/ __Vtemp_5;
        VlWide<3>/*95:0*/ __Vtemp_6;
        VlWide<3>/*95:0*/ __Vtemp_7;
        VlWide<3>/*95:0*/ __Vtemp_8;
        // Body
        vlSelfRef.o3 = (0x000000ffffffffffULL & (vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1 
                                                 | ((QData)((IData)(
                                                                    VL_ONEHOT_W(5, vlSelfRef.rnd01__DOT__w8))) 
                                                    + (QData)((IData)(
                                                                      VL_COUNTONES_I(
                                                                                ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0) 
                                                                                <= 
                                                                                (1U 
                                                                                & (IData)(
                                                                                (vlSelfRef.i5 
                                                                                >> 
                                                                                (0x0000001fU 
                                                                                & vlSelfRef.rnd01__DOT__r3[0U])))))))))));
        if ((1U & VL_SHIFTL_III(1,1,8, (~ (0U != ((
                                                   ((vlSelfRef.rnd01__DOT__w8[0U] 
                                                     | vlSelfRef.rnd01__DOT__w8[1U]) 
                                                    | vlSelfRef.rnd01__DOT__w8[2U]) 
                                                   | vlSelfRef.rnd01__DOT__w8[3U]) 
                                                  | vlSelfRef.rnd01__DOT__w8[4U]))), 
                                (0x000000ffU & vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[0U])))) {
            __Vtemp_3[0U] = (0x0000007fU & (IData)(vlSelfRef.i2));
            __Vtemp_3[1U] = 0U;
            __Vtemp_3[2U] = 0U;
            __Vtemp_4[0U] = ((0U != (0x00ffffffU & (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1))) 
                             & ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_3) 
                                >> (1U & (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1))));
            __Vtemp_4[1U] = 0U;
            __Vtemp_4[2U] = 0U;
            VL_ADD_W(3, __Vtemp_2, __Vtemp_3, __Vtemp_4);
        } else {
            __Vtemp_5[0U] = 0x00000094U;
            __Vtemp_5[1U] = 0U;
            __Vtemp_5[2U] = 0U;
            __Vtemp_6[0U] = VL_ONEHOT_Q((0x000000ffffffffffULL 
                                         & VL_DIVS_QQQ(40, 
                                                       (0x000000ffffffffffULL 
                                                        & vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1), 
                                                       (0x000000ffffffffffULL 
                                                        & VL_EXTENDS_QI(40,32, vlSelfRef.rnd01__DOT__r0)))));
            __Vtemp_6[1U] = 0U;
            __Vtemp_6[2U] = 0U;
            VL_ADD_W(3, __Vtemp_2, __Vtemp_5, __Vtemp_6);
        }
        vlSelfRef.rnd01__DOT__w11[0U] = __Vtemp_2[0U];
        vlSelfRef.rnd01__DOT__w11[1U] = __Vtemp_2[1U];
        vlSelfRef.rnd01__DOT__w11[2U] = (1U & __Vtemp_2[2U]);
        if (vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0) {
            __Vtemp_8[0U] = vlSelfRef.__VdfgRegularize_h6e95ff9d_0_5;
            __Vtemp_8[1U] = 0U;
            __Vtemp_8[2U] = 0U;
            VL_SHIFTL_WWI(96,96,8, __Vtemp_7, __Vtemp_8, 
                          (0x000000ffU & vlSelfRef.rnd01__DOT__w11[0U]));
        } else {
            __Vtemp_7[0U] = (IData)(vlSelfRef.i1);
            __Vtemp_7[1U] = (IData)((vlSelfRef.i1 >> 0x00000020U));
            __Vtemp_7[2U] = 0U;
        }
        vlSelfRef.o4[0U] = (__Vtemp_7[0U] ^ (IData)(
                                                    (((QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[1U])) 
                                                      << 0x00000020U) 
                                                     | (QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[0U])))));
        vlSelfRef.o4[1U] = (__Vtemp_7[1U] ^ (IData)(
                                                    ((((QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[1U])) 
                                                       << 0x00000020U) 
                                                      | (QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[0U]))) 
                                                     >> 0x00000020U)));
        vlSelfRef.o4[2U] = __Vtemp_7[2U];
    }

    VLCU_D static void Vrnd01___024root___eval_ico(Vrnd01___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        if ((0x0000000000000020ULL & vlSelfRef.__VicoTriggered[0U])) {
            {
                // Inlined CFunc: _ico_sequent__TOP__0
                vlSelfRef.o2 = (IData)((0x000000ffffffffffULL 
                                        & VL_DIVS_QQQ(40, 
                                                      (0x000000ffffffffffULL 
                                                       & VL_EXTENDS_QI(40,32, vlSelfRef.rnd01__DOT__r0)), vlSelfRef.i4)));
            }
        }
        if ((2ULL & vlSelfRef.__VicoTriggered[0U])) {
            {
                // Inlined CFunc: _ico_sequent__TOP__1
                VlWide<4>/*127:0*/ __Vinline_0__ico_sequent__TOP__1___Vtemp_1;
                VlWide<4>/*127:0*/ __Vinline_0__ico_sequent__TOP__1___Vtemp_3;
                VlWide<4>/*127:0*/ __Vinline_0__ico_sequent__TOP__1___Vtemp_4;
                VlWide<4>/*127:0*/ __Vinline_0__ico_sequent__TOP__1___Vtemp_5;
                __Vinline_0__ico_sequent__TOP__1___Vtemp_1[0U] = 0U;
                __Vinline_0__ico_sequent__TOP__1___Vtemp_1[1U] = 0U;
                __Vinline_0__ico_sequent__TOP__1___Vtemp_1[2U] = 0U;
                __Vinline_0__ico_sequent__TOP__1___Vtemp_1[3U] = 0U;
                __Vinline_0__ico_sequent__TOP__1___Vtemp_3[0U] 
                    = vlSelfRef.i0[0U];
                __Vinline_0__ico_sequent__TOP__1___Vtemp_3[1U] 
                    = vlSelfRef.i0[1U];
                __Vinline_0__ico_sequent__TOP__1___Vtemp_3[2U] 
                    = vlSelfRef.i0[2U];
                __Vinline_0__ico_sequent__TOP__1___Vtemp_3[3U] = 0U;
                VL_MUL_W(4, __Vinline_0__ico_sequent__TOP__1___Vtemp_4, vlSelfRef.rnd01__DOT__r1, __Vinline_0__ico_sequent__TOP__1___Vtemp_3);
                __Vinline_0__ico_sequent__TOP__1___Vtemp_5[0U] 
                    = __Vinline_0__ico_sequent__TOP__1___Vtemp_4[0U];
                __Vinline_0__ico_sequent__TOP__1___Vtemp_5[1U] 
                    = __Vinline_0__ico_sequent__TOP__1___Vtemp_4[1U];
                __Vinline_0__ico_sequent__TOP__1___Vtemp_5[2U] 
                    = __Vinline_0__ico_sequent__TOP__1___Vtemp_4[2U];
                __Vinline_0__ico_sequent__TOP__1___Vtemp_5[3U] 
                    = (0x0000000fU & __Vinline_0__ico_sequent__TOP__1___Vtemp_4[3U]);
                vlSelfRef.__VdfgRegularize_h6e95ff9d_0_3 
                    = VL_LTS_IWW(100, __Vinline_0__ico_sequent__TOP__1___Vtemp_1, __Vinline_0__ico_sequent__TOP__1___Vtemp_5);
            }
        }
        if ((8ULL & vlSelfRef.__VicoTriggered[0U])) {
            {
                // Inlined CFunc: _ico_sequent__TOP__2
                VlWide<4>/*127:0*/ __Vinline_0__ico_sequent__TOP__2___Vtemp_3;
                VlWide<4>/*127:0*/ __Vinline_0__ico_sequent__TOP__2___Vtemp_4;
                VlWide<4>/*127:0*/ __Vinline_0__ico_sequent__TOP__2___Vtemp_5;
                VlWide<4>/*127:0*/ __Vinline_0__ico_sequent__TOP__2___Vtemp_6;
                VlWide<4>/*127:0*/ __Vinline_0__ico_sequent__TOP__2___Vtemp_8;
                VlWide<4>/*127:0*/ __Vinline_0__ico_sequent__TOP__2___Vtemp_9;
                __Vinline_0__ico_sequent__TOP__2___Vtemp_3[0U] 
                    = vlSelfRef.i2;
                __Vinline_0__ico_sequent__TOP__2___Vtemp_3[1U] = 0U;
                __Vinline_0__ico_sequent__TOP__2___Vtemp_3[2U] = 0U;
                __Vinline_0__ico_sequent__TOP__2___Vtemp_3[3U] = 0U;
                __Vinline_0__ico_sequent__TOP__2___Vtemp_4[0U] 
                    = vlSelfRef.rnd01__DOT__r0;
                __Vinline_0__ico_sequent__TOP__2___Vtemp_4[1U] = 0U;
                __Vinline_0__ico_sequent__TOP__2___Vtemp_4[2U] = 0U;
                __Vinline_0__ico_sequent__TOP__2___Vtemp_4[3U] = 0U;
                VL_MODDIV_WWW(100, __Vinline_0__ico_sequent__TOP__2___Vtemp_5, __Vinline_0__ico_sequent__TOP__2___Vtemp_3, __Vinline_0__ico_sequent__TOP__2___Vtemp_4);
                __Vinline_0__ico_sequent__TOP__2___Vtemp_6[0U] 
                    = __Vinline_0__ico_sequent__TOP__2___Vtemp_5[0U];
                __Vinline_0__ico_sequent__TOP__2___Vtemp_6[1U] 
                    = __Vinline_0__ico_sequent__TOP__2___Vtemp_5[1U];
                __Vinline_0__ico_sequent__TOP__2___Vtemp_6[2U] 
                    = __Vinline_0__ico_sequent__TOP__2___Vtemp_5[2U];
                __Vinline_0__ico_sequent__TOP__2___Vtemp_6[3U] 
                    = (0x0000000fU & __Vinline_0__ico_sequent__TOP__2___Vtemp_5[3U]);
                __Vinline_0__ico_sequent__TOP__2___Vtemp_8[0U] 
                    = vlSelfRef.rnd01__DOT__r1[0U];
                __Vinline_0__ico_sequent__TOP__2___Vtemp_8[1U] 
                    = (0x00008000U ^ vlSelfRef.rnd01__DOT__r1[1U]);
                __Vinline_0__ico_sequent__TOP__2___Vtemp_8[2U] 
                    = vlSelfRef.rnd01__DOT__r1[2U];
                __Vinline_0__ico_sequent__TOP__2___Vtemp_8[3U] 
                    = vlSelfRef.rnd01__DOT__r1[3U];
                VL_MUL_W(4, __Vinline_0__ico_sequent__TOP__2___Vtemp_9, __Vinline_0__ico_sequent__TOP__2___Vtemp_6, __Vinline_0__ico_sequent__TOP__2___Vtemp_8);
                vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[0U] 
                    = __Vinline_0__ico_sequent__TOP__2___Vtemp_9[0U];
                vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[1U] 
                    = __Vinline_0__ico_sequent__TOP__2___Vtemp_9[1U];
                vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[2U] 
                    = __Vinline_0__ico_sequent__TOP__2___Vtemp_9[2U];
                vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[3U] 
                    = (0x0000000fU & __Vinline_0__ico_sequent__TOP__2___Vtemp_9[3U]);
            }
        }
        if ((0x000000000000001cULL & vlSelfRef.__VicoTriggered[0U])) {
            {
                // Inlined CFunc: _ico_comb__TOP__0
                VlWide<4>/*127:0*/ __Vinline_0__ico_comb__TOP__0___Vtemp_1;
                VlWide<4>/*127:0*/ __Vinline_0__ico_comb__TOP__0___Vtemp_4;
                VlWide<4>/*127:0*/ __Vinline_0__ico_comb__TOP__0___Vtemp_5;
                VlWide<4>/*127:0*/ __Vinline_0__ico_comb__TOP__0___Vtemp_6;
                VlWide<4>/*127:0*/ __Vinline_0__ico_comb__TOP__0___Vtemp_7;
                VlWide<4>/*127:0*/ __Vinline_0__ico_comb__TOP__0___Vtemp_8;
                VL_EXTENDS_WI(128,1, __Vinline_0__ico_comb__TOP__0___Vtemp_1, 
                              (1U & (~ ((vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2
                                         [(1U & ((IData)(vlSelfRef.i2) 
                                                 >> 5U))] 
                                         >> (0x0000001fU 
                                             & (IData)(vlSelfRef.i2))) 
                                        | (0U != (((QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[1U])) 
                                                   << 0x00000020U) 
                                                  | (QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[0U]))))))));
                __Vinline_0__ico_comb__TOP__0___Vtemp_4[0U] 
                    = (vlSelfRef.rnd01__DOT__r1[0U] 
                       | vlSelfRef.i3);
                __Vinline_0__ico_comb__TOP__0___Vtemp_4[1U] 
                    = vlSelfRef.rnd01__DOT__r1[1U];
                __Vinline_0__ico_comb__TOP__0___Vtemp_4[2U] 
                    = vlSelfRef.rnd01__DOT__r1[2U];
                __Vinline_0__ico_comb__TOP__0___Vtemp_4[3U] 
                    = vlSelfRef.rnd01__DOT__r1[3U];
                __Vinline_0__ico_comb__TOP__0___Vtemp_5[0U] 
                    = (IData)(vlSelfRef.i1);
                __Vinline_0__ico_comb__TOP__0___Vtemp_5[1U] 
                    = (IData)((vlSelfRef.i1 >> 0x00000020U));
                __Vinline_0__ico_comb__TOP__0___Vtemp_5[2U] = 0U;
                __Vinline_0__ico_comb__TOP__0___Vtemp_5[3U] = 0U;
                __Vinline_0__ico_comb__TOP__0___Vtemp_6[0U] = 0x33da3873U;
                __Vinline_0__ico_comb__TOP__0___Vtemp_6[1U] = 0xf87305faU;
                __Vinline_0__ico_comb__TOP__0___Vtemp_6[2U] = 0x96158e5fU;
                __Vinline_0__ico_comb__TOP__0___Vtemp_6[3U] = 0x665c8735U;
                VL_SUB_W(4, __Vinline_0__ico_comb__TOP__0___Vtemp_7, __Vinline_0__ico_comb__TOP__0___Vtemp_5, __Vinline_0__ico_comb__TOP__0___Vtemp_6);
                VL_SUB_W(4, __Vinline_0__ico_comb__TOP__0___Vtemp_8, __Vinline_0__ico_comb__TOP__0___Vtemp_4, __Vinline_0__ico_comb__TOP__0___Vtemp_7);
                vlSelfRef.__VdfgRegularize_h6e95ff9d_0_4 
                    = VL_GTS_IWW(128, __Vinline_0__ico_comb__TOP__0___Vtemp_1, __Vinline_0__ico_comb__TOP__0___Vtemp_8);
            }
        }
        if ((0x0000000000000078ULL & vlSelfRef.__VicoTriggered[0U])) {
            {
                // Inlined CFunc: _ico_comb__TOP__1
                VlWide<5>/*159:0*/ __Vinline_0__ico_comb__TOP__1___Vtemp_2;
                VlWide<5>/*159:0*/ __Vinline_0__ico_comb__TOP__1___Vtemp_3;
                VlWide<5>/*159:0*/ __Vinline_0__ico_comb__TOP__1___Vtemp_5;
                VlWide<5>/*159:0*/ __Vinline_0__ico_comb__TOP__1___Vtemp_6;
                VL_EXTENDS_WQ(150,64, __Vinline_0__ico_comb__TOP__1___Vtemp_2, 
                              (((QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[1U])) 
                                << 0x00000020U) | (QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[0U]))));
                __Vinline_0__ico_comb__TOP__1___Vtemp_3[0U] 
                    = __Vinline_0__ico_comb__TOP__1___Vtemp_2[0U];
                __Vinline_0__ico_comb__TOP__1___Vtemp_3[1U] 
                    = __Vinline_0__ico_comb__TOP__1___Vtemp_2[1U];
                __Vinline_0__ico_comb__TOP__1___Vtemp_3[2U] 
                    = __Vinline_0__ico_comb__TOP__1___Vtemp_2[2U];
                __Vinline_0__ico_comb__TOP__1___Vtemp_3[3U] 
                    = __Vinline_0__ico_comb__TOP__1___Vtemp_2[3U];
                __Vinline_0__ico_comb__TOP__1___Vtemp_3[4U] 
                    = (0x003fffffU & __Vinline_0__ico_comb__TOP__1___Vtemp_2[4U]);
                VL_SHIFTRS_WWI(150,150,8, __Vinline_0__ico_comb__TOP__1___Vtemp_5, vlSelfRef.rnd01__DOT__r3, 
                               (0x000000ffU & vlSelfRef.i3));
                __Vinline_0__ico_comb__TOP__1___Vtemp_6[0U] 
                    = __Vinline_0__ico_comb__TOP__1___Vtemp_5[0U];
                __Vinline_0__ico_comb__TOP__1___Vtemp_6[1U] 
                    = __Vinline_0__ico_comb__TOP__1___Vtemp_5[1U];
                __Vinline_0__ico_comb__TOP__1___Vtemp_6[2U] 
                    = __Vinline_0__ico_comb__TOP__1___Vtemp_5[2U];
                __Vinline_0__ico_comb__TOP__1___Vtemp_6[3U] 
                    = __Vinline_0__ico_comb__TOP__1___Vtemp_5[3U];
                __Vinline_0__ico_comb__TOP__1___Vtemp_6[4U] 
                    = (0x003fffffU & __Vinline_0__ico_comb__TOP__1___Vtemp_5[4U]);
                vlSelfRef.__VdfgRegularize_h6e95ff9d_0_5 
                    = VL_GTS_III(31, (0x7fffffffU & 
                                      VL_EXTENDS_II(31,1, 
                                                    VL_LTS_IWW(150, __Vinline_0__ico_comb__TOP__1___Vtemp_3, __Vinline_0__ico_comb__TOP__1___Vtemp_6))), 
                                 (0x7fffffffU & (VL_MULS_III(31, vlSelfRef.i3, 
                                                             (0x7fffffffU 
                                                              & VL_EXTENDS_II(31,8, (IData)(vlSelfRef.i2)))) 
                                                 << 
                                                 (vlSelfRef.i4 
                                                  > vlSelfRef.i5))));
            }
        }
        if ((0x000000000000002cULL & vlSelfRef.__VicoTriggered[0U])) {
            {
                // Inlined CFunc: _ico_comb__TOP__2
                VlWide<5>/*159:0*/ __Vinline_0__ico_comb__TOP__2___Vtemp_2;
                VlWide<5>/*159:0*/ __Vinline_0__ico_comb__TOP__2___Vtemp_3;
                vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1 
                    = (0x0000ffffffffffffULL & (vlSelfRef.i1 
                                                * (QData)((IData)(
                                                                  ((VL_SHIFTR_III(28,28,8, 
                                                                                (vlSelfRef.rnd01__DOT__r0 
                                                                                >> 4U), (IData)(vlSelfRef.i2)) 
                                                                    << 1U) 
                                                                   | ((((QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[1U])) 
                                                                        << 0x00000020U) 
                                                                       | (QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[0U]))) 
                                                                      <= vlSelfRef.i4))))));
                VL_EXTENDS_WI(150,24, __Vinline_0__ico_comb__TOP__2___Vtemp_2, 
                              (0x00ffffffU & (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1)));
                __Vinline_0__ico_comb__TOP__2___Vtemp_3[0U] 
                    = __Vinline_0__ico_comb__TOP__2___Vtemp_2[0U];
                __Vinline_0__ico_comb__TOP__2___Vtemp_3[1U] 
                    = __Vinline_0__ico_comb__TOP__2___Vtemp_2[1U];
                __Vinline_0__ico_comb__TOP__2___Vtemp_3[2U] 
                    = __Vinline_0__ico_comb__TOP__2___Vtemp_2[2U];
                __Vinline_0__ico_comb__TOP__2___Vtemp_3[3U] 
                    = __Vinline_0__ico_comb__TOP__2___Vtemp_2[3U];
                __Vinline_0__ico_comb__TOP__2___Vtemp_3[4U] 
                    = (0x003fffffU & __Vinline_0__ico_comb__TOP__2___Vtemp_2[4U]);
                vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0 
                    = VL_GTES_IWW(150, vlSelfRef.rnd01__DOT__r3, __Vinline_0__ico_comb__TOP__2___Vtemp_3);
            }
        }
        if ((0x000000000000007aULL & vlSelfRef.__VicoTriggered[0U])) {
            {
                // Inlined CFunc: _ico_comb__TOP__3
                VlWide<8>/*255:0*/ __Vinline_0__ico_comb__TOP__3___Vtemp_2;
                VlWide<8>/*255:0*/ __Vinline_0__ico_comb__TOP__3___Vtemp_7;
                VlWide<8>/*255:0*/ __Vinline_0__ico_comb__TOP__3___Vtemp_8;
                VlWide<8>/*255:0*/ __Vinline_0__ico_comb__TOP__3___Vtemp_9;
                VL_SHIFTR_WWI(250,250,31, __Vinline_0__ico_comb__TOP__3___Vtemp_2, Vrnd01__ConstPool__CONST_h113dda8b_0, vlSelfRef.i3);
                __Vinline_0__ico_comb__TOP__3___Vtemp_7[0U] 
                    = (0x0000007fU & (vlSelfRef.i3 
                                      >> 2U));
                __Vinline_0__ico_comb__TOP__3___Vtemp_7[1U] = 0U;
                __Vinline_0__ico_comb__TOP__3___Vtemp_7[2U] = 0U;
                __Vinline_0__ico_comb__TOP__3___Vtemp_7[3U] = 0U;
                __Vinline_0__ico_comb__TOP__3___Vtemp_7[4U] = 0U;
                __Vinline_0__ico_comb__TOP__3___Vtemp_7[5U] = 0U;
                __Vinline_0__ico_comb__TOP__3___Vtemp_7[6U] = 0U;
                __Vinline_0__ico_comb__TOP__3___Vtemp_7[7U] = 0U;
                __Vinline_0__ico_comb__TOP__3___Vtemp_8[0U] 
                    = vlSelfRef.__VdfgRegularize_h6e95ff9d_0_3;
                __Vinline_0__ico_comb__TOP__3___Vtemp_8[1U] = 0U;
                __Vinline_0__ico_comb__TOP__3___Vtemp_8[2U] = 0U;
                __Vinline_0__ico_comb__TOP__3___Vtemp_8[3U] = 0U;
                __Vinline_0__ico_comb__TOP__3___Vtemp_8[4U] = 0U;
                __Vinline_0__ico_comb__TOP__3___Vtemp_8[5U] = 0U;
                __Vinline_0__ico_comb__TOP__3___Vtemp_8[6U] = 0U;
                __Vinline_0__ico_comb__TOP__3___Vtemp_8[7U] = 0U;
                VL_MUL_W(8, __Vinline_0__ico_comb__TOP__3___Vtemp_9, __Vinline_0__ico_comb__TOP__3___Vtemp_7, __Vinline_0__ico_comb__TOP__3___Vtemp_8);
                vlSelfRef.rnd01__DOT__w8[0U] = ((~ 
                                                 (__Vinline_0__ico_comb__TOP__3___Vtemp_2[0U] 
                                                  ^ 
                                                  (0x0000000fU 
                                                   & ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_5) 
                                                      >> 
                                                      (0x0000000fU 
                                                       & (IData)(vlSelfRef.i4)))))) 
                                                ^ __Vinline_0__ico_comb__TOP__3___Vtemp_9[0U]);
                vlSelfRef.rnd01__DOT__w8[1U] = ((~ __Vinline_0__ico_comb__TOP__3___Vtemp_2[1U]) 
                                                ^ __Vinline_0__ico_comb__TOP__3___Vtemp_9[1U]);
                vlSelfRef.rnd01__DOT__w8[2U] = ((~ __Vinline_0__ico_comb__TOP__3___Vtemp_2[2U]) 
                                                ^ __Vinline_0__ico_comb__TOP__3___Vtemp_9[2U]);
                vlSelfRef.rnd01__DOT__w8[3U] = ((~ __Vinline_0__ico_comb__TOP__3___Vtemp_2[3U]) 
                                                ^ __Vinline_0__ico_comb__TOP__3___Vtemp_9[3U]);
                vlSelfRef.rnd01__DOT__w8[4U] = (0x003fffffU 
                                                & ((~ __Vinline_0__ico_comb__TOP__3___Vtemp_2[4U]) 
                                                   ^ __Vinline_0__ico_comb__TOP__3___Vtemp_9[4U]));
            }
        }
        if ((0x000000000000003eULL & vlSelfRef.__VicoTriggered[0U])) {
            {
                // Inlined CFunc: _ico_comb__TOP__4
                VlWide<5>/*159:0*/ __Vinline_0__ico_comb__TOP__4___Vtemp_1;
                __Vinline_0__ico_comb__TOP__4___Vtemp_1[0U] = 0U;
                __Vinline_0__ico_comb__TOP__4___Vtemp_1[1U] = 0U;
                __Vinline_0__ico_comb__TOP__4___Vtemp_1[2U] = 0U;
                __Vinline_0__ico_comb__TOP__4___Vtemp_1[3U] = 0U;
                __Vinline_0__ico_comb__TOP__4___Vtemp_1[4U] = 0U;
                vlSelfRef.o1 = (0x0000ffffU & ((0U 
                                                != 
                                                (((QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[1U])) 
                                                  << 0x00000020U) 
                                                 | (QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[0U]))))
                                                ? VL_MODDIV_III(32, vlSelfRef.i3, 
                                                                (0x00ffffffU 
                                                                 & (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1)))
                                                : (~ 
                                                   ((VL_COUNTONES_I(
                                                                    (1U 
                                                                     & (vlSelfRef.i0
                                                                        [
                                                                        (1U 
                                                                         & (vlSelfRef.rnd01__DOT__r3[0U] 
                                                                            >> 5U))] 
                                                                        >> 
                                                                        (0x0000001fU 
                                                                         & vlSelfRef.rnd01__DOT__r3[0U])))) 
                                                     + 
                                                     VL_ONEHOT_I(
                                                                 (0x000007ffU 
                                                                  & (vlSelfRef.i3 
                                                                     >> 0x00000011U)))) 
                                                    ^ 
                                                    VL_GTE_W(5, __Vinline_0__ico_comb__TOP__4___Vtemp_1, vlSelfRef.rnd01__DOT__r3)))));
            }
        }
        if ((0x000000000000006cULL & vlSelfRef.__VicoTriggered[0U])) {
            {
                // Inlined CFunc: _ico_comb__TOP__5
                vlSelfRef.o0 = (1U & ((~ (vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[0U] 
                                          ^ (IData)(
                                                    VL_SHIFTL_QQI(64,64,8, vlSelfRef.i5, 
                                                                  (0x000000ffU 
                                                                   & (IData)(vlSelfRef.i1)))))) 
                                      ^ (IData)((1ULL 
                                                 >> (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)))));
            }
        }
        if ((0x000000000000007eULL & vlSelfRef.__VicoTriggered[0U])) {
            Vrnd01___024root___ico_comb__TOP__6(vlSelf, vlCtx);
        }
    }

    VLCU_D static bool Vrnd01___024root___eval_phase__ico(Vrnd01___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VicoExecute;
        // Body
        {
            // Inlined CFunc: _eval_triggers_vec__ico
            vlSelfRef.__VicoTriggered[0U] = (QData)((IData)(
                                                            ((((vlSelfRef.i5 
                                                                != vlSelfRef.__Vtrigprevexpr___TOP__i5__0) 
                                                               << 6U) 
                                                              | (((vlSelfRef.i4 
                                                                   != vlSelfRef.__Vtrigprevexpr___TOP__i4__0) 
                                                                  << 5U) 
                                                                 | ((vlSelfRef.i3 
                                                                     != vlSelfRef.__Vtrigprevexpr___TOP__i3__0) 
                                                                    << 4U))) 
                                                             | (((((IData)(vlSelfRef.i2) 
                                                                   != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__i2__0)) 
                                                                  << 3U) 
                                                                 | ((vlSelfRef.i1 
                                                                     != vlSelfRef.__Vtrigprevexpr___TOP__i1__0) 
                                                                    << 2U)) 
                                                                | (((0U 
                                                                     != 
                                                                     (((vlSelfRef.i0[0U] 
                                                                        ^ vlSelfRef.__Vtrigprevexpr___TOP__i0__0[0U]) 
                                                                       | (vlSelfRef.i0[1U] 
                                                                          ^ vlSelfRef.__Vtrigprevexpr___TOP__i0__0[1U])) 
                                                                      | (vlSelfRef.i0[2U] 
                                                                         ^ vlSelfRef.__Vtrigprevexpr___TOP__i0__0[2U]))) 
                                                                    << 1U) 
                                                                   | ((IData)(vlSelfRef.clk) 
                                                                      != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__clk__0)))))));
            vlSelfRef.__Vtrigprevexpr___TOP__clk__0 
                = vlSelfRef.clk;
            vlSelfRef.__Vtrigprevexpr___TOP__i0__0[0U] 
                = vlSelfRef.i0[0U];
            vlSelfRef.__Vtrigprevexpr___TOP__i0__0[1U] 
                = vlSelfRef.i0[1U];
            vlSelfRef.__Vtrigprevexpr___TOP__i0__0[2U] 
                = vlSelfRef.i0[2U];
            vlSelfRef.__Vtrigprevexpr___TOP__i1__0 
                = vlSelfRef.i1;
            vlSelfRef.__Vtrigprevexpr___TOP__i2__0 
                = vlSelfRef.i2;
            vlSelfRef.__Vtrigprevexpr___TOP__i3__0 
                = vlSelfRef.i3;
            vlSelfRef.__Vtrigprevexpr___TOP__i4__0 
                = vlSelfRef.i4;
            vlSelfRef.__Vtrigprevexpr___TOP__i5__0 
                = vlSelfRef.i5;
            if (VL_UNLIKELY(((1U & (~ (IData)(vlSelfRef.__VicoDidInit)))))) {
                vlSelfRef.__VicoDidInit = 1U;
                vlSelfRef.__VicoTriggered[0U] = (1ULL 
                                                 | vlSelfRef.__VicoTriggered[0U]);
                vlSelfRef.__VicoTriggered[0U] = (2ULL 
                                                 | vlSelfRef.__VicoTriggered[0U]);
                vlSelfRef.__VicoTriggered[0U] = (4ULL 
                                                 | vlSelfRef.__VicoTriggered[0U]);
                vlSelfRef.__VicoTriggered[0U] = (8ULL 
                                                 | vlSelfRef.__VicoTriggered[0U]);
                vlSelfRef.__VicoTriggered[0U] = (0x0000000000000010ULL 
                                                 | vlSelfRef.__VicoTriggered[0U]);
                vlSelfRef.__VicoTriggered[0U] = (0x0000000000000020ULL 
                                                 | vlSelfRef.__VicoTriggered[0U]);
                vlSelfRef.__VicoTriggered[0U] = (0x0000000000000040ULL 
                                                 | vlSelfRef.__VicoTriggered[0U]);
            }
        }
#ifdef VL_DEBUG
        if (VL_UNLIKELY(vlSymsp->_vm_contextp__->debug())) {
            Vrnd01___024root___dump_triggers__ico(vlSelfRef.__VicoTriggered, "ico"s);
        }
#endif
        __VicoExecute = Vrnd01___024root___trigger_anySet__ico(vlSelfRef.__VicoTriggered);
        if (__VicoExecute) {
            Vrnd01___024root___eval_ico(vlSelf, vlCtx);
        }
        return (__VicoExecute);
    }

    VLCU_D static bool Vrnd01___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vrnd01___024root___nba_sequent__TOP__0(Vrnd01___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        VlUnpacked<IData/*23:0*/, 16> rnd01__DOT__mem;
        for (int __Vi0 = 0; __Vi0 < 16; ++__Vi0) {
            rnd01__DOT__mem[__Vi0] = 0;
        }
        VlWide<4>/*99:0*/ __Vdly__rnd01__DOT__r1;
        VL_ZERO_W(100, __Vdly__rnd01__DOT__r1);
        QData/*63:0*/ __Vdly__rnd01__DOT__r2;
        __Vdly__rnd01__DOT__r2 = 0;
        VlWide<5>/*149:0*/ __Vdly__rnd01__DOT__r3;
        VL_ZERO_W(150, __Vdly__rnd01__DOT__r3);
        IData/*23:0*/ __VdlyVal__rnd01__DOT__mem__v0;
        __VdlyVal__rnd01__DOT__mem__v0 = 0;
        CData/*3:0*/ __VdlyDim0__rnd01__DOT__mem__v0;
        __VdlyDim0__rnd01__DOT__mem__v0 = 0;
        VlWide<4>/*127:0*/ __Vtemp_2;
        VlWide<5>/*159:0*/ __Vtemp_5;
        VlWide<5>/*159:0*/ __Vtemp_7;
        VlWide<5>/*159:0*/ __Vtemp_8;
        VlWide<5>/*159:0*/ __Vtemp_9;
        VlWide<5>/*159:0*/ __Vtemp_10;
        VlWide<5>/*159:0*/ __Vtemp_11;
        VlWide<4>/*127:0*/ __Vtemp_12;
        VlWide<4>/*127:0*/ __Vtemp_14;
        VlWide<4>/*127:0*/ __Vtemp_15;
        VlWide<4>/*127:0*/ __Vtemp_16;
        VlWide<4>/*127:0*/ __Vtemp_19;
        VlWide<4>/*127:0*/ __Vtemp_20;
        VlWide<4>/*127:0*/ __Vtemp_21;
        VlWide<4>/*127:0*/ __Vtemp_22;
        VlWide<4>/*127:0*/ __Vtemp_24;
        VlWide<4>/*127:0*/ __Vtemp_25;
        VlWide<4>/*127:0*/ __Vtemp_26;
        VlWide<4>/*127:0*/ __Vtemp_29;
        VlWide<4>/*127:0*/ __Vtemp_30;
        VlWide<4>/*127:0*/ __Vtemp_31;
        VlWide<4>/*127:0*/ __Vtemp_32;
        VlWide<4>/*127:0*/ __Vtemp_33;
        VlWide<5>/*159:0*/ __Vtemp_35;
        VlWide<5>/*159:0*/ __Vtemp_36;
        VlWide<5>/*159:0*/ __Vtemp_38;
        VlWide<5>/*159:0*/ __Vtemp_39;
        VlWide<8>/*255:0*/ __Vtemp_41;
        VlWide<8>/*255:0*/ __Vtemp_46;
        VlWide<8>/*255:0*/ __Vtemp_47;
        VlWide<8>/*255:0*/ __Vtemp_48;
        VlWide<5>/*159:0*/ __Vtemp_52;
        VlWide<5>/*159:0*/ __Vtemp_54;
        VlWide<5>/*159:0*/ __Vtemp_55;
        VlWide<3>/*95:0*/ __Vtemp_57;
        VlWide<3>/*95:0*/ __Vtemp_58;
        VlWide<3>/*95:0*/ __Vtemp_59;
        VlWide<3>/*95:0*/ __Vtemp_60;
        VlWide<3>/*95:0*/ __Vtemp_61;
        VlWide<3>/*95:0*/ __Vtemp_62;
        VlWide<3>/*95:0*/ __Vtemp_63;
        // Body
        __Vdly__rnd01__DOT__r2 = vlSelfRef.rnd01__DOT__r2;
        __Vdly__rnd01__DOT__r3[0U] = vlSelfRef.rnd01__DOT__r3[0U];
        __Vdly__rnd01__DOT__r3[1U] = vlSelfRef.rnd01__DOT__r3[1U];
        __Vdly__rnd01__DOT__r3[2U] = vlSelfRef.rnd01__DOT__r3[2U];
        __Vdly__rnd01__DOT__r3[3U] = vlSelfRef.rnd01__DOT__r3[3U];
        __Vdly__rnd01__DOT__r3[4U] = vlSelfRef.rnd01__DOT__r3[4U];
        __Vdly__rnd01__DOT__r1[0U] = vlSelfRef.rnd01__DOT__r1[0U];
        __Vdly__rnd01__DOT__r1[1U] = vlSelfRef.rnd01__DOT__r1[1U];
        __Vdly__rnd01__DOT__r1[2U] = vlSelfRef.rnd01__DOT__r1[2U];
        __Vdly__rnd01__DOT__r1[3U] = vlSelfRef.rnd01__DOT__r1[3U];
        __Vdly__rnd01__DOT__r1[0U] = ((0xfffffffeU 
                                       & __Vdly__rnd01__DOT__r1[0U]) 
                                      | VL_GTES_III(2, 
                                                    ((1U 
                                                      & (vlSelfRef.rnd01__DOT__r3
                                                         [
                                                         (3U 
                                                          & ((IData)(vlSelfRef.rnd01__DOT__r2) 
                                                             >> 5U))] 
                                                         >> 
                                                         (0x0000001fU 
                                                          & (IData)(vlSelfRef.rnd01__DOT__r2)))) 
                                                     | (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_3)), 
                                                    (3U 
                                                     & VL_EXTENDS_II(2,1, 
                                                                     (1U 
                                                                      == vlSelfRef.rnd01__DOT__r0)))));
        __Vdly__rnd01__DOT__r1[0U] = (1U & __Vdly__rnd01__DOT__r1[0U]);
        __Vdly__rnd01__DOT__r1[1U] = 0U;
        __Vdly__rnd01__DOT__r1[2U] = 0U;
        __Vdly__rnd01__DOT__r1[3U] = 0U;
        if (VL_LTES_III(14, (0x00003fffU & VL_EXTENDS_II(14,8, 
                                                         (0x000000ffU 
                                                          & VL_SHIFTRS_III(8,8,8, (IData)(vlSelfRef.i2), (IData)(vlSelfRef.i2))))), 
                        VL_SHIFTR_IIQ(14,14,48, (0x00003fffU 
                                                 & (vlSelfRef.rnd01__DOT__w11[0U] 
                                                    >> 0x00000011U)), vlSelfRef.i1))) {
            __Vtemp_2[0U] = vlSelfRef.i0[0U];
            __Vtemp_2[1U] = vlSelfRef.i0[1U];
            __Vtemp_2[2U] = vlSelfRef.i0[2U];
            __Vtemp_2[3U] = 0U;
        } else if ((0U != VL_SHIFTR_QQI(64,64,8, (((QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[1U])) 
                                                   << 0x00000020U) 
                                                  | (QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[0U]))), 
                                        (0x000000ffU 
                                         & (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1))))) {
            __Vtemp_2[0U] = ((1U & (((vlSelfRef.rnd01__DOT__r0 
                                      ^ (1U & (~ (0U 
                                                  != 
                                                  (0x0003ffffU 
                                                   & (vlSelfRef.rnd01__DOT__r3[1U] 
                                                      >> 1U)))))) 
                                     >> (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_4)) 
                                    >> (7U & (IData)(vlSelfRef.i4))))
                              ? (IData)(vlSelfRef.i2)
                              : 1U);
            __Vtemp_2[1U] = 0U;
            __Vtemp_2[2U] = 0U;
            __Vtemp_2[3U] = 0U;
        } else {
            __Vtemp_2[0U] = vlSelfRef.rnd01__DOT__r1[0U];
            __Vtemp_2[1U] = vlSelfRef.rnd01__DOT__r1[1U];
        }
        __Vdly__rnd01__DOT__r2 = (((QData)((IData)(__Vtemp_2[1U])) 
                                   << 0x00000020U) 
                                  | (QData)((IData)(__Vtemp_2[0U])));
        if ((0U != ((((vlSelfRef.rnd01__DOT__w8[0U] 
                       | vlSelfRef.rnd01__DOT__w8[1U]) 
                      | vlSelfRef.rnd01__DOT__w8[2U]) 
                     | vlSelfRef.rnd01__DOT__w8[3U]) 
                    | vlSelfRef.rnd01__DOT__w8[4U]))) {
            __Vtemp_5[0U] = vlSelfRef.__VdfgRegularize_h6e95ff9d_0_3;
            __Vtemp_5[1U] = 0U;
            __Vtemp_5[2U] = 0U;
            __Vtemp_5[3U] = 0U;
        } else {
            __Vtemp_5[0U] = vlSelfRef.rnd01__DOT__r1[0U];
            __Vtemp_5[1U] = vlSelfRef.rnd01__DOT__r1[1U];
            __Vtemp_5[2U] = vlSelfRef.rnd01__DOT__r1[2U];
            __Vtemp_5[3U] = vlSelfRef.rnd01__DOT__r1[3U];
        }
        __Vtemp_5[4U] = 0U;
        __Vtemp_7[0U] = ((1U & (IData)((vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1 
                                        >> 0x00000010U))) 
                         | vlSelfRef.rnd01__DOT__r3[0U]);
        __Vtemp_7[1U] = vlSelfRef.rnd01__DOT__r3[1U];
        __Vtemp_7[2U] = vlSelfRef.rnd01__DOT__r3[2U];
        __Vtemp_7[3U] = vlSelfRef.rnd01__DOT__r3[3U];
        __Vtemp_7[4U] = vlSelfRef.rnd01__DOT__r3[4U];
        VL_MUL_W(5, __Vtemp_8, __Vtemp_5, __Vtemp_7);
        __Vtemp_9[0U] = __Vtemp_8[0U];
        __Vtemp_9[1U] = __Vtemp_8[1U];
        __Vtemp_9[2U] = __Vtemp_8[2U];
        __Vtemp_9[3U] = __Vtemp_8[3U];
        __Vtemp_9[4U] = (0x003fffffU & __Vtemp_8[4U]);
        VL_SHIFTRS_WWI(150,150,8, __Vtemp_10, __Vtemp_9, 
                       (0x000000ffU & (IData)(vlSelfRef.rnd01__DOT__r2)));
        __Vdly__rnd01__DOT__r3[0U] = __Vtemp_10[0U];
        __Vdly__rnd01__DOT__r3[1U] = __Vtemp_10[1U];
        __Vdly__rnd01__DOT__r3[2U] = __Vtemp_10[2U];
        __Vdly__rnd01__DOT__r3[3U] = __Vtemp_10[3U];
        __Vdly__rnd01__DOT__r3[4U] = (0x003fffffU & __Vtemp_10[4U]);
        if ((0U != vlSelfRef.i3)) {
            __Vtemp_11[0U] = (IData)((((QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[1U])) 
                                       << 0x00000020U) 
                                      | (QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[0U]))));
            __Vtemp_11[1U] = (IData)(((((QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[1U])) 
                                        << 0x00000020U) 
                                       | (QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[0U]))) 
                                      >> 0x00000020U));
            __Vtemp_11[2U] = 0U;
            __Vtemp_11[3U] = 0U;
            __Vtemp_11[4U] = 0U;
        } else {
            __Vtemp_11[0U] = vlSelfRef.rnd01__DOT__r3[0U];
        }
        __VdlyVal__rnd01__DOT__mem__v0 = (0x00ffffffU 
                                          & (__Vtemp_11[0U] 
                                             ^ VL_GTES_III(31, 0x40000000U, 
                                                           (0x7fffffffU 
                                                            & VL_EXTENDS_II(31,17, 
                                                                            (0x0001ffffU 
                                                                             & (IData)(
                                                                                (vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1 
                                                                                >> 0x00000016U))))))));
        __VdlyDim0__rnd01__DOT__mem__v0 = (0x0000000fU 
                                           & (IData)(vlSelfRef.i4));
        vlSelfRef.rnd01__DOT__r2 = __Vdly__rnd01__DOT__r2;
        rnd01__DOT__mem[__VdlyDim0__rnd01__DOT__mem__v0] 
            = __VdlyVal__rnd01__DOT__mem__v0;
        vlSelfRef.rnd01__DOT__r3[0U] = __Vdly__rnd01__DOT__r3[0U];
        vlSelfRef.rnd01__DOT__r3[1U] = __Vdly__rnd01__DOT__r3[1U];
        vlSelfRef.rnd01__DOT__r3[2U] = __Vdly__rnd01__DOT__r3[2U];
        vlSelfRef.rnd01__DOT__r3[3U] = __Vdly__rnd01__DOT__r3[3U];
        vlSelfRef.rnd01__DOT__r3[4U] = __Vdly__rnd01__DOT__r3[4U];
        vlSelfRef.rnd01__DOT__r1[0U] = __Vdly__rnd01__DOT__r1[0U];
        vlSelfRef.rnd01__DOT__r1[1U] = __Vdly__rnd01__DOT__r1[1U];
        vlSelfRef.rnd01__DOT__r1[2U] = __Vdly__rnd01__DOT__r1[2U];
        vlSelfRef.rnd01__DOT__r1[3U] = __Vdly__rnd01__DOT__r1[3U];
        __Vtemp_12[0U] = 0U;
        __Vtemp_12[1U] = 0U;
        __Vtemp_12[2U] = 0U;
        __Vtemp_12[3U] = 0U;
        __Vtemp_14[0U] = vlSelfRef.i0[0U];
        __Vtemp_14[1U] = vlSelfRef.i0[1U];
        __Vtemp_14[2U] = vlSelfRef.i0[2U];
        __Vtemp_14[3U] = 0U;
        VL_MUL_W(4, __Vtemp_15, vlSelfRef.rnd01__DOT__r1, __Vtemp_14);
        __Vtemp_16[0U] = __Vtemp_15[0U];
        __Vtemp_16[1U] = __Vtemp_15[1U];
        __Vtemp_16[2U] = __Vtemp_15[2U];
        __Vtemp_16[3U] = (0x0000000fU & __Vtemp_15[3U]);
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_3 = 
            VL_LTS_IWW(100, __Vtemp_12, __Vtemp_16);
        vlSelfRef.rnd01__DOT__r0 = (~ (0x00ffffffU 
                                       & (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1)));
        vlSelfRef.o2 = (IData)((0x000000ffffffffffULL 
                                & VL_DIVS_QQQ(40, (0x000000ffffffffffULL 
                                                   & VL_EXTENDS_QI(40,32, vlSelfRef.rnd01__DOT__r0)), vlSelfRef.i4)));
        __Vtemp_19[0U] = vlSelfRef.i2;
        __Vtemp_19[1U] = 0U;
        __Vtemp_19[2U] = 0U;
        __Vtemp_19[3U] = 0U;
        __Vtemp_20[0U] = vlSelfRef.rnd01__DOT__r0;
        __Vtemp_20[1U] = 0U;
        __Vtemp_20[2U] = 0U;
        __Vtemp_20[3U] = 0U;
        VL_MODDIV_WWW(100, __Vtemp_21, __Vtemp_19, __Vtemp_20);
        __Vtemp_22[0U] = __Vtemp_21[0U];
        __Vtemp_22[1U] = __Vtemp_21[1U];
        __Vtemp_22[2U] = __Vtemp_21[2U];
        __Vtemp_22[3U] = (0x0000000fU & __Vtemp_21[3U]);
        __Vtemp_24[0U] = vlSelfRef.rnd01__DOT__r1[0U];
        __Vtemp_24[1U] = (0x00008000U ^ vlSelfRef.rnd01__DOT__r1[1U]);
        __Vtemp_24[2U] = vlSelfRef.rnd01__DOT__r1[2U];
        __Vtemp_24[3U] = vlSelfRef.rnd01__DOT__r1[3U];
        VL_MUL_W(4, __Vtemp_25, __Vtemp_22, __Vtemp_24);
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[0U] 
            = __Vtemp_25[0U];
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[1U] 
            = __Vtemp_25[1U];
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[2U] 
            = __Vtemp_25[2U];
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[3U] 
            = (0x0000000fU & __Vtemp_25[3U]);
        VL_EXTENDS_WI(128,1, __Vtemp_26, (1U & (~ (
                                                   (vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2
                                                    [
                                                    (1U 
                                                     & ((IData)(vlSelfRef.i2) 
                                                        >> 5U))] 
                                                    >> 
                                                    (0x0000001fU 
                                                     & (IData)(vlSelfRef.i2))) 
                                                   | (0U 
                                                      != 
                                                      (((QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[1U])) 
                                                        << 0x00000020U) 
                                                       | (QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[0U]))))))));
        __Vtemp_29[0U] = (vlSelfRef.rnd01__DOT__r1[0U] 
                          | vlSelfRef.i3);
        __Vtemp_29[1U] = vlSelfRef.rnd01__DOT__r1[1U];
        __Vtemp_29[2U] = vlSelfRef.rnd01__DOT__r1[2U];
        __Vtemp_29[3U] = vlSelfRef.rnd01__DOT__r1[3U];
        __Vtemp_30[0U] = (IData)(vlSelfRef.i1);
        __Vtemp_30[1U] = (IData)((vlSelfRef.i1 >> 0x00000020U));
        __Vtemp_30[2U] = 0U;
        __Vtemp_30[3U] = 0U;
        __Vtemp_31[0U] = 0x33da3873U;
        __Vtemp_31[1U] = 0xf87305faU;
        __Vtemp_31[2U] = 0x96158e5fU;
        __Vtemp_31[3U] = 0x665c8735U;
        VL_SUB_W(4, __Vtemp_32, __Vtemp_30, __Vtemp_31);
        VL_SUB_W(4, __Vtemp_33, __Vtemp_29, __Vtemp_32);
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_4 = 
            VL_GTS_IWW(128, __Vtemp_26, __Vtemp_33);
        VL_EXTENDS_WQ(150,64, __Vtemp_35, (((QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[1U])) 
                                            << 0x00000020U) 
                                           | (QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[0U]))));
        __Vtemp_36[0U] = __Vtemp_35[0U];
        __Vtemp_36[1U] = __Vtemp_35[1U];
        __Vtemp_36[2U] = __Vtemp_35[2U];
        __Vtemp_36[3U] = __Vtemp_35[3U];
        __Vtemp_36[4U] = (0x003fffffU & __Vtemp_35[4U]);
        VL_SHIFTRS_WWI(150,150,8, __Vtemp_38, vlSelfRef.rnd01__DOT__r3, 
                       (0x000000ffU & vlSelfRef.i3));
        __Vtemp_39[0U] = __Vtemp_38[0U];
        __Vtemp_39[1U] = __Vtemp_38[1U];
        __Vtemp_39[2U] = __Vtemp_38[2U];
        __Vtemp_39[3U] = __Vtemp_38[3U];
        __Vtemp_39[4U] = (0x003fffffU & __Vtemp_38[4U]);
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_5 = 
            VL_GTS_III(31, (0x7fffffffU & VL_EXTENDS_II(31,1, 
                                                        VL_LTS_IWW(150, __Vtemp_36, __Vtemp_39))), 
                       (0x7fffffffU & (VL_MULS_III(31, vlSelfRef.i3, 
                                                   (0x7fffffffU 
                                                    & VL_EXTENDS_II(31,8, (IData)(vlSelfRef.i2)))) 
                                       << (vlSelfRef.i4 
                                           > vlSelfRef.i5))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1 = 
            (0x0000ffffffffffffULL & (vlSelfRef.i1 
                                      * (QData)((IData)(
                                                        ((VL_SHIFTR_III(28,28,8, 
                                                                        (vlSelfRef.rnd01__DOT__r0 
                                                                         >> 4U), (IData)(vlSelfRef.i2)) 
                                                          << 1U) 
                                                         | ((((QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[1U])) 
                                                              << 0x00000020U) 
                                                             | (QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[0U]))) 
                                                            <= vlSelfRef.i4))))));
        VL_SHIFTR_WWI(250,250,31, __Vtemp_41, Vrnd01__ConstPool__CONST_h113dda8b_0, vlSelfRef.i3);
        __Vtemp_46[0U] = (0x0000007fU & (vlSelfRef.i3 
                                         >> 2U));
        __Vtemp_46[1U] = 0U;
        __Vtemp_46[2U] = 0U;
        __Vtemp_46[3U] = 0U;
        __Vtemp_46[4U] = 0U;
        __Vtemp_46[5U] = 0U;
        __Vtemp_46[6U] = 0U;
        __Vtemp_46[7U] = 0U;
        __Vtemp_47[0U] = vlSelfRef.__VdfgRegularize_h6e95ff9d_0_3;
        __Vtemp_47[1U] = 0U;
        __Vtemp_47[2U] = 0U;
        __Vtemp_47[3U] = 0U;
        __Vtemp_47[4U] = 0U;
        __Vtemp_47[5U] = 0U;
        __Vtemp_47[6U] = 0U;
        __Vtemp_47[7U] = 0U;
        VL_MUL_W(8, __Vtemp_48, __Vtemp_46, __Vtemp_47);
        vlSelfRef.rnd01__DOT__w8[0U] = ((~ (__Vtemp_41[0U] 
                                            ^ (0x0000000fU 
                                               & ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_5) 
                                                  >> 
                                                  (0x0000000fU 
                                                   & (IData)(vlSelfRef.i4)))))) 
                                        ^ __Vtemp_48[0U]);
        vlSelfRef.rnd01__DOT__w8[1U] = ((~ __Vtemp_41[1U]) 
                                        ^ __Vtemp_48[1U]);
        vlSelfRef.rnd01__DOT__w8[2U] = ((~ __Vtemp_41[2U]) 
                                        ^ __Vtemp_48[2U]);
        vlSelfRef.rnd01__DOT__w8[3U] = ((~ __Vtemp_41[3U]) 
                                        ^ __Vtemp_48[3U]);
        vlSelfRef.rnd01__DOT__w8[4U] = (0x003fffffU 
                                        & ((~ __Vtemp_41[4U]) 
                                           ^ __Vtemp_48[4U]));
        __Vtemp_52[0U] = 0U;
        __Vtemp_52[1U] = 0U;
        __Vtemp_52[2U] = 0U;
        __Vtemp_52[3U] = 0U;
        __Vtemp_52[4U] = 0U;
        vlSelfRef.o1 = (0x0000ffffU & ((0U != (((QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[1U])) 
                                                << 0x00000020U) 
                                               | (QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[0U]))))
                                        ? VL_MODDIV_III(32, vlSelfRef.i3, 
                                                        (0x00ffffffU 
                                                         & (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1)))
                                        : (~ ((VL_COUNTONES_I(
                                                              (1U 
                                                               & (vlSelfRef.i0
                                                                  [
                                                                  (1U 
                                                                   & (vlSelfRef.rnd01__DOT__r3[0U] 
                                                                      >> 5U))] 
                                                                  >> 
                                                                  (0x0000001fU 
                                                                   & vlSelfRef.rnd01__DOT__r3[0U])))) 
                                               + VL_ONEHOT_I(
                                                             (0x000007ffU 
                                                              & (vlSelfRef.i3 
                                                                 >> 0x00000011U)))) 
                                              ^ VL_GTE_W(5, __Vtemp_52, vlSelfRef.rnd01__DOT__r3)))));
        VL_EXTENDS_WI(150,24, __Vtemp_54, (0x00ffffffU 
                                           & (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1)));
        __Vtemp_55[0U] = __Vtemp_54[0U];
        __Vtemp_55[1U] = __Vtemp_54[1U];
        __Vtemp_55[2U] = __Vtemp_54[2U];
        __Vtemp_55[3U] = __Vtemp_54[3U];
        __Vtemp_55[4U] = (0x003fffffU & __Vtemp_54[4U]);
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0 = 
            VL_GTES_IWW(150, vlSelfRef.rnd01__DOT__r3, __Vtemp_55);
        if ((1U & VL_SHIFTL_III(1,1,8, (~ (0U != ((
                                                   ((vlSelfRef.rnd01__DOT__w8[0U] 
                                                     | vlSelfRef.rnd01__DOT__w8[1U]) 
                                                    | vlSelfRef.rnd01__DOT__w8[2U]) 
                                                   | vlSelfRef.rnd01__DOT__w8[3U]) 
                                                  | vlSelfRef.rnd01__DOT__w8[4U]))), 
                                (0x000000ffU & vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[0U])))) {
            __Vtemp_58[0U] = (0x0000007fU & (IData)(vlSelfRef.i2));
            __Vtemp_58[1U] = 0U;
            __Vtemp_58[2U] = 0U;
            __Vtemp_59[0U] = ((0U != (0x00ffffffU & (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1))) 
                              & ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_3) 
                                 >> (1U & (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1))));
            __Vtemp_59[1U] = 0U;
            __Vtemp_59[2U] = 0U;
            VL_ADD_W(3, __Vtemp_57, __Vtemp_58, __Vtemp_59);
        } else {
            __Vtemp_60[0U] = 0x00000094U;
            __Vtemp_60[1U] = 0U;
            __Vtemp_60[2U] = 0U;
            __Vtemp_61[0U] = VL_ONEHOT_Q((0x000000ffffffffffULL 
                                          & VL_DIVS_QQQ(40, 
                                                        (0x000000ffffffffffULL 
                                                         & vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1), 
                                                        (0x000000ffffffffffULL 
                                                         & VL_EXTENDS_QI(40,32, vlSelfRef.rnd01__DOT__r0)))));
            __Vtemp_61[1U] = 0U;
            __Vtemp_61[2U] = 0U;
            VL_ADD_W(3, __Vtemp_57, __Vtemp_60, __Vtemp_61);
        }
        vlSelfRef.rnd01__DOT__w11[0U] = __Vtemp_57[0U];
        vlSelfRef.rnd01__DOT__w11[1U] = __Vtemp_57[1U];
        vlSelfRef.rnd01__DOT__w11[2U] = (1U & __Vtemp_57[2U]);
        vlSelfRef.o0 = (1U & ((~ (vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[0U] 
                                  ^ (IData)(VL_SHIFTL_QQI(64,64,8, vlSelfRef.i5, 
                                                          (0x000000ffU 
                                                           & (IData)(vlSelfRef.i1)))))) 
                              ^ (IData)((1ULL >> (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)))));
        vlSelfRef.o3 = (0x000000ffffffffffULL & (vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1 
                                                 | ((QData)((IData)(
                                                                    VL_ONEHOT_W(5, vlSelfRef.rnd01__DOT__w8))) 
                                                    + (QData)((IData)(
                                                                      VL_COUNTONES_I(
                                                                                ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0) 
                                                                                <= 
                                                                                (1U 
                                                                                & (IData)(
                                                                                (vlSelfRef.i5 
                                                                                >> 
                                                                                (0x0000001fU 
                                                                                & vlSelfRef.rnd01__DOT__r3[0U])))))))))));
        if (vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0) {
            __Vtemp_63[0U] = vlSelfRef.__VdfgRegularize_h6e95ff9d_0_5;
            __Vtemp_63[1U] = 0U;
            __Vtemp_63[2U] = 0U;
            VL_SHIFTL_WWI(96,96,8, __Vtemp_62, __Vtemp_63, 
                          (0x000000ffU & vlSelfRef.rnd01__DOT__w11[0U]));
        } else {
            __Vtemp_62[0U] = (IData)(vlSelfRef.i1);
            __Vtemp_62[1U] = (IData)((vlSelfRef.i1 
                                      >> 0x00000020U));
            __Vtemp_62[2U] = 0U;
        }
        vlSelfRef.o4[0U] = (__Vtemp_62[0U] ^ (IData)(
                                                     (((QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[1U])) 
                                                       << 0x00000020U) 
                                                      | (QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[0U])))));
        vlSelfRef.o4[1U] = (__Vtemp_62[1U] ^ (IData)(
                                                     ((((QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[1U])) 
                                                        << 0x00000020U) 
                                                       | (QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2[0U]))) 
                                                      >> 0x00000020U)));
        vlSelfRef.o4[2U] = __Vtemp_62[2U];
    }

    VLCU_D static void Vrnd01___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vrnd01___024root___eval_phase__act(Vrnd01___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        {
            // Inlined CFunc: _eval_triggers_vec__act
            vlSelfRef.__VactTriggered[0U] = (QData)((IData)(
                                                            ((IData)(vlSelfRef.clk) 
                                                             & (~ (IData)(vlSelfRef.__Vtrigprevexpr___TOP__clk__1)))));
            vlSelfRef.__Vtrigprevexpr___TOP__clk__1 
                = vlSelfRef.clk;
        }
#ifdef VL_DEBUG
        if (VL_UNLIKELY(vlSymsp->_vm_contextp__->debug())) {
            Vrnd01___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vrnd01___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vrnd01___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vrnd01___024root___eval_phase__nba(Vrnd01___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vrnd01___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    Vrnd01___024root___nba_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
            Vrnd01___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vrnd01___024root___eval(Vrnd01___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        IData/*31:0*/ __VicoIterCount;
        IData/*31:0*/ __VnbaIterCount;
        // Body
        __VicoIterCount = 0U;
        do {
            if (VL_UNLIKELY(((0x00002710U < __VicoIterCount)))) {
#ifdef VL_DEBUG
                Vrnd01___024root___dump_triggers__ico(vlSelfRef.__VicoTriggered, "ico"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/rnd01.v", 2, "", "DIDNOTCONVERGE: Input combinational region did not converge after '--converge-limit' of 10000 tries");
            }
            __VicoIterCount = ((IData)(1U) + __VicoIterCount);
            vlSelfRef.__VicoPhaseResult = Vrnd01___024root___eval_phase__ico(vlSelf, vlCtx);
        } while (vlSelfRef.__VicoPhaseResult);
        __VnbaIterCount = 0U;
        do {
            if (VL_UNLIKELY(((0x00002710U < __VnbaIterCount)))) {
#ifdef VL_DEBUG
                Vrnd01___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/rnd01.v", 2, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vrnd01___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("../verilator_b200/csrc/designs/torture/rnd01.v", 2, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vrnd01___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vrnd01___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vrnd01___024root___ctor_var_reset(Vrnd01___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        VL_SCOPED_RAND_RESET_W(96, vlSelf->i0, __VscopeHash, 8471475882151657015ull);
        vlSelf->i1 = VL_SCOPED_RAND_RESET_Q(48, __VscopeHash, 2165040029148572678ull);
        vlSelf->i2 = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 12628911202397369737ull);
        vlSelf->i3 = VL_SCOPED_RAND_RESET_I(31, __VscopeHash, 6886176223853197501ull);
        vlSelf->i4 = VL_SCOPED_RAND_RESET_Q(40, __VscopeHash, 5103504462683158156ull);
        vlSelf->i5 = VL_SCOPED_RAND_RESET_Q(33, __VscopeHash, 10954576514065331241ull);
        vlSelf->o0 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 4709259436872190324ull);
        vlSelf->o1 = VL_SCOPED_RAND_RESET_I(16, __VscopeHash, 16871796271195187576ull);
        vlSelf->o2 = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 6221000330064061286ull);
        vlSelf->o3 = VL_SCOPED_RAND_RESET_Q(40, __VscopeHash, 17472377812510545679ull);
        VL_SCOPED_RAND_RESET_W(96, vlSelf->o4, __VscopeHash, 12832999154864337794ull);
        vlSelf->rnd01__DOT__r0 = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 14030323845612525961ull);
        VL_SCOPED_RAND_RESET_W(100, vlSelf->rnd01__DOT__r1, __VscopeHash, 5128380883843462818ull);
        vlSelf->rnd01__DOT__r2 = VL_SCOPED_RAND_RESET_Q(64, __VscopeHash, 1098015726095441435ull);
        VL_SCOPED_RAND_RESET_W(150, vlSelf->rnd01__DOT__r3, __VscopeHash, 226595549040757621ull);
        VL_SCOPED_RAND_RESET_W(150, vlSelf->rnd01__DOT__w8, __VscopeHash, 2031618261185045626ull);
        VL_SCOPED_RAND_RESET_W(65, vlSelf->rnd01__DOT__w11, __VscopeHash, 6685290939348945952ull);
        vlSelf->__VdfgRegularize_h6e95ff9d_0_0 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_1 = 0;
        VL_ZERO_RESET_W(100, vlSelf->__VdfgRegularize_h6e95ff9d_0_2);
        vlSelf->__VdfgRegularize_h6e95ff9d_0_3 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_4 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_5 = 0;
        for (int __Vi0 = 0; __Vi0 < 1; ++__Vi0) {
            vlSelf->__VstlTriggered[__Vi0] = 0;
        }
        for (int __Vi0 = 0; __Vi0 < 2; ++__Vi0) {
            vlSelf->__VicoTriggered[__Vi0] = 0;
        }
        vlSelf->__Vtrigprevexpr___TOP__clk__0 = 0;
        VL_ZERO_RESET_W(96, vlSelf->__Vtrigprevexpr___TOP__i0__0);
        vlSelf->__Vtrigprevexpr___TOP__i1__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__i2__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__i3__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__i4__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__i5__0 = 0;
        vlSelf->__VicoDidInit = 0;
        for (int __Vi0 = 0; __Vi0 < 1; ++__Vi0) {
            vlSelf->__VactTriggered[__Vi0] = 0;
        }
        vlSelf->__Vtrigprevexpr___TOP__clk__1 = 0;
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
        Vrnd01___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vrnd01___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vrnd01___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vrnd01___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vrnd01___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vrnd01___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vrnd01__cuda::info() {
    typedef Vrnd01__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"i0", "i0", VLCU_MEMBER_IN, 4, 0, 3, 96, 95, 0, T::OFF_i0},
        {"i1", "i1", VLCU_MEMBER_IN, 8, 0, 1, 48, 47, 0, T::OFF_i1},
        {"i2", "i2", VLCU_MEMBER_IN, 1, 0, 1, 8, 7, 0, T::OFF_i2},
        {"i3", "i3", VLCU_MEMBER_IN, 4, 0, 1, 31, 30, 0, T::OFF_i3},
        {"i4", "i4", VLCU_MEMBER_IN, 8, 0, 1, 40, 39, 0, T::OFF_i4},
        {"i5", "i5", VLCU_MEMBER_IN, 8, 0, 1, 33, 32, 0, T::OFF_i5},
        {"o0", "o0", VLCU_MEMBER_OUT, 1, 0, 1, 1, 0, 0, T::OFF_o0},
        {"o1", "o1", VLCU_MEMBER_OUT, 2, 0, 1, 16, 15, 0, T::OFF_o1},
        {"o2", "o2", VLCU_MEMBER_OUT, 4, 0, 1, 32, 31, 0, T::OFF_o2},
        {"o3", "o3", VLCU_MEMBER_OUT, 8, 0, 1, 40, 39, 0, T::OFF_o3},
        {"o4", "o4", VLCU_MEMBER_OUT, 4, 0, 3, 96, 95, 0, T::OFF_o4},
        {"__VdfgRegularize_h6e95ff9d_0_0", "__VdfgRegularize_h6e95ff9d_0_0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_0},
        {"__VdfgRegularize_h6e95ff9d_0_3", "__VdfgRegularize_h6e95ff9d_0_3", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_3},
        {"__VdfgRegularize_h6e95ff9d_0_4", "__VdfgRegularize_h6e95ff9d_0_4", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_4},
        {"__VdfgRegularize_h6e95ff9d_0_5", "__VdfgRegularize_h6e95ff9d_0_5", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_5},
        {"__VstlFirstIteration", "__VstlFirstIteration", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlFirstIteration},
        {"__VstlPhaseResult", "__VstlPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__Vtrigprevexpr___TOP__i2__0", "__Vtrigprevexpr___TOP__i2__0", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF___Vtrigprevexpr___TOP__i2__0},
        {"__VicoDidInit", "__VicoDidInit", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VicoDidInit},
        {"__VicoPhaseResult", "__VicoPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VicoPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__1", "__Vtrigprevexpr___TOP__clk__1", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__1},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"rnd01__DOT__r0", "rnd01.r0", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rnd01__DOT__r0},
        {"rnd01__DOT__r1", "rnd01.r1", VLCU_MEMBER_STATE, 4, 0, 4, 100, 99, 0, T::OFF_rnd01__DOT__r1},
        {"rnd01__DOT__r3", "rnd01.r3", VLCU_MEMBER_STATE, 4, 0, 5, 150, 149, 0, T::OFF_rnd01__DOT__r3},
        {"rnd01__DOT__w8", "rnd01.w8", VLCU_MEMBER_STATE, 4, 0, 5, 150, 149, 0, T::OFF_rnd01__DOT__w8},
        {"rnd01__DOT__w11", "rnd01.w11", VLCU_MEMBER_STATE, 4, 0, 3, 65, 64, 0, T::OFF_rnd01__DOT__w11},
        {"__VdfgRegularize_h6e95ff9d_0_2", "__VdfgRegularize_h6e95ff9d_0_2", VLCU_MEMBER_STATE, 4, 0, 4, 100, 99, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_2},
        {"__Vtrigprevexpr___TOP__i0__0", "__Vtrigprevexpr___TOP__i0__0", VLCU_MEMBER_STATE, 4, 0, 3, 96, 95, 0, T::OFF___Vtrigprevexpr___TOP__i0__0},
        {"__Vtrigprevexpr___TOP__i3__0", "__Vtrigprevexpr___TOP__i3__0", VLCU_MEMBER_STATE, 4, 0, 1, 31, 30, 0, T::OFF___Vtrigprevexpr___TOP__i3__0},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"rnd01__DOT__r2", "rnd01.r2", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF_rnd01__DOT__r2},
        {"__VdfgRegularize_h6e95ff9d_0_1", "__VdfgRegularize_h6e95ff9d_0_1", VLCU_MEMBER_STATE, 8, 0, 1, 48, 47, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_1},
        {"__Vtrigprevexpr___TOP__i1__0", "__Vtrigprevexpr___TOP__i1__0", VLCU_MEMBER_STATE, 8, 0, 1, 48, 47, 0, T::OFF___Vtrigprevexpr___TOP__i1__0},
        {"__Vtrigprevexpr___TOP__i4__0", "__Vtrigprevexpr___TOP__i4__0", VLCU_MEMBER_STATE, 8, 0, 1, 40, 39, 0, T::OFF___Vtrigprevexpr___TOP__i4__0},
        {"__Vtrigprevexpr___TOP__i5__0", "__Vtrigprevexpr___TOP__i5__0", VLCU_MEMBER_STATE, 8, 0, 1, 33, 32, 0, T::OFF___Vtrigprevexpr___TOP__i5__0},
        {"__VstlTriggered", "__VstlTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VstlTriggered},
        {"__VicoTriggered", "__VicoTriggered", VLCU_MEMBER_STATE, 8, 0, 2, 64, 63, 0, T::OFF___VicoTriggered},
        {"__VactTriggered", "__VactTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VactTriggered},
        {"__VnbaTriggered", "__VnbaTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VnbaTriggered},
    };
    static const char* const messages[] = {nullptr};
    static const char* const plusargs[] = {nullptr};
    static const VlcuTraceDecl traceDecls[] = {
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0}};  // (terminator, not counted)
    static const uint32_t traceOrder[] = {0};  // codes in dump order (+ terminator)
    static const uint32_t saveOrder[] = {0, 3, 7, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 8, 1, 4, 9, 11, 25, 26, 27, 28, 29, 30, 31, 32, 33, 2, 5, 6, 10, 34, 35, 36, 37, 38, 39, 40, 41, 42, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "rnd01", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 0, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 43, 0, 0x72d6802b2c7e727aULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vrnd01__cuda VLCU_TESTBENCHES);
