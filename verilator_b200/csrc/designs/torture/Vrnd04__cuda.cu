// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vrnd04"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vrnd04___024root final {
    CData/*0:0*/ clk;
    VlWide<5>/*149:0*/ i0;
    IData/*31:0*/ i1;
    QData/*47:0*/ i2;
    SData/*15:0*/ i3;
    VlWide<5>/*129:0*/ i4;
    VlWide<4>/*127:0*/ i5;
    CData/*7:0*/ o0;
    SData/*15:0*/ o1;
    IData/*31:0*/ o2;
    QData/*39:0*/ o3;
    VlWide<3>/*95:0*/ o4;
    CData/*0:0*/ rnd04__DOT__r2;
    CData/*7:0*/ rnd04__DOT__r3;
    CData/*1:0*/ rnd04__DOT__w2;
    CData/*7:0*/ rnd04__DOT__w4;
    CData/*1:0*/ rnd04__DOT__w6;
    CData/*0:0*/ __VdfgRegularize_h6e95ff9d_0_0;
    CData/*0:0*/ __VstlFirstIteration;
    CData/*0:0*/ __VstlPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*0:0*/ __VicoDidInit;
    CData/*0:0*/ __VicoPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__1;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    SData/*15:0*/ rnd04__DOT__w7;
    SData/*15:0*/ __Vtrigprevexpr___TOP__i3__0;
    VlWide<3>/*95:0*/ rnd04__DOT__r0;
    IData/*31:0*/ rnd04__DOT__r1;
    VlWide<3>/*95:0*/ rnd04__DOT__w10;
    VlWide<5>/*149:0*/ __Vtrigprevexpr___TOP__i0__0;
    IData/*31:0*/ __Vtrigprevexpr___TOP__i1__0;
    VlWide<5>/*129:0*/ __Vtrigprevexpr___TOP__i4__0;
    VlWide<4>/*127:0*/ __Vtrigprevexpr___TOP__i5__0;
    IData/*31:0*/ __VactIterCount;
    QData/*63:0*/ rnd04__DOT__w1;
    QData/*39:0*/ rnd04__DOT__w8;
    QData/*47:0*/ __Vtrigprevexpr___TOP__i2__0;
    VlcuSharedArray<IData/*23:0*/, 16> rnd04__DOT__mem;  // run-time indexed: shared memory [index][lane]
    VlUnpacked<QData/*63:0*/, 1> __VstlTriggered;
    VlUnpacked<QData/*63:0*/, 2> __VicoTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vrnd04__cuda final {
    typedef Vrnd04___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 320,
        OFF_i0 = 80,
        OFF_i1 = 100,
        OFF_i2 = 0,
        OFF_i3 = 312,
        OFF_i4 = 104,
        OFF_i5 = 124,
        OFF_o0 = 321,
        OFF_o1 = 314,
        OFF_o2 = 140,
        OFF_o3 = 8,
        OFF_o4 = 144,
        OFF_rnd04__DOT__r2 = 322,
        OFF_rnd04__DOT__r3 = 323,
        OFF_rnd04__DOT__w2 = 324,
        OFF_rnd04__DOT__w4 = 325,
        OFF_rnd04__DOT__w6 = 326,
        OFF___VdfgRegularize_h6e95ff9d_0_0 = 327,
        OFF___VstlFirstIteration = 328,
        OFF___VstlPhaseResult = 329,
        OFF___Vtrigprevexpr___TOP__clk__0 = 330,
        OFF___VicoDidInit = 331,
        OFF___VicoPhaseResult = 332,
        OFF___Vtrigprevexpr___TOP__clk__1 = 333,
        OFF___VactPhaseResult = 334,
        OFF___VnbaPhaseResult = 335,
        OFF_rnd04__DOT__w7 = 316,
        OFF___Vtrigprevexpr___TOP__i3__0 = 318,
        OFF_rnd04__DOT__r0 = 156,
        OFF_rnd04__DOT__r1 = 168,
        OFF_rnd04__DOT__w10 = 172,
        OFF___Vtrigprevexpr___TOP__i0__0 = 184,
        OFF___Vtrigprevexpr___TOP__i1__0 = 204,
        OFF___Vtrigprevexpr___TOP__i4__0 = 208,
        OFF___Vtrigprevexpr___TOP__i5__0 = 228,
        OFF___VactIterCount = 244,
        OFF_rnd04__DOT__w1 = 16,
        OFF_rnd04__DOT__w8 = 24,
        OFF___Vtrigprevexpr___TOP__i2__0 = 32,
        OFF_rnd04__DOT__mem = 248,
        OFF___VstlTriggered = 40,
        OFF___VicoTriggered = 48,
        OFF___VactTriggered = 64,
        OFF___VnbaTriggered = 72,
        kStateBytes = 336,  // S_persist
        kTouchedReadBytes = 255,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 195,  // S_state_wr
    };
    struct Locals final {
    };
    struct Shared final {
        IData/*23:0*/ rnd04__DOT__mem[16][VLCU_BLOCK];
    };
    VLCU_D static void bind(Root& s, Locals& l, Shared& sh, const VlcuState& st, uint64_t i) {
        (void)l; (void)sh; (void)st; (void)i;
        s.vlNamep = "TOP";
        s.rnd04__DOT__mem.bind(sh.rnd04__DOT__mem, threadIdx.x);
    }
    VLCU_D static void load(Root& s, const VlcuState& st, uint64_t i) {
        s.clk = st.ld<VlcuMem8>(OFF_clk, i);
        st.ldW(OFF_i0, i, s.i0);
        s.i1 = st.ld<IData>(OFF_i1, i);
        s.i2 = st.ld<QData>(OFF_i2, i);
        s.i3 = st.ld<VlcuMem16>(OFF_i3, i);
        st.ldW(OFF_i4, i, s.i4);
        st.ldW(OFF_i5, i, s.i5);
        s.o0 = st.ld<VlcuMem8>(OFF_o0, i);
        s.o1 = st.ld<VlcuMem16>(OFF_o1, i);
        s.o2 = st.ld<IData>(OFF_o2, i);
        s.o3 = st.ld<QData>(OFF_o3, i);
        st.ldW(OFF_o4, i, s.o4);
        s.rnd04__DOT__r2 = st.ld<VlcuMem8>(OFF_rnd04__DOT__r2, i);
        s.rnd04__DOT__r3 = st.ld<VlcuMem8>(OFF_rnd04__DOT__r3, i);
        s.rnd04__DOT__w2 = st.ld<VlcuMem8>(OFF_rnd04__DOT__w2, i);
        s.rnd04__DOT__w4 = st.ld<VlcuMem8>(OFF_rnd04__DOT__w4, i);
        s.rnd04__DOT__w6 = st.ld<VlcuMem8>(OFF_rnd04__DOT__w6, i);
        s.__VdfgRegularize_h6e95ff9d_0_0 = st.ld<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_0, i);
        s.__VstlFirstIteration = st.ld<VlcuMem8>(OFF___VstlFirstIteration, i);
        s.__VstlPhaseResult = st.ld<VlcuMem8>(OFF___VstlPhaseResult, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__VicoDidInit = st.ld<VlcuMem8>(OFF___VicoDidInit, i);
        s.__VicoPhaseResult = st.ld<VlcuMem8>(OFF___VicoPhaseResult, i);
        s.__Vtrigprevexpr___TOP__clk__1 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__1, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        s.rnd04__DOT__w7 = st.ld<VlcuMem16>(OFF_rnd04__DOT__w7, i);
        s.__Vtrigprevexpr___TOP__i3__0 = st.ld<VlcuMem16>(OFF___Vtrigprevexpr___TOP__i3__0, i);
        st.ldW(OFF_rnd04__DOT__r0, i, s.rnd04__DOT__r0);
        s.rnd04__DOT__r1 = st.ld<IData>(OFF_rnd04__DOT__r1, i);
        st.ldW(OFF_rnd04__DOT__w10, i, s.rnd04__DOT__w10);
        st.ldW(OFF___Vtrigprevexpr___TOP__i0__0, i, s.__Vtrigprevexpr___TOP__i0__0);
        s.__Vtrigprevexpr___TOP__i1__0 = st.ld<IData>(OFF___Vtrigprevexpr___TOP__i1__0, i);
        st.ldW(OFF___Vtrigprevexpr___TOP__i4__0, i, s.__Vtrigprevexpr___TOP__i4__0);
        st.ldW(OFF___Vtrigprevexpr___TOP__i5__0, i, s.__Vtrigprevexpr___TOP__i5__0);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        s.rnd04__DOT__w1 = st.ld<QData>(OFF_rnd04__DOT__w1, i);
        s.rnd04__DOT__w8 = st.ld<QData>(OFF_rnd04__DOT__w8, i);
        s.__Vtrigprevexpr___TOP__i2__0 = st.ld<QData>(OFF___Vtrigprevexpr___TOP__i2__0, i);
        for (uint32_t e = 0; e < 16; ++e) s.rnd04__DOT__mem[e] = st.ldE<IData>(OFF_rnd04__DOT__mem, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VstlTriggered[e] = st.ldE<QData>(OFF___VstlTriggered, e, i);
        for (uint32_t e = 0; e < 2; ++e) s.__VicoTriggered[e] = st.ldE<QData>(OFF___VicoTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.stW(OFF_i0, i, s.i0);
        st.st<IData>(OFF_i1, i, s.i1);
        st.st<QData>(OFF_i2, i, s.i2);
        st.st<VlcuMem16>(OFF_i3, i, s.i3);
        st.stW(OFF_i4, i, s.i4);
        st.stW(OFF_i5, i, s.i5);
        st.st<VlcuMem8>(OFF_o0, i, s.o0);
        st.st<VlcuMem16>(OFF_o1, i, s.o1);
        st.st<IData>(OFF_o2, i, s.o2);
        st.st<QData>(OFF_o3, i, s.o3);
        st.stW(OFF_o4, i, s.o4);
        st.st<VlcuMem8>(OFF_rnd04__DOT__r2, i, s.rnd04__DOT__r2);
        st.st<VlcuMem8>(OFF_rnd04__DOT__r3, i, s.rnd04__DOT__r3);
        st.st<VlcuMem8>(OFF_rnd04__DOT__w2, i, s.rnd04__DOT__w2);
        st.st<VlcuMem8>(OFF_rnd04__DOT__w4, i, s.rnd04__DOT__w4);
        st.st<VlcuMem8>(OFF_rnd04__DOT__w6, i, s.rnd04__DOT__w6);
        st.st<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_0, i, s.__VdfgRegularize_h6e95ff9d_0_0);
        st.st<VlcuMem8>(OFF___VstlFirstIteration, i, s.__VstlFirstIteration);
        st.st<VlcuMem8>(OFF___VstlPhaseResult, i, s.__VstlPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___VicoDidInit, i, s.__VicoDidInit);
        st.st<VlcuMem8>(OFF___VicoPhaseResult, i, s.__VicoPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__1, i, s.__Vtrigprevexpr___TOP__clk__1);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.st<VlcuMem16>(OFF_rnd04__DOT__w7, i, s.rnd04__DOT__w7);
        st.st<VlcuMem16>(OFF___Vtrigprevexpr___TOP__i3__0, i, s.__Vtrigprevexpr___TOP__i3__0);
        st.stW(OFF_rnd04__DOT__r0, i, s.rnd04__DOT__r0);
        st.st<IData>(OFF_rnd04__DOT__r1, i, s.rnd04__DOT__r1);
        st.stW(OFF_rnd04__DOT__w10, i, s.rnd04__DOT__w10);
        st.stW(OFF___Vtrigprevexpr___TOP__i0__0, i, s.__Vtrigprevexpr___TOP__i0__0);
        st.st<IData>(OFF___Vtrigprevexpr___TOP__i1__0, i, s.__Vtrigprevexpr___TOP__i1__0);
        st.stW(OFF___Vtrigprevexpr___TOP__i4__0, i, s.__Vtrigprevexpr___TOP__i4__0);
        st.stW(OFF___Vtrigprevexpr___TOP__i5__0, i, s.__Vtrigprevexpr___TOP__i5__0);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
        st.st<QData>(OFF_rnd04__DOT__w1, i, s.rnd04__DOT__w1);
        st.st<QData>(OFF_rnd04__DOT__w8, i, s.rnd04__DOT__w8);
        st.st<QData>(OFF___Vtrigprevexpr___TOP__i2__0, i, s.__Vtrigprevexpr___TOP__i2__0);
        for (uint32_t e = 0; e < 16; ++e) st.stE<IData>(OFF_rnd04__DOT__mem, e, i, s.rnd04__DOT__mem[e]);
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
        s.i0[3] = rowp[4 * stride + i] & 0xffffffffU;
        s.i0[4] = rowp[5 * stride + i] & 0x3fffffU;
        s.i1 = static_cast<IData>(rowp[6 * stride + i] & 0xffffffffU);
        s.i2 = (static_cast<QData>(rowp[8 * stride + i] & 0xffffU) << 32) | rowp[7 * stride + i];
        s.i3 = static_cast<VlcuMem16>(rowp[9 * stride + i] & 0xffffU);
        s.i4[0] = rowp[10 * stride + i] & 0xffffffffU;
        s.i4[1] = rowp[11 * stride + i] & 0xffffffffU;
        s.i4[2] = rowp[12 * stride + i] & 0xffffffffU;
        s.i4[3] = rowp[13 * stride + i] & 0xffffffffU;
        s.i4[4] = rowp[14 * stride + i] & 0x3U;
        s.i5[0] = rowp[15 * stride + i] & 0xffffffffU;
        s.i5[1] = rowp[16 * stride + i] & 0xffffffffU;
        s.i5[2] = rowp[17 * stride + i] & 0xffffffffU;
        s.i5[3] = rowp[18 * stride + i] & 0xffffffffU;
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
    enum : uint32_t { kInWords = 19, kOutWords = 8, kInBytes = 71, kOutBytes = 27, kTraceWords = 0 };

    VLCU_D static void Vrnd04___024root___eval_static(Vrnd04___024root* vlSelf, VlcuContext& vlCtx) {
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
        vlSelfRef.__Vtrigprevexpr___TOP__i0__0[3U] 
            = vlSelfRef.i0[3U];
        vlSelfRef.__Vtrigprevexpr___TOP__i0__0[4U] 
            = vlSelfRef.i0[4U];
        vlSelfRef.__Vtrigprevexpr___TOP__i1__0 = vlSelfRef.i1;
        vlSelfRef.__Vtrigprevexpr___TOP__i2__0 = vlSelfRef.i2;
        vlSelfRef.__Vtrigprevexpr___TOP__i3__0 = vlSelfRef.i3;
        vlSelfRef.__Vtrigprevexpr___TOP__i4__0[0U] 
            = vlSelfRef.i4[0U];
        vlSelfRef.__Vtrigprevexpr___TOP__i4__0[1U] 
            = vlSelfRef.i4[1U];
        vlSelfRef.__Vtrigprevexpr___TOP__i4__0[2U] 
            = vlSelfRef.i4[2U];
        vlSelfRef.__Vtrigprevexpr___TOP__i4__0[3U] 
            = vlSelfRef.i4[3U];
        vlSelfRef.__Vtrigprevexpr___TOP__i4__0[4U] 
            = vlSelfRef.i4[4U];
        vlSelfRef.__Vtrigprevexpr___TOP__i5__0[0U] 
            = vlSelfRef.i5[0U];
        vlSelfRef.__Vtrigprevexpr___TOP__i5__0[1U] 
            = vlSelfRef.i5[1U];
        vlSelfRef.__Vtrigprevexpr___TOP__i5__0[2U] 
            = vlSelfRef.i5[2U];
        vlSelfRef.__Vtrigprevexpr___TOP__i5__0[3U] 
            = vlSelfRef.i5[3U];
        vlSelfRef.__Vtrigprevexpr___TOP__clk__1 = vlSelfRef.clk;
    }

    VLCU_D static void Vrnd04___024root___eval_initial(Vrnd04___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vrnd04___024root___eval_final(Vrnd04___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vrnd04___024root___eval_settle(Vrnd04___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vrnd04___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/rnd04.v", 2, "", "DIDNOTCONVERGE: Settle region did not converge after '--converge-limit' of 10000 tries");
            }
            __VstlIterCount = ((IData)(1U) + __VstlIterCount);
            vlSelfRef.__VstlPhaseResult = Vrnd04___024root___eval_phase__stl(vlSelf, vlCtx);
            vlSelfRef.__VstlFirstIteration = 0U;
        } while (vlSelfRef.__VstlPhaseResult);
    }

    VLCU_D static bool Vrnd04___024root___trigger_anySet__stl(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vrnd04___024root___stl_sequent__TOP__0(Vrnd04___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        VlWide<5>/*159:0*/ __Vtemp_7;
        VlWide<3>/*95:0*/ __Vtemp_9;
        VlWide<3>/*95:0*/ __Vtemp_10;
        VlWide<3>/*95:0*/ __Vtemp_11;
        // Body
        vlSelfRef.o1 = (0x0000ffffU & ((IData)(0x0f28U) 
                                       - (vlSelfRef.i4[0U] 
                                          - vlSelfRef.rnd04__DOT__r0[0U])));
        vlSelfRef.o4[0U] = ((0xfffffffeU & vlSelfRef.o4[0U]) 
                            | ((0x00000020U & ((~ (0U 
                                                   != 
                                                   (0x0000007fU 
                                                    & (vlSelfRef.i5[0U] 
                                                       >> 2U)))) 
                                               << 5U)) 
                               >= (0x00ffffffU & (~ 
                                                  (1U 
                                                   ^ 
                                                   VL_DIV_III(24, (IData)(vlSelfRef.rnd04__DOT__r2), vlSelfRef.rnd04__DOT__mem
                                                              [vlSelfRef.rnd04__DOT__r2]))))));
        vlSelfRef.o4[0U] = (1U & vlSelfRef.o4[0U]);
        vlSelfRef.o4[1U] = 0U;
        vlSelfRef.o4[2U] = 0U;
        __Vtemp_7[0U] = (vlSelfRef.i4[0U] | ((vlSelfRef.rnd04__DOT__r0[0U] 
                                              ^ (IData)(vlSelfRef.rnd04__DOT__r3)) 
                                             | (1U 
                                                & (IData)(
                                                          (vlSelfRef.i2 
                                                           >> 
                                                           (0x0000001fU 
                                                            & vlSelfRef.i5[0U]))))));
        __Vtemp_7[1U] = (vlSelfRef.i4[1U] | vlSelfRef.rnd04__DOT__r0[1U]);
        __Vtemp_7[2U] = (vlSelfRef.i4[2U] | vlSelfRef.rnd04__DOT__r0[2U]);
        __Vtemp_7[3U] = vlSelfRef.i4[3U];
        __Vtemp_7[4U] = vlSelfRef.i4[4U];
        vlSelfRef.rnd04__DOT__w7 = (0x0000ffffU & (
                                                   VL_COUNTONES_I(
                                                                  ((IData)(vlSelfRef.rnd04__DOT__r2) 
                                                                   >= 
                                                                   VL_COUNTONES_I(
                                                                                (1U 
                                                                                & (vlSelfRef.i1 
                                                                                >> 
                                                                                (0x0000001fU 
                                                                                & vlSelfRef.i5[0U])))))) 
                                                   + 
                                                   VL_ONEHOT_W(5, __Vtemp_7)));
        vlSelfRef.rnd04__DOT__w1 = ((QData)((IData)(
                                                    VL_ONEHOT_I(
                                                                (0x0000ffffU 
                                                                 & (~ 
                                                                    VL_SHIFTRS_III(16,16,8, (IData)(vlSelfRef.i3), 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.i2)))))))) 
                                    + (QData)((IData)(
                                                      VL_COUNTONES_I(
                                                                     (1U 
                                                                      & (~ 
                                                                         (0U 
                                                                          != (IData)(vlSelfRef.i3))))))));
        vlSelfRef.rnd04__DOT__w2 = (3U & ((IData)(1U) 
                                          - VL_SHIFTR_III(2,8,8, (IData)(vlSelfRef.rnd04__DOT__r3), 
                                                          (0x000000ffU 
                                                           & vlSelfRef.rnd04__DOT__r0[0U]))));
        vlSelfRef.rnd04__DOT__w4 = (0x000000ffU & VL_SHIFTRS_III(8,8,8, 
                                                                 (0x000000ffU 
                                                                  & VL_EXTENDS_II(8,1, 
                                                                                VL_SHIFTR_III(1,1,8, 
                                                                                VL_LTS_IQQ(64, 1ULL, vlSelfRef.rnd04__DOT__w1), 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.rnd04__DOT__w1))))), 
                                                                 (0x000000ffU 
                                                                  & (IData)(vlSelfRef.rnd04__DOT__w1))));
        vlSelfRef.o0 = (0x000000ffU & (((IData)(1U) 
                                        - vlSelfRef.i5[0U]) 
                                       ^ (vlSelfRef.i5[0U] 
                                          + (IData)(vlSelfRef.rnd04__DOT__w2))));
        vlSelfRef.rnd04__DOT__w6 = ((1U & (vlSelfRef.i1 
                                           >> (0x0000001fU 
                                               & vlSelfRef.rnd04__DOT__r0[0U]))) 
                                    | ((0x0000ffffU 
                                        & (~ ((IData)(vlSelfRef.i3) 
                                              ^ (IData)(vlSelfRef.rnd04__DOT__w2)))) 
                                       == (0x0000001fU 
                                           & vlSelfRef.i1)));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0 = 
            (VL_COUNTONES_I((0x0000000fU & (((0U == 
                                              (0x0000001fU 
                                               & vlSelfRef.i5[0U]))
                                              ? 0U : 
                                             (vlSelfRef.i4
                                              [(((IData)(3U) 
                                                 + 
                                                 (0x0000003fU 
                                                  & vlSelfRef.i5[0U])) 
                                                >> 5U)] 
                                              << ((IData)(0x00000020U) 
                                                  - 
                                                  (0x0000001fU 
                                                   & vlSelfRef.i5[0U])))) 
                                            | (vlSelfRef.i4
                                               [(1U 
                                                 & (vlSelfRef.i5[0U] 
                                                    >> 5U))] 
                                               >> (0x0000001fU 
                                                   & vlSelfRef.i5[0U]))))) 
             < VL_SHIFTR_III(32,32,8, VL_SHIFTR_III(32,8,8, (IData)(vlSelfRef.rnd04__DOT__r3), (IData)(vlSelfRef.rnd04__DOT__w4)), 
                             (0x000000ffU & vlSelfRef.i5[0U])));
        __Vtemp_9[0U] = ((vlSelfRef.i0[1U] << 0x00000017U) 
                         | (vlSelfRef.i0[0U] >> 9U));
        __Vtemp_9[1U] = ((vlSelfRef.i0[2U] << 0x00000017U) 
                         | (vlSelfRef.i0[1U] >> 9U));
        __Vtemp_9[2U] = ((vlSelfRef.i0[3U] << 0x00000017U) 
                         | (vlSelfRef.i0[2U] >> 9U));
        __Vtemp_10[0U] = (1U & ((IData)(vlSelfRef.rnd04__DOT__w6) 
                                >> (1U & vlSelfRef.i5[0U])));
        __Vtemp_10[1U] = 0U;
        __Vtemp_10[2U] = 0U;
        VL_SUB_W(3, __Vtemp_11, __Vtemp_9, __Vtemp_10);
        vlSelfRef.rnd04__DOT__w10[0U] = ((0xfffffffeU 
                                          & vlSelfRef.rnd04__DOT__w10[0U]) 
                                         | (1U & (~ 
                                                  (VL_REDXOR_32(
                                                                VL_COUNTONES_I(
                                                                               (1U 
                                                                                & (~ 
                                                                                (0U 
                                                                                != 
                                                                                (0x0000007fU 
                                                                                & (vlSelfRef.i5[0U] 
                                                                                >> 2U))))))) 
                                                   ^ 
                                                   VL_REDXOR_32(
                                                                ((__Vtemp_11[0U] 
                                                                  ^ __Vtemp_11[1U]) 
                                                                 ^ 
                                                                 (0x0007ffffU 
                                                                  & __Vtemp_11[2U])))))));
        vlSelfRef.rnd04__DOT__w10[0U] = (1U & vlSelfRef.rnd04__DOT__w10[0U]);
        vlSelfRef.rnd04__DOT__w10[1U] = 0U;
        vlSelfRef.rnd04__DOT__w10[2U] = 0U;
        vlSelfRef.o3 = (0x000000ffffffffffULL & VL_MODDIVS_QQQ(40, 
                                                               (0x000000ffffffffffULL 
                                                                & VL_EXTENDS_QI(40,2, (IData)(vlSelfRef.rnd04__DOT__w2))), 
                                                               (0x000000ffffffffffULL 
                                                                & ((QData)((IData)(
                                                                                (1U 
                                                                                & ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0) 
                                                                                >> (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0))))) 
                                                                   | ((((QData)((IData)(vlSelfRef.i0[1U])) 
                                                                        << 0x00000020U) 
                                                                       | (QData)((IData)(vlSelfRef.i0[0U]))) 
                                                                      | ((QData)((IData)(vlSelfRef.i1)) 
                                                                         + (QData)((IData)(
                                                                                ((IData)(vlSelfRef.rnd04__DOT__w6) 
                                                                                <= 
                                                                                (1U 
                                                                                & vlSelfRef.i1))))))))));
        vlSelfRef.rnd04__DOT__w8 = (QData)((IData)(
                                                   ((VL_SHIFTR_III(1,1,8, 
                                                                   VL_LTS_III(9, 0x0051U, 
                                                                              (0x000001ffU 
                                                                               & VL_EXTENDS_II(9,1, 
                                                                                (1U 
                                                                                & (vlSelfRef.i1 
                                                                                >> (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)))))), 
                                                                   (0x000000ffU 
                                                                    & (IData)(vlSelfRef.rnd04__DOT__w7))) 
                                                     << 5U) 
                                                    | (0x0000001fU 
                                                       & vlSelfRef.i1))));
        vlSelfRef.o2 = (IData)((0x000000ffffffffffULL 
                                & VL_DIV_QQQ(40, vlSelfRef.rnd04__DOT__w8, (QData)((IData)(
                                                                                (1U 
                                                                                & (~ 
                                                                                (0U 
                                                                                != 
                                                                                (0x0000007fU 
                                                                                & (vlSelfRef.i5[0U] 
                                                                                >> 2U))))))))));
    }

    VLCU_D static bool Vrnd04___024root___eval_phase__stl(Vrnd04___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vrnd04___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
        }
#endif
        __VstlExecute = Vrnd04___024root___trigger_anySet__stl(vlSelfRef.__VstlTriggered);
        if (__VstlExecute) {
            {
                // Inlined CFunc: _eval_stl
                if ((1ULL & vlSelfRef.__VstlTriggered[0U])) {
                    Vrnd04___024root___stl_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
        }
        return (__VstlExecute);
    }

    VLCU_D static void Vrnd04___024root___eval_triggers_vec__ico(Vrnd04___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlSelfRef.__VicoTriggered[0U] = (QData)((IData)(
                                                        ((((0U 
                                                            != 
                                                            ((((vlSelfRef.i5[0U] 
                                                                ^ vlSelfRef.__Vtrigprevexpr___TOP__i5__0[0U]) 
                                                               | (vlSelfRef.i5[1U] 
                                                                  ^ vlSelfRef.__Vtrigprevexpr___TOP__i5__0[1U])) 
                                                              | (vlSelfRef.i5[2U] 
                                                                 ^ vlSelfRef.__Vtrigprevexpr___TOP__i5__0[2U])) 
                                                             | (vlSelfRef.i5[3U] 
                                                                ^ vlSelfRef.__Vtrigprevexpr___TOP__i5__0[3U]))) 
                                                           << 6U) 
                                                          | (((0U 
                                                               != 
                                                               (((((vlSelfRef.i4[0U] 
                                                                    ^ vlSelfRef.__Vtrigprevexpr___TOP__i4__0[0U]) 
                                                                   | (vlSelfRef.i4[1U] 
                                                                      ^ vlSelfRef.__Vtrigprevexpr___TOP__i4__0[1U])) 
                                                                  | (vlSelfRef.i4[2U] 
                                                                     ^ vlSelfRef.__Vtrigprevexpr___TOP__i4__0[2U])) 
                                                                 | (vlSelfRef.i4[3U] 
                                                                    ^ vlSelfRef.__Vtrigprevexpr___TOP__i4__0[3U])) 
                                                                | (vlSelfRef.i4[4U] 
                                                                   ^ vlSelfRef.__Vtrigprevexpr___TOP__i4__0[4U]))) 
                                                              << 5U) 
                                                             | (((IData)(vlSelfRef.i3) 
                                                                 != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__i3__0)) 
                                                                << 4U))) 
                                                         | ((((vlSelfRef.i2 
                                                               != vlSelfRef.__Vtrigprevexpr___TOP__i2__0) 
                                                              << 3U) 
                                                             | ((vlSelfRef.i1 
                                                                 != vlSelfRef.__Vtrigprevexpr___TOP__i1__0) 
                                                                << 2U)) 
                                                            | (((0U 
                                                                 != 
                                                                 (((((vlSelfRef.i0[0U] 
                                                                      ^ vlSelfRef.__Vtrigprevexpr___TOP__i0__0[0U]) 
                                                                     | (vlSelfRef.i0[1U] 
                                                                        ^ vlSelfRef.__Vtrigprevexpr___TOP__i0__0[1U])) 
                                                                    | (vlSelfRef.i0[2U] 
                                                                       ^ vlSelfRef.__Vtrigprevexpr___TOP__i0__0[2U])) 
                                                                   | (vlSelfRef.i0[3U] 
                                                                      ^ vlSelfRef.__Vtrigprevexpr___TOP__i0__0[3U])) 
                                                                  | (vlSelfRef.i0[4U] 
                                                                     ^ vlSelfRef.__Vtrigprevexpr___TOP__i0__0[4U]))) 
                                                                << 1U) 
                                                               | ((IData)(vlSelfRef.clk) 
                                                                  != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__clk__0)))))));
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
        vlSelfRef.__Vtrigprevexpr___TOP__i0__0[0U] 
            = vlSelfRef.i0[0U];
        vlSelfRef.__Vtrigprevexpr___TOP__i0__0[1U] 
            = vlSelfRef.i0[1U];
        vlSelfRef.__Vtrigprevexpr___TOP__i0__0[2U] 
            = vlSelfRef.i0[2U];
        vlSelfRef.__Vtrigprevexpr___TOP__i0__0[3U] 
            = vlSelfRef.i0[3U];
        vlSelfRef.__Vtrigprevexpr___TOP__i0__0[4U] 
            = vlSelfRef.i0[4U];
        vlSelfRef.__Vtrigprevexpr___TOP__i1__0 = vlSelfRef.i1;
        vlSelfRef.__Vtrigprevexpr___TOP__i2__0 = vlSelfRef.i2;
        vlSelfRef.__Vtrigprevexpr___TOP__i3__0 = vlSelfRef.i3;
        vlSelfRef.__Vtrigprevexpr___TOP__i4__0[0U] 
            = vlSelfRef.i4[0U];
        vlSelfRef.__Vtrigprevexpr___TOP__i4__0[1U] 
            = vlSelfRef.i4[1U];
        vlSelfRef.__Vtrigprevexpr___TOP__i4__0[2U] 
            = vlSelfRef.i4[2U];
        vlSelfRef.__Vtrigprevexpr___TOP__i4__0[3U] 
            = vlSelfRef.i4[3U];
        vlSelfRef.__Vtrigprevexpr___TOP__i4__0[4U] 
            = vlSelfRef.i4[4U];
        vlSelfRef.__Vtrigprevexpr___TOP__i5__0[0U] 
            = vlSelfRef.i5[0U];
        vlSelfRef.__Vtrigprevexpr___TOP__i5__0[1U] 
            = vlSelfRef.i5[1U];
        vlSelfRef.__Vtrigprevexpr___TOP__i5__0[2U] 
            = vlSelfRef.i5[2U];
        vlSelfRef.__Vtrigprevexpr___TOP__i5__0[3U] 
            = vlSelfRef.i5[3U];
        if (VL_UNLIKELY(((1U & (~ (IData)(vlSelfRef.__VicoDidInit)))))) {
            vlSelfRef.__VicoDidInit = 1U;
            vlSelfRef.__VicoTriggered[0U] = (1ULL | vlSelfRef.__VicoTriggered[0U]);
            vlSelfRef.__VicoTriggered[0U] = (2ULL | vlSelfRef.__VicoTriggered[0U]);
            vlSelfRef.__VicoTriggered[0U] = (4ULL | vlSelfRef.__VicoTriggered[0U]);
            vlSelfRef.__VicoTriggered[0U] = (8ULL | vlSelfRef.__VicoTriggered[0U]);
            vlSelfRef.__VicoTriggered[0U] = (0x0000000000000010ULL 
                                             | vlSelfRef.__VicoTriggered[0U]);
            vlSelfRef.__VicoTriggered[0U] = (0x0000000000000020ULL 
                                             | vlSelfRef.__VicoTriggered[0U]);
            vlSelfRef.__VicoTriggered[0U] = (0x0000000000000040ULL 
                                             | vlSelfRef.__VicoTriggered[0U]);
        }
    }

    VLCU_D static bool Vrnd04___024root___trigger_anySet__ico(const VlUnpacked<QData/*63:0*/, 2> &in) {
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

    VLCU_D static void Vrnd04___024root___eval_ico(Vrnd04___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        if ((0x0000000000000040ULL & vlSelfRef.__VicoTriggered[0U])) {
            {
                // Inlined CFunc: _ico_sequent__TOP__0
                vlSelfRef.o0 = (0x000000ffU & (((IData)(1U) 
                                                - vlSelfRef.i5[0U]) 
                                               ^ (vlSelfRef.i5[0U] 
                                                  + (IData)(vlSelfRef.rnd04__DOT__w2))));
                vlSelfRef.o4[0U] = ((0xfffffffeU & vlSelfRef.o4[0U]) 
                                    | ((0x00000020U 
                                        & ((~ (0U != 
                                               (0x0000007fU 
                                                & (vlSelfRef.i5[0U] 
                                                   >> 2U)))) 
                                           << 5U)) 
                                       >= (0x00ffffffU 
                                           & (~ (1U 
                                                 ^ 
                                                 VL_DIV_III(24, (IData)(vlSelfRef.rnd04__DOT__r2), vlSelfRef.rnd04__DOT__mem
                                                            [vlSelfRef.rnd04__DOT__r2]))))));
                vlSelfRef.o4[0U] = (1U & vlSelfRef.o4[0U]);
                vlSelfRef.o4[1U] = 0U;
                vlSelfRef.o4[2U] = 0U;
            }
        }
        if ((0x0000000000000020ULL & vlSelfRef.__VicoTriggered[0U])) {
            {
                // Inlined CFunc: _ico_sequent__TOP__1
                vlSelfRef.o1 = (0x0000ffffU & ((IData)(0x0f28U) 
                                               - (vlSelfRef.i4[0U] 
                                                  - vlSelfRef.rnd04__DOT__r0[0U])));
            }
        }
        if ((0x0000000000000014ULL & vlSelfRef.__VicoTriggered[0U])) {
            {
                // Inlined CFunc: _ico_comb__TOP__0
                vlSelfRef.rnd04__DOT__w6 = ((1U & (vlSelfRef.i1 
                                                   >> 
                                                   (0x0000001fU 
                                                    & vlSelfRef.rnd04__DOT__r0[0U]))) 
                                            | ((0x0000ffffU 
                                                & (~ 
                                                   ((IData)(vlSelfRef.i3) 
                                                    ^ (IData)(vlSelfRef.rnd04__DOT__w2)))) 
                                               == (0x0000001fU 
                                                   & vlSelfRef.i1)));
            }
        }
        if ((0x000000000000006cULL & vlSelfRef.__VicoTriggered[0U])) {
            {
                // Inlined CFunc: _ico_comb__TOP__1
                VlWide<5>/*159:0*/ __Vinline_0__ico_comb__TOP__1___Vtemp_6;
                __Vinline_0__ico_comb__TOP__1___Vtemp_6[0U] 
                    = (vlSelfRef.i4[0U] | ((vlSelfRef.rnd04__DOT__r0[0U] 
                                            ^ (IData)(vlSelfRef.rnd04__DOT__r3)) 
                                           | (1U & (IData)(
                                                           (vlSelfRef.i2 
                                                            >> 
                                                            (0x0000001fU 
                                                             & vlSelfRef.i5[0U]))))));
                __Vinline_0__ico_comb__TOP__1___Vtemp_6[1U] 
                    = (vlSelfRef.i4[1U] | vlSelfRef.rnd04__DOT__r0[1U]);
                __Vinline_0__ico_comb__TOP__1___Vtemp_6[2U] 
                    = (vlSelfRef.i4[2U] | vlSelfRef.rnd04__DOT__r0[2U]);
                __Vinline_0__ico_comb__TOP__1___Vtemp_6[3U] 
                    = vlSelfRef.i4[3U];
                __Vinline_0__ico_comb__TOP__1___Vtemp_6[4U] 
                    = vlSelfRef.i4[4U];
                vlSelfRef.rnd04__DOT__w7 = (0x0000ffffU 
                                            & (VL_COUNTONES_I(
                                                              ((IData)(vlSelfRef.rnd04__DOT__r2) 
                                                               >= 
                                                               VL_COUNTONES_I(
                                                                              (1U 
                                                                               & (vlSelfRef.i1 
                                                                                >> 
                                                                                (0x0000001fU 
                                                                                & vlSelfRef.i5[0U])))))) 
                                               + VL_ONEHOT_W(5, __Vinline_0__ico_comb__TOP__1___Vtemp_6)));
            }
        }
        if ((0x0000000000000018ULL & vlSelfRef.__VicoTriggered[0U])) {
            {
                // Inlined CFunc: _ico_comb__TOP__2
                vlSelfRef.rnd04__DOT__w1 = ((QData)((IData)(
                                                            VL_ONEHOT_I(
                                                                        (0x0000ffffU 
                                                                         & (~ 
                                                                            VL_SHIFTRS_III(16,16,8, (IData)(vlSelfRef.i3), 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.i2)))))))) 
                                            + (QData)((IData)(
                                                              VL_COUNTONES_I(
                                                                             (1U 
                                                                              & (~ 
                                                                                (0U 
                                                                                != (IData)(vlSelfRef.i3))))))));
                vlSelfRef.rnd04__DOT__w4 = (0x000000ffU 
                                            & VL_SHIFTRS_III(8,8,8, 
                                                             (0x000000ffU 
                                                              & VL_EXTENDS_II(8,1, 
                                                                              VL_SHIFTR_III(1,1,8, 
                                                                                VL_LTS_IQQ(64, 1ULL, vlSelfRef.rnd04__DOT__w1), 
                                                                                (0x000000ffU 
                                                                                & (IData)(vlSelfRef.rnd04__DOT__w1))))), 
                                                             (0x000000ffU 
                                                              & (IData)(vlSelfRef.rnd04__DOT__w1))));
            }
        }
        if ((0x0000000000000056ULL & vlSelfRef.__VicoTriggered[0U])) {
            {
                // Inlined CFunc: _ico_comb__TOP__3
                VlWide<3>/*95:0*/ __Vinline_0__ico_comb__TOP__3___Vtemp_2;
                VlWide<3>/*95:0*/ __Vinline_0__ico_comb__TOP__3___Vtemp_3;
                VlWide<3>/*95:0*/ __Vinline_0__ico_comb__TOP__3___Vtemp_4;
                __Vinline_0__ico_comb__TOP__3___Vtemp_2[0U] 
                    = ((vlSelfRef.i0[1U] << 0x00000017U) 
                       | (vlSelfRef.i0[0U] >> 9U));
                __Vinline_0__ico_comb__TOP__3___Vtemp_2[1U] 
                    = ((vlSelfRef.i0[2U] << 0x00000017U) 
                       | (vlSelfRef.i0[1U] >> 9U));
                __Vinline_0__ico_comb__TOP__3___Vtemp_2[2U] 
                    = ((vlSelfRef.i0[3U] << 0x00000017U) 
                       | (vlSelfRef.i0[2U] >> 9U));
                __Vinline_0__ico_comb__TOP__3___Vtemp_3[0U] 
                    = (1U & ((IData)(vlSelfRef.rnd04__DOT__w6) 
                             >> (1U & vlSelfRef.i5[0U])));
                __Vinline_0__ico_comb__TOP__3___Vtemp_3[1U] = 0U;
                __Vinline_0__ico_comb__TOP__3___Vtemp_3[2U] = 0U;
                VL_SUB_W(3, __Vinline_0__ico_comb__TOP__3___Vtemp_4, __Vinline_0__ico_comb__TOP__3___Vtemp_2, __Vinline_0__ico_comb__TOP__3___Vtemp_3);
                vlSelfRef.rnd04__DOT__w10[0U] = ((0xfffffffeU 
                                                  & vlSelfRef.rnd04__DOT__w10[0U]) 
                                                 | (1U 
                                                    & (~ 
                                                       (VL_REDXOR_32(
                                                                     VL_COUNTONES_I(
                                                                                (1U 
                                                                                & (~ 
                                                                                (0U 
                                                                                != 
                                                                                (0x0000007fU 
                                                                                & (vlSelfRef.i5[0U] 
                                                                                >> 2U))))))) 
                                                        ^ 
                                                        VL_REDXOR_32(
                                                                     ((__Vinline_0__ico_comb__TOP__3___Vtemp_4[0U] 
                                                                       ^ __Vinline_0__ico_comb__TOP__3___Vtemp_4[1U]) 
                                                                      ^ 
                                                                      (0x0007ffffU 
                                                                       & __Vinline_0__ico_comb__TOP__3___Vtemp_4[2U])))))));
                vlSelfRef.rnd04__DOT__w10[0U] = (1U 
                                                 & vlSelfRef.rnd04__DOT__w10[0U]);
                vlSelfRef.rnd04__DOT__w10[1U] = 0U;
                vlSelfRef.rnd04__DOT__w10[2U] = 0U;
            }
        }
        if ((0x0000000000000078ULL & vlSelfRef.__VicoTriggered[0U])) {
            {
                // Inlined CFunc: _ico_comb__TOP__4
                vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0 
                    = (VL_COUNTONES_I((0x0000000fU 
                                       & (((0U == (0x0000001fU 
                                                   & vlSelfRef.i5[0U]))
                                            ? 0U : 
                                           (vlSelfRef.i4
                                            [(((IData)(3U) 
                                               + (0x0000003fU 
                                                  & vlSelfRef.i5[0U])) 
                                              >> 5U)] 
                                            << ((IData)(0x00000020U) 
                                                - (0x0000001fU 
                                                   & vlSelfRef.i5[0U])))) 
                                          | (vlSelfRef.i4
                                             [(1U & 
                                               (vlSelfRef.i5[0U] 
                                                >> 5U))] 
                                             >> (0x0000001fU 
                                                 & vlSelfRef.i5[0U]))))) 
                       < VL_SHIFTR_III(32,32,8, VL_SHIFTR_III(32,8,8, (IData)(vlSelfRef.rnd04__DOT__r3), (IData)(vlSelfRef.rnd04__DOT__w4)), 
                                       (0x000000ffU 
                                        & vlSelfRef.i5[0U])));
            }
        }
        if ((0x000000000000007eULL & vlSelfRef.__VicoTriggered[0U])) {
            {
                // Inlined CFunc: _ico_comb__TOP__5
                vlSelfRef.o3 = (0x000000ffffffffffULL 
                                & VL_MODDIVS_QQQ(40, 
                                                 (0x000000ffffffffffULL 
                                                  & VL_EXTENDS_QI(40,2, (IData)(vlSelfRef.rnd04__DOT__w2))), 
                                                 (0x000000ffffffffffULL 
                                                  & ((QData)((IData)(
                                                                     (1U 
                                                                      & ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0) 
                                                                         >> (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0))))) 
                                                     | ((((QData)((IData)(vlSelfRef.i0[1U])) 
                                                          << 0x00000020U) 
                                                         | (QData)((IData)(vlSelfRef.i0[0U]))) 
                                                        | ((QData)((IData)(vlSelfRef.i1)) 
                                                           + (QData)((IData)(
                                                                             ((IData)(vlSelfRef.rnd04__DOT__w6) 
                                                                              <= 
                                                                              (1U 
                                                                               & vlSelfRef.i1))))))))));
            }
        }
        if ((0x000000000000007cULL & vlSelfRef.__VicoTriggered[0U])) {
            {
                // Inlined CFunc: _ico_comb__TOP__6
                vlSelfRef.rnd04__DOT__w8 = (QData)((IData)(
                                                           ((VL_SHIFTR_III(1,1,8, 
                                                                           VL_LTS_III(9, 0x0051U, 
                                                                                (0x000001ffU 
                                                                                & VL_EXTENDS_II(9,1, 
                                                                                (1U 
                                                                                & (vlSelfRef.i1 
                                                                                >> (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)))))), 
                                                                           (0x000000ffU 
                                                                            & (IData)(vlSelfRef.rnd04__DOT__w7))) 
                                                             << 5U) 
                                                            | (0x0000001fU 
                                                               & vlSelfRef.i1))));
                vlSelfRef.o2 = (IData)((0x000000ffffffffffULL 
                                        & VL_DIV_QQQ(40, vlSelfRef.rnd04__DOT__w8, (QData)((IData)(
                                                                                (1U 
                                                                                & (~ 
                                                                                (0U 
                                                                                != 
                                                                                (0x0000007fU 
                                                                                & (vlSelfRef.i5[0U] 
                                                                                >> 2U))))))))));
            }
        }
    }

    VLCU_D static bool Vrnd04___024root___eval_phase__ico(Vrnd04___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VicoExecute;
        // Body
        Vrnd04___024root___eval_triggers_vec__ico(vlSelf, vlCtx);
#ifdef VL_DEBUG
        if (VL_UNLIKELY(vlSymsp->_vm_contextp__->debug())) {
            Vrnd04___024root___dump_triggers__ico(vlSelfRef.__VicoTriggered, "ico"s);
        }
#endif
        __VicoExecute = Vrnd04___024root___trigger_anySet__ico(vlSelfRef.__VicoTriggered);
        if (__VicoExecute) {
            Vrnd04___024root___eval_ico(vlSelf, vlCtx);
        }
        return (__VicoExecute);
    }

    VLCU_D static bool Vrnd04___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vrnd04___024root___nba_sequent__TOP__0(Vrnd04___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        VlWide<3>/*95:0*/ __Vdly__rnd04__DOT__r0;
        VL_ZERO_W(96, __Vdly__rnd04__DOT__r0);
        IData/*31:0*/ __Vdly__rnd04__DOT__r1;
        __Vdly__rnd04__DOT__r1 = 0;
        IData/*23:0*/ __VdlyVal__rnd04__DOT__mem__v0;
        __VdlyVal__rnd04__DOT__mem__v0 = 0;
        CData/*3:0*/ __VdlyDim0__rnd04__DOT__mem__v0;
        __VdlyDim0__rnd04__DOT__mem__v0 = 0;
        VlWide<3>/*95:0*/ __Vtemp_2;
        VlWide<3>/*95:0*/ __Vtemp_3;
        VlWide<3>/*95:0*/ __Vtemp_4;
        VlWide<3>/*95:0*/ __Vtemp_7;
        VlWide<3>/*95:0*/ __Vtemp_8;
        VlWide<3>/*95:0*/ __Vtemp_9;
        VlWide<3>/*95:0*/ __Vtemp_10;
        VlWide<3>/*95:0*/ __Vtemp_11;
        VlWide<4>/*127:0*/ __Vtemp_12;
        VlWide<5>/*159:0*/ __Vtemp_19;
        VlWide<3>/*95:0*/ __Vtemp_21;
        VlWide<3>/*95:0*/ __Vtemp_22;
        VlWide<3>/*95:0*/ __Vtemp_23;
        // Body
        __Vdly__rnd04__DOT__r1 = vlSelfRef.rnd04__DOT__r1;
        __Vdly__rnd04__DOT__r0[0U] = vlSelfRef.rnd04__DOT__r0[0U];
        __Vdly__rnd04__DOT__r0[1U] = vlSelfRef.rnd04__DOT__r0[1U];
        __Vdly__rnd04__DOT__r0[2U] = vlSelfRef.rnd04__DOT__r0[2U];
        if ((0U != ((1U & (vlSelfRef.i1 >> (3U & vlSelfRef.rnd04__DOT__w10[0U]))) 
                    | (IData)(vlSelfRef.rnd04__DOT__r3)))) {
            __Vtemp_2[0U] = (((IData)(vlSelfRef.rnd04__DOT__w1) 
                              << 1U) | (1U & (vlSelfRef.i1 
                                              >> (IData)(vlSelfRef.rnd04__DOT__w2))));
            __Vtemp_2[1U] = (((IData)(vlSelfRef.rnd04__DOT__w1) 
                              >> 0x0000001fU) | ((IData)(
                                                         (vlSelfRef.rnd04__DOT__w1 
                                                          >> 0x00000020U)) 
                                                 << 1U));
            __Vtemp_2[2U] = ((IData)((vlSelfRef.rnd04__DOT__w1 
                                      >> 0x00000020U)) 
                             >> 0x0000001fU);
            __Vtemp_3[0U] = 0x00dbb8c8U;
            __Vtemp_3[1U] = 0U;
            __Vtemp_3[2U] = 0U;
            VL_SUB_W(3, __Vtemp_4, __Vtemp_2, __Vtemp_3);
            __Vdly__rnd04__DOT__r0[0U] = (~ (__Vtemp_4[0U] 
                                             ^ VL_LTES_IQQ(40, 0x000000ffffffffffULL, 
                                                           (0x000000ffffffffffULL 
                                                            & (vlSelfRef.rnd04__DOT__w8 
                                                               * (QData)((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)))))));
            __Vdly__rnd04__DOT__r0[1U] = (~ __Vtemp_4[1U]);
            __Vdly__rnd04__DOT__r0[2U] = (~ __Vtemp_4[2U]);
        } else {
            __Vtemp_7[0U] = vlSelfRef.rnd04__DOT__r2;
            __Vtemp_7[1U] = 0U;
            __Vtemp_7[2U] = 0U;
            __Vtemp_8[0U] = (1U & (vlSelfRef.i5[2U] 
                                   >> 0x0000000eU));
            __Vtemp_8[1U] = 0U;
            __Vtemp_8[2U] = 0U;
            VL_SUB_W(3, __Vtemp_9, __Vtemp_7, __Vtemp_8);
            __Vtemp_10[0U] = (0x0000000fU & (vlSelfRef.rnd04__DOT__r1 
                                             >> (0x0000000fU 
                                                 & vlSelfRef.i0[0U])));
            __Vtemp_10[1U] = 0U;
            __Vtemp_10[2U] = 0U;
            VL_ADD_W(3, __Vdly__rnd04__DOT__r0, __Vtemp_9, __Vtemp_10);
        }
        VL_EXTENDS_WI(96,1, __Vtemp_11, (IData)(vlSelfRef.rnd04__DOT__r2));
        __Vdly__rnd04__DOT__r1 = (IData)((0x0000ffffffffffffULL 
                                          & (VL_SHIFTR_QQI(48,48,8, vlSelfRef.i2, 
                                                           (0x000000ffU 
                                                            & vlSelfRef.i4[0U])) 
                                             * (0x0000ffffffffffffULL 
                                                & (((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                     ? (QData)((IData)(
                                                                       (3U 
                                                                        & (IData)(vlSelfRef.rnd04__DOT__w4))))
                                                     : (QData)((IData)(vlSelfRef.rnd04__DOT__w4))) 
                                                   - (QData)((IData)(
                                                                     VL_LTES_IWW(96, __Vtemp_11, vlSelfRef.rnd04__DOT__w10))))))));
        __VdlyVal__rnd04__DOT__mem__v0 = (0x00ffffffU 
                                          & (~ (VL_SHIFTL_III(24,24,8, 
                                                              (1U 
                                                               & ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0) 
                                                                  >> 
                                                                  (1U 
                                                                   & (~ 
                                                                      (0U 
                                                                       != 
                                                                       (0x0000007fU 
                                                                        & (vlSelfRef.i5[0U] 
                                                                           >> 2U))))))), 
                                                              (0x000000ffU 
                                                               & vlSelfRef.i1)) 
                                                ^ (0x0000001fU 
                                                   & vlSelfRef.i1))));
        __VdlyDim0__rnd04__DOT__mem__v0 = (0x0000000fU 
                                           & vlSelfRef.rnd04__DOT__r1);
        if ((0U != (VL_COUNTONES_W(4, vlSelfRef.i5) 
                    + VL_ONEHOT_I((0x01ffffffU & (IData)(
                                                         (vlSelfRef.i2 
                                                          >> 0x0aU))))))) {
            __Vtemp_12[0U] = vlSelfRef.i5[0U];
            __Vtemp_12[1U] = vlSelfRef.i5[1U];
            __Vtemp_12[2U] = vlSelfRef.i5[2U];
            __Vtemp_12[3U] = vlSelfRef.i5[3U];
        } else {
            __Vtemp_12[0U] = ((0U != (0x0000000fU & 
                                      ((IData)(vlSelfRef.rnd04__DOT__w7) 
                                       >> (7U & vlSelfRef.rnd04__DOT__r0[0U]))))
                               ? (IData)(vlSelfRef.rnd04__DOT__w4)
                               : (IData)(vlSelfRef.rnd04__DOT__r3));
        }
        vlSelfRef.rnd04__DOT__r2 = (1U & ((0U != (7U 
                                                  & VL_SHIFTRS_III(3,3,8, 
                                                                   (7U 
                                                                    & (vlSelfRef.i5[0U] 
                                                                       >> 3U)), 
                                                                   (0x000000ffU 
                                                                    & vlSelfRef.i5[0U]))))
                                           ? vlSelfRef.i0[0U]
                                           : __Vtemp_12[0U]));
        vlSelfRef.rnd04__DOT__r1 = __Vdly__rnd04__DOT__r1;
        vlSelfRef.rnd04__DOT__mem[__VdlyDim0__rnd04__DOT__mem__v0] 
            = __VdlyVal__rnd04__DOT__mem__v0;
        vlSelfRef.rnd04__DOT__r0[0U] = __Vdly__rnd04__DOT__r0[0U];
        vlSelfRef.rnd04__DOT__r0[1U] = __Vdly__rnd04__DOT__r0[1U];
        vlSelfRef.rnd04__DOT__r0[2U] = __Vdly__rnd04__DOT__r0[2U];
        vlSelfRef.o4[0U] = ((0xfffffffeU & vlSelfRef.o4[0U]) 
                            | ((0x00000020U & ((~ (0U 
                                                   != 
                                                   (0x0000007fU 
                                                    & (vlSelfRef.i5[0U] 
                                                       >> 2U)))) 
                                               << 5U)) 
                               >= (0x00ffffffU & (~ 
                                                  (1U 
                                                   ^ 
                                                   VL_DIV_III(24, (IData)(vlSelfRef.rnd04__DOT__r2), vlSelfRef.rnd04__DOT__mem
                                                              [vlSelfRef.rnd04__DOT__r2]))))));
        vlSelfRef.o4[0U] = (1U & vlSelfRef.o4[0U]);
        vlSelfRef.o4[1U] = 0U;
        vlSelfRef.o4[2U] = 0U;
        vlSelfRef.o1 = (0x0000ffffU & ((IData)(0x0f28U) 
                                       - (vlSelfRef.i4[0U] 
                                          - vlSelfRef.rnd04__DOT__r0[0U])));
        vlSelfRef.rnd04__DOT__r3 = (0x000000ffU & (~ (IData)(
                                                             VL_SHIFTRS_QQW(64,64,150, vlSelfRef.rnd04__DOT__w1, vlSelfRef.i0))));
        __Vtemp_19[0U] = (vlSelfRef.i4[0U] | ((vlSelfRef.rnd04__DOT__r0[0U] 
                                               ^ (IData)(vlSelfRef.rnd04__DOT__r3)) 
                                              | (1U 
                                                 & (IData)(
                                                           (vlSelfRef.i2 
                                                            >> 
                                                            (0x0000001fU 
                                                             & vlSelfRef.i5[0U]))))));
        __Vtemp_19[1U] = (vlSelfRef.i4[1U] | vlSelfRef.rnd04__DOT__r0[1U]);
        __Vtemp_19[2U] = (vlSelfRef.i4[2U] | vlSelfRef.rnd04__DOT__r0[2U]);
        __Vtemp_19[3U] = vlSelfRef.i4[3U];
        __Vtemp_19[4U] = vlSelfRef.i4[4U];
        vlSelfRef.rnd04__DOT__w7 = (0x0000ffffU & (
                                                   VL_COUNTONES_I(
                                                                  ((IData)(vlSelfRef.rnd04__DOT__r2) 
                                                                   >= 
                                                                   VL_COUNTONES_I(
                                                                                (1U 
                                                                                & (vlSelfRef.i1 
                                                                                >> 
                                                                                (0x0000001fU 
                                                                                & vlSelfRef.i5[0U])))))) 
                                                   + 
                                                   VL_ONEHOT_W(5, __Vtemp_19)));
        vlSelfRef.rnd04__DOT__w2 = (3U & ((IData)(1U) 
                                          - VL_SHIFTR_III(2,8,8, (IData)(vlSelfRef.rnd04__DOT__r3), 
                                                          (0x000000ffU 
                                                           & vlSelfRef.rnd04__DOT__r0[0U]))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0 = 
            (VL_COUNTONES_I((0x0000000fU & (((0U == 
                                              (0x0000001fU 
                                               & vlSelfRef.i5[0U]))
                                              ? 0U : 
                                             (vlSelfRef.i4
                                              [(((IData)(3U) 
                                                 + 
                                                 (0x0000003fU 
                                                  & vlSelfRef.i5[0U])) 
                                                >> 5U)] 
                                              << ((IData)(0x00000020U) 
                                                  - 
                                                  (0x0000001fU 
                                                   & vlSelfRef.i5[0U])))) 
                                            | (vlSelfRef.i4
                                               [(1U 
                                                 & (vlSelfRef.i5[0U] 
                                                    >> 5U))] 
                                               >> (0x0000001fU 
                                                   & vlSelfRef.i5[0U]))))) 
             < VL_SHIFTR_III(32,32,8, VL_SHIFTR_III(32,8,8, (IData)(vlSelfRef.rnd04__DOT__r3), (IData)(vlSelfRef.rnd04__DOT__w4)), 
                             (0x000000ffU & vlSelfRef.i5[0U])));
        vlSelfRef.o0 = (0x000000ffU & (((IData)(1U) 
                                        - vlSelfRef.i5[0U]) 
                                       ^ (vlSelfRef.i5[0U] 
                                          + (IData)(vlSelfRef.rnd04__DOT__w2))));
        vlSelfRef.rnd04__DOT__w6 = ((1U & (vlSelfRef.i1 
                                           >> (0x0000001fU 
                                               & vlSelfRef.rnd04__DOT__r0[0U]))) 
                                    | ((0x0000ffffU 
                                        & (~ ((IData)(vlSelfRef.i3) 
                                              ^ (IData)(vlSelfRef.rnd04__DOT__w2)))) 
                                       == (0x0000001fU 
                                           & vlSelfRef.i1)));
        vlSelfRef.rnd04__DOT__w8 = (QData)((IData)(
                                                   ((VL_SHIFTR_III(1,1,8, 
                                                                   VL_LTS_III(9, 0x0051U, 
                                                                              (0x000001ffU 
                                                                               & VL_EXTENDS_II(9,1, 
                                                                                (1U 
                                                                                & (vlSelfRef.i1 
                                                                                >> (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)))))), 
                                                                   (0x000000ffU 
                                                                    & (IData)(vlSelfRef.rnd04__DOT__w7))) 
                                                     << 5U) 
                                                    | (0x0000001fU 
                                                       & vlSelfRef.i1))));
        __Vtemp_21[0U] = ((vlSelfRef.i0[1U] << 0x00000017U) 
                          | (vlSelfRef.i0[0U] >> 9U));
        __Vtemp_21[1U] = ((vlSelfRef.i0[2U] << 0x00000017U) 
                          | (vlSelfRef.i0[1U] >> 9U));
        __Vtemp_21[2U] = ((vlSelfRef.i0[3U] << 0x00000017U) 
                          | (vlSelfRef.i0[2U] >> 9U));
        __Vtemp_22[0U] = (1U & ((IData)(vlSelfRef.rnd04__DOT__w6) 
                                >> (1U & vlSelfRef.i5[0U])));
        __Vtemp_22[1U] = 0U;
        __Vtemp_22[2U] = 0U;
        VL_SUB_W(3, __Vtemp_23, __Vtemp_21, __Vtemp_22);
        vlSelfRef.rnd04__DOT__w10[0U] = ((0xfffffffeU 
                                          & vlSelfRef.rnd04__DOT__w10[0U]) 
                                         | (1U & (~ 
                                                  (VL_REDXOR_32(
                                                                VL_COUNTONES_I(
                                                                               (1U 
                                                                                & (~ 
                                                                                (0U 
                                                                                != 
                                                                                (0x0000007fU 
                                                                                & (vlSelfRef.i5[0U] 
                                                                                >> 2U))))))) 
                                                   ^ 
                                                   VL_REDXOR_32(
                                                                ((__Vtemp_23[0U] 
                                                                  ^ __Vtemp_23[1U]) 
                                                                 ^ 
                                                                 (0x0007ffffU 
                                                                  & __Vtemp_23[2U])))))));
        vlSelfRef.rnd04__DOT__w10[0U] = (1U & vlSelfRef.rnd04__DOT__w10[0U]);
        vlSelfRef.rnd04__DOT__w10[1U] = 0U;
        vlSelfRef.rnd04__DOT__w10[2U] = 0U;
        vlSelfRef.o3 = (0x000000ffffffffffULL & VL_MODDIVS_QQQ(40, 
                                                               (0x000000ffffffffffULL 
                                                                & VL_EXTENDS_QI(40,2, (IData)(vlSelfRef.rnd04__DOT__w2))), 
                                                               (0x000000ffffffffffULL 
                                                                & ((QData)((IData)(
                                                                                (1U 
                                                                                & ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0) 
                                                                                >> (IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0))))) 
                                                                   | ((((QData)((IData)(vlSelfRef.i0[1U])) 
                                                                        << 0x00000020U) 
                                                                       | (QData)((IData)(vlSelfRef.i0[0U]))) 
                                                                      | ((QData)((IData)(vlSelfRef.i1)) 
                                                                         + (QData)((IData)(
                                                                                ((IData)(vlSelfRef.rnd04__DOT__w6) 
                                                                                <= 
                                                                                (1U 
                                                                                & vlSelfRef.i1))))))))));
        vlSelfRef.o2 = (IData)((0x000000ffffffffffULL 
                                & VL_DIV_QQQ(40, vlSelfRef.rnd04__DOT__w8, (QData)((IData)(
                                                                                (1U 
                                                                                & (~ 
                                                                                (0U 
                                                                                != 
                                                                                (0x0000007fU 
                                                                                & (vlSelfRef.i5[0U] 
                                                                                >> 2U))))))))));
    }

    VLCU_D static void Vrnd04___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vrnd04___024root___eval_phase__act(Vrnd04___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vrnd04___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vrnd04___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vrnd04___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vrnd04___024root___eval_phase__nba(Vrnd04___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vrnd04___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    Vrnd04___024root___nba_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
            Vrnd04___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vrnd04___024root___eval(Vrnd04___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vrnd04___024root___dump_triggers__ico(vlSelfRef.__VicoTriggered, "ico"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/rnd04.v", 2, "", "DIDNOTCONVERGE: Input combinational region did not converge after '--converge-limit' of 10000 tries");
            }
            __VicoIterCount = ((IData)(1U) + __VicoIterCount);
            vlSelfRef.__VicoPhaseResult = Vrnd04___024root___eval_phase__ico(vlSelf, vlCtx);
        } while (vlSelfRef.__VicoPhaseResult);
        __VnbaIterCount = 0U;
        do {
            if (VL_UNLIKELY(((0x00002710U < __VnbaIterCount)))) {
#ifdef VL_DEBUG
                Vrnd04___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/rnd04.v", 2, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vrnd04___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("../verilator_b200/csrc/designs/torture/rnd04.v", 2, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vrnd04___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vrnd04___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vrnd04___024root___ctor_var_reset(Vrnd04___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        VL_SCOPED_RAND_RESET_W(150, vlSelf->i0, __VscopeHash, 8471475882151657015ull);
        vlSelf->i1 = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 2165040029148572678ull);
        vlSelf->i2 = VL_SCOPED_RAND_RESET_Q(48, __VscopeHash, 12628911202397369737ull);
        vlSelf->i3 = VL_SCOPED_RAND_RESET_I(16, __VscopeHash, 6886176223853197501ull);
        VL_SCOPED_RAND_RESET_W(130, vlSelf->i4, __VscopeHash, 5103504462683158156ull);
        VL_SCOPED_RAND_RESET_W(128, vlSelf->i5, __VscopeHash, 10954576514065331241ull);
        vlSelf->o0 = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 4709259436872190324ull);
        vlSelf->o1 = VL_SCOPED_RAND_RESET_I(16, __VscopeHash, 16871796271195187576ull);
        vlSelf->o2 = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 6221000330064061286ull);
        vlSelf->o3 = VL_SCOPED_RAND_RESET_Q(40, __VscopeHash, 17472377812510545679ull);
        VL_SCOPED_RAND_RESET_W(96, vlSelf->o4, __VscopeHash, 12832999154864337794ull);
        VL_SCOPED_RAND_RESET_W(96, vlSelf->rnd04__DOT__r0, __VscopeHash, 1074958518393649346ull);
        vlSelf->rnd04__DOT__r1 = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 10214369410696512347ull);
        vlSelf->rnd04__DOT__r2 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 4680981040456695697ull);
        vlSelf->rnd04__DOT__r3 = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 4618886325969487856ull);
        for (int __Vi0 = 0; __Vi0 < 16; ++__Vi0) {
            vlSelf->rnd04__DOT__mem[__Vi0] = VL_SCOPED_RAND_RESET_I(24, __VscopeHash, 11410535578637631199ull);
        }
        vlSelf->rnd04__DOT__w1 = VL_SCOPED_RAND_RESET_Q(64, __VscopeHash, 18237956165772048314ull);
        vlSelf->rnd04__DOT__w2 = VL_SCOPED_RAND_RESET_I(2, __VscopeHash, 14912084142498577363ull);
        vlSelf->rnd04__DOT__w4 = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 1320292084247447348ull);
        vlSelf->rnd04__DOT__w6 = VL_SCOPED_RAND_RESET_I(2, __VscopeHash, 16026830252656010298ull);
        vlSelf->rnd04__DOT__w7 = VL_SCOPED_RAND_RESET_I(16, __VscopeHash, 3406617755771234909ull);
        vlSelf->rnd04__DOT__w8 = VL_SCOPED_RAND_RESET_Q(40, __VscopeHash, 16607627940230700726ull);
        VL_SCOPED_RAND_RESET_W(96, vlSelf->rnd04__DOT__w10, __VscopeHash, 8869122030822033676ull);
        vlSelf->__VdfgRegularize_h6e95ff9d_0_0 = 0;
        for (int __Vi0 = 0; __Vi0 < 1; ++__Vi0) {
            vlSelf->__VstlTriggered[__Vi0] = 0;
        }
        for (int __Vi0 = 0; __Vi0 < 2; ++__Vi0) {
            vlSelf->__VicoTriggered[__Vi0] = 0;
        }
        vlSelf->__Vtrigprevexpr___TOP__clk__0 = 0;
        VL_ZERO_RESET_W(150, vlSelf->__Vtrigprevexpr___TOP__i0__0);
        vlSelf->__Vtrigprevexpr___TOP__i1__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__i2__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__i3__0 = 0;
        VL_ZERO_RESET_W(130, vlSelf->__Vtrigprevexpr___TOP__i4__0);
        VL_ZERO_RESET_W(128, vlSelf->__Vtrigprevexpr___TOP__i5__0);
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
        Vrnd04___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vrnd04___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vrnd04___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vrnd04___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vrnd04___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vrnd04___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vrnd04__cuda::info() {
    typedef Vrnd04__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"i0", "i0", VLCU_MEMBER_IN, 4, 0, 5, 150, 149, 0, T::OFF_i0},
        {"i1", "i1", VLCU_MEMBER_IN, 4, 0, 1, 32, 31, 0, T::OFF_i1},
        {"i2", "i2", VLCU_MEMBER_IN, 8, 0, 1, 48, 47, 0, T::OFF_i2},
        {"i3", "i3", VLCU_MEMBER_IN, 2, 0, 1, 16, 15, 0, T::OFF_i3},
        {"i4", "i4", VLCU_MEMBER_IN, 4, 0, 5, 130, 129, 0, T::OFF_i4},
        {"i5", "i5", VLCU_MEMBER_IN, 4, 0, 4, 128, 127, 0, T::OFF_i5},
        {"o0", "o0", VLCU_MEMBER_OUT, 1, 0, 1, 8, 7, 0, T::OFF_o0},
        {"o1", "o1", VLCU_MEMBER_OUT, 2, 0, 1, 16, 15, 0, T::OFF_o1},
        {"o2", "o2", VLCU_MEMBER_OUT, 4, 0, 1, 32, 31, 0, T::OFF_o2},
        {"o3", "o3", VLCU_MEMBER_OUT, 8, 0, 1, 40, 39, 0, T::OFF_o3},
        {"o4", "o4", VLCU_MEMBER_OUT, 4, 0, 3, 96, 95, 0, T::OFF_o4},
        {"rnd04__DOT__r2", "rnd04.r2", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_rnd04__DOT__r2},
        {"rnd04__DOT__r3", "rnd04.r3", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF_rnd04__DOT__r3},
        {"rnd04__DOT__w2", "rnd04.w2", VLCU_MEMBER_STATE, 1, 0, 1, 2, 1, 0, T::OFF_rnd04__DOT__w2},
        {"rnd04__DOT__w4", "rnd04.w4", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF_rnd04__DOT__w4},
        {"rnd04__DOT__w6", "rnd04.w6", VLCU_MEMBER_STATE, 1, 0, 1, 2, 1, 0, T::OFF_rnd04__DOT__w6},
        {"__VdfgRegularize_h6e95ff9d_0_0", "__VdfgRegularize_h6e95ff9d_0_0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_0},
        {"__VstlFirstIteration", "__VstlFirstIteration", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlFirstIteration},
        {"__VstlPhaseResult", "__VstlPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__VicoDidInit", "__VicoDidInit", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VicoDidInit},
        {"__VicoPhaseResult", "__VicoPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VicoPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__1", "__Vtrigprevexpr___TOP__clk__1", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__1},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"rnd04__DOT__w7", "rnd04.w7", VLCU_MEMBER_STATE, 2, 0, 1, 16, 15, 0, T::OFF_rnd04__DOT__w7},
        {"__Vtrigprevexpr___TOP__i3__0", "__Vtrigprevexpr___TOP__i3__0", VLCU_MEMBER_STATE, 2, 0, 1, 16, 15, 0, T::OFF___Vtrigprevexpr___TOP__i3__0},
        {"rnd04__DOT__r0", "rnd04.r0", VLCU_MEMBER_STATE, 4, 0, 3, 96, 95, 0, T::OFF_rnd04__DOT__r0},
        {"rnd04__DOT__r1", "rnd04.r1", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_rnd04__DOT__r1},
        {"rnd04__DOT__w10", "rnd04.w10", VLCU_MEMBER_STATE, 4, 0, 3, 96, 95, 0, T::OFF_rnd04__DOT__w10},
        {"__Vtrigprevexpr___TOP__i0__0", "__Vtrigprevexpr___TOP__i0__0", VLCU_MEMBER_STATE, 4, 0, 5, 150, 149, 0, T::OFF___Vtrigprevexpr___TOP__i0__0},
        {"__Vtrigprevexpr___TOP__i1__0", "__Vtrigprevexpr___TOP__i1__0", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___Vtrigprevexpr___TOP__i1__0},
        {"__Vtrigprevexpr___TOP__i4__0", "__Vtrigprevexpr___TOP__i4__0", VLCU_MEMBER_STATE, 4, 0, 5, 130, 129, 0, T::OFF___Vtrigprevexpr___TOP__i4__0},
        {"__Vtrigprevexpr___TOP__i5__0", "__Vtrigprevexpr___TOP__i5__0", VLCU_MEMBER_STATE, 4, 0, 4, 128, 127, 0, T::OFF___Vtrigprevexpr___TOP__i5__0},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"rnd04__DOT__w1", "rnd04.w1", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF_rnd04__DOT__w1},
        {"rnd04__DOT__w8", "rnd04.w8", VLCU_MEMBER_STATE, 8, 0, 1, 40, 39, 0, T::OFF_rnd04__DOT__w8},
        {"__Vtrigprevexpr___TOP__i2__0", "__Vtrigprevexpr___TOP__i2__0", VLCU_MEMBER_STATE, 8, 0, 1, 48, 47, 0, T::OFF___Vtrigprevexpr___TOP__i2__0},
        {"rnd04__DOT__mem", "rnd04.mem", VLCU_MEMBER_STATE, 4, 0, 16, 24, 23, 0, T::OFF_rnd04__DOT__mem},
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
    static const uint32_t saveOrder[] = {0, 7, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 4, 8, 26, 27, 1, 2, 5, 6, 9, 11, 28, 29, 30, 31, 32, 33, 34, 35, 3, 10, 36, 37, 38, 39, 40, 41, 42, 43, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "rnd04", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 0, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 44, 0, 0xd189c098c351952cULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vrnd04__cuda VLCU_TESTBENCHES);
