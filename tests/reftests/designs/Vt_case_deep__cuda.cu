// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt_case_deep"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt_case_deep___024root final {
    CData/*0:0*/ clk;
    CData/*0:0*/ __VdfgRegularize_h6e95ff9d_0_17;
    CData/*0:0*/ __VdfgRegularize_h6e95ff9d_0_18;
    CData/*0:0*/ __VdfgRegularize_h6e95ff9d_0_19;
    CData/*0:0*/ __VdfgRegularize_h6e95ff9d_0_20;
    CData/*0:0*/ __VdfgRegularize_h6e95ff9d_0_21;
    CData/*0:0*/ __VdfgRegularize_h6e95ff9d_0_22;
    CData/*0:0*/ __VdfgRegularize_h6e95ff9d_0_23;
    CData/*0:0*/ __VdfgRegularize_h6e95ff9d_0_24;
    CData/*0:0*/ __VdfgRegularize_h6e95ff9d_0_25;
    CData/*0:0*/ __VdfgRegularize_h6e95ff9d_0_26;
    CData/*0:0*/ __VdfgRegularize_h6e95ff9d_0_27;
    CData/*0:0*/ __VdfgRegularize_h6e95ff9d_0_28;
    CData/*0:0*/ __VdfgRegularize_h6e95ff9d_0_29;
    CData/*0:0*/ __VdfgRegularize_h6e95ff9d_0_30;
    CData/*0:0*/ __VdfgRegularize_h6e95ff9d_0_31;
    CData/*0:0*/ __VdfgRegularize_h6e95ff9d_0_32;
    CData/*0:0*/ __VdfgRegularize_h6e95ff9d_0_33;
    CData/*0:0*/ __VdfgRegularize_h6e95ff9d_0_34;
    CData/*0:0*/ __VdfgRegularize_h6e95ff9d_0_35;
    CData/*0:0*/ __VdfgRegularize_h6e95ff9d_0_36;
    CData/*0:0*/ __VdfgRegularize_h6e95ff9d_0_37;
    CData/*0:0*/ __VdfgRegularize_h6e95ff9d_0_38;
    CData/*0:0*/ __VdfgRegularize_h6e95ff9d_0_39;
    CData/*0:0*/ __VdfgRegularize_h6e95ff9d_0_40;
    CData/*0:0*/ __VstlFirstIteration;
    CData/*0:0*/ __VstlPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    IData/*31:0*/ t__DOT__cyc;
    IData/*23:0*/ __VdfgRegularize_h6e95ff9d_0_0;
    IData/*23:0*/ __VdfgRegularize_h6e95ff9d_0_1;
    IData/*23:0*/ __VdfgRegularize_h6e95ff9d_0_2;
    IData/*23:0*/ __VdfgRegularize_h6e95ff9d_0_3;
    IData/*23:0*/ __VdfgRegularize_h6e95ff9d_0_4;
    IData/*23:0*/ __VdfgRegularize_h6e95ff9d_0_5;
    IData/*23:0*/ __VdfgRegularize_h6e95ff9d_0_6;
    IData/*23:0*/ __VdfgRegularize_h6e95ff9d_0_7;
    IData/*23:0*/ __VdfgRegularize_h6e95ff9d_0_8;
    IData/*23:0*/ __VdfgRegularize_h6e95ff9d_0_9;
    IData/*23:0*/ __VdfgRegularize_h6e95ff9d_0_10;
    IData/*23:0*/ __VdfgRegularize_h6e95ff9d_0_11;
    IData/*23:0*/ __VdfgRegularize_h6e95ff9d_0_12;
    IData/*23:0*/ __VdfgRegularize_h6e95ff9d_0_13;
    IData/*23:0*/ __VdfgRegularize_h6e95ff9d_0_14;
    IData/*23:0*/ __VdfgRegularize_h6e95ff9d_0_15;
    IData/*23:0*/ __VdfgRegularize_h6e95ff9d_0_16;
    IData/*31:0*/ __VactIterCount;
    QData/*63:0*/ t__DOT__crc;
    QData/*63:0*/ t__DOT__sum;
    VlUnpacked<QData/*63:0*/, 1> __VstlTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vt_case_deep__cuda final {
    typedef Vt_case_deep___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 116,
        OFF___VdfgRegularize_h6e95ff9d_0_17 = 117,
        OFF___VdfgRegularize_h6e95ff9d_0_18 = 118,
        OFF___VdfgRegularize_h6e95ff9d_0_19 = 119,
        OFF___VdfgRegularize_h6e95ff9d_0_20 = 120,
        OFF___VdfgRegularize_h6e95ff9d_0_21 = 121,
        OFF___VdfgRegularize_h6e95ff9d_0_22 = 122,
        OFF___VdfgRegularize_h6e95ff9d_0_23 = 123,
        OFF___VdfgRegularize_h6e95ff9d_0_24 = 124,
        OFF___VdfgRegularize_h6e95ff9d_0_25 = 125,
        OFF___VdfgRegularize_h6e95ff9d_0_26 = 126,
        OFF___VdfgRegularize_h6e95ff9d_0_27 = 127,
        OFF___VdfgRegularize_h6e95ff9d_0_28 = 128,
        OFF___VdfgRegularize_h6e95ff9d_0_29 = 129,
        OFF___VdfgRegularize_h6e95ff9d_0_30 = 130,
        OFF___VdfgRegularize_h6e95ff9d_0_31 = 131,
        OFF___VdfgRegularize_h6e95ff9d_0_32 = 132,
        OFF___VdfgRegularize_h6e95ff9d_0_33 = 133,
        OFF___VdfgRegularize_h6e95ff9d_0_34 = 134,
        OFF___VdfgRegularize_h6e95ff9d_0_35 = 135,
        OFF___VdfgRegularize_h6e95ff9d_0_36 = 136,
        OFF___VdfgRegularize_h6e95ff9d_0_37 = 137,
        OFF___VdfgRegularize_h6e95ff9d_0_38 = 138,
        OFF___VdfgRegularize_h6e95ff9d_0_39 = 139,
        OFF___VdfgRegularize_h6e95ff9d_0_40 = 140,
        OFF___VstlFirstIteration = 141,
        OFF___VstlPhaseResult = 142,
        OFF___Vtrigprevexpr___TOP__clk__0 = 143,
        OFF___VactPhaseResult = 144,
        OFF___VnbaPhaseResult = 145,
        OFF_t__DOT__cyc = 40,
        OFF___VdfgRegularize_h6e95ff9d_0_0 = 44,
        OFF___VdfgRegularize_h6e95ff9d_0_1 = 48,
        OFF___VdfgRegularize_h6e95ff9d_0_2 = 52,
        OFF___VdfgRegularize_h6e95ff9d_0_3 = 56,
        OFF___VdfgRegularize_h6e95ff9d_0_4 = 60,
        OFF___VdfgRegularize_h6e95ff9d_0_5 = 64,
        OFF___VdfgRegularize_h6e95ff9d_0_6 = 68,
        OFF___VdfgRegularize_h6e95ff9d_0_7 = 72,
        OFF___VdfgRegularize_h6e95ff9d_0_8 = 76,
        OFF___VdfgRegularize_h6e95ff9d_0_9 = 80,
        OFF___VdfgRegularize_h6e95ff9d_0_10 = 84,
        OFF___VdfgRegularize_h6e95ff9d_0_11 = 88,
        OFF___VdfgRegularize_h6e95ff9d_0_12 = 92,
        OFF___VdfgRegularize_h6e95ff9d_0_13 = 96,
        OFF___VdfgRegularize_h6e95ff9d_0_14 = 100,
        OFF___VdfgRegularize_h6e95ff9d_0_15 = 104,
        OFF___VdfgRegularize_h6e95ff9d_0_16 = 108,
        OFF___VactIterCount = 112,
        OFF_t__DOT__crc = 0,
        OFF_t__DOT__sum = 8,
        OFF___VstlTriggered = 16,
        OFF___VactTriggered = 24,
        OFF___VnbaTriggered = 32,
        kStateBytes = 152,  // S_persist
        kTouchedReadBytes = 136,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 135,  // S_state_wr
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
        s.__VdfgRegularize_h6e95ff9d_0_17 = st.ld<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_17, i);
        s.__VdfgRegularize_h6e95ff9d_0_18 = st.ld<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_18, i);
        s.__VdfgRegularize_h6e95ff9d_0_19 = st.ld<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_19, i);
        s.__VdfgRegularize_h6e95ff9d_0_20 = st.ld<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_20, i);
        s.__VdfgRegularize_h6e95ff9d_0_21 = st.ld<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_21, i);
        s.__VdfgRegularize_h6e95ff9d_0_22 = st.ld<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_22, i);
        s.__VdfgRegularize_h6e95ff9d_0_23 = st.ld<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_23, i);
        s.__VdfgRegularize_h6e95ff9d_0_24 = st.ld<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_24, i);
        s.__VdfgRegularize_h6e95ff9d_0_25 = st.ld<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_25, i);
        s.__VdfgRegularize_h6e95ff9d_0_26 = st.ld<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_26, i);
        s.__VdfgRegularize_h6e95ff9d_0_27 = st.ld<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_27, i);
        s.__VdfgRegularize_h6e95ff9d_0_28 = st.ld<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_28, i);
        s.__VdfgRegularize_h6e95ff9d_0_29 = st.ld<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_29, i);
        s.__VdfgRegularize_h6e95ff9d_0_30 = st.ld<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_30, i);
        s.__VdfgRegularize_h6e95ff9d_0_31 = st.ld<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_31, i);
        s.__VdfgRegularize_h6e95ff9d_0_32 = st.ld<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_32, i);
        s.__VdfgRegularize_h6e95ff9d_0_33 = st.ld<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_33, i);
        s.__VdfgRegularize_h6e95ff9d_0_34 = st.ld<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_34, i);
        s.__VdfgRegularize_h6e95ff9d_0_35 = st.ld<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_35, i);
        s.__VdfgRegularize_h6e95ff9d_0_36 = st.ld<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_36, i);
        s.__VdfgRegularize_h6e95ff9d_0_37 = st.ld<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_37, i);
        s.__VdfgRegularize_h6e95ff9d_0_38 = st.ld<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_38, i);
        s.__VdfgRegularize_h6e95ff9d_0_39 = st.ld<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_39, i);
        s.__VdfgRegularize_h6e95ff9d_0_40 = st.ld<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_40, i);
        s.__VstlFirstIteration = st.ld<VlcuMem8>(OFF___VstlFirstIteration, i);
        s.__VstlPhaseResult = st.ld<VlcuMem8>(OFF___VstlPhaseResult, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        s.t__DOT__cyc = st.ld<IData>(OFF_t__DOT__cyc, i);
        s.__VdfgRegularize_h6e95ff9d_0_0 = st.ld<IData>(OFF___VdfgRegularize_h6e95ff9d_0_0, i);
        s.__VdfgRegularize_h6e95ff9d_0_1 = st.ld<IData>(OFF___VdfgRegularize_h6e95ff9d_0_1, i);
        s.__VdfgRegularize_h6e95ff9d_0_2 = st.ld<IData>(OFF___VdfgRegularize_h6e95ff9d_0_2, i);
        s.__VdfgRegularize_h6e95ff9d_0_3 = st.ld<IData>(OFF___VdfgRegularize_h6e95ff9d_0_3, i);
        s.__VdfgRegularize_h6e95ff9d_0_4 = st.ld<IData>(OFF___VdfgRegularize_h6e95ff9d_0_4, i);
        s.__VdfgRegularize_h6e95ff9d_0_5 = st.ld<IData>(OFF___VdfgRegularize_h6e95ff9d_0_5, i);
        s.__VdfgRegularize_h6e95ff9d_0_6 = st.ld<IData>(OFF___VdfgRegularize_h6e95ff9d_0_6, i);
        s.__VdfgRegularize_h6e95ff9d_0_7 = st.ld<IData>(OFF___VdfgRegularize_h6e95ff9d_0_7, i);
        s.__VdfgRegularize_h6e95ff9d_0_8 = st.ld<IData>(OFF___VdfgRegularize_h6e95ff9d_0_8, i);
        s.__VdfgRegularize_h6e95ff9d_0_9 = st.ld<IData>(OFF___VdfgRegularize_h6e95ff9d_0_9, i);
        s.__VdfgRegularize_h6e95ff9d_0_10 = st.ld<IData>(OFF___VdfgRegularize_h6e95ff9d_0_10, i);
        s.__VdfgRegularize_h6e95ff9d_0_11 = st.ld<IData>(OFF___VdfgRegularize_h6e95ff9d_0_11, i);
        s.__VdfgRegularize_h6e95ff9d_0_12 = st.ld<IData>(OFF___VdfgRegularize_h6e95ff9d_0_12, i);
        s.__VdfgRegularize_h6e95ff9d_0_13 = st.ld<IData>(OFF___VdfgRegularize_h6e95ff9d_0_13, i);
        s.__VdfgRegularize_h6e95ff9d_0_14 = st.ld<IData>(OFF___VdfgRegularize_h6e95ff9d_0_14, i);
        s.__VdfgRegularize_h6e95ff9d_0_15 = st.ld<IData>(OFF___VdfgRegularize_h6e95ff9d_0_15, i);
        s.__VdfgRegularize_h6e95ff9d_0_16 = st.ld<IData>(OFF___VdfgRegularize_h6e95ff9d_0_16, i);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        s.t__DOT__crc = st.ld<QData>(OFF_t__DOT__crc, i);
        s.t__DOT__sum = st.ld<QData>(OFF_t__DOT__sum, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VstlTriggered[e] = st.ldE<QData>(OFF___VstlTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.st<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_17, i, s.__VdfgRegularize_h6e95ff9d_0_17);
        st.st<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_18, i, s.__VdfgRegularize_h6e95ff9d_0_18);
        st.st<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_19, i, s.__VdfgRegularize_h6e95ff9d_0_19);
        st.st<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_20, i, s.__VdfgRegularize_h6e95ff9d_0_20);
        st.st<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_21, i, s.__VdfgRegularize_h6e95ff9d_0_21);
        st.st<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_22, i, s.__VdfgRegularize_h6e95ff9d_0_22);
        st.st<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_23, i, s.__VdfgRegularize_h6e95ff9d_0_23);
        st.st<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_24, i, s.__VdfgRegularize_h6e95ff9d_0_24);
        st.st<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_25, i, s.__VdfgRegularize_h6e95ff9d_0_25);
        st.st<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_26, i, s.__VdfgRegularize_h6e95ff9d_0_26);
        st.st<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_27, i, s.__VdfgRegularize_h6e95ff9d_0_27);
        st.st<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_28, i, s.__VdfgRegularize_h6e95ff9d_0_28);
        st.st<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_29, i, s.__VdfgRegularize_h6e95ff9d_0_29);
        st.st<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_30, i, s.__VdfgRegularize_h6e95ff9d_0_30);
        st.st<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_31, i, s.__VdfgRegularize_h6e95ff9d_0_31);
        st.st<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_32, i, s.__VdfgRegularize_h6e95ff9d_0_32);
        st.st<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_33, i, s.__VdfgRegularize_h6e95ff9d_0_33);
        st.st<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_34, i, s.__VdfgRegularize_h6e95ff9d_0_34);
        st.st<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_35, i, s.__VdfgRegularize_h6e95ff9d_0_35);
        st.st<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_36, i, s.__VdfgRegularize_h6e95ff9d_0_36);
        st.st<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_37, i, s.__VdfgRegularize_h6e95ff9d_0_37);
        st.st<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_38, i, s.__VdfgRegularize_h6e95ff9d_0_38);
        st.st<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_39, i, s.__VdfgRegularize_h6e95ff9d_0_39);
        st.st<VlcuMem8>(OFF___VdfgRegularize_h6e95ff9d_0_40, i, s.__VdfgRegularize_h6e95ff9d_0_40);
        st.st<VlcuMem8>(OFF___VstlFirstIteration, i, s.__VstlFirstIteration);
        st.st<VlcuMem8>(OFF___VstlPhaseResult, i, s.__VstlPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.st<IData>(OFF_t__DOT__cyc, i, s.t__DOT__cyc);
        st.st<IData>(OFF___VdfgRegularize_h6e95ff9d_0_0, i, s.__VdfgRegularize_h6e95ff9d_0_0);
        st.st<IData>(OFF___VdfgRegularize_h6e95ff9d_0_1, i, s.__VdfgRegularize_h6e95ff9d_0_1);
        st.st<IData>(OFF___VdfgRegularize_h6e95ff9d_0_2, i, s.__VdfgRegularize_h6e95ff9d_0_2);
        st.st<IData>(OFF___VdfgRegularize_h6e95ff9d_0_3, i, s.__VdfgRegularize_h6e95ff9d_0_3);
        st.st<IData>(OFF___VdfgRegularize_h6e95ff9d_0_4, i, s.__VdfgRegularize_h6e95ff9d_0_4);
        st.st<IData>(OFF___VdfgRegularize_h6e95ff9d_0_5, i, s.__VdfgRegularize_h6e95ff9d_0_5);
        st.st<IData>(OFF___VdfgRegularize_h6e95ff9d_0_6, i, s.__VdfgRegularize_h6e95ff9d_0_6);
        st.st<IData>(OFF___VdfgRegularize_h6e95ff9d_0_7, i, s.__VdfgRegularize_h6e95ff9d_0_7);
        st.st<IData>(OFF___VdfgRegularize_h6e95ff9d_0_8, i, s.__VdfgRegularize_h6e95ff9d_0_8);
        st.st<IData>(OFF___VdfgRegularize_h6e95ff9d_0_9, i, s.__VdfgRegularize_h6e95ff9d_0_9);
        st.st<IData>(OFF___VdfgRegularize_h6e95ff9d_0_10, i, s.__VdfgRegularize_h6e95ff9d_0_10);
        st.st<IData>(OFF___VdfgRegularize_h6e95ff9d_0_11, i, s.__VdfgRegularize_h6e95ff9d_0_11);
        st.st<IData>(OFF___VdfgRegularize_h6e95ff9d_0_12, i, s.__VdfgRegularize_h6e95ff9d_0_12);
        st.st<IData>(OFF___VdfgRegularize_h6e95ff9d_0_13, i, s.__VdfgRegularize_h6e95ff9d_0_13);
        st.st<IData>(OFF___VdfgRegularize_h6e95ff9d_0_14, i, s.__VdfgRegularize_h6e95ff9d_0_14);
        st.st<IData>(OFF___VdfgRegularize_h6e95ff9d_0_15, i, s.__VdfgRegularize_h6e95ff9d_0_15);
        st.st<IData>(OFF___VdfgRegularize_h6e95ff9d_0_16, i, s.__VdfgRegularize_h6e95ff9d_0_16);
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

    VLCU_D static void Vt_case_deep___024root___eval_static(Vt_case_deep___024root* vlSelf, VlcuContext& vlCtx) {
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

    VLCU_D static void Vt_case_deep___024root___eval_initial(Vt_case_deep___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_case_deep___024root___eval_final(Vt_case_deep___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_case_deep___024root___eval_settle(Vt_case_deep___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt_case_deep___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
#endif
                VL_FATAL_MT("/root/reference/test_regress/t/t_case_deep.v", 7, "", "DIDNOTCONVERGE: Settle region did not converge after '--converge-limit' of 10000 tries");
            }
            __VstlIterCount = ((IData)(1U) + __VstlIterCount);
            vlSelfRef.__VstlPhaseResult = Vt_case_deep___024root___eval_phase__stl(vlSelf, vlCtx);
            vlSelfRef.__VstlFirstIteration = 0U;
        } while (vlSelfRef.__VstlPhaseResult);
    }

    VLCU_D static bool Vt_case_deep___024root___trigger_anySet__stl(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt_case_deep___024root___stl_sequent__TOP__0(Vt_case_deep___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0 = 
            (0x00fffff0U & (IData)((vlSelfRef.t__DOT__crc 
                                    >> 8U)));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1 = 
            (0x00ffffe0U & (IData)((vlSelfRef.t__DOT__crc 
                                    >> 8U)));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2 = 
            (0x00ffffc0U & (IData)((vlSelfRef.t__DOT__crc 
                                    >> 8U)));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_3 = 
            (0x00ffff00U & (IData)((vlSelfRef.t__DOT__crc 
                                    >> 8U)));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_4 = 
            (0x00ffff80U & (IData)((vlSelfRef.t__DOT__crc 
                                    >> 8U)));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_5 = 
            (0x00fffc00U & (IData)((vlSelfRef.t__DOT__crc 
                                    >> 8U)));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_6 = 
            (0x00fffe00U & (IData)((vlSelfRef.t__DOT__crc 
                                    >> 8U)));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_7 = 
            (0x00ffe000U & (IData)((vlSelfRef.t__DOT__crc 
                                    >> 8U)));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_8 = 
            (0x00fff800U & (IData)((vlSelfRef.t__DOT__crc 
                                    >> 8U)));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_9 = 
            (0x00ffc000U & (IData)((vlSelfRef.t__DOT__crc 
                                    >> 8U)));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_10 = 
            (0x00fff000U & (IData)((vlSelfRef.t__DOT__crc 
                                    >> 8U)));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_11 = 
            (0x00f00000U & (IData)((vlSelfRef.t__DOT__crc 
                                    >> 8U)));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_12 = 
            (0x00fc0000U & (IData)((vlSelfRef.t__DOT__crc 
                                    >> 8U)));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_13 = 
            (0x00fe0000U & (IData)((vlSelfRef.t__DOT__crc 
                                    >> 8U)));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_14 = 
            (0x00ff0000U & (IData)((vlSelfRef.t__DOT__crc 
                                    >> 8U)));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_15 = 
            (0x00ff8000U & (IData)((vlSelfRef.t__DOT__crc 
                                    >> 8U)));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_16 = 
            (0x00f80000U & (IData)((vlSelfRef.t__DOT__crc 
                                    >> 8U)));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_17 = 
            (2U == (2U & (IData)((vlSelfRef.t__DOT__crc 
                                  >> 0x00000020U))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_18 = 
            ((0x0024f8d0U == (0x00fffff0U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x0079a6b0U == (0x00fffff0U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | ((0x00510110U == (0x00fffff0U & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                   | ((0x00f5fee0U == (0x00fffff0U 
                                       & (IData)((vlSelfRef.t__DOT__crc 
                                                  >> 8U)))) 
                      | ((0x0082fb30U == (0x00fffff0U 
                                          & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
                         | ((0x00047a70U == (0x00fffff0U 
                                             & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                            | ((0x00fd3ec0U == (0x00fffff0U 
                                                & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                               | (0x002b8f00U == (0x00fffff0U 
                                                  & (IData)(
                                                            (vlSelfRef.t__DOT__crc 
                                                             >> 8U)))))))))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_19 = 
            ((0x00f0aef0U == (0x00fffff0U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x00249aa0U == (0x00fffff0U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | ((0x00f16c30U == (0x00fffff0U & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                   | ((0x00572330U == (0x00fffff0U 
                                       & (IData)((vlSelfRef.t__DOT__crc 
                                                  >> 8U)))) 
                      | ((0x00781150U == (0x00fffff0U 
                                          & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
                         | ((0x0025acd0U == (0x00fffff0U 
                                             & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                            | ((0x00b2c390U == (0x00fffff0U 
                                                & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                               | (0x00b0a830U == (0x00fffff0U 
                                                  & (IData)(
                                                            (vlSelfRef.t__DOT__crc 
                                                             >> 8U)))))))))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_20 = 
            ((0x00506fe0U == (0x00fffff0U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x00070820U == (0x00fffff0U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | ((0x002a0c80U == (0x00fffff0U & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                   | ((0x004e52e0U == (0x00fffff0U 
                                       & (IData)((vlSelfRef.t__DOT__crc 
                                                  >> 8U)))) 
                      | ((0x00f00800U == (0x00fffff0U 
                                          & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
                         | ((0x004d6790U == (0x00fffff0U 
                                             & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                            | ((0x00f71d50U == (0x00fffff0U 
                                                & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                               | (0x0079a760U == (0x00fffff0U 
                                                  & (IData)(
                                                            (vlSelfRef.t__DOT__crc 
                                                             >> 8U)))))))))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_21 = 
            ((0x00521d60U == (0x00fffff0U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x00eaaef0U == (0x00fffff0U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | ((0x00fe4e30U == (0x00fffff0U & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                   | ((0x00573ef0U == (0x00fffff0U 
                                       & (IData)((vlSelfRef.t__DOT__crc 
                                                  >> 8U)))) 
                      | ((0x00b1ffd0U == (0x00fffff0U 
                                          & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
                         | ((0x00b13300U == (0x00fffff0U 
                                             & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                            | ((0x00f463d0U == (0x00fffff0U 
                                                & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                               | (0x002bba30U == (0x00fffff0U 
                                                  & (IData)(
                                                            (vlSelfRef.t__DOT__crc 
                                                             >> 8U)))))))))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_22 = 
            ((0x00b19770U == (0x00fffff0U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x00bcde80U == (0x00fffff0U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | ((0x00515d00U == (0x00fffff0U & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                   | ((0x0054f8e0U == (0x00fffff0U 
                                       & (IData)((vlSelfRef.t__DOT__crc 
                                                  >> 8U)))) 
                      | ((0x00fe6670U == (0x00fffff0U 
                                          & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
                         | ((0x00f7f510U == (0x00fffff0U 
                                             & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                            | ((0x00b05e00U == (0x00fffff0U 
                                                & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                               | (0x004f8450U == (0x00fffff0U 
                                                  & (IData)(
                                                            (vlSelfRef.t__DOT__crc 
                                                             >> 8U)))))))))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_23 = 
            ((0x0056f100U == (0x00fffff0U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x00b7cc90U == (0x00fffff0U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | ((0x00fdb250U == (0x00fffff0U & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                   | ((0x00b05210U == (0x00fffff0U 
                                       & (IData)((vlSelfRef.t__DOT__crc 
                                                  >> 8U)))) 
                      | ((0x00b26b60U == (0x00fffff0U 
                                          & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
                         | ((0x00794180U == (0x00fffff0U 
                                             & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                            | ((0x00f62db0U == (0x00fffff0U 
                                                & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                               | (0x005090b0U == (0x00fffff0U 
                                                  & (IData)(
                                                            (vlSelfRef.t__DOT__crc 
                                                             >> 8U)))))))))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_24 = 
            ((0x0022f6d0U == (0x00fffff0U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x0082aad0U == (0x00fffff0U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | (0x00b294d0U == (0x00fffff0U & (IData)(
                                                         (vlSelfRef.t__DOT__crc 
                                                          >> 8U))))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_25 = 
            ((0x0082c680U == (0x00ffffe0U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x0025cec0U == (0x00ffffe0U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | ((0x005641a0U == (0x00ffffe0U & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                   | ((0x00509560U == (0x00ffffe0U 
                                       & (IData)((vlSelfRef.t__DOT__crc 
                                                  >> 8U)))) 
                      | ((0x00f7e9a0U == (0x00ffffe0U 
                                          & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
                         | ((0x00b22ac0U == (0x00ffffe0U 
                                             & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                            | ((0x00b04420U == (0x00ffffe0U 
                                                & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                               | (0x00eb1320U == (0x00ffffe0U 
                                                  & (IData)(
                                                            (vlSelfRef.t__DOT__crc 
                                                             >> 8U)))))))))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_26 = 
            ((0x00b7aa40U == (0x00ffffc0U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x00f72f80U == (0x00ffffc0U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | ((0x00523400U == (0x00ffffc0U & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                   | ((0x0024f640U == (0x00ffffc0U 
                                       & (IData)((vlSelfRef.t__DOT__crc 
                                                  >> 8U)))) 
                      | ((0x00532240U == (0x00ffffc0U 
                                          & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
                         | ((0x00eb0e00U == (0x00ffffc0U 
                                             & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                            | ((0x00eb3140U == (0x00ffffc0U 
                                                & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                               | (0x00508e40U == (0x00ffffc0U 
                                                  & (IData)(
                                                            (vlSelfRef.t__DOT__crc 
                                                             >> 8U)))))))))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_27 = 
            ((0x004f5200U == (0x00ffffc0U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x00eae480U == (0x00ffffc0U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | ((0x00249e00U == (0x00ffffc0U & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                   | (0x00bc6f40U == (0x00ffffc0U & (IData)(
                                                            (vlSelfRef.t__DOT__crc 
                                                             >> 8U)))))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_28 = 
            ((0x0022c600U == (0x00ffff80U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x0022b500U == (0x00ffff80U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | ((0x00fef000U == (0x00ffff80U & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                   | ((0x0020e800U == (0x00ffff80U 
                                       & (IData)((vlSelfRef.t__DOT__crc 
                                                  >> 8U)))) 
                      | ((0x00277480U == (0x00ffff80U 
                                          & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
                         | ((0x00fee180U == (0x00ffff80U 
                                             & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                            | (0x00f1a200U == (0x00ffff80U 
                                               & (IData)(
                                                         (vlSelfRef.t__DOT__crc 
                                                          >> 8U))))))))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_29 = 
            ((0x00f6b200U == (0x00ffff00U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x00246400U == (0x00ffff00U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | ((0x00577400U == (0x00ffff00U & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                   | ((0x00568f00U == (0x00ffff00U 
                                       & (IData)((vlSelfRef.t__DOT__crc 
                                                  >> 8U)))) 
                      | ((0x00219f00U == (0x00ffff00U 
                                          & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
                         | ((0x00548500U == (0x00ffff00U 
                                             & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                            | ((0x00eac000U == (0x00ffff00U 
                                                & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                               | (0x00f03200U == (0x00ffff00U 
                                                  & (IData)(
                                                            (vlSelfRef.t__DOT__crc 
                                                             >> 8U)))))))))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_30 = 
            ((0x00240500U == (0x00ffff00U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | (0x0024a900U == (0x00ffff00U & (IData)(
                                                      (vlSelfRef.t__DOT__crc 
                                                       >> 8U)))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_31 = 
            ((0x00bd9600U == (0x00fffe00U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x00bd1800U == (0x00fffe00U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | ((0x0083ce00U == (0x00fffe00U & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                   | (0x002ab000U == (0x00fffe00U & (IData)(
                                                            (vlSelfRef.t__DOT__crc 
                                                             >> 8U)))))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_32 = 
            ((0x00548c00U == (0x00fffc00U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x00b0c800U == (0x00fffc00U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | ((0x00bd3c00U == (0x00fffc00U & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                   | ((0x00b05400U == (0x00fffc00U 
                                       & (IData)((vlSelfRef.t__DOT__crc 
                                                  >> 8U)))) 
                      | ((0x002b3c00U == (0x00fffc00U 
                                          & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
                         | (0x0020b000U == (0x00fffc00U 
                                            & (IData)(
                                                      (vlSelfRef.t__DOT__crc 
                                                       >> 8U)))))))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_33 = 
            ((0x00ebe800U == (0x00fff800U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x004d7000U == (0x00fff800U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | ((0x00fed800U == (0x00fff800U & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                   | ((0x0056c800U == (0x00fff800U 
                                       & (IData)((vlSelfRef.t__DOT__crc 
                                                  >> 8U)))) 
                      | ((0x00216000U == (0x00fff800U 
                                          & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
                         | ((0x00ff6800U == (0x00fff800U 
                                             & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                            | (0x00bcb000U == (0x00fff800U 
                                               & (IData)(
                                                         (vlSelfRef.t__DOT__crc 
                                                          >> 8U))))))))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_34 = 
            ((0x00fec000U == (0x00fff000U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x0083e000U == (0x00fff000U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | ((0x00832000U == (0x00fff000U & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                   | ((0x00839000U == (0x00fff000U 
                                       & (IData)((vlSelfRef.t__DOT__crc 
                                                  >> 8U)))) 
                      | ((0x0052f000U == (0x00fff000U 
                                          & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
                         | ((0x0020c000U == (0x00fff000U 
                                             & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                            | (0x0007f000U == (0x00fff000U 
                                               & (IData)(
                                                         (vlSelfRef.t__DOT__crc 
                                                          >> 8U))))))))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_35 = 
            ((0x00b74000U == (0x00ffe000U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x00f06000U == (0x00ffe000U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | ((0x004f6000U == (0x00ffe000U & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                   | ((0x002b6000U == (0x00ffe000U 
                                       & (IData)((vlSelfRef.t__DOT__crc 
                                                  >> 8U)))) 
                      | ((0x00528000U == (0x00ffe000U 
                                          & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
                         | ((0x00f78000U == (0x00ffe000U 
                                             & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                            | ((0x00272000U == (0x00ffe000U 
                                                & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                               | (0x00b14000U == (0x00ffe000U 
                                                  & (IData)(
                                                            (vlSelfRef.t__DOT__crc 
                                                             >> 8U)))))))))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_36 = 
            ((0x00f54000U == (0x00ffc000U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x00220000U == (0x00ffc000U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | ((0x00534000U == (0x00ffc000U & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                   | ((0x00250000U == (0x00ffc000U 
                                       & (IData)((vlSelfRef.t__DOT__crc 
                                                  >> 8U)))) 
                      | ((0x00b38000U == (0x00ffc000U 
                                          & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
                         | ((0x00f0c000U == (0x00ffc000U 
                                             & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                            | ((0x00500000U == (0x00ffc000U 
                                                & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                               | (0x00fc0000U == (0x00ffc000U 
                                                  & (IData)(
                                                            (vlSelfRef.t__DOT__crc 
                                                             >> 8U)))))))))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_37 = 
            ((0x004e8000U == (0x00ff8000U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x00200000U == (0x00ff8000U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | ((0x00578000U == (0x00ff8000U & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                   | ((0x00550000U == (0x00ff8000U 
                                       & (IData)((vlSelfRef.t__DOT__crc 
                                                  >> 8U)))) 
                      | ((0x004d8000U == (0x00ff8000U 
                                          & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
                         | ((0x00518000U == (0x00ff8000U 
                                             & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                            | (0x00540000U == (0x00ff8000U 
                                               & (IData)(
                                                         (vlSelfRef.t__DOT__crc 
                                                          >> 8U))))))))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_38 = 
            ((0x007c0000U == (0x00ff0000U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x00060000U == (0x00ff0000U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | ((0x00050000U == (0x00ff0000U & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                   | ((0x004c0000U == (0x00ff0000U 
                                       & (IData)((vlSelfRef.t__DOT__crc 
                                                  >> 8U)))) 
                      | ((0x00b60000U == (0x00ff0000U 
                                          & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
                         | (0x00260000U == (0x00ff0000U 
                                            & (IData)(
                                                      (vlSelfRef.t__DOT__crc 
                                                       >> 8U)))))))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_39 = 
            ((0x007a0000U == (0x00fe0000U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x00280000U == (0x00fe0000U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | ((0x007e0000U == (0x00fe0000U & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                   | ((0x00b40000U == (0x00fe0000U 
                                       & (IData)((vlSelfRef.t__DOT__crc 
                                                  >> 8U)))) 
                      | (0x00800000U == (0x00fe0000U 
                                         & (IData)(
                                                   (vlSelfRef.t__DOT__crc 
                                                    >> 8U))))))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_40 = 
            ((0x00840000U == (0x00fc0000U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x00f80000U == (0x00fc0000U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | (0x00480000U == (0x00fc0000U & (IData)(
                                                         (vlSelfRef.t__DOT__crc 
                                                          >> 8U))))));
    }

    VLCU_D static bool Vt_case_deep___024root___eval_phase__stl(Vt_case_deep___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt_case_deep___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
        }
#endif
        __VstlExecute = Vt_case_deep___024root___trigger_anySet__stl(vlSelfRef.__VstlTriggered);
        if (__VstlExecute) {
            {
                // Inlined CFunc: _eval_stl
                if ((1ULL & vlSelfRef.__VstlTriggered[0U])) {
                    Vt_case_deep___024root___stl_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
        }
        return (__VstlExecute);
    }

    VLCU_D static bool Vt_case_deep___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt_case_deep___024root___nba_sequent__TOP__0(Vt_case_deep___024root* vlSelf, VlcuContext& vlCtx) {
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
        IData/*31:0*/ __Vdeeptemp_h5279618a__0;
        // Body
        __Vdly__t__DOT__cyc = vlSelfRef.t__DOT__cyc;
        __Vdly__t__DOT__sum = vlSelfRef.t__DOT__sum;
        __Vdly__t__DOT__crc = vlSelfRef.t__DOT__crc;
        __Vdly__t__DOT__cyc = ((IData)(1U) + vlSelfRef.t__DOT__cyc);
        __Vdly__t__DOT__crc = ((vlSelfRef.t__DOT__crc 
                                << 1U) | (QData)((IData)(
                                                         (1U 
                                                          & VL_REDXOR_64(
                                                                         (0x8000000000000005ULL 
                                                                          & vlSelfRef.t__DOT__crc))))));
        __Vdeeptemp_h5279618a__0 = ((0x00bd1800U == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_6)
                                     ? 0x00000056U : 
                                    ((0x0083ce00U == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_6)
                                      ? 0x00000057U
                                      : ((0x002ab000U 
                                          == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_6)
                                          ? 0x00000058U
                                          : ((0x00fe4200U 
                                              == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_6)
                                              ? 0x00000059U
                                              : ((0x00258400U 
                                                  == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_6)
                                                  ? 0x0000005aU
                                                  : 
                                                 ((0x00799600U 
                                                   == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_6)
                                                   ? 0x0000005bU
                                                   : 
                                                  ((0x00fff400U 
                                                    == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_6)
                                                    ? 0x0000005cU
                                                    : 
                                                   ((0x00bca600U 
                                                     == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_6)
                                                     ? 0x0000005dU
                                                     : 
                                                    ((0x00830e00U 
                                                      == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_6)
                                                      ? 0x0000005eU
                                                      : 
                                                     ((0x00240500U 
                                                       == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_3)
                                                       ? 0x0000005fU
                                                       : 
                                                      ((0x0024a900U 
                                                        == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_3)
                                                        ? 0x00000060U
                                                        : 
                                                       ((0x00f6b200U 
                                                         == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_3)
                                                         ? 0x00000061U
                                                         : 
                                                        ((0x00246400U 
                                                          == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_3)
                                                          ? 0x00000062U
                                                          : 
                                                         ((0x00577400U 
                                                           == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_3)
                                                           ? 0x00000063U
                                                           : 
                                                          ((0x00568f00U 
                                                            == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_3)
                                                            ? 0x00000064U
                                                            : 
                                                           ((0x00219f00U 
                                                             == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_3)
                                                             ? 0x00000065U
                                                             : 
                                                            ((0x00548500U 
                                                              == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_3)
                                                              ? 0x00000066U
                                                              : 
                                                             ((0x00eac000U 
                                                               == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_3)
                                                               ? 0x00000067U
                                                               : 
                                                              ((0x00f03200U 
                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_3)
                                                                ? 0x00000068U
                                                                : 
                                                               ((0x00788d00U 
                                                                 == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_3)
                                                                 ? 0x00000069U
                                                                 : 
                                                                ((0x0022c600U 
                                                                  == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_4)
                                                                  ? 0x0000006aU
                                                                  : 
                                                                 ((0x0022b500U 
                                                                   == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_4)
                                                                   ? 0x0000006bU
                                                                   : 
                                                                  ((0x00fef000U 
                                                                    == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_4)
                                                                    ? 0x0000006cU
                                                                    : 
                                                                   ((0x0020e800U 
                                                                     == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_4)
                                                                     ? 0x0000006dU
                                                                     : 
                                                                    ((0x00277480U 
                                                                      == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_4)
                                                                      ? 0x0000006eU
                                                                      : 
                                                                     ((0x00fee180U 
                                                                       == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_4)
                                                                       ? 0x0000006fU
                                                                       : 
                                                                      ((0x00f1a200U 
                                                                        == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_4)
                                                                        ? 0x00000070U
                                                                        : 
                                                                       ((0x00ebae80U 
                                                                         == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_4)
                                                                         ? 0x00000071U
                                                                         : 
                                                                        ((0x00509580U 
                                                                          == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_4)
                                                                          ? 0x00000072U
                                                                          : 
                                                                         ((0x0024cc80U 
                                                                           == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_4)
                                                                           ? 0x00000073U
                                                                           : 
                                                                          ((0x004e5400U 
                                                                            == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_4)
                                                                            ? 0x00000074U
                                                                            : 
                                                                           ((0x004f5200U 
                                                                             == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2)
                                                                             ? 0x00000075U
                                                                             : 
                                                                            ((0x00eae480U 
                                                                              == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2)
                                                                              ? 0x00000076U
                                                                              : 
                                                                             ((0x00249e00U 
                                                                               == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2)
                                                                               ? 0x00000077U
                                                                               : 
                                                                              ((0x00bc6f40U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2)
                                                                                ? 0x00000078U
                                                                                : 
                                                                               ((0x00b7aa40U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2)
                                                                                 ? 0x00000079U
                                                                                 : 
                                                                                ((0x00f72f80U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2)
                                                                                 ? 0x0000007aU
                                                                                 : 
                                                                                ((0x00523400U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2)
                                                                                 ? 0x0000007bU
                                                                                 : 
                                                                                ((0x0024f640U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2)
                                                                                 ? 0x0000007cU
                                                                                 : 
                                                                                ((0x00532240U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2)
                                                                                 ? 0x0000007dU
                                                                                 : 
                                                                                ((0x00eb0e00U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2)
                                                                                 ? 0x0000007eU
                                                                                 : 
                                                                                ((0x00eb3140U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2)
                                                                                 ? 0x0000007fU
                                                                                 : 
                                                                                ((0x00508e40U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2)
                                                                                 ? 0x00000080U
                                                                                 : 
                                                                                ((0x0082c680U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1)
                                                                                 ? 0x00000081U
                                                                                 : 
                                                                                ((0x0025cec0U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1)
                                                                                 ? 0x00000082U
                                                                                 : 
                                                                                ((0x005641a0U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1)
                                                                                 ? 0x00000083U
                                                                                 : 
                                                                                ((0x00509560U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1)
                                                                                 ? 0x00000084U
                                                                                 : 
                                                                                ((0x00f7e9a0U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1)
                                                                                 ? 0x00000085U
                                                                                 : 
                                                                                ((0x00b22ac0U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1)
                                                                                 ? 0x00000086U
                                                                                 : 
                                                                                ((0x00b04420U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1)
                                                                                 ? 0x00000087U
                                                                                 : 
                                                                                ((0x00eb1320U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1)
                                                                                 ? 0x00000088U
                                                                                 : 
                                                                                ((0x0025d7c0U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1)
                                                                                 ? 0x00000089U
                                                                                 : 
                                                                                ((0x00246780U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1)
                                                                                 ? 0x0000008aU
                                                                                 : 
                                                                                ((0x00b2e0a0U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1)
                                                                                 ? 0x0000008bU
                                                                                 : 
                                                                                ((0x00b114a0U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1)
                                                                                 ? 0x0000008cU
                                                                                 : 
                                                                                ((0x00fe4760U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1)
                                                                                 ? 0x0000008dU
                                                                                 : 
                                                                                ((0x0022f6d0U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x0000008eU
                                                                                 : 
                                                                                ((0x0082aad0U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x0000008fU
                                                                                 : 
                                                                                ((0x00b294d0U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x00000090U
                                                                                 : 
                                                                                ((0x0056f100U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x00000091U
                                                                                 : 
                                                                                ((0x00b7cc90U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x00000092U
                                                                                 : 
                                                                                ((0x00fdb250U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x00000093U
                                                                                 : 
                                                                                ((0x00b05210U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x00000094U
                                                                                 : 
                                                                                ((0x00b26b60U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x00000095U
                                                                                 : 
                                                                                ((0x00794180U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x00000096U
                                                                                 : 
                                                                                ((0x00f62db0U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x00000097U
                                                                                 : 
                                                                                ((0x005090b0U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x00000098U
                                                                                 : 
                                                                                ((0x00b19770U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x00000099U
                                                                                 : 
                                                                                ((0x00bcde80U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x0000009aU
                                                                                 : 
                                                                                ((0x00515d00U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x0000009bU
                                                                                 : 
                                                                                ((0x0054f8e0U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x0000009cU
                                                                                 : 
                                                                                ((0x00fe6670U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x0000009dU
                                                                                 : 
                                                                                ((0x00f7f510U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x0000009eU
                                                                                 : 
                                                                                ((0x00b05e00U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x0000009fU
                                                                                 : 
                                                                                ((0x004f8450U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x000000a0U
                                                                                 : 
                                                                                ((0x00521d60U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x000000a1U
                                                                                 : 
                                                                                ((0x00eaaef0U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x000000a2U
                                                                                 : 
                                                                                ((0x00fe4e30U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x000000a3U
                                                                                 : 
                                                                                ((0x00573ef0U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x000000a4U
                                                                                 : 
                                                                                ((0x00b1ffd0U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x000000a5U
                                                                                 : 
                                                                                ((0x00b13300U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x000000a6U
                                                                                 : 
                                                                                ((0x00f463d0U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x000000a7U
                                                                                 : 
                                                                                ((0x002bba30U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x000000a8U
                                                                                 : 
                                                                                ((0x00506fe0U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x000000a9U
                                                                                 : 
                                                                                ((0x00070820U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x000000aaU
                                                                                 : 
                                                                                ((0x002a0c80U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x000000abU
                                                                                 : 
                                                                                ((0x004e52e0U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x000000acU
                                                                                 : 
                                                                                ((0x00f00800U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x000000adU
                                                                                 : 
                                                                                ((0x004d6790U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x000000aeU
                                                                                 : 
                                                                                ((0x00f71d50U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x000000afU
                                                                                 : 
                                                                                ((0x0079a760U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x000000b0U
                                                                                 : 
                                                                                ((0x00f0aef0U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x000000b1U
                                                                                 : 
                                                                                ((0x00249aa0U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x000000b2U
                                                                                 : 
                                                                                ((0x00f16c30U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x000000b3U
                                                                                 : 
                                                                                ((0x00572330U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x000000b4U
                                                                                 : 
                                                                                ((0x00781150U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x000000b5U
                                                                                 : 
                                                                                ((0x0025acd0U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x000000b6U
                                                                                 : 
                                                                                ((0x00b2c390U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x000000b7U
                                                                                 : 
                                                                                ((0x00b0a830U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x000000b8U
                                                                                 : 
                                                                                ((0x0024f8d0U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x000000b9U
                                                                                 : 
                                                                                ((0x0079a6b0U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x000000baU
                                                                                 : 
                                                                                ((0x00510110U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x000000bbU
                                                                                 : 
                                                                                ((0x00f5fee0U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x000000bcU
                                                                                 : 
                                                                                ((0x0082fb30U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x000000bdU
                                                                                 : 
                                                                                ((0x00047a70U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x000000beU
                                                                                 : 
                                                                                ((0x00fd3ec0U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x000000bfU
                                                                                 : 
                                                                                ((0x002b8f00U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x000000c0U
                                                                                 : 
                                                                                ((0x0024e590U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x000000c1U
                                                                                 : 
                                                                                ((0x0082a040U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x000000c2U
                                                                                 : 
                                                                                ((0x00f123c0U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x000000c3U
                                                                                 : 
                                                                                ((0x00bc6990U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x000000c4U
                                                                                 : 
                                                                                (((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_17)
                                                                                 ? 0x000000c5U
                                                                                 : 
                                                                                ((1U 
                                                                                == 
                                                                                (3U 
                                                                                & (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000020U))))
                                                                                 ? 0x000000c6U
                                                                                 : 
                                                                                (0x000000c7U 
                                                                                & (- (IData)(
                                                                                (0U 
                                                                                == 
                                                                                (3U 
                                                                                & (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000020U))))))))) 
                                                                                & (- (IData)(
                                                                                (0U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_12)))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))));
        __Vdly__t__DOT__sum = ((((QData)((IData)(((0x00000020U 
                                                   & (((0x00a00000U 
                                                        == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_11)
                                                        ? (IData)(
                                                                  (vlSelfRef.t__DOT__crc 
                                                                   >> 0x0000001bU))
                                                        : 
                                                       ((0x00c00000U 
                                                         == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_11)
                                                         ? (IData)(
                                                                   (vlSelfRef.t__DOT__crc 
                                                                    >> 0x0000001bU))
                                                         : 
                                                        ((0x00600000U 
                                                          == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_11)
                                                          ? (IData)(
                                                                    (vlSelfRef.t__DOT__crc 
                                                                     >> 0x0000001bU))
                                                          : 
                                                         ((0x00900000U 
                                                           == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_11)
                                                           ? (IData)(
                                                                     (vlSelfRef.t__DOT__crc 
                                                                      >> 0x0000001bU))
                                                           : 
                                                          ((0x00d00000U 
                                                            == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_11)
                                                            ? (IData)(
                                                                      (vlSelfRef.t__DOT__crc 
                                                                       >> 0x0000001bU))
                                                            : 
                                                           ((0x00300000U 
                                                             == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_11)
                                                             ? (IData)(
                                                                       (vlSelfRef.t__DOT__crc 
                                                                        >> 0x0000001bU))
                                                             : 
                                                            ((0x00100000U 
                                                              == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_11)
                                                              ? (IData)(
                                                                        (vlSelfRef.t__DOT__crc 
                                                                         >> 0x0000001bU))
                                                              : 
                                                             ((0x00880000U 
                                                               == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_16)
                                                               ? (IData)(
                                                                         (vlSelfRef.t__DOT__crc 
                                                                          >> 0x0000001aU))
                                                               : 
                                                              ((0x00700000U 
                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_16)
                                                                ? (IData)(
                                                                          (vlSelfRef.t__DOT__crc 
                                                                           >> 0x0000001aU))
                                                                : 
                                                               ((0x00400000U 
                                                                 == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_16)
                                                                 ? (IData)(
                                                                           (vlSelfRef.t__DOT__crc 
                                                                            >> 0x0000001aU))
                                                                 : 
                                                                ((0x00080000U 
                                                                  == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_16)
                                                                  ? (IData)(
                                                                            (vlSelfRef.t__DOT__crc 
                                                                             >> 0x0000001aU))
                                                                  : 
                                                                 ((0x00e00000U 
                                                                   == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_16)
                                                                   ? (IData)(
                                                                             (vlSelfRef.t__DOT__crc 
                                                                              >> 0x0000001aU))
                                                                   : 
                                                                  ((0x00580000U 
                                                                    == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_16)
                                                                    ? (IData)(
                                                                              (vlSelfRef.t__DOT__crc 
                                                                               >> 0x0000001aU))
                                                                    : 
                                                                   ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_40)
                                                                     ? (IData)(
                                                                               (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000019U))
                                                                     : 
                                                                    ((0x002c0000U 
                                                                      == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_12)
                                                                      ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000019U))
                                                                      : 
                                                                     ((0x00b80000U 
                                                                       == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_12)
                                                                       ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000019U))
                                                                       : 
                                                                      ((0x00ec0000U 
                                                                        == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_12)
                                                                        ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000019U))
                                                                        : 
                                                                       ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_39)
                                                                         ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000018U))
                                                                         : 
                                                                        ((0x00be0000U 
                                                                          == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_13)
                                                                          ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000018U))
                                                                          : 
                                                                         ((0x00e80000U 
                                                                           == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_13)
                                                                           ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000018U))
                                                                           : 
                                                                          ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_38)
                                                                            ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000017U))
                                                                            : 
                                                                           ((0x00f20000U 
                                                                             == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_14)
                                                                             ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000017U))
                                                                             : 
                                                                            ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_37)
                                                                              ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000016U))
                                                                              : 
                                                                             ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_36)
                                                                               ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000015U))
                                                                               : 
                                                                              ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_35)
                                                                                ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000014U))
                                                                                : 
                                                                               ((0x00820000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_7)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000014U))
                                                                                 : 
                                                                                ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_34)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000013U))
                                                                                 : 
                                                                                ((0x007d4000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_10)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000013U))
                                                                                 : 
                                                                                ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_33)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000012U))
                                                                                 : 
                                                                                ((0x0055d000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_8)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000012U))
                                                                                 : 
                                                                                ((0x00f69000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_8)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000012U))
                                                                                 : 
                                                                                ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_32)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000011U))
                                                                                 : 
                                                                                ((0x00fe5c00U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_5)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000011U))
                                                                                 : 
                                                                                ((0x00b28000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_5)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000011U))
                                                                                 : 
                                                                                ((0x00eba000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_5)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000011U))
                                                                                 : 
                                                                                ((0x00507c00U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_5)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000011U))
                                                                                 : 
                                                                                ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_31)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000010U))
                                                                                 : 
                                                                                ((0x00fe4200U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_6)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000010U))
                                                                                 : 
                                                                                ((0x00258400U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_6)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000010U))
                                                                                 : 
                                                                                ((0x00799600U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_6)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000010U))
                                                                                 : 
                                                                                ((0x00fff400U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_6)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000010U))
                                                                                 : 
                                                                                ((0x00bca600U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_6)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000010U))
                                                                                 : 
                                                                                ((0x00830e00U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_6)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000010U))
                                                                                 : 
                                                                                ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_30)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000000fU))
                                                                                 : 
                                                                                ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_29)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000000fU))
                                                                                 : 
                                                                                ((0x00788d00U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_3)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000000fU))
                                                                                 : 
                                                                                ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_28)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000000eU))
                                                                                 : 
                                                                                ((0x00ebae80U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_4)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000000eU))
                                                                                 : 
                                                                                ((0x00509580U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_4)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000000eU))
                                                                                 : 
                                                                                ((0x0024cc80U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_4)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000000eU))
                                                                                 : 
                                                                                ((0x004e5400U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_4)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000000eU))
                                                                                 : 
                                                                                ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_27)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000000dU))
                                                                                 : 
                                                                                ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_26)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000000dU))
                                                                                 : 
                                                                                ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_25)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000000cU))
                                                                                 : 
                                                                                ((0x0025d7c0U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000000cU))
                                                                                 : 
                                                                                ((0x00246780U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000000cU))
                                                                                 : 
                                                                                ((0x00b2e0a0U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000000cU))
                                                                                 : 
                                                                                ((0x00b114a0U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000000cU))
                                                                                 : 
                                                                                ((0x00fe4760U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000000cU))
                                                                                 : 
                                                                                ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_24)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000000bU))
                                                                                 : 
                                                                                ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_23)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000000bU))
                                                                                 : 
                                                                                ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_22)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000000bU))
                                                                                 : 
                                                                                ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_21)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000000bU))
                                                                                 : 
                                                                                ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_20)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000000bU))
                                                                                 : 
                                                                                ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_19)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000000bU))
                                                                                 : 
                                                                                ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_18)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000000bU))
                                                                                 : 
                                                                                ((0x0024e590U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000000bU))
                                                                                 : 
                                                                                ((0x0082a040U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000000bU))
                                                                                 : 
                                                                                ((0x00f123c0U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000000bU))
                                                                                 : 
                                                                                ((0x00bc6990U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0) 
                                                                                & (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x0000000bU)))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))) 
                                                      << 5U)) 
                                                  | ((0x00a00000U 
                                                      == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_11)
                                                      ? 4U
                                                      : 
                                                     ((0x00c00000U 
                                                       == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_11)
                                                       ? 4U
                                                       : 
                                                      ((0x00600000U 
                                                        == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_11)
                                                        ? 4U
                                                        : 
                                                       ((0x00900000U 
                                                         == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_11)
                                                         ? 4U
                                                         : 
                                                        ((0x00d00000U 
                                                          == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_11)
                                                          ? 4U
                                                          : 
                                                         ((0x00300000U 
                                                           == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_11)
                                                           ? 4U
                                                           : 
                                                          ((0x00100000U 
                                                            == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_11)
                                                            ? 4U
                                                            : 
                                                           ((0x00880000U 
                                                             == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_16)
                                                             ? 5U
                                                             : 
                                                            ((0x00700000U 
                                                              == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_16)
                                                              ? 5U
                                                              : 
                                                             ((0x00400000U 
                                                               == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_16)
                                                               ? 5U
                                                               : 
                                                              ((0x00080000U 
                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_16)
                                                                ? 5U
                                                                : 
                                                               ((0x00e00000U 
                                                                 == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_16)
                                                                 ? 5U
                                                                 : 
                                                                ((0x00580000U 
                                                                  == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_16)
                                                                  ? 5U
                                                                  : 
                                                                 ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_40)
                                                                   ? 6U
                                                                   : 
                                                                  ((0x002c0000U 
                                                                    == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_12)
                                                                    ? 6U
                                                                    : 
                                                                   ((0x00b80000U 
                                                                     == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_12)
                                                                     ? 6U
                                                                     : 
                                                                    ((0x00ec0000U 
                                                                      == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_12)
                                                                      ? 6U
                                                                      : 
                                                                     ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_39)
                                                                       ? 7U
                                                                       : 
                                                                      ((0x00be0000U 
                                                                        == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_13)
                                                                        ? 7U
                                                                        : 
                                                                       ((0x00e80000U 
                                                                         == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_13)
                                                                         ? 7U
                                                                         : 
                                                                        ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_38)
                                                                          ? 8U
                                                                          : 
                                                                         ((0x00f20000U 
                                                                           == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_14)
                                                                           ? 8U
                                                                           : 
                                                                          ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_37)
                                                                            ? 9U
                                                                            : 
                                                                           ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_36)
                                                                             ? 0x0000000aU
                                                                             : 
                                                                            ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_35)
                                                                              ? 0x0000000bU
                                                                              : 
                                                                             ((0x00820000U 
                                                                               == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_7)
                                                                               ? 0x0000000bU
                                                                               : 
                                                                              ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_34)
                                                                                ? 0x0000000cU
                                                                                : 
                                                                               ((0x007d4000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_10)
                                                                                 ? 0x0000000cU
                                                                                 : 
                                                                                ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_33)
                                                                                 ? 0x0000000dU
                                                                                 : 
                                                                                ((0x0055d000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_8)
                                                                                 ? 0x0000000dU
                                                                                 : 
                                                                                ((0x00f69000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_8)
                                                                                 ? 0x0000000dU
                                                                                 : 
                                                                                ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_32)
                                                                                 ? 0x0000000eU
                                                                                 : 
                                                                                ((0x00fe5c00U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_5)
                                                                                 ? 0x0000000eU
                                                                                 : 
                                                                                ((0x00b28000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_5)
                                                                                 ? 0x0000000eU
                                                                                 : 
                                                                                ((0x00eba000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_5)
                                                                                 ? 0x0000000eU
                                                                                 : 
                                                                                ((0x00507c00U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_5)
                                                                                 ? 0x0000000eU
                                                                                 : 
                                                                                ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_31)
                                                                                 ? 0x0000000fU
                                                                                 : 
                                                                                ((0x00fe4200U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_6)
                                                                                 ? 0x0000000fU
                                                                                 : 
                                                                                ((0x00258400U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_6)
                                                                                 ? 0x0000000fU
                                                                                 : 
                                                                                ((0x00799600U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_6)
                                                                                 ? 0x0000000fU
                                                                                 : 
                                                                                ((0x00fff400U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_6)
                                                                                 ? 0x0000000fU
                                                                                 : 
                                                                                ((0x00bca600U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_6)
                                                                                 ? 0x0000000fU
                                                                                 : 
                                                                                ((0x00830e00U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_6)
                                                                                 ? 0x0000000fU
                                                                                 : 
                                                                                ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_30)
                                                                                 ? 0x00000010U
                                                                                 : 
                                                                                ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_29)
                                                                                 ? 0x00000010U
                                                                                 : 
                                                                                ((0x00788d00U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_3)
                                                                                 ? 0x00000010U
                                                                                 : 
                                                                                ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_28)
                                                                                 ? 0x00000011U
                                                                                 : 
                                                                                ((0x00ebae80U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_4)
                                                                                 ? 0x00000011U
                                                                                 : 
                                                                                ((0x00509580U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_4)
                                                                                 ? 0x00000011U
                                                                                 : 
                                                                                ((0x0024cc80U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_4)
                                                                                 ? 0x00000011U
                                                                                 : 
                                                                                ((0x004e5400U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_4)
                                                                                 ? 0x00000011U
                                                                                 : 
                                                                                ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_27)
                                                                                 ? 0x00000012U
                                                                                 : 
                                                                                ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_26)
                                                                                 ? 0x00000012U
                                                                                 : 
                                                                                ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_25)
                                                                                 ? 0x00000013U
                                                                                 : 
                                                                                ((0x0025d7c0U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1)
                                                                                 ? 0x00000013U
                                                                                 : 
                                                                                ((0x00246780U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1)
                                                                                 ? 0x00000013U
                                                                                 : 
                                                                                ((0x00b2e0a0U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1)
                                                                                 ? 0x00000013U
                                                                                 : 
                                                                                ((0x00b114a0U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1)
                                                                                 ? 0x00000013U
                                                                                 : 
                                                                                ((0x00fe4760U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1)
                                                                                 ? 0x00000013U
                                                                                 : 
                                                                                ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_24)
                                                                                 ? 0x00000014U
                                                                                 : 
                                                                                ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_23)
                                                                                 ? 0x00000014U
                                                                                 : 
                                                                                ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_22)
                                                                                 ? 0x00000014U
                                                                                 : 
                                                                                ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_21)
                                                                                 ? 0x00000014U
                                                                                 : 
                                                                                ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_20)
                                                                                 ? 0x00000014U
                                                                                 : 
                                                                                ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_19)
                                                                                 ? 0x00000014U
                                                                                 : 
                                                                                ((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_18)
                                                                                 ? 0x00000014U
                                                                                 : 
                                                                                ((0x0024e590U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x00000014U
                                                                                 : 
                                                                                ((0x0082a040U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x00000014U
                                                                                 : 
                                                                                ((0x00f123c0U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x00000014U
                                                                                 : 
                                                                                ((0x00bc6990U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0)
                                                                                 ? 0x00000014U
                                                                                 : 
                                                                                (0x0000001fU 
                                                                                & (((IData)(vlSelfRef.__VdfgRegularize_h6e95ff9d_0_17)
                                                                                 ? 0x12U
                                                                                 : 
                                                                                ((1U 
                                                                                == 
                                                                                (3U 
                                                                                & (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000020U))))
                                                                                 ? 0x13U
                                                                                 : 
                                                                                (0x13U 
                                                                                | (- (IData)(
                                                                                (0U 
                                                                                != 
                                                                                (3U 
                                                                                & (IData)(
                                                                                (vlSelfRef.t__DOT__crc 
                                                                                >> 0x00000020U))))))))) 
                                                                                | (- (IData)(
                                                                                (0U 
                                                                                != vlSelfRef.__VdfgRegularize_h6e95ff9d_0_12)))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))) 
                                 << 0x00000020U) | (QData)((IData)(
                                                                   (((0x00a00000U 
                                                                      == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_11) 
                                                                     | ((0x00c00000U 
                                                                         == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_11) 
                                                                        | ((0x00600000U 
                                                                            == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_11) 
                                                                           | ((0x00900000U 
                                                                               == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_11) 
                                                                              | ((0x00d00000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_11) 
                                                                                | ((0x00300000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_11) 
                                                                                | ((0x00100000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_11) 
                                                                                | (0x00880000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_16))))))))
                                                                     ? 
                                                                    (((0x00c00000U 
                                                                       == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_11)
                                                                       ? 1U
                                                                       : 
                                                                      ((0x00600000U 
                                                                        == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_11)
                                                                        ? 2U
                                                                        : 
                                                                       ((0x00900000U 
                                                                         == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_11)
                                                                         ? 3U
                                                                         : 
                                                                        ((0x00d00000U 
                                                                          == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_11)
                                                                          ? 4U
                                                                          : 
                                                                         ((0x00300000U 
                                                                           == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_11)
                                                                           ? 5U
                                                                           : 
                                                                          ((0x00100000U 
                                                                            == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_11)
                                                                            ? 6U
                                                                            : 7U)))))) 
                                                                     & (- (IData)(
                                                                                (0x00a00000U 
                                                                                != vlSelfRef.__VdfgRegularize_h6e95ff9d_0_11))))
                                                                     : 
                                                                    ((0x00700000U 
                                                                      == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_16)
                                                                      ? 8U
                                                                      : 
                                                                     ((0x00400000U 
                                                                       == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_16)
                                                                       ? 9U
                                                                       : 
                                                                      ((0x00080000U 
                                                                        == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_16)
                                                                        ? 0x0000000aU
                                                                        : 
                                                                       ((0x00e00000U 
                                                                         == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_16)
                                                                         ? 0x0000000bU
                                                                         : 
                                                                        ((0x00580000U 
                                                                          == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_16)
                                                                          ? 0x0000000cU
                                                                          : 
                                                                         ((0x00840000U 
                                                                           == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_12)
                                                                           ? 0x0000000dU
                                                                           : 
                                                                          ((0x00f80000U 
                                                                            == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_12)
                                                                            ? 0x0000000eU
                                                                            : 
                                                                           ((0x00480000U 
                                                                             == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_12)
                                                                             ? 0x0000000fU
                                                                             : 
                                                                            ((0x002c0000U 
                                                                              == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_12)
                                                                              ? 0x00000010U
                                                                              : 
                                                                             ((0x00b80000U 
                                                                               == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_12)
                                                                               ? 0x00000011U
                                                                               : 
                                                                              ((0x00ec0000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_12)
                                                                                ? 0x00000012U
                                                                                : 
                                                                               ((0x007a0000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_13)
                                                                                 ? 0x00000014U
                                                                                 : 
                                                                                ((0x00280000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_13)
                                                                                 ? 0x00000015U
                                                                                 : 
                                                                                ((0x007e0000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_13)
                                                                                 ? 0x00000016U
                                                                                 : 
                                                                                ((0x00b40000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_13)
                                                                                 ? 0x00000017U
                                                                                 : 
                                                                                ((0x00800000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_13)
                                                                                 ? 0x00000018U
                                                                                 : 
                                                                                ((0x00be0000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_13)
                                                                                 ? 0x00000019U
                                                                                 : 
                                                                                ((0x00e80000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_13)
                                                                                 ? 0x0000001aU
                                                                                 : 
                                                                                ((0x007c0000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_14)
                                                                                 ? 0x0000001bU
                                                                                 : 
                                                                                ((0x00060000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_14)
                                                                                 ? 0x0000001cU
                                                                                 : 
                                                                                ((0x00050000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_14)
                                                                                 ? 0x0000001dU
                                                                                 : 
                                                                                ((0x004c0000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_14)
                                                                                 ? 0x0000001eU
                                                                                 : 
                                                                                ((0x00b60000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_14)
                                                                                 ? 0x0000001fU
                                                                                 : 
                                                                                ((0x00260000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_14)
                                                                                 ? 0x00000020U
                                                                                 : 
                                                                                ((0x00f20000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_14)
                                                                                 ? 0x00000021U
                                                                                 : 
                                                                                ((0x004e8000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_15)
                                                                                 ? 0x00000022U
                                                                                 : 
                                                                                ((0x00200000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_15)
                                                                                 ? 0x00000023U
                                                                                 : 
                                                                                ((0x00578000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_15)
                                                                                 ? 0x00000024U
                                                                                 : 
                                                                                ((0x00550000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_15)
                                                                                 ? 0x00000025U
                                                                                 : 
                                                                                ((0x004d8000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_15)
                                                                                 ? 0x00000026U
                                                                                 : 
                                                                                ((0x00518000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_15)
                                                                                 ? 0x00000027U
                                                                                 : 
                                                                                ((0x00540000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_15)
                                                                                 ? 0x00000028U
                                                                                 : 
                                                                                ((0x00f54000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_9)
                                                                                 ? 0x00000029U
                                                                                 : 
                                                                                ((0x00220000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_9)
                                                                                 ? 0x0000002aU
                                                                                 : 
                                                                                ((0x00534000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_9)
                                                                                 ? 0x0000002bU
                                                                                 : 
                                                                                ((0x00250000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_9)
                                                                                 ? 0x0000002cU
                                                                                 : 
                                                                                ((0x00b38000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_9)
                                                                                 ? 0x0000002dU
                                                                                 : 
                                                                                ((0x00f0c000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_9)
                                                                                 ? 0x0000002eU
                                                                                 : 
                                                                                ((0x00500000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_9)
                                                                                 ? 0x0000002fU
                                                                                 : 
                                                                                ((0x00fc0000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_9)
                                                                                 ? 0x00000030U
                                                                                 : 
                                                                                ((0x00b74000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_7)
                                                                                 ? 0x00000031U
                                                                                 : 
                                                                                ((0x00f06000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_7)
                                                                                 ? 0x00000032U
                                                                                 : 
                                                                                ((0x004f6000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_7)
                                                                                 ? 0x00000033U
                                                                                 : 
                                                                                ((0x002b6000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_7)
                                                                                 ? 0x00000034U
                                                                                 : 
                                                                                ((0x00528000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_7)
                                                                                 ? 0x00000035U
                                                                                 : 
                                                                                ((0x00f78000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_7)
                                                                                 ? 0x00000036U
                                                                                 : 
                                                                                ((0x00272000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_7)
                                                                                 ? 0x00000037U
                                                                                 : 
                                                                                ((0x00b14000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_7)
                                                                                 ? 0x00000038U
                                                                                 : 
                                                                                ((0x00820000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_7)
                                                                                 ? 0x00000039U
                                                                                 : 
                                                                                ((0x00fec000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_10)
                                                                                 ? 0x0000003aU
                                                                                 : 
                                                                                ((0x0083e000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_10)
                                                                                 ? 0x0000003bU
                                                                                 : 
                                                                                ((0x00832000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_10)
                                                                                 ? 0x0000003cU
                                                                                 : 
                                                                                ((0x00839000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_10)
                                                                                 ? 0x0000003dU
                                                                                 : 
                                                                                ((0x0052f000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_10)
                                                                                 ? 0x0000003eU
                                                                                 : 
                                                                                ((0x0020c000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_10)
                                                                                 ? 0x0000003fU
                                                                                 : 
                                                                                ((0x0007f000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_10)
                                                                                 ? 0x00000040U
                                                                                 : 
                                                                                ((0x007d4000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_10)
                                                                                 ? 0x00000041U
                                                                                 : 
                                                                                ((0x00ebe800U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_8)
                                                                                 ? 0x00000042U
                                                                                 : 
                                                                                ((0x004d7000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_8)
                                                                                 ? 0x00000043U
                                                                                 : 
                                                                                ((0x00fed800U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_8)
                                                                                 ? 0x00000044U
                                                                                 : 
                                                                                ((0x0056c800U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_8)
                                                                                 ? 0x00000045U
                                                                                 : 
                                                                                ((0x00216000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_8)
                                                                                 ? 0x00000046U
                                                                                 : 
                                                                                ((0x00ff6800U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_8)
                                                                                 ? 0x00000047U
                                                                                 : 
                                                                                ((0x00bcb000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_8)
                                                                                 ? 0x00000048U
                                                                                 : 
                                                                                ((0x0055d000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_8)
                                                                                 ? 0x00000049U
                                                                                 : 
                                                                                ((0x00f69000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_8)
                                                                                 ? 0x0000004aU
                                                                                 : 
                                                                                ((0x00548c00U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_5)
                                                                                 ? 0x0000004bU
                                                                                 : 
                                                                                ((0x00b0c800U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_5)
                                                                                 ? 0x0000004cU
                                                                                 : 
                                                                                ((0x00bd3c00U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_5)
                                                                                 ? 0x0000004dU
                                                                                 : 
                                                                                ((0x00b05400U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_5)
                                                                                 ? 0x0000004eU
                                                                                 : 
                                                                                ((0x002b3c00U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_5)
                                                                                 ? 0x0000004fU
                                                                                 : 
                                                                                ((0x0020b000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_5)
                                                                                 ? 0x00000050U
                                                                                 : 
                                                                                ((0x00fe5c00U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_5)
                                                                                 ? 0x00000051U
                                                                                 : 
                                                                                ((0x00b28000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_5)
                                                                                 ? 0x00000052U
                                                                                 : 
                                                                                ((0x00eba000U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_5)
                                                                                 ? 0x00000053U
                                                                                 : 
                                                                                ((0x00507c00U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_5)
                                                                                 ? 0x00000054U
                                                                                 : 
                                                                                ((0x00bd9600U 
                                                                                == vlSelfRef.__VdfgRegularize_h6e95ff9d_0_6)
                                                                                 ? 0x00000055U
                                                                                 : __Vdeeptemp_h5279618a__0))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))))) 
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
                    VL_STOP_MT("/root/reference/test_regress/t/t_case_deep.v", 60, "");
                }
                if (VL_UNLIKELY(((0x5537fa30d49bf865ULL 
                                  != vlSelfRef.t__DOT__sum)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_case_deep.v", 61, "");
                }
                vlCtx.writef(1U, 0U);
                VL_FINISH_MT("/root/reference/test_regress/t/t_case_deep.v", 63, "");
            }
        }
        vlSelfRef.t__DOT__cyc = __Vdly__t__DOT__cyc;
        vlSelfRef.t__DOT__sum = __Vdly__t__DOT__sum;
        vlSelfRef.t__DOT__crc = __Vdly__t__DOT__crc;
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_0 = 
            (0x00fffff0U & (IData)((vlSelfRef.t__DOT__crc 
                                    >> 8U)));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_1 = 
            (0x00ffffe0U & (IData)((vlSelfRef.t__DOT__crc 
                                    >> 8U)));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_2 = 
            (0x00ffffc0U & (IData)((vlSelfRef.t__DOT__crc 
                                    >> 8U)));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_3 = 
            (0x00ffff00U & (IData)((vlSelfRef.t__DOT__crc 
                                    >> 8U)));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_4 = 
            (0x00ffff80U & (IData)((vlSelfRef.t__DOT__crc 
                                    >> 8U)));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_5 = 
            (0x00fffc00U & (IData)((vlSelfRef.t__DOT__crc 
                                    >> 8U)));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_6 = 
            (0x00fffe00U & (IData)((vlSelfRef.t__DOT__crc 
                                    >> 8U)));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_7 = 
            (0x00ffe000U & (IData)((vlSelfRef.t__DOT__crc 
                                    >> 8U)));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_8 = 
            (0x00fff800U & (IData)((vlSelfRef.t__DOT__crc 
                                    >> 8U)));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_9 = 
            (0x00ffc000U & (IData)((vlSelfRef.t__DOT__crc 
                                    >> 8U)));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_10 = 
            (0x00fff000U & (IData)((vlSelfRef.t__DOT__crc 
                                    >> 8U)));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_11 = 
            (0x00f00000U & (IData)((vlSelfRef.t__DOT__crc 
                                    >> 8U)));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_12 = 
            (0x00fc0000U & (IData)((vlSelfRef.t__DOT__crc 
                                    >> 8U)));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_13 = 
            (0x00fe0000U & (IData)((vlSelfRef.t__DOT__crc 
                                    >> 8U)));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_14 = 
            (0x00ff0000U & (IData)((vlSelfRef.t__DOT__crc 
                                    >> 8U)));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_15 = 
            (0x00ff8000U & (IData)((vlSelfRef.t__DOT__crc 
                                    >> 8U)));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_16 = 
            (0x00f80000U & (IData)((vlSelfRef.t__DOT__crc 
                                    >> 8U)));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_17 = 
            (2U == (2U & (IData)((vlSelfRef.t__DOT__crc 
                                  >> 0x00000020U))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_18 = 
            ((0x0024f8d0U == (0x00fffff0U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x0079a6b0U == (0x00fffff0U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | ((0x00510110U == (0x00fffff0U & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                   | ((0x00f5fee0U == (0x00fffff0U 
                                       & (IData)((vlSelfRef.t__DOT__crc 
                                                  >> 8U)))) 
                      | ((0x0082fb30U == (0x00fffff0U 
                                          & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
                         | ((0x00047a70U == (0x00fffff0U 
                                             & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                            | ((0x00fd3ec0U == (0x00fffff0U 
                                                & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                               | (0x002b8f00U == (0x00fffff0U 
                                                  & (IData)(
                                                            (vlSelfRef.t__DOT__crc 
                                                             >> 8U)))))))))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_19 = 
            ((0x00f0aef0U == (0x00fffff0U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x00249aa0U == (0x00fffff0U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | ((0x00f16c30U == (0x00fffff0U & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                   | ((0x00572330U == (0x00fffff0U 
                                       & (IData)((vlSelfRef.t__DOT__crc 
                                                  >> 8U)))) 
                      | ((0x00781150U == (0x00fffff0U 
                                          & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
                         | ((0x0025acd0U == (0x00fffff0U 
                                             & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                            | ((0x00b2c390U == (0x00fffff0U 
                                                & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                               | (0x00b0a830U == (0x00fffff0U 
                                                  & (IData)(
                                                            (vlSelfRef.t__DOT__crc 
                                                             >> 8U)))))))))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_20 = 
            ((0x00506fe0U == (0x00fffff0U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x00070820U == (0x00fffff0U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | ((0x002a0c80U == (0x00fffff0U & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                   | ((0x004e52e0U == (0x00fffff0U 
                                       & (IData)((vlSelfRef.t__DOT__crc 
                                                  >> 8U)))) 
                      | ((0x00f00800U == (0x00fffff0U 
                                          & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
                         | ((0x004d6790U == (0x00fffff0U 
                                             & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                            | ((0x00f71d50U == (0x00fffff0U 
                                                & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                               | (0x0079a760U == (0x00fffff0U 
                                                  & (IData)(
                                                            (vlSelfRef.t__DOT__crc 
                                                             >> 8U)))))))))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_21 = 
            ((0x00521d60U == (0x00fffff0U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x00eaaef0U == (0x00fffff0U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | ((0x00fe4e30U == (0x00fffff0U & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                   | ((0x00573ef0U == (0x00fffff0U 
                                       & (IData)((vlSelfRef.t__DOT__crc 
                                                  >> 8U)))) 
                      | ((0x00b1ffd0U == (0x00fffff0U 
                                          & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
                         | ((0x00b13300U == (0x00fffff0U 
                                             & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                            | ((0x00f463d0U == (0x00fffff0U 
                                                & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                               | (0x002bba30U == (0x00fffff0U 
                                                  & (IData)(
                                                            (vlSelfRef.t__DOT__crc 
                                                             >> 8U)))))))))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_22 = 
            ((0x00b19770U == (0x00fffff0U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x00bcde80U == (0x00fffff0U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | ((0x00515d00U == (0x00fffff0U & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                   | ((0x0054f8e0U == (0x00fffff0U 
                                       & (IData)((vlSelfRef.t__DOT__crc 
                                                  >> 8U)))) 
                      | ((0x00fe6670U == (0x00fffff0U 
                                          & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
                         | ((0x00f7f510U == (0x00fffff0U 
                                             & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                            | ((0x00b05e00U == (0x00fffff0U 
                                                & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                               | (0x004f8450U == (0x00fffff0U 
                                                  & (IData)(
                                                            (vlSelfRef.t__DOT__crc 
                                                             >> 8U)))))))))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_23 = 
            ((0x0056f100U == (0x00fffff0U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x00b7cc90U == (0x00fffff0U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | ((0x00fdb250U == (0x00fffff0U & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                   | ((0x00b05210U == (0x00fffff0U 
                                       & (IData)((vlSelfRef.t__DOT__crc 
                                                  >> 8U)))) 
                      | ((0x00b26b60U == (0x00fffff0U 
                                          & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
                         | ((0x00794180U == (0x00fffff0U 
                                             & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                            | ((0x00f62db0U == (0x00fffff0U 
                                                & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                               | (0x005090b0U == (0x00fffff0U 
                                                  & (IData)(
                                                            (vlSelfRef.t__DOT__crc 
                                                             >> 8U)))))))))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_24 = 
            ((0x0022f6d0U == (0x00fffff0U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x0082aad0U == (0x00fffff0U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | (0x00b294d0U == (0x00fffff0U & (IData)(
                                                         (vlSelfRef.t__DOT__crc 
                                                          >> 8U))))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_25 = 
            ((0x0082c680U == (0x00ffffe0U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x0025cec0U == (0x00ffffe0U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | ((0x005641a0U == (0x00ffffe0U & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                   | ((0x00509560U == (0x00ffffe0U 
                                       & (IData)((vlSelfRef.t__DOT__crc 
                                                  >> 8U)))) 
                      | ((0x00f7e9a0U == (0x00ffffe0U 
                                          & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
                         | ((0x00b22ac0U == (0x00ffffe0U 
                                             & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                            | ((0x00b04420U == (0x00ffffe0U 
                                                & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                               | (0x00eb1320U == (0x00ffffe0U 
                                                  & (IData)(
                                                            (vlSelfRef.t__DOT__crc 
                                                             >> 8U)))))))))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_26 = 
            ((0x00b7aa40U == (0x00ffffc0U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x00f72f80U == (0x00ffffc0U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | ((0x00523400U == (0x00ffffc0U & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                   | ((0x0024f640U == (0x00ffffc0U 
                                       & (IData)((vlSelfRef.t__DOT__crc 
                                                  >> 8U)))) 
                      | ((0x00532240U == (0x00ffffc0U 
                                          & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
                         | ((0x00eb0e00U == (0x00ffffc0U 
                                             & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                            | ((0x00eb3140U == (0x00ffffc0U 
                                                & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                               | (0x00508e40U == (0x00ffffc0U 
                                                  & (IData)(
                                                            (vlSelfRef.t__DOT__crc 
                                                             >> 8U)))))))))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_27 = 
            ((0x004f5200U == (0x00ffffc0U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x00eae480U == (0x00ffffc0U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | ((0x00249e00U == (0x00ffffc0U & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                   | (0x00bc6f40U == (0x00ffffc0U & (IData)(
                                                            (vlSelfRef.t__DOT__crc 
                                                             >> 8U)))))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_28 = 
            ((0x0022c600U == (0x00ffff80U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x0022b500U == (0x00ffff80U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | ((0x00fef000U == (0x00ffff80U & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                   | ((0x0020e800U == (0x00ffff80U 
                                       & (IData)((vlSelfRef.t__DOT__crc 
                                                  >> 8U)))) 
                      | ((0x00277480U == (0x00ffff80U 
                                          & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
                         | ((0x00fee180U == (0x00ffff80U 
                                             & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                            | (0x00f1a200U == (0x00ffff80U 
                                               & (IData)(
                                                         (vlSelfRef.t__DOT__crc 
                                                          >> 8U))))))))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_29 = 
            ((0x00f6b200U == (0x00ffff00U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x00246400U == (0x00ffff00U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | ((0x00577400U == (0x00ffff00U & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                   | ((0x00568f00U == (0x00ffff00U 
                                       & (IData)((vlSelfRef.t__DOT__crc 
                                                  >> 8U)))) 
                      | ((0x00219f00U == (0x00ffff00U 
                                          & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
                         | ((0x00548500U == (0x00ffff00U 
                                             & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                            | ((0x00eac000U == (0x00ffff00U 
                                                & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                               | (0x00f03200U == (0x00ffff00U 
                                                  & (IData)(
                                                            (vlSelfRef.t__DOT__crc 
                                                             >> 8U)))))))))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_30 = 
            ((0x00240500U == (0x00ffff00U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | (0x0024a900U == (0x00ffff00U & (IData)(
                                                      (vlSelfRef.t__DOT__crc 
                                                       >> 8U)))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_31 = 
            ((0x00bd9600U == (0x00fffe00U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x00bd1800U == (0x00fffe00U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | ((0x0083ce00U == (0x00fffe00U & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                   | (0x002ab000U == (0x00fffe00U & (IData)(
                                                            (vlSelfRef.t__DOT__crc 
                                                             >> 8U)))))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_32 = 
            ((0x00548c00U == (0x00fffc00U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x00b0c800U == (0x00fffc00U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | ((0x00bd3c00U == (0x00fffc00U & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                   | ((0x00b05400U == (0x00fffc00U 
                                       & (IData)((vlSelfRef.t__DOT__crc 
                                                  >> 8U)))) 
                      | ((0x002b3c00U == (0x00fffc00U 
                                          & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
                         | (0x0020b000U == (0x00fffc00U 
                                            & (IData)(
                                                      (vlSelfRef.t__DOT__crc 
                                                       >> 8U)))))))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_33 = 
            ((0x00ebe800U == (0x00fff800U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x004d7000U == (0x00fff800U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | ((0x00fed800U == (0x00fff800U & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                   | ((0x0056c800U == (0x00fff800U 
                                       & (IData)((vlSelfRef.t__DOT__crc 
                                                  >> 8U)))) 
                      | ((0x00216000U == (0x00fff800U 
                                          & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
                         | ((0x00ff6800U == (0x00fff800U 
                                             & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                            | (0x00bcb000U == (0x00fff800U 
                                               & (IData)(
                                                         (vlSelfRef.t__DOT__crc 
                                                          >> 8U))))))))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_34 = 
            ((0x00fec000U == (0x00fff000U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x0083e000U == (0x00fff000U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | ((0x00832000U == (0x00fff000U & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                   | ((0x00839000U == (0x00fff000U 
                                       & (IData)((vlSelfRef.t__DOT__crc 
                                                  >> 8U)))) 
                      | ((0x0052f000U == (0x00fff000U 
                                          & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
                         | ((0x0020c000U == (0x00fff000U 
                                             & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                            | (0x0007f000U == (0x00fff000U 
                                               & (IData)(
                                                         (vlSelfRef.t__DOT__crc 
                                                          >> 8U))))))))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_35 = 
            ((0x00b74000U == (0x00ffe000U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x00f06000U == (0x00ffe000U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | ((0x004f6000U == (0x00ffe000U & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                   | ((0x002b6000U == (0x00ffe000U 
                                       & (IData)((vlSelfRef.t__DOT__crc 
                                                  >> 8U)))) 
                      | ((0x00528000U == (0x00ffe000U 
                                          & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
                         | ((0x00f78000U == (0x00ffe000U 
                                             & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                            | ((0x00272000U == (0x00ffe000U 
                                                & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                               | (0x00b14000U == (0x00ffe000U 
                                                  & (IData)(
                                                            (vlSelfRef.t__DOT__crc 
                                                             >> 8U)))))))))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_36 = 
            ((0x00f54000U == (0x00ffc000U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x00220000U == (0x00ffc000U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | ((0x00534000U == (0x00ffc000U & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                   | ((0x00250000U == (0x00ffc000U 
                                       & (IData)((vlSelfRef.t__DOT__crc 
                                                  >> 8U)))) 
                      | ((0x00b38000U == (0x00ffc000U 
                                          & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
                         | ((0x00f0c000U == (0x00ffc000U 
                                             & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                            | ((0x00500000U == (0x00ffc000U 
                                                & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                               | (0x00fc0000U == (0x00ffc000U 
                                                  & (IData)(
                                                            (vlSelfRef.t__DOT__crc 
                                                             >> 8U)))))))))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_37 = 
            ((0x004e8000U == (0x00ff8000U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x00200000U == (0x00ff8000U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | ((0x00578000U == (0x00ff8000U & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                   | ((0x00550000U == (0x00ff8000U 
                                       & (IData)((vlSelfRef.t__DOT__crc 
                                                  >> 8U)))) 
                      | ((0x004d8000U == (0x00ff8000U 
                                          & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
                         | ((0x00518000U == (0x00ff8000U 
                                             & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                            | (0x00540000U == (0x00ff8000U 
                                               & (IData)(
                                                         (vlSelfRef.t__DOT__crc 
                                                          >> 8U))))))))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_38 = 
            ((0x007c0000U == (0x00ff0000U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x00060000U == (0x00ff0000U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | ((0x00050000U == (0x00ff0000U & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                   | ((0x004c0000U == (0x00ff0000U 
                                       & (IData)((vlSelfRef.t__DOT__crc 
                                                  >> 8U)))) 
                      | ((0x00b60000U == (0x00ff0000U 
                                          & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
                         | (0x00260000U == (0x00ff0000U 
                                            & (IData)(
                                                      (vlSelfRef.t__DOT__crc 
                                                       >> 8U)))))))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_39 = 
            ((0x007a0000U == (0x00fe0000U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x00280000U == (0x00fe0000U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | ((0x007e0000U == (0x00fe0000U & (IData)(
                                                          (vlSelfRef.t__DOT__crc 
                                                           >> 8U)))) 
                   | ((0x00b40000U == (0x00fe0000U 
                                       & (IData)((vlSelfRef.t__DOT__crc 
                                                  >> 8U)))) 
                      | (0x00800000U == (0x00fe0000U 
                                         & (IData)(
                                                   (vlSelfRef.t__DOT__crc 
                                                    >> 8U))))))));
        vlSelfRef.__VdfgRegularize_h6e95ff9d_0_40 = 
            ((0x00840000U == (0x00fc0000U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 8U)))) 
             | ((0x00f80000U == (0x00fc0000U & (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 8U)))) 
                | (0x00480000U == (0x00fc0000U & (IData)(
                                                         (vlSelfRef.t__DOT__crc 
                                                          >> 8U))))));
    }

    VLCU_D static void Vt_case_deep___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt_case_deep___024root___eval_phase__act(Vt_case_deep___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt_case_deep___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vt_case_deep___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vt_case_deep___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vt_case_deep___024root___eval_phase__nba(Vt_case_deep___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vt_case_deep___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    Vt_case_deep___024root___nba_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
            Vt_case_deep___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vt_case_deep___024root___eval(Vt_case_deep___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt_case_deep___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("/root/reference/test_regress/t/t_case_deep.v", 7, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vt_case_deep___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("/root/reference/test_regress/t/t_case_deep.v", 7, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vt_case_deep___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vt_case_deep___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vt_case_deep___024root___ctor_var_reset(Vt_case_deep___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        vlSelf->t__DOT__crc = VL_SCOPED_RAND_RESET_Q(64, __VscopeHash, 17680824324711171631ull);
        vlSelf->t__DOT__sum = VL_SCOPED_RAND_RESET_Q(64, __VscopeHash, 1514072735762163239ull);
        vlSelf->__VdfgRegularize_h6e95ff9d_0_0 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_1 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_2 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_3 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_4 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_5 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_6 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_7 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_8 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_9 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_10 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_11 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_12 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_13 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_14 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_15 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_16 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_17 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_18 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_19 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_20 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_21 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_22 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_23 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_24 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_25 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_26 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_27 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_28 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_29 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_30 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_31 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_32 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_33 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_34 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_35 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_36 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_37 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_38 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_39 = 0;
        vlSelf->__VdfgRegularize_h6e95ff9d_0_40 = 0;
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
        Vt_case_deep___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt_case_deep___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt_case_deep___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt_case_deep___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vt_case_deep___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt_case_deep___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt_case_deep__cuda::info() {
    typedef Vt_case_deep__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"__VdfgRegularize_h6e95ff9d_0_17", "__VdfgRegularize_h6e95ff9d_0_17", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_17},
        {"__VdfgRegularize_h6e95ff9d_0_18", "__VdfgRegularize_h6e95ff9d_0_18", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_18},
        {"__VdfgRegularize_h6e95ff9d_0_19", "__VdfgRegularize_h6e95ff9d_0_19", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_19},
        {"__VdfgRegularize_h6e95ff9d_0_20", "__VdfgRegularize_h6e95ff9d_0_20", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_20},
        {"__VdfgRegularize_h6e95ff9d_0_21", "__VdfgRegularize_h6e95ff9d_0_21", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_21},
        {"__VdfgRegularize_h6e95ff9d_0_22", "__VdfgRegularize_h6e95ff9d_0_22", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_22},
        {"__VdfgRegularize_h6e95ff9d_0_23", "__VdfgRegularize_h6e95ff9d_0_23", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_23},
        {"__VdfgRegularize_h6e95ff9d_0_24", "__VdfgRegularize_h6e95ff9d_0_24", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_24},
        {"__VdfgRegularize_h6e95ff9d_0_25", "__VdfgRegularize_h6e95ff9d_0_25", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_25},
        {"__VdfgRegularize_h6e95ff9d_0_26", "__VdfgRegularize_h6e95ff9d_0_26", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_26},
        {"__VdfgRegularize_h6e95ff9d_0_27", "__VdfgRegularize_h6e95ff9d_0_27", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_27},
        {"__VdfgRegularize_h6e95ff9d_0_28", "__VdfgRegularize_h6e95ff9d_0_28", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_28},
        {"__VdfgRegularize_h6e95ff9d_0_29", "__VdfgRegularize_h6e95ff9d_0_29", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_29},
        {"__VdfgRegularize_h6e95ff9d_0_30", "__VdfgRegularize_h6e95ff9d_0_30", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_30},
        {"__VdfgRegularize_h6e95ff9d_0_31", "__VdfgRegularize_h6e95ff9d_0_31", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_31},
        {"__VdfgRegularize_h6e95ff9d_0_32", "__VdfgRegularize_h6e95ff9d_0_32", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_32},
        {"__VdfgRegularize_h6e95ff9d_0_33", "__VdfgRegularize_h6e95ff9d_0_33", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_33},
        {"__VdfgRegularize_h6e95ff9d_0_34", "__VdfgRegularize_h6e95ff9d_0_34", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_34},
        {"__VdfgRegularize_h6e95ff9d_0_35", "__VdfgRegularize_h6e95ff9d_0_35", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_35},
        {"__VdfgRegularize_h6e95ff9d_0_36", "__VdfgRegularize_h6e95ff9d_0_36", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_36},
        {"__VdfgRegularize_h6e95ff9d_0_37", "__VdfgRegularize_h6e95ff9d_0_37", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_37},
        {"__VdfgRegularize_h6e95ff9d_0_38", "__VdfgRegularize_h6e95ff9d_0_38", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_38},
        {"__VdfgRegularize_h6e95ff9d_0_39", "__VdfgRegularize_h6e95ff9d_0_39", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_39},
        {"__VdfgRegularize_h6e95ff9d_0_40", "__VdfgRegularize_h6e95ff9d_0_40", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_40},
        {"__VstlFirstIteration", "__VstlFirstIteration", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlFirstIteration},
        {"__VstlPhaseResult", "__VstlPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"t__DOT__cyc", "t.cyc", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__cyc},
        {"__VdfgRegularize_h6e95ff9d_0_0", "__VdfgRegularize_h6e95ff9d_0_0", VLCU_MEMBER_STATE, 4, 0, 1, 24, 23, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_0},
        {"__VdfgRegularize_h6e95ff9d_0_1", "__VdfgRegularize_h6e95ff9d_0_1", VLCU_MEMBER_STATE, 4, 0, 1, 24, 23, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_1},
        {"__VdfgRegularize_h6e95ff9d_0_2", "__VdfgRegularize_h6e95ff9d_0_2", VLCU_MEMBER_STATE, 4, 0, 1, 24, 23, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_2},
        {"__VdfgRegularize_h6e95ff9d_0_3", "__VdfgRegularize_h6e95ff9d_0_3", VLCU_MEMBER_STATE, 4, 0, 1, 24, 23, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_3},
        {"__VdfgRegularize_h6e95ff9d_0_4", "__VdfgRegularize_h6e95ff9d_0_4", VLCU_MEMBER_STATE, 4, 0, 1, 24, 23, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_4},
        {"__VdfgRegularize_h6e95ff9d_0_5", "__VdfgRegularize_h6e95ff9d_0_5", VLCU_MEMBER_STATE, 4, 0, 1, 24, 23, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_5},
        {"__VdfgRegularize_h6e95ff9d_0_6", "__VdfgRegularize_h6e95ff9d_0_6", VLCU_MEMBER_STATE, 4, 0, 1, 24, 23, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_6},
        {"__VdfgRegularize_h6e95ff9d_0_7", "__VdfgRegularize_h6e95ff9d_0_7", VLCU_MEMBER_STATE, 4, 0, 1, 24, 23, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_7},
        {"__VdfgRegularize_h6e95ff9d_0_8", "__VdfgRegularize_h6e95ff9d_0_8", VLCU_MEMBER_STATE, 4, 0, 1, 24, 23, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_8},
        {"__VdfgRegularize_h6e95ff9d_0_9", "__VdfgRegularize_h6e95ff9d_0_9", VLCU_MEMBER_STATE, 4, 0, 1, 24, 23, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_9},
        {"__VdfgRegularize_h6e95ff9d_0_10", "__VdfgRegularize_h6e95ff9d_0_10", VLCU_MEMBER_STATE, 4, 0, 1, 24, 23, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_10},
        {"__VdfgRegularize_h6e95ff9d_0_11", "__VdfgRegularize_h6e95ff9d_0_11", VLCU_MEMBER_STATE, 4, 0, 1, 24, 23, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_11},
        {"__VdfgRegularize_h6e95ff9d_0_12", "__VdfgRegularize_h6e95ff9d_0_12", VLCU_MEMBER_STATE, 4, 0, 1, 24, 23, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_12},
        {"__VdfgRegularize_h6e95ff9d_0_13", "__VdfgRegularize_h6e95ff9d_0_13", VLCU_MEMBER_STATE, 4, 0, 1, 24, 23, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_13},
        {"__VdfgRegularize_h6e95ff9d_0_14", "__VdfgRegularize_h6e95ff9d_0_14", VLCU_MEMBER_STATE, 4, 0, 1, 24, 23, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_14},
        {"__VdfgRegularize_h6e95ff9d_0_15", "__VdfgRegularize_h6e95ff9d_0_15", VLCU_MEMBER_STATE, 4, 0, 1, 24, 23, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_15},
        {"__VdfgRegularize_h6e95ff9d_0_16", "__VdfgRegularize_h6e95ff9d_0_16", VLCU_MEMBER_STATE, 4, 0, 1, 24, 23, 0, T::OFF___VdfgRegularize_h6e95ff9d_0_16},
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
    static const uint32_t saveOrder[] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 36, 37, 38, 39, 40, 41, 42, 43, 44, 45, 46, 47, 48, 49, 50, 51, 52, 53, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "t", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 2, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 54, 0, 0x77787360b8d7627fULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt_case_deep__cuda VLCU_TESTBENCHES);
