// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt_math_signed"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt_math_signed___024root final {
    CData/*0:0*/ clk;
    CData/*4:0*/ t__DOT__b;
    CData/*0:0*/ t__DOT__w97__DOT__i1;
    CData/*0:0*/ t__DOT__w96__DOT__i1;
    CData/*0:0*/ t__DOT__w95__DOT__i1;
    CData/*0:0*/ t__DOT__w65__DOT__i1;
    CData/*0:0*/ t__DOT__w64__DOT__i1;
    CData/*0:0*/ t__DOT__w63__DOT__i1;
    CData/*0:0*/ t__DOT__w33__DOT__i1;
    CData/*0:0*/ t__DOT__w32__DOT__i1;
    CData/*0:0*/ t__DOT__w31__DOT__i1;
    CData/*0:0*/ t__DOT__w1__DOT__i1;
    CData/*0:0*/ __VstlFirstIteration;
    CData/*0:0*/ __VstlPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    SData/*15:0*/ t__DOT__a;
    SData/*15:0*/ t__DOT__sr;
    SData/*15:0*/ t__DOT__srs;
    SData/*15:0*/ t__DOT__sl;
    SData/*15:0*/ t__DOT__b_s;
    IData/*31:0*/ t__DOT__cyc;
    VlWide<3>/*64:0*/ t__DOT__w97__DOT__i65;
    VlWide<4>/*96:0*/ t__DOT__w97__DOT__i1extp;
    VlWide<4>/*96:0*/ t__DOT__w97__DOT__i63ext;
    VlWide<4>/*96:0*/ t__DOT__w97__DOT__i65ext;
    IData/*31:0*/ t__DOT__w97__DOT__cyc;
    VlWide<3>/*64:0*/ t__DOT__w96__DOT__i65;
    VlWide<3>/*95:0*/ t__DOT__w96__DOT__i1extp;
    VlWide<3>/*95:0*/ t__DOT__w96__DOT__i63ext;
    VlWide<3>/*95:0*/ t__DOT__w96__DOT__i65ext;
    IData/*31:0*/ t__DOT__w96__DOT__cyc;
    VlWide<3>/*64:0*/ t__DOT__w95__DOT__i65;
    VlWide<3>/*94:0*/ t__DOT__w95__DOT__i1extp;
    VlWide<3>/*94:0*/ t__DOT__w95__DOT__i63ext;
    VlWide<3>/*94:0*/ t__DOT__w95__DOT__i65ext;
    IData/*31:0*/ t__DOT__w95__DOT__cyc;
    VlWide<3>/*64:0*/ t__DOT__w65__DOT__i65;
    VlWide<3>/*64:0*/ t__DOT__w65__DOT__i1extp;
    VlWide<3>/*64:0*/ t__DOT__w65__DOT__i63ext;
    IData/*31:0*/ t__DOT__w65__DOT__cyc;
    VlWide<3>/*64:0*/ t__DOT__w64__DOT__i65;
    IData/*31:0*/ t__DOT__w64__DOT__cyc;
    VlWide<3>/*64:0*/ t__DOT__w63__DOT__i65;
    IData/*31:0*/ t__DOT__w63__DOT__cyc;
    VlWide<3>/*64:0*/ t__DOT__w33__DOT__i65;
    IData/*31:0*/ t__DOT__w33__DOT__cyc;
    VlWide<3>/*64:0*/ t__DOT__w32__DOT__i65;
    IData/*31:0*/ t__DOT__w32__DOT__i1extp;
    IData/*31:0*/ t__DOT__w32__DOT__cyc;
    VlWide<3>/*64:0*/ t__DOT__w31__DOT__i65;
    IData/*30:0*/ t__DOT__w31__DOT__i1extp;
    IData/*31:0*/ t__DOT__w31__DOT__cyc;
    VlWide<3>/*64:0*/ t__DOT__w1__DOT__i65;
    IData/*31:0*/ t__DOT__w1__DOT__cyc;
    IData/*31:0*/ __VactIterCount;
    QData/*62:0*/ t__DOT__w97__DOT__i63;
    QData/*62:0*/ t__DOT__w96__DOT__i63;
    QData/*62:0*/ t__DOT__w95__DOT__i63;
    QData/*62:0*/ t__DOT__w65__DOT__i63;
    QData/*62:0*/ t__DOT__w64__DOT__i63;
    QData/*63:0*/ t__DOT__w64__DOT__i1extp;
    QData/*63:0*/ t__DOT__w64__DOT__i63ext;
    QData/*63:0*/ t__DOT__w64__DOT__i65ext;
    QData/*62:0*/ t__DOT__w63__DOT__i63;
    QData/*62:0*/ t__DOT__w63__DOT__i1extp;
    QData/*62:0*/ t__DOT__w63__DOT__i65ext;
    QData/*62:0*/ t__DOT__w33__DOT__i63;
    QData/*32:0*/ t__DOT__w33__DOT__i1extp;
    QData/*32:0*/ t__DOT__w33__DOT__i65ext;
    QData/*62:0*/ t__DOT__w32__DOT__i63;
    QData/*62:0*/ t__DOT__w31__DOT__i63;
    QData/*62:0*/ t__DOT__w1__DOT__i63;
    VlUnpacked<QData/*63:0*/, 1> __VstlTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vt_math_signed__cuda final {
    typedef Vt_math_signed___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 490,
        OFF_t__DOT__b = 491,
        OFF_t__DOT__w97__DOT__i1 = 492,
        OFF_t__DOT__w96__DOT__i1 = 493,
        OFF_t__DOT__w95__DOT__i1 = 494,
        OFF_t__DOT__w65__DOT__i1 = 495,
        OFF_t__DOT__w64__DOT__i1 = 496,
        OFF_t__DOT__w63__DOT__i1 = 497,
        OFF_t__DOT__w33__DOT__i1 = 498,
        OFF_t__DOT__w32__DOT__i1 = 499,
        OFF_t__DOT__w31__DOT__i1 = 500,
        OFF_t__DOT__w1__DOT__i1 = 501,
        OFF___VstlFirstIteration = 502,
        OFF___VstlPhaseResult = 503,
        OFF___Vtrigprevexpr___TOP__clk__0 = 504,
        OFF___VactPhaseResult = 505,
        OFF___VnbaPhaseResult = 506,
        OFF_t__DOT__a = 480,
        OFF_t__DOT__sr = 482,
        OFF_t__DOT__srs = 484,
        OFF_t__DOT__sl = 486,
        OFF_t__DOT__b_s = 488,
        OFF_t__DOT__cyc = 160,
        OFF_t__DOT__w97__DOT__i65 = 164,
        OFF_t__DOT__w97__DOT__i1extp = 176,
        OFF_t__DOT__w97__DOT__i63ext = 192,
        OFF_t__DOT__w97__DOT__i65ext = 208,
        OFF_t__DOT__w97__DOT__cyc = 224,
        OFF_t__DOT__w96__DOT__i65 = 228,
        OFF_t__DOT__w96__DOT__i1extp = 240,
        OFF_t__DOT__w96__DOT__i63ext = 252,
        OFF_t__DOT__w96__DOT__i65ext = 264,
        OFF_t__DOT__w96__DOT__cyc = 276,
        OFF_t__DOT__w95__DOT__i65 = 280,
        OFF_t__DOT__w95__DOT__i1extp = 292,
        OFF_t__DOT__w95__DOT__i63ext = 304,
        OFF_t__DOT__w95__DOT__i65ext = 316,
        OFF_t__DOT__w95__DOT__cyc = 328,
        OFF_t__DOT__w65__DOT__i65 = 332,
        OFF_t__DOT__w65__DOT__i1extp = 344,
        OFF_t__DOT__w65__DOT__i63ext = 356,
        OFF_t__DOT__w65__DOT__cyc = 368,
        OFF_t__DOT__w64__DOT__i65 = 372,
        OFF_t__DOT__w64__DOT__cyc = 384,
        OFF_t__DOT__w63__DOT__i65 = 388,
        OFF_t__DOT__w63__DOT__cyc = 400,
        OFF_t__DOT__w33__DOT__i65 = 404,
        OFF_t__DOT__w33__DOT__cyc = 416,
        OFF_t__DOT__w32__DOT__i65 = 420,
        OFF_t__DOT__w32__DOT__i1extp = 432,
        OFF_t__DOT__w32__DOT__cyc = 436,
        OFF_t__DOT__w31__DOT__i65 = 440,
        OFF_t__DOT__w31__DOT__i1extp = 452,
        OFF_t__DOT__w31__DOT__cyc = 456,
        OFF_t__DOT__w1__DOT__i65 = 460,
        OFF_t__DOT__w1__DOT__cyc = 472,
        OFF___VactIterCount = 476,
        OFF_t__DOT__w97__DOT__i63 = 0,
        OFF_t__DOT__w96__DOT__i63 = 8,
        OFF_t__DOT__w95__DOT__i63 = 16,
        OFF_t__DOT__w65__DOT__i63 = 24,
        OFF_t__DOT__w64__DOT__i63 = 32,
        OFF_t__DOT__w64__DOT__i1extp = 40,
        OFF_t__DOT__w64__DOT__i63ext = 48,
        OFF_t__DOT__w64__DOT__i65ext = 56,
        OFF_t__DOT__w63__DOT__i63 = 64,
        OFF_t__DOT__w63__DOT__i1extp = 72,
        OFF_t__DOT__w63__DOT__i65ext = 80,
        OFF_t__DOT__w33__DOT__i63 = 88,
        OFF_t__DOT__w33__DOT__i1extp = 96,
        OFF_t__DOT__w33__DOT__i65ext = 104,
        OFF_t__DOT__w32__DOT__i63 = 112,
        OFF_t__DOT__w31__DOT__i63 = 120,
        OFF_t__DOT__w1__DOT__i63 = 128,
        OFF___VstlTriggered = 136,
        OFF___VactTriggered = 144,
        OFF___VnbaTriggered = 152,
        kStateBytes = 512,  // S_persist
        kTouchedReadBytes = 497,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 496,  // S_state_wr
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
        s.t__DOT__b = st.ld<VlcuMem8>(OFF_t__DOT__b, i);
        s.t__DOT__w97__DOT__i1 = st.ld<VlcuMem8>(OFF_t__DOT__w97__DOT__i1, i);
        s.t__DOT__w96__DOT__i1 = st.ld<VlcuMem8>(OFF_t__DOT__w96__DOT__i1, i);
        s.t__DOT__w95__DOT__i1 = st.ld<VlcuMem8>(OFF_t__DOT__w95__DOT__i1, i);
        s.t__DOT__w65__DOT__i1 = st.ld<VlcuMem8>(OFF_t__DOT__w65__DOT__i1, i);
        s.t__DOT__w64__DOT__i1 = st.ld<VlcuMem8>(OFF_t__DOT__w64__DOT__i1, i);
        s.t__DOT__w63__DOT__i1 = st.ld<VlcuMem8>(OFF_t__DOT__w63__DOT__i1, i);
        s.t__DOT__w33__DOT__i1 = st.ld<VlcuMem8>(OFF_t__DOT__w33__DOT__i1, i);
        s.t__DOT__w32__DOT__i1 = st.ld<VlcuMem8>(OFF_t__DOT__w32__DOT__i1, i);
        s.t__DOT__w31__DOT__i1 = st.ld<VlcuMem8>(OFF_t__DOT__w31__DOT__i1, i);
        s.t__DOT__w1__DOT__i1 = st.ld<VlcuMem8>(OFF_t__DOT__w1__DOT__i1, i);
        s.__VstlFirstIteration = st.ld<VlcuMem8>(OFF___VstlFirstIteration, i);
        s.__VstlPhaseResult = st.ld<VlcuMem8>(OFF___VstlPhaseResult, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        s.t__DOT__a = st.ld<VlcuMem16>(OFF_t__DOT__a, i);
        s.t__DOT__sr = st.ld<VlcuMem16>(OFF_t__DOT__sr, i);
        s.t__DOT__srs = st.ld<VlcuMem16>(OFF_t__DOT__srs, i);
        s.t__DOT__sl = st.ld<VlcuMem16>(OFF_t__DOT__sl, i);
        s.t__DOT__b_s = st.ld<VlcuMem16>(OFF_t__DOT__b_s, i);
        s.t__DOT__cyc = st.ld<IData>(OFF_t__DOT__cyc, i);
        st.ldW(OFF_t__DOT__w97__DOT__i65, i, s.t__DOT__w97__DOT__i65);
        st.ldW(OFF_t__DOT__w97__DOT__i1extp, i, s.t__DOT__w97__DOT__i1extp);
        st.ldW(OFF_t__DOT__w97__DOT__i63ext, i, s.t__DOT__w97__DOT__i63ext);
        st.ldW(OFF_t__DOT__w97__DOT__i65ext, i, s.t__DOT__w97__DOT__i65ext);
        s.t__DOT__w97__DOT__cyc = st.ld<IData>(OFF_t__DOT__w97__DOT__cyc, i);
        st.ldW(OFF_t__DOT__w96__DOT__i65, i, s.t__DOT__w96__DOT__i65);
        st.ldW(OFF_t__DOT__w96__DOT__i1extp, i, s.t__DOT__w96__DOT__i1extp);
        st.ldW(OFF_t__DOT__w96__DOT__i63ext, i, s.t__DOT__w96__DOT__i63ext);
        st.ldW(OFF_t__DOT__w96__DOT__i65ext, i, s.t__DOT__w96__DOT__i65ext);
        s.t__DOT__w96__DOT__cyc = st.ld<IData>(OFF_t__DOT__w96__DOT__cyc, i);
        st.ldW(OFF_t__DOT__w95__DOT__i65, i, s.t__DOT__w95__DOT__i65);
        st.ldW(OFF_t__DOT__w95__DOT__i1extp, i, s.t__DOT__w95__DOT__i1extp);
        st.ldW(OFF_t__DOT__w95__DOT__i63ext, i, s.t__DOT__w95__DOT__i63ext);
        st.ldW(OFF_t__DOT__w95__DOT__i65ext, i, s.t__DOT__w95__DOT__i65ext);
        s.t__DOT__w95__DOT__cyc = st.ld<IData>(OFF_t__DOT__w95__DOT__cyc, i);
        st.ldW(OFF_t__DOT__w65__DOT__i65, i, s.t__DOT__w65__DOT__i65);
        st.ldW(OFF_t__DOT__w65__DOT__i1extp, i, s.t__DOT__w65__DOT__i1extp);
        st.ldW(OFF_t__DOT__w65__DOT__i63ext, i, s.t__DOT__w65__DOT__i63ext);
        s.t__DOT__w65__DOT__cyc = st.ld<IData>(OFF_t__DOT__w65__DOT__cyc, i);
        st.ldW(OFF_t__DOT__w64__DOT__i65, i, s.t__DOT__w64__DOT__i65);
        s.t__DOT__w64__DOT__cyc = st.ld<IData>(OFF_t__DOT__w64__DOT__cyc, i);
        st.ldW(OFF_t__DOT__w63__DOT__i65, i, s.t__DOT__w63__DOT__i65);
        s.t__DOT__w63__DOT__cyc = st.ld<IData>(OFF_t__DOT__w63__DOT__cyc, i);
        st.ldW(OFF_t__DOT__w33__DOT__i65, i, s.t__DOT__w33__DOT__i65);
        s.t__DOT__w33__DOT__cyc = st.ld<IData>(OFF_t__DOT__w33__DOT__cyc, i);
        st.ldW(OFF_t__DOT__w32__DOT__i65, i, s.t__DOT__w32__DOT__i65);
        s.t__DOT__w32__DOT__i1extp = st.ld<IData>(OFF_t__DOT__w32__DOT__i1extp, i);
        s.t__DOT__w32__DOT__cyc = st.ld<IData>(OFF_t__DOT__w32__DOT__cyc, i);
        st.ldW(OFF_t__DOT__w31__DOT__i65, i, s.t__DOT__w31__DOT__i65);
        s.t__DOT__w31__DOT__i1extp = st.ld<IData>(OFF_t__DOT__w31__DOT__i1extp, i);
        s.t__DOT__w31__DOT__cyc = st.ld<IData>(OFF_t__DOT__w31__DOT__cyc, i);
        st.ldW(OFF_t__DOT__w1__DOT__i65, i, s.t__DOT__w1__DOT__i65);
        s.t__DOT__w1__DOT__cyc = st.ld<IData>(OFF_t__DOT__w1__DOT__cyc, i);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        s.t__DOT__w97__DOT__i63 = st.ld<QData>(OFF_t__DOT__w97__DOT__i63, i);
        s.t__DOT__w96__DOT__i63 = st.ld<QData>(OFF_t__DOT__w96__DOT__i63, i);
        s.t__DOT__w95__DOT__i63 = st.ld<QData>(OFF_t__DOT__w95__DOT__i63, i);
        s.t__DOT__w65__DOT__i63 = st.ld<QData>(OFF_t__DOT__w65__DOT__i63, i);
        s.t__DOT__w64__DOT__i63 = st.ld<QData>(OFF_t__DOT__w64__DOT__i63, i);
        s.t__DOT__w64__DOT__i1extp = st.ld<QData>(OFF_t__DOT__w64__DOT__i1extp, i);
        s.t__DOT__w64__DOT__i63ext = st.ld<QData>(OFF_t__DOT__w64__DOT__i63ext, i);
        s.t__DOT__w64__DOT__i65ext = st.ld<QData>(OFF_t__DOT__w64__DOT__i65ext, i);
        s.t__DOT__w63__DOT__i63 = st.ld<QData>(OFF_t__DOT__w63__DOT__i63, i);
        s.t__DOT__w63__DOT__i1extp = st.ld<QData>(OFF_t__DOT__w63__DOT__i1extp, i);
        s.t__DOT__w63__DOT__i65ext = st.ld<QData>(OFF_t__DOT__w63__DOT__i65ext, i);
        s.t__DOT__w33__DOT__i63 = st.ld<QData>(OFF_t__DOT__w33__DOT__i63, i);
        s.t__DOT__w33__DOT__i1extp = st.ld<QData>(OFF_t__DOT__w33__DOT__i1extp, i);
        s.t__DOT__w33__DOT__i65ext = st.ld<QData>(OFF_t__DOT__w33__DOT__i65ext, i);
        s.t__DOT__w32__DOT__i63 = st.ld<QData>(OFF_t__DOT__w32__DOT__i63, i);
        s.t__DOT__w31__DOT__i63 = st.ld<QData>(OFF_t__DOT__w31__DOT__i63, i);
        s.t__DOT__w1__DOT__i63 = st.ld<QData>(OFF_t__DOT__w1__DOT__i63, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VstlTriggered[e] = st.ldE<QData>(OFF___VstlTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.st<VlcuMem8>(OFF_t__DOT__b, i, s.t__DOT__b);
        st.st<VlcuMem8>(OFF_t__DOT__w97__DOT__i1, i, s.t__DOT__w97__DOT__i1);
        st.st<VlcuMem8>(OFF_t__DOT__w96__DOT__i1, i, s.t__DOT__w96__DOT__i1);
        st.st<VlcuMem8>(OFF_t__DOT__w95__DOT__i1, i, s.t__DOT__w95__DOT__i1);
        st.st<VlcuMem8>(OFF_t__DOT__w65__DOT__i1, i, s.t__DOT__w65__DOT__i1);
        st.st<VlcuMem8>(OFF_t__DOT__w64__DOT__i1, i, s.t__DOT__w64__DOT__i1);
        st.st<VlcuMem8>(OFF_t__DOT__w63__DOT__i1, i, s.t__DOT__w63__DOT__i1);
        st.st<VlcuMem8>(OFF_t__DOT__w33__DOT__i1, i, s.t__DOT__w33__DOT__i1);
        st.st<VlcuMem8>(OFF_t__DOT__w32__DOT__i1, i, s.t__DOT__w32__DOT__i1);
        st.st<VlcuMem8>(OFF_t__DOT__w31__DOT__i1, i, s.t__DOT__w31__DOT__i1);
        st.st<VlcuMem8>(OFF_t__DOT__w1__DOT__i1, i, s.t__DOT__w1__DOT__i1);
        st.st<VlcuMem8>(OFF___VstlFirstIteration, i, s.__VstlFirstIteration);
        st.st<VlcuMem8>(OFF___VstlPhaseResult, i, s.__VstlPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.st<VlcuMem16>(OFF_t__DOT__a, i, s.t__DOT__a);
        st.st<VlcuMem16>(OFF_t__DOT__sr, i, s.t__DOT__sr);
        st.st<VlcuMem16>(OFF_t__DOT__srs, i, s.t__DOT__srs);
        st.st<VlcuMem16>(OFF_t__DOT__sl, i, s.t__DOT__sl);
        st.st<VlcuMem16>(OFF_t__DOT__b_s, i, s.t__DOT__b_s);
        st.st<IData>(OFF_t__DOT__cyc, i, s.t__DOT__cyc);
        st.stW(OFF_t__DOT__w97__DOT__i65, i, s.t__DOT__w97__DOT__i65);
        st.stW(OFF_t__DOT__w97__DOT__i1extp, i, s.t__DOT__w97__DOT__i1extp);
        st.stW(OFF_t__DOT__w97__DOT__i63ext, i, s.t__DOT__w97__DOT__i63ext);
        st.stW(OFF_t__DOT__w97__DOT__i65ext, i, s.t__DOT__w97__DOT__i65ext);
        st.st<IData>(OFF_t__DOT__w97__DOT__cyc, i, s.t__DOT__w97__DOT__cyc);
        st.stW(OFF_t__DOT__w96__DOT__i65, i, s.t__DOT__w96__DOT__i65);
        st.stW(OFF_t__DOT__w96__DOT__i1extp, i, s.t__DOT__w96__DOT__i1extp);
        st.stW(OFF_t__DOT__w96__DOT__i63ext, i, s.t__DOT__w96__DOT__i63ext);
        st.stW(OFF_t__DOT__w96__DOT__i65ext, i, s.t__DOT__w96__DOT__i65ext);
        st.st<IData>(OFF_t__DOT__w96__DOT__cyc, i, s.t__DOT__w96__DOT__cyc);
        st.stW(OFF_t__DOT__w95__DOT__i65, i, s.t__DOT__w95__DOT__i65);
        st.stW(OFF_t__DOT__w95__DOT__i1extp, i, s.t__DOT__w95__DOT__i1extp);
        st.stW(OFF_t__DOT__w95__DOT__i63ext, i, s.t__DOT__w95__DOT__i63ext);
        st.stW(OFF_t__DOT__w95__DOT__i65ext, i, s.t__DOT__w95__DOT__i65ext);
        st.st<IData>(OFF_t__DOT__w95__DOT__cyc, i, s.t__DOT__w95__DOT__cyc);
        st.stW(OFF_t__DOT__w65__DOT__i65, i, s.t__DOT__w65__DOT__i65);
        st.stW(OFF_t__DOT__w65__DOT__i1extp, i, s.t__DOT__w65__DOT__i1extp);
        st.stW(OFF_t__DOT__w65__DOT__i63ext, i, s.t__DOT__w65__DOT__i63ext);
        st.st<IData>(OFF_t__DOT__w65__DOT__cyc, i, s.t__DOT__w65__DOT__cyc);
        st.stW(OFF_t__DOT__w64__DOT__i65, i, s.t__DOT__w64__DOT__i65);
        st.st<IData>(OFF_t__DOT__w64__DOT__cyc, i, s.t__DOT__w64__DOT__cyc);
        st.stW(OFF_t__DOT__w63__DOT__i65, i, s.t__DOT__w63__DOT__i65);
        st.st<IData>(OFF_t__DOT__w63__DOT__cyc, i, s.t__DOT__w63__DOT__cyc);
        st.stW(OFF_t__DOT__w33__DOT__i65, i, s.t__DOT__w33__DOT__i65);
        st.st<IData>(OFF_t__DOT__w33__DOT__cyc, i, s.t__DOT__w33__DOT__cyc);
        st.stW(OFF_t__DOT__w32__DOT__i65, i, s.t__DOT__w32__DOT__i65);
        st.st<IData>(OFF_t__DOT__w32__DOT__i1extp, i, s.t__DOT__w32__DOT__i1extp);
        st.st<IData>(OFF_t__DOT__w32__DOT__cyc, i, s.t__DOT__w32__DOT__cyc);
        st.stW(OFF_t__DOT__w31__DOT__i65, i, s.t__DOT__w31__DOT__i65);
        st.st<IData>(OFF_t__DOT__w31__DOT__i1extp, i, s.t__DOT__w31__DOT__i1extp);
        st.st<IData>(OFF_t__DOT__w31__DOT__cyc, i, s.t__DOT__w31__DOT__cyc);
        st.stW(OFF_t__DOT__w1__DOT__i65, i, s.t__DOT__w1__DOT__i65);
        st.st<IData>(OFF_t__DOT__w1__DOT__cyc, i, s.t__DOT__w1__DOT__cyc);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
        st.st<QData>(OFF_t__DOT__w97__DOT__i63, i, s.t__DOT__w97__DOT__i63);
        st.st<QData>(OFF_t__DOT__w96__DOT__i63, i, s.t__DOT__w96__DOT__i63);
        st.st<QData>(OFF_t__DOT__w95__DOT__i63, i, s.t__DOT__w95__DOT__i63);
        st.st<QData>(OFF_t__DOT__w65__DOT__i63, i, s.t__DOT__w65__DOT__i63);
        st.st<QData>(OFF_t__DOT__w64__DOT__i63, i, s.t__DOT__w64__DOT__i63);
        st.st<QData>(OFF_t__DOT__w64__DOT__i1extp, i, s.t__DOT__w64__DOT__i1extp);
        st.st<QData>(OFF_t__DOT__w64__DOT__i63ext, i, s.t__DOT__w64__DOT__i63ext);
        st.st<QData>(OFF_t__DOT__w64__DOT__i65ext, i, s.t__DOT__w64__DOT__i65ext);
        st.st<QData>(OFF_t__DOT__w63__DOT__i63, i, s.t__DOT__w63__DOT__i63);
        st.st<QData>(OFF_t__DOT__w63__DOT__i1extp, i, s.t__DOT__w63__DOT__i1extp);
        st.st<QData>(OFF_t__DOT__w63__DOT__i65ext, i, s.t__DOT__w63__DOT__i65ext);
        st.st<QData>(OFF_t__DOT__w33__DOT__i63, i, s.t__DOT__w33__DOT__i63);
        st.st<QData>(OFF_t__DOT__w33__DOT__i1extp, i, s.t__DOT__w33__DOT__i1extp);
        st.st<QData>(OFF_t__DOT__w33__DOT__i65ext, i, s.t__DOT__w33__DOT__i65ext);
        st.st<QData>(OFF_t__DOT__w32__DOT__i63, i, s.t__DOT__w32__DOT__i63);
        st.st<QData>(OFF_t__DOT__w31__DOT__i63, i, s.t__DOT__w31__DOT__i63);
        st.st<QData>(OFF_t__DOT__w1__DOT__i63, i, s.t__DOT__w1__DOT__i63);
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

    VLCU_D static void Vt_math_signed___024root___eval_static(Vt_math_signed___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
    }

    VLCU_D static void Vt_math_signed___024root___eval_initial(Vt_math_signed___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        {
            // Inlined CFunc: _eval_initial__TOP
            vlSelfRef.t__DOT__cyc = 0U;
            vlSelfRef.t__DOT__w97__DOT__cyc = 0U;
            vlSelfRef.t__DOT__w96__DOT__cyc = 0U;
            vlSelfRef.t__DOT__w95__DOT__cyc = 0U;
            vlSelfRef.t__DOT__w65__DOT__cyc = 0U;
            vlSelfRef.t__DOT__w64__DOT__cyc = 0U;
            vlSelfRef.t__DOT__w63__DOT__cyc = 0U;
            vlSelfRef.t__DOT__w33__DOT__cyc = 0U;
            vlSelfRef.t__DOT__w32__DOT__cyc = 0U;
            vlSelfRef.t__DOT__w31__DOT__cyc = 0U;
            vlSelfRef.t__DOT__w1__DOT__cyc = 0U;
        }
    }

    VLCU_D static void Vt_math_signed___024root___eval_final(Vt_math_signed___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_math_signed___024root___eval_settle(Vt_math_signed___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt_math_signed___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
#endif
                VL_FATAL_MT("/root/reference/test_regress/t/t_math_signed.v", 7, "", "DIDNOTCONVERGE: Settle region did not converge after '--converge-limit' of 10000 tries");
            }
            __VstlIterCount = ((IData)(1U) + __VstlIterCount);
            vlSelfRef.__VstlPhaseResult = Vt_math_signed___024root___eval_phase__stl(vlSelf, vlCtx);
            vlSelfRef.__VstlFirstIteration = 0U;
        } while (vlSelfRef.__VstlPhaseResult);
    }

    VLCU_D static bool Vt_math_signed___024root___trigger_anySet__stl(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt_math_signed___024root___stl_sequent__TOP__0(Vt_math_signed___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        VlWide<4>/*127:0*/ __Vtemp_2;
        VlWide<4>/*127:0*/ __Vtemp_4;
        VlWide<4>/*127:0*/ __Vtemp_6;
        VlWide<3>/*95:0*/ __Vtemp_8;
        VlWide<3>/*95:0*/ __Vtemp_10;
        VlWide<3>/*95:0*/ __Vtemp_12;
        VlWide<3>/*95:0*/ __Vtemp_14;
        VlWide<3>/*95:0*/ __Vtemp_16;
        // Body
        vlSelfRef.t__DOT__b_s = (0x0000ffffU & VL_SHIFTRS_III(16,16,32, 
                                                              (0x0000ffffU 
                                                               & VL_EXTENDS_II(16,5, (IData)(vlSelfRef.t__DOT__b))), 4U));
        VL_EXTENDS_WI(97,1, __Vtemp_2, (IData)(vlSelfRef.t__DOT__w97__DOT__i1));
        vlSelfRef.t__DOT__w97__DOT__i1extp[0U] = __Vtemp_2[0U];
        vlSelfRef.t__DOT__w97__DOT__i1extp[1U] = __Vtemp_2[1U];
        vlSelfRef.t__DOT__w97__DOT__i1extp[2U] = __Vtemp_2[2U];
        vlSelfRef.t__DOT__w97__DOT__i1extp[3U] = (1U 
                                                  & __Vtemp_2[3U]);
        VL_EXTENDS_WQ(97,63, __Vtemp_4, vlSelfRef.t__DOT__w97__DOT__i63);
        vlSelfRef.t__DOT__w97__DOT__i63ext[0U] = __Vtemp_4[0U];
        vlSelfRef.t__DOT__w97__DOT__i63ext[1U] = __Vtemp_4[1U];
        vlSelfRef.t__DOT__w97__DOT__i63ext[2U] = __Vtemp_4[2U];
        vlSelfRef.t__DOT__w97__DOT__i63ext[3U] = (1U 
                                                  & __Vtemp_4[3U]);
        VL_EXTENDS_WW(97,65, __Vtemp_6, vlSelfRef.t__DOT__w97__DOT__i65);
        vlSelfRef.t__DOT__w97__DOT__i65ext[0U] = __Vtemp_6[0U];
        vlSelfRef.t__DOT__w97__DOT__i65ext[1U] = __Vtemp_6[1U];
        vlSelfRef.t__DOT__w97__DOT__i65ext[2U] = __Vtemp_6[2U];
        vlSelfRef.t__DOT__w97__DOT__i65ext[3U] = (1U 
                                                  & __Vtemp_6[3U]);
        VL_EXTENDS_WI(96,1, vlSelfRef.t__DOT__w96__DOT__i1extp, (IData)(vlSelfRef.t__DOT__w96__DOT__i1));
        VL_EXTENDS_WQ(96,63, vlSelfRef.t__DOT__w96__DOT__i63ext, vlSelfRef.t__DOT__w96__DOT__i63);
        VL_EXTENDS_WW(96,65, vlSelfRef.t__DOT__w96__DOT__i65ext, vlSelfRef.t__DOT__w96__DOT__i65);
        VL_EXTENDS_WI(95,1, __Vtemp_8, (IData)(vlSelfRef.t__DOT__w95__DOT__i1));
        vlSelfRef.t__DOT__w95__DOT__i1extp[0U] = __Vtemp_8[0U];
        vlSelfRef.t__DOT__w95__DOT__i1extp[1U] = __Vtemp_8[1U];
        vlSelfRef.t__DOT__w95__DOT__i1extp[2U] = (0x7fffffffU 
                                                  & __Vtemp_8[2U]);
        VL_EXTENDS_WQ(95,63, __Vtemp_10, vlSelfRef.t__DOT__w95__DOT__i63);
        vlSelfRef.t__DOT__w95__DOT__i63ext[0U] = __Vtemp_10[0U];
        vlSelfRef.t__DOT__w95__DOT__i63ext[1U] = __Vtemp_10[1U];
        vlSelfRef.t__DOT__w95__DOT__i63ext[2U] = (0x7fffffffU 
                                                  & __Vtemp_10[2U]);
        VL_EXTENDS_WW(95,65, __Vtemp_12, vlSelfRef.t__DOT__w95__DOT__i65);
        vlSelfRef.t__DOT__w95__DOT__i65ext[0U] = __Vtemp_12[0U];
        vlSelfRef.t__DOT__w95__DOT__i65ext[1U] = __Vtemp_12[1U];
        vlSelfRef.t__DOT__w95__DOT__i65ext[2U] = (0x7fffffffU 
                                                  & __Vtemp_12[2U]);
        VL_EXTENDS_WI(65,1, __Vtemp_14, (IData)(vlSelfRef.t__DOT__w65__DOT__i1));
        vlSelfRef.t__DOT__w65__DOT__i1extp[0U] = __Vtemp_14[0U];
        vlSelfRef.t__DOT__w65__DOT__i1extp[1U] = __Vtemp_14[1U];
        vlSelfRef.t__DOT__w65__DOT__i1extp[2U] = (1U 
                                                  & __Vtemp_14[2U]);
        VL_EXTENDS_WQ(65,63, __Vtemp_16, vlSelfRef.t__DOT__w65__DOT__i63);
        vlSelfRef.t__DOT__w65__DOT__i63ext[0U] = __Vtemp_16[0U];
        vlSelfRef.t__DOT__w65__DOT__i63ext[1U] = __Vtemp_16[1U];
        vlSelfRef.t__DOT__w65__DOT__i63ext[2U] = (1U 
                                                  & __Vtemp_16[2U]);
        vlSelfRef.t__DOT__w64__DOT__i1extp = VL_EXTENDS_QI(64,1, (IData)(vlSelfRef.t__DOT__w64__DOT__i1));
        vlSelfRef.t__DOT__w64__DOT__i63ext = VL_EXTENDS_QQ(64,63, vlSelfRef.t__DOT__w64__DOT__i63);
        vlSelfRef.t__DOT__w64__DOT__i65ext = (((QData)((IData)(vlSelfRef.t__DOT__w64__DOT__i65[1U])) 
                                               << 0x00000020U) 
                                              | (QData)((IData)(vlSelfRef.t__DOT__w64__DOT__i65[0U])));
        vlSelfRef.t__DOT__w63__DOT__i1extp = (0x7fffffffffffffffULL 
                                              & VL_EXTENDS_QI(63,1, (IData)(vlSelfRef.t__DOT__w63__DOT__i1)));
        vlSelfRef.t__DOT__w63__DOT__i65ext = (0x7fffffffffffffffULL 
                                              & (((QData)((IData)(vlSelfRef.t__DOT__w63__DOT__i65[1U])) 
                                                  << 0x00000020U) 
                                                 | (QData)((IData)(vlSelfRef.t__DOT__w63__DOT__i65[0U]))));
        vlSelfRef.t__DOT__w33__DOT__i1extp = (0x00000001ffffffffULL 
                                              & VL_EXTENDS_QI(33,1, (IData)(vlSelfRef.t__DOT__w33__DOT__i1)));
        vlSelfRef.t__DOT__w33__DOT__i65ext = (0x00000001ffffffffULL 
                                              & (((QData)((IData)(vlSelfRef.t__DOT__w33__DOT__i65[1U])) 
                                                  << 0x00000020U) 
                                                 | (QData)((IData)(vlSelfRef.t__DOT__w33__DOT__i65[0U]))));
        vlSelfRef.t__DOT__w32__DOT__i1extp = VL_EXTENDS_II(32,1, (IData)(vlSelfRef.t__DOT__w32__DOT__i1));
        vlSelfRef.t__DOT__w31__DOT__i1extp = (0x7fffffffU 
                                              & VL_EXTENDS_II(31,1, (IData)(vlSelfRef.t__DOT__w31__DOT__i1)));
        vlSelfRef.t__DOT__sr = ((IData)(vlSelfRef.t__DOT__a) 
                                >> (IData)(vlSelfRef.t__DOT__b));
        vlSelfRef.t__DOT__srs = (0x0000ffffU & VL_SHIFTRS_III(16,16,5, (IData)(vlSelfRef.t__DOT__a), (IData)(vlSelfRef.t__DOT__b)));
        vlSelfRef.t__DOT__sl = (0x0000ffffU & ((IData)(vlSelfRef.t__DOT__a) 
                                               << (IData)(vlSelfRef.t__DOT__b)));
    }

    VLCU_D static bool Vt_math_signed___024root___eval_phase__stl(Vt_math_signed___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt_math_signed___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
        }
#endif
        __VstlExecute = Vt_math_signed___024root___trigger_anySet__stl(vlSelfRef.__VstlTriggered);
        if (__VstlExecute) {
            {
                // Inlined CFunc: _eval_stl
                if ((1ULL & vlSelfRef.__VstlTriggered[0U])) {
                    Vt_math_signed___024root___stl_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
        }
        return (__VstlExecute);
    }

    VLCU_D static bool Vt_math_signed___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt_math_signed___024root___nba_sequent__TOP__0(Vt_math_signed___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        IData/*31:0*/ __Vdly__t__DOT__cyc;
        __Vdly__t__DOT__cyc = 0;
        CData/*4:0*/ __Vdly__t__DOT__b;
        __Vdly__t__DOT__b = 0;
        IData/*31:0*/ __Vdly__t__DOT__w97__DOT__cyc;
        __Vdly__t__DOT__w97__DOT__cyc = 0;
        IData/*31:0*/ __Vdly__t__DOT__w96__DOT__cyc;
        __Vdly__t__DOT__w96__DOT__cyc = 0;
        IData/*31:0*/ __Vdly__t__DOT__w95__DOT__cyc;
        __Vdly__t__DOT__w95__DOT__cyc = 0;
        IData/*31:0*/ __Vdly__t__DOT__w65__DOT__cyc;
        __Vdly__t__DOT__w65__DOT__cyc = 0;
        VlWide<3>/*64:0*/ __Vdly__t__DOT__w65__DOT__i65;
        VL_ZERO_W(65, __Vdly__t__DOT__w65__DOT__i65);
        IData/*31:0*/ __Vdly__t__DOT__w64__DOT__cyc;
        __Vdly__t__DOT__w64__DOT__cyc = 0;
        IData/*31:0*/ __Vdly__t__DOT__w63__DOT__cyc;
        __Vdly__t__DOT__w63__DOT__cyc = 0;
        QData/*62:0*/ __Vdly__t__DOT__w63__DOT__i63;
        __Vdly__t__DOT__w63__DOT__i63 = 0;
        IData/*31:0*/ __Vdly__t__DOT__w33__DOT__cyc;
        __Vdly__t__DOT__w33__DOT__cyc = 0;
        QData/*62:0*/ __Vdly__t__DOT__w33__DOT__i63;
        __Vdly__t__DOT__w33__DOT__i63 = 0;
        IData/*31:0*/ __Vdly__t__DOT__w32__DOT__cyc;
        __Vdly__t__DOT__w32__DOT__cyc = 0;
        QData/*62:0*/ __Vdly__t__DOT__w32__DOT__i63;
        __Vdly__t__DOT__w32__DOT__i63 = 0;
        VlWide<3>/*64:0*/ __Vdly__t__DOT__w32__DOT__i65;
        VL_ZERO_W(65, __Vdly__t__DOT__w32__DOT__i65);
        IData/*31:0*/ __Vdly__t__DOT__w31__DOT__cyc;
        __Vdly__t__DOT__w31__DOT__cyc = 0;
        QData/*62:0*/ __Vdly__t__DOT__w31__DOT__i63;
        __Vdly__t__DOT__w31__DOT__i63 = 0;
        VlWide<3>/*64:0*/ __Vdly__t__DOT__w31__DOT__i65;
        VL_ZERO_W(65, __Vdly__t__DOT__w31__DOT__i65);
        IData/*31:0*/ __Vdly__t__DOT__w1__DOT__cyc;
        __Vdly__t__DOT__w1__DOT__cyc = 0;
        CData/*0:0*/ __Vdly__t__DOT__w1__DOT__i1;
        __Vdly__t__DOT__w1__DOT__i1 = 0;
        QData/*62:0*/ __Vdly__t__DOT__w1__DOT__i63;
        __Vdly__t__DOT__w1__DOT__i63 = 0;
        VlWide<3>/*64:0*/ __Vdly__t__DOT__w1__DOT__i65;
        VL_ZERO_W(65, __Vdly__t__DOT__w1__DOT__i65);
        VlWide<3>/*95:0*/ __Vtemp_68;
        VlWide<3>/*95:0*/ __Vtemp_70;
        VlWide<4>/*127:0*/ __Vtemp_72;
        VlWide<4>/*127:0*/ __Vtemp_74;
        VlWide<4>/*127:0*/ __Vtemp_76;
        VlWide<3>/*95:0*/ __Vtemp_78;
        VlWide<3>/*95:0*/ __Vtemp_80;
        VlWide<3>/*95:0*/ __Vtemp_82;
        // Body
        __Vdly__t__DOT__w1__DOT__cyc = vlSelfRef.t__DOT__w1__DOT__cyc;
        __Vdly__t__DOT__w1__DOT__i1 = vlSelfRef.t__DOT__w1__DOT__i1;
        __Vdly__t__DOT__w1__DOT__i63 = vlSelfRef.t__DOT__w1__DOT__i63;
        __Vdly__t__DOT__w1__DOT__i65[0U] = vlSelfRef.t__DOT__w1__DOT__i65[0U];
        __Vdly__t__DOT__w1__DOT__i65[1U] = vlSelfRef.t__DOT__w1__DOT__i65[1U];
        __Vdly__t__DOT__w1__DOT__i65[2U] = vlSelfRef.t__DOT__w1__DOT__i65[2U];
        __Vdly__t__DOT__w32__DOT__cyc = vlSelfRef.t__DOT__w32__DOT__cyc;
        __Vdly__t__DOT__w32__DOT__i63 = vlSelfRef.t__DOT__w32__DOT__i63;
        __Vdly__t__DOT__w32__DOT__i65[0U] = vlSelfRef.t__DOT__w32__DOT__i65[0U];
        __Vdly__t__DOT__w32__DOT__i65[1U] = vlSelfRef.t__DOT__w32__DOT__i65[1U];
        __Vdly__t__DOT__w32__DOT__i65[2U] = vlSelfRef.t__DOT__w32__DOT__i65[2U];
        __Vdly__t__DOT__w31__DOT__cyc = vlSelfRef.t__DOT__w31__DOT__cyc;
        __Vdly__t__DOT__w31__DOT__i63 = vlSelfRef.t__DOT__w31__DOT__i63;
        __Vdly__t__DOT__w31__DOT__i65[0U] = vlSelfRef.t__DOT__w31__DOT__i65[0U];
        __Vdly__t__DOT__w31__DOT__i65[1U] = vlSelfRef.t__DOT__w31__DOT__i65[1U];
        __Vdly__t__DOT__w31__DOT__i65[2U] = vlSelfRef.t__DOT__w31__DOT__i65[2U];
        __Vdly__t__DOT__w65__DOT__cyc = vlSelfRef.t__DOT__w65__DOT__cyc;
        __Vdly__t__DOT__w65__DOT__i65[0U] = vlSelfRef.t__DOT__w65__DOT__i65[0U];
        __Vdly__t__DOT__w65__DOT__i65[1U] = vlSelfRef.t__DOT__w65__DOT__i65[1U];
        __Vdly__t__DOT__w65__DOT__i65[2U] = vlSelfRef.t__DOT__w65__DOT__i65[2U];
        __Vdly__t__DOT__w63__DOT__cyc = vlSelfRef.t__DOT__w63__DOT__cyc;
        __Vdly__t__DOT__w63__DOT__i63 = vlSelfRef.t__DOT__w63__DOT__i63;
        __Vdly__t__DOT__w33__DOT__cyc = vlSelfRef.t__DOT__w33__DOT__cyc;
        __Vdly__t__DOT__w33__DOT__i63 = vlSelfRef.t__DOT__w33__DOT__i63;
        __Vdly__t__DOT__w97__DOT__cyc = vlSelfRef.t__DOT__w97__DOT__cyc;
        __Vdly__t__DOT__w96__DOT__cyc = vlSelfRef.t__DOT__w96__DOT__cyc;
        __Vdly__t__DOT__w95__DOT__cyc = vlSelfRef.t__DOT__w95__DOT__cyc;
        __Vdly__t__DOT__w64__DOT__cyc = vlSelfRef.t__DOT__w64__DOT__cyc;
        __Vdly__t__DOT__cyc = vlSelfRef.t__DOT__cyc;
        __Vdly__t__DOT__b = vlSelfRef.t__DOT__b;
        __Vdly__t__DOT__w1__DOT__cyc = ((IData)(1U) 
                                        + vlSelfRef.t__DOT__w1__DOT__cyc);
        __Vdly__t__DOT__w1__DOT__i1 = (1U & vlSelfRef.t__DOT__w1__DOT__cyc);
        __Vdly__t__DOT__w1__DOT__i63 = (0x7fffffffffffffffULL 
                                        & (- (QData)((IData)(
                                                             (1U 
                                                              & vlSelfRef.t__DOT__w1__DOT__cyc)))));
        __Vdly__t__DOT__w1__DOT__i65[0U] = (- (IData)(
                                                      (1U 
                                                       & vlSelfRef.t__DOT__w1__DOT__cyc)));
        __Vdly__t__DOT__w1__DOT__i65[1U] = (- (IData)(
                                                      (1U 
                                                       & vlSelfRef.t__DOT__w1__DOT__cyc)));
        __Vdly__t__DOT__w1__DOT__i65[2U] = (1U & (- (IData)(
                                                            (1U 
                                                             & vlSelfRef.t__DOT__w1__DOT__cyc))));
        if ((1U == vlSelfRef.t__DOT__w1__DOT__cyc)) {
            if (VL_UNLIKELY((vlSelfRef.t__DOT__w1__DOT__i1))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 201, "");
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 202, "");
            }
            if (VL_UNLIKELY(((1U & (IData)(vlSelfRef.t__DOT__w1__DOT__i63))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 203, "");
            }
            if (VL_UNLIKELY(((1U & vlSelfRef.t__DOT__w1__DOT__i65[0U])))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 204, "");
            }
        } else if ((2U == vlSelfRef.t__DOT__w1__DOT__cyc)) {
            if (VL_UNLIKELY(((1U & (~ (IData)(vlSelfRef.t__DOT__w1__DOT__i1)))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 207, "");
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 208, "");
            }
            if (VL_UNLIKELY(((1U & (~ (IData)(vlSelfRef.t__DOT__w1__DOT__i63)))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 209, "");
            }
            if (VL_UNLIKELY(((1U & (~ vlSelfRef.t__DOT__w1__DOT__i65[0U]))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 210, "");
            }
        }
        __Vdly__t__DOT__w32__DOT__cyc = ((IData)(1U) 
                                         + vlSelfRef.t__DOT__w32__DOT__cyc);
        vlSelfRef.t__DOT__w32__DOT__i1 = (1U & vlSelfRef.t__DOT__w32__DOT__cyc);
        __Vdly__t__DOT__w32__DOT__i63 = (0x7fffffffffffffffULL 
                                         & (- (QData)((IData)(
                                                              (1U 
                                                               & vlSelfRef.t__DOT__w32__DOT__cyc)))));
        __Vdly__t__DOT__w32__DOT__i65[0U] = (- (IData)(
                                                       (1U 
                                                        & vlSelfRef.t__DOT__w32__DOT__cyc)));
        __Vdly__t__DOT__w32__DOT__i65[1U] = (- (IData)(
                                                       (1U 
                                                        & vlSelfRef.t__DOT__w32__DOT__cyc)));
        __Vdly__t__DOT__w32__DOT__i65[2U] = (1U & (- (IData)(
                                                             (1U 
                                                              & vlSelfRef.t__DOT__w32__DOT__cyc))));
        if ((1U == vlSelfRef.t__DOT__w32__DOT__cyc)) {
            if (VL_UNLIKELY(((0U != vlSelfRef.t__DOT__w32__DOT__i1extp)))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 201, "");
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 202, "");
            }
            if (VL_UNLIKELY(((0U != (IData)(vlSelfRef.t__DOT__w32__DOT__i63))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 203, "");
            }
            if (VL_UNLIKELY(((0U != vlSelfRef.t__DOT__w32__DOT__i65[0U])))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 204, "");
            }
        } else if ((2U == vlSelfRef.t__DOT__w32__DOT__cyc)) {
            if (VL_UNLIKELY(((0xffffffffU != vlSelfRef.t__DOT__w32__DOT__i1extp)))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 207, "");
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 208, "");
            }
            if (VL_UNLIKELY(((0xffffffffU != (IData)(vlSelfRef.t__DOT__w32__DOT__i63))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 209, "");
            }
            if (VL_UNLIKELY(((0xffffffffU != vlSelfRef.t__DOT__w32__DOT__i65[0U])))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 210, "");
            }
        }
        __Vdly__t__DOT__w31__DOT__cyc = ((IData)(1U) 
                                         + vlSelfRef.t__DOT__w31__DOT__cyc);
        vlSelfRef.t__DOT__w31__DOT__i1 = (1U & vlSelfRef.t__DOT__w31__DOT__cyc);
        __Vdly__t__DOT__w31__DOT__i63 = (0x7fffffffffffffffULL 
                                         & (- (QData)((IData)(
                                                              (1U 
                                                               & vlSelfRef.t__DOT__w31__DOT__cyc)))));
        __Vdly__t__DOT__w31__DOT__i65[0U] = (- (IData)(
                                                       (1U 
                                                        & vlSelfRef.t__DOT__w31__DOT__cyc)));
        __Vdly__t__DOT__w31__DOT__i65[1U] = (- (IData)(
                                                       (1U 
                                                        & vlSelfRef.t__DOT__w31__DOT__cyc)));
        __Vdly__t__DOT__w31__DOT__i65[2U] = (1U & (- (IData)(
                                                             (1U 
                                                              & vlSelfRef.t__DOT__w31__DOT__cyc))));
        if ((1U == vlSelfRef.t__DOT__w31__DOT__cyc)) {
            if (VL_UNLIKELY(((0U != vlSelfRef.t__DOT__w31__DOT__i1extp)))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 201, "");
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 202, "");
            }
            if (VL_UNLIKELY(((0U != (0x7fffffffU & (IData)(vlSelfRef.t__DOT__w31__DOT__i63)))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 203, "");
            }
            if (VL_UNLIKELY(((0U != (0x7fffffffU & vlSelfRef.t__DOT__w31__DOT__i65[0U]))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 204, "");
            }
        } else if ((2U == vlSelfRef.t__DOT__w31__DOT__cyc)) {
            if (VL_UNLIKELY(((0x7fffffffU != vlSelfRef.t__DOT__w31__DOT__i1extp)))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 207, "");
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 208, "");
            }
            if (VL_UNLIKELY(((0x7fffffffU != (0x7fffffffU 
                                              & (IData)(vlSelfRef.t__DOT__w31__DOT__i63)))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 209, "");
            }
            if (VL_UNLIKELY(((0x7fffffffU != (0x7fffffffU 
                                              & vlSelfRef.t__DOT__w31__DOT__i65[0U]))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 210, "");
            }
        }
        __Vdly__t__DOT__w65__DOT__cyc = ((IData)(1U) 
                                         + vlSelfRef.t__DOT__w65__DOT__cyc);
        vlSelfRef.t__DOT__w65__DOT__i1 = (1U & vlSelfRef.t__DOT__w65__DOT__cyc);
        vlSelfRef.t__DOT__w65__DOT__i63 = (0x7fffffffffffffffULL 
                                           & (- (QData)((IData)(
                                                                (1U 
                                                                 & vlSelfRef.t__DOT__w65__DOT__cyc)))));
        __Vdly__t__DOT__w65__DOT__i65[0U] = (- (IData)(
                                                       (1U 
                                                        & vlSelfRef.t__DOT__w65__DOT__cyc)));
        __Vdly__t__DOT__w65__DOT__i65[1U] = (- (IData)(
                                                       (1U 
                                                        & vlSelfRef.t__DOT__w65__DOT__cyc)));
        __Vdly__t__DOT__w65__DOT__i65[2U] = (1U & (- (IData)(
                                                             (1U 
                                                              & vlSelfRef.t__DOT__w65__DOT__cyc))));
        if ((1U == vlSelfRef.t__DOT__w65__DOT__cyc)) {
            if (VL_UNLIKELY(((0U != ((vlSelfRef.t__DOT__w65__DOT__i1extp[0U] 
                                      | vlSelfRef.t__DOT__w65__DOT__i1extp[1U]) 
                                     | vlSelfRef.t__DOT__w65__DOT__i1extp[2U]))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 201, "");
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 202, "");
            }
            if (VL_UNLIKELY(((0U != ((vlSelfRef.t__DOT__w65__DOT__i63ext[0U] 
                                      | vlSelfRef.t__DOT__w65__DOT__i63ext[1U]) 
                                     | vlSelfRef.t__DOT__w65__DOT__i63ext[2U]))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 203, "");
            }
            if (VL_UNLIKELY(((0U != ((vlSelfRef.t__DOT__w65__DOT__i65[0U] 
                                      | vlSelfRef.t__DOT__w65__DOT__i65[1U]) 
                                     | vlSelfRef.t__DOT__w65__DOT__i65[2U]))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 204, "");
            }
        } else if ((2U == vlSelfRef.t__DOT__w65__DOT__cyc)) {
            if (VL_UNLIKELY(((0U != (((~ vlSelfRef.t__DOT__w65__DOT__i1extp[0U]) 
                                      | (~ vlSelfRef.t__DOT__w65__DOT__i1extp[1U])) 
                                     | (1U ^ vlSelfRef.t__DOT__w65__DOT__i1extp[2U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 207, "");
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 208, "");
            }
            if (VL_UNLIKELY(((0U != (((~ vlSelfRef.t__DOT__w65__DOT__i63ext[0U]) 
                                      | (~ vlSelfRef.t__DOT__w65__DOT__i63ext[1U])) 
                                     | (1U ^ vlSelfRef.t__DOT__w65__DOT__i63ext[2U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 209, "");
            }
            if (VL_UNLIKELY(((0U != (((~ vlSelfRef.t__DOT__w65__DOT__i65[0U]) 
                                      | (~ vlSelfRef.t__DOT__w65__DOT__i65[1U])) 
                                     | (1U ^ vlSelfRef.t__DOT__w65__DOT__i65[2U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 210, "");
            }
        }
        __Vdly__t__DOT__w63__DOT__cyc = ((IData)(1U) 
                                         + vlSelfRef.t__DOT__w63__DOT__cyc);
        vlSelfRef.t__DOT__w63__DOT__i1 = (1U & vlSelfRef.t__DOT__w63__DOT__cyc);
        __Vdly__t__DOT__w63__DOT__i63 = (0x7fffffffffffffffULL 
                                         & (- (QData)((IData)(
                                                              (1U 
                                                               & vlSelfRef.t__DOT__w63__DOT__cyc)))));
        vlSelfRef.t__DOT__w63__DOT__i65[0U] = (- (IData)(
                                                         (1U 
                                                          & vlSelfRef.t__DOT__w63__DOT__cyc)));
        vlSelfRef.t__DOT__w63__DOT__i65[1U] = (- (IData)(
                                                         (1U 
                                                          & vlSelfRef.t__DOT__w63__DOT__cyc)));
        vlSelfRef.t__DOT__w63__DOT__i65[2U] = (1U & 
                                               (- (IData)(
                                                          (1U 
                                                           & vlSelfRef.t__DOT__w63__DOT__cyc))));
        if ((1U == vlSelfRef.t__DOT__w63__DOT__cyc)) {
            if (VL_UNLIKELY(((0ULL != vlSelfRef.t__DOT__w63__DOT__i1extp)))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 201, "");
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 202, "");
            }
            if (VL_UNLIKELY(((0ULL != vlSelfRef.t__DOT__w63__DOT__i63)))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 203, "");
            }
            if (VL_UNLIKELY(((0ULL != vlSelfRef.t__DOT__w63__DOT__i65ext)))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 204, "");
            }
        } else if ((2U == vlSelfRef.t__DOT__w63__DOT__cyc)) {
            if (VL_UNLIKELY(((0x7fffffffffffffffULL 
                              != vlSelfRef.t__DOT__w63__DOT__i1extp)))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 207, "");
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 208, "");
            }
            if (VL_UNLIKELY(((0x7fffffffffffffffULL 
                              != vlSelfRef.t__DOT__w63__DOT__i63)))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 209, "");
            }
            if (VL_UNLIKELY(((0x7fffffffffffffffULL 
                              != vlSelfRef.t__DOT__w63__DOT__i65ext)))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 210, "");
            }
        }
        __Vdly__t__DOT__w33__DOT__cyc = ((IData)(1U) 
                                         + vlSelfRef.t__DOT__w33__DOT__cyc);
        vlSelfRef.t__DOT__w33__DOT__i1 = (1U & vlSelfRef.t__DOT__w33__DOT__cyc);
        __Vdly__t__DOT__w33__DOT__i63 = (0x7fffffffffffffffULL 
                                         & (- (QData)((IData)(
                                                              (1U 
                                                               & vlSelfRef.t__DOT__w33__DOT__cyc)))));
        vlSelfRef.t__DOT__w33__DOT__i65[0U] = (- (IData)(
                                                         (1U 
                                                          & vlSelfRef.t__DOT__w33__DOT__cyc)));
        vlSelfRef.t__DOT__w33__DOT__i65[1U] = (- (IData)(
                                                         (1U 
                                                          & vlSelfRef.t__DOT__w33__DOT__cyc)));
        vlSelfRef.t__DOT__w33__DOT__i65[2U] = (1U & 
                                               (- (IData)(
                                                          (1U 
                                                           & vlSelfRef.t__DOT__w33__DOT__cyc))));
        if ((1U == vlSelfRef.t__DOT__w33__DOT__cyc)) {
            if (VL_UNLIKELY(((0ULL != vlSelfRef.t__DOT__w33__DOT__i1extp)))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 201, "");
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 202, "");
            }
            if (VL_UNLIKELY(((0ULL != (0x00000001ffffffffULL 
                                       & vlSelfRef.t__DOT__w33__DOT__i63))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 203, "");
            }
            if (VL_UNLIKELY(((0ULL != vlSelfRef.t__DOT__w33__DOT__i65ext)))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 204, "");
            }
        } else if ((2U == vlSelfRef.t__DOT__w33__DOT__cyc)) {
            if (VL_UNLIKELY(((0x00000001ffffffffULL 
                              != vlSelfRef.t__DOT__w33__DOT__i1extp)))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 207, "");
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 208, "");
            }
            if (VL_UNLIKELY(((0x00000001ffffffffULL 
                              != (0x00000001ffffffffULL 
                                  & vlSelfRef.t__DOT__w33__DOT__i63))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 209, "");
            }
            if (VL_UNLIKELY(((0x00000001ffffffffULL 
                              != vlSelfRef.t__DOT__w33__DOT__i65ext)))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 210, "");
            }
        }
        __Vdly__t__DOT__w97__DOT__cyc = ((IData)(1U) 
                                         + vlSelfRef.t__DOT__w97__DOT__cyc);
        vlSelfRef.t__DOT__w97__DOT__i1 = (1U & vlSelfRef.t__DOT__w97__DOT__cyc);
        vlSelfRef.t__DOT__w97__DOT__i63 = (0x7fffffffffffffffULL 
                                           & (- (QData)((IData)(
                                                                (1U 
                                                                 & vlSelfRef.t__DOT__w97__DOT__cyc)))));
        vlSelfRef.t__DOT__w97__DOT__i65[0U] = (- (IData)(
                                                         (1U 
                                                          & vlSelfRef.t__DOT__w97__DOT__cyc)));
        vlSelfRef.t__DOT__w97__DOT__i65[1U] = (- (IData)(
                                                         (1U 
                                                          & vlSelfRef.t__DOT__w97__DOT__cyc)));
        vlSelfRef.t__DOT__w97__DOT__i65[2U] = (1U & 
                                               (- (IData)(
                                                          (1U 
                                                           & vlSelfRef.t__DOT__w97__DOT__cyc))));
        if ((1U == vlSelfRef.t__DOT__w97__DOT__cyc)) {
            if (VL_UNLIKELY(((0U != (((vlSelfRef.t__DOT__w97__DOT__i1extp[0U] 
                                       | vlSelfRef.t__DOT__w97__DOT__i1extp[1U]) 
                                      | vlSelfRef.t__DOT__w97__DOT__i1extp[2U]) 
                                     | vlSelfRef.t__DOT__w97__DOT__i1extp[3U]))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 201, "");
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 202, "");
            }
            if (VL_UNLIKELY(((0U != (((vlSelfRef.t__DOT__w97__DOT__i63ext[0U] 
                                       | vlSelfRef.t__DOT__w97__DOT__i63ext[1U]) 
                                      | vlSelfRef.t__DOT__w97__DOT__i63ext[2U]) 
                                     | vlSelfRef.t__DOT__w97__DOT__i63ext[3U]))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 203, "");
            }
            if (VL_UNLIKELY(((0U != (((vlSelfRef.t__DOT__w97__DOT__i65ext[0U] 
                                       | vlSelfRef.t__DOT__w97__DOT__i65ext[1U]) 
                                      | vlSelfRef.t__DOT__w97__DOT__i65ext[2U]) 
                                     | vlSelfRef.t__DOT__w97__DOT__i65ext[3U]))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 204, "");
            }
        } else if ((2U == vlSelfRef.t__DOT__w97__DOT__cyc)) {
            if (VL_UNLIKELY(((0U != ((((~ vlSelfRef.t__DOT__w97__DOT__i1extp[0U]) 
                                       | (~ vlSelfRef.t__DOT__w97__DOT__i1extp[1U])) 
                                      | (~ vlSelfRef.t__DOT__w97__DOT__i1extp[2U])) 
                                     | (1U ^ vlSelfRef.t__DOT__w97__DOT__i1extp[3U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 207, "");
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 208, "");
            }
            if (VL_UNLIKELY(((0U != ((((~ vlSelfRef.t__DOT__w97__DOT__i63ext[0U]) 
                                       | (~ vlSelfRef.t__DOT__w97__DOT__i63ext[1U])) 
                                      | (~ vlSelfRef.t__DOT__w97__DOT__i63ext[2U])) 
                                     | (1U ^ vlSelfRef.t__DOT__w97__DOT__i63ext[3U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 209, "");
            }
            if (VL_UNLIKELY(((0U != ((((~ vlSelfRef.t__DOT__w97__DOT__i65ext[0U]) 
                                       | (~ vlSelfRef.t__DOT__w97__DOT__i65ext[1U])) 
                                      | (~ vlSelfRef.t__DOT__w97__DOT__i65ext[2U])) 
                                     | (1U ^ vlSelfRef.t__DOT__w97__DOT__i65ext[3U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 210, "");
            }
        }
        __Vdly__t__DOT__w96__DOT__cyc = ((IData)(1U) 
                                         + vlSelfRef.t__DOT__w96__DOT__cyc);
        vlSelfRef.t__DOT__w96__DOT__i1 = (1U & vlSelfRef.t__DOT__w96__DOT__cyc);
        vlSelfRef.t__DOT__w96__DOT__i63 = (0x7fffffffffffffffULL 
                                           & (- (QData)((IData)(
                                                                (1U 
                                                                 & vlSelfRef.t__DOT__w96__DOT__cyc)))));
        vlSelfRef.t__DOT__w96__DOT__i65[0U] = (- (IData)(
                                                         (1U 
                                                          & vlSelfRef.t__DOT__w96__DOT__cyc)));
        vlSelfRef.t__DOT__w96__DOT__i65[1U] = (- (IData)(
                                                         (1U 
                                                          & vlSelfRef.t__DOT__w96__DOT__cyc)));
        vlSelfRef.t__DOT__w96__DOT__i65[2U] = (1U & 
                                               (- (IData)(
                                                          (1U 
                                                           & vlSelfRef.t__DOT__w96__DOT__cyc))));
        if ((1U == vlSelfRef.t__DOT__w96__DOT__cyc)) {
            if (VL_UNLIKELY(((0U != ((vlSelfRef.t__DOT__w96__DOT__i1extp[0U] 
                                      | vlSelfRef.t__DOT__w96__DOT__i1extp[1U]) 
                                     | vlSelfRef.t__DOT__w96__DOT__i1extp[2U]))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 201, "");
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 202, "");
            }
            if (VL_UNLIKELY(((0U != ((vlSelfRef.t__DOT__w96__DOT__i63ext[0U] 
                                      | vlSelfRef.t__DOT__w96__DOT__i63ext[1U]) 
                                     | vlSelfRef.t__DOT__w96__DOT__i63ext[2U]))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 203, "");
            }
            if (VL_UNLIKELY(((0U != ((vlSelfRef.t__DOT__w96__DOT__i65ext[0U] 
                                      | vlSelfRef.t__DOT__w96__DOT__i65ext[1U]) 
                                     | vlSelfRef.t__DOT__w96__DOT__i65ext[2U]))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 204, "");
            }
        } else if ((2U == vlSelfRef.t__DOT__w96__DOT__cyc)) {
            if (VL_UNLIKELY(((0U != (((~ vlSelfRef.t__DOT__w96__DOT__i1extp[0U]) 
                                      | (~ vlSelfRef.t__DOT__w96__DOT__i1extp[1U])) 
                                     | (~ vlSelfRef.t__DOT__w96__DOT__i1extp[2U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 207, "");
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 208, "");
            }
            if (VL_UNLIKELY(((0U != (((~ vlSelfRef.t__DOT__w96__DOT__i63ext[0U]) 
                                      | (~ vlSelfRef.t__DOT__w96__DOT__i63ext[1U])) 
                                     | (~ vlSelfRef.t__DOT__w96__DOT__i63ext[2U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 209, "");
            }
            if (VL_UNLIKELY(((0U != (((~ vlSelfRef.t__DOT__w96__DOT__i65ext[0U]) 
                                      | (~ vlSelfRef.t__DOT__w96__DOT__i65ext[1U])) 
                                     | (~ vlSelfRef.t__DOT__w96__DOT__i65ext[2U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 210, "");
            }
        }
        __Vdly__t__DOT__w95__DOT__cyc = ((IData)(1U) 
                                         + vlSelfRef.t__DOT__w95__DOT__cyc);
        vlSelfRef.t__DOT__w95__DOT__i1 = (1U & vlSelfRef.t__DOT__w95__DOT__cyc);
        vlSelfRef.t__DOT__w95__DOT__i63 = (0x7fffffffffffffffULL 
                                           & (- (QData)((IData)(
                                                                (1U 
                                                                 & vlSelfRef.t__DOT__w95__DOT__cyc)))));
        vlSelfRef.t__DOT__w95__DOT__i65[0U] = (- (IData)(
                                                         (1U 
                                                          & vlSelfRef.t__DOT__w95__DOT__cyc)));
        vlSelfRef.t__DOT__w95__DOT__i65[1U] = (- (IData)(
                                                         (1U 
                                                          & vlSelfRef.t__DOT__w95__DOT__cyc)));
        vlSelfRef.t__DOT__w95__DOT__i65[2U] = (1U & 
                                               (- (IData)(
                                                          (1U 
                                                           & vlSelfRef.t__DOT__w95__DOT__cyc))));
        if ((1U == vlSelfRef.t__DOT__w95__DOT__cyc)) {
            if (VL_UNLIKELY(((0U != ((vlSelfRef.t__DOT__w95__DOT__i1extp[0U] 
                                      | vlSelfRef.t__DOT__w95__DOT__i1extp[1U]) 
                                     | vlSelfRef.t__DOT__w95__DOT__i1extp[2U]))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 201, "");
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 202, "");
            }
            if (VL_UNLIKELY(((0U != ((vlSelfRef.t__DOT__w95__DOT__i63ext[0U] 
                                      | vlSelfRef.t__DOT__w95__DOT__i63ext[1U]) 
                                     | vlSelfRef.t__DOT__w95__DOT__i63ext[2U]))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 203, "");
            }
            if (VL_UNLIKELY(((0U != ((vlSelfRef.t__DOT__w95__DOT__i65ext[0U] 
                                      | vlSelfRef.t__DOT__w95__DOT__i65ext[1U]) 
                                     | vlSelfRef.t__DOT__w95__DOT__i65ext[2U]))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 204, "");
            }
        } else if ((2U == vlSelfRef.t__DOT__w95__DOT__cyc)) {
            if (VL_UNLIKELY(((0U != (((~ vlSelfRef.t__DOT__w95__DOT__i1extp[0U]) 
                                      | (~ vlSelfRef.t__DOT__w95__DOT__i1extp[1U])) 
                                     | (0x7fffffffU 
                                        ^ vlSelfRef.t__DOT__w95__DOT__i1extp[2U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 207, "");
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 208, "");
            }
            if (VL_UNLIKELY(((0U != (((~ vlSelfRef.t__DOT__w95__DOT__i63ext[0U]) 
                                      | (~ vlSelfRef.t__DOT__w95__DOT__i63ext[1U])) 
                                     | (0x7fffffffU 
                                        ^ vlSelfRef.t__DOT__w95__DOT__i63ext[2U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 209, "");
            }
            if (VL_UNLIKELY(((0U != (((~ vlSelfRef.t__DOT__w95__DOT__i65ext[0U]) 
                                      | (~ vlSelfRef.t__DOT__w95__DOT__i65ext[1U])) 
                                     | (0x7fffffffU 
                                        ^ vlSelfRef.t__DOT__w95__DOT__i65ext[2U])))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 210, "");
            }
        }
        __Vdly__t__DOT__w64__DOT__cyc = ((IData)(1U) 
                                         + vlSelfRef.t__DOT__w64__DOT__cyc);
        vlSelfRef.t__DOT__w64__DOT__i1 = (1U & vlSelfRef.t__DOT__w64__DOT__cyc);
        vlSelfRef.t__DOT__w64__DOT__i63 = (0x7fffffffffffffffULL 
                                           & (- (QData)((IData)(
                                                                (1U 
                                                                 & vlSelfRef.t__DOT__w64__DOT__cyc)))));
        vlSelfRef.t__DOT__w64__DOT__i65[0U] = (- (IData)(
                                                         (1U 
                                                          & vlSelfRef.t__DOT__w64__DOT__cyc)));
        vlSelfRef.t__DOT__w64__DOT__i65[1U] = (- (IData)(
                                                         (1U 
                                                          & vlSelfRef.t__DOT__w64__DOT__cyc)));
        vlSelfRef.t__DOT__w64__DOT__i65[2U] = (1U & 
                                               (- (IData)(
                                                          (1U 
                                                           & vlSelfRef.t__DOT__w64__DOT__cyc))));
        if ((1U == vlSelfRef.t__DOT__w64__DOT__cyc)) {
            if (VL_UNLIKELY(((0ULL != vlSelfRef.t__DOT__w64__DOT__i1extp)))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 201, "");
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 202, "");
            }
            if (VL_UNLIKELY(((0ULL != vlSelfRef.t__DOT__w64__DOT__i63ext)))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 203, "");
            }
            if (VL_UNLIKELY(((0ULL != vlSelfRef.t__DOT__w64__DOT__i65ext)))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 204, "");
            }
        } else if ((2U == vlSelfRef.t__DOT__w64__DOT__cyc)) {
            if (VL_UNLIKELY(((0xffffffffffffffffULL 
                              != vlSelfRef.t__DOT__w64__DOT__i1extp)))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 207, "");
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 208, "");
            }
            if (VL_UNLIKELY(((0xffffffffffffffffULL 
                              != vlSelfRef.t__DOT__w64__DOT__i63ext)))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 209, "");
            }
            if (VL_UNLIKELY(((0xffffffffffffffffULL 
                              != vlSelfRef.t__DOT__w64__DOT__i65ext)))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 210, "");
            }
        }
        __Vdly__t__DOT__cyc = ((IData)(1U) + vlSelfRef.t__DOT__cyc);
        if (((((((((0U == vlSelfRef.t__DOT__cyc) | 
                   (1U == vlSelfRef.t__DOT__cyc)) | 
                  (2U == vlSelfRef.t__DOT__cyc)) | 
                 (3U == vlSelfRef.t__DOT__cyc)) | (4U 
                                                   == vlSelfRef.t__DOT__cyc)) 
               | (5U == vlSelfRef.t__DOT__cyc)) | (6U 
                                                   == vlSelfRef.t__DOT__cyc)) 
             | (9U == vlSelfRef.t__DOT__cyc))) {
            if ((0U == vlSelfRef.t__DOT__cyc)) {
                vlSelfRef.t__DOT__a = 0x8b1bU;
                __Vdly__t__DOT__b = 0x1fU;
            } else if ((1U == vlSelfRef.t__DOT__cyc)) {
                vlSelfRef.t__DOT__a = 0x8b1bU;
                __Vdly__t__DOT__b = 1U;
            } else if ((2U == vlSelfRef.t__DOT__cyc)) {
                vlSelfRef.t__DOT__a = 0x8b1bU;
                __Vdly__t__DOT__b = 0x1eU;
                if (VL_UNLIKELY(((2U != vlSelfRef.t__DOT__cyc)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 122, "");
                }
                if (VL_UNLIKELY(((1U != (vlSelfRef.t__DOT__cyc 
                                         >> 1U))))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 123, "");
                }
                if (VL_UNLIKELY(((2U != vlSelfRef.t__DOT__cyc)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 124, "");
                }
                if (VL_UNLIKELY(((4U != (vlSelfRef.t__DOT__cyc 
                                         << 1U))))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 125, "");
                }
                if (VL_UNLIKELY(((6U != ((IData)(3U) 
                                         * vlSelfRef.t__DOT__cyc))))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 126, "");
                }
                if (VL_UNLIKELY(((2U != vlSelfRef.t__DOT__cyc)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 127, "");
                }
                if (VL_UNLIKELY(((4U != VL_MULS_III(32, (IData)(2U), vlSelfRef.t__DOT__cyc))))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 128, "");
                }
                if (VL_UNLIKELY(((6U != VL_MULS_III(32, (IData)(3U), vlSelfRef.t__DOT__cyc))))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 129, "");
                }
            } else if ((3U == vlSelfRef.t__DOT__cyc)) {
                if (VL_UNLIKELY(((3U != vlSelfRef.t__DOT__cyc)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 134, "");
                }
                vlSelfRef.t__DOT__a = 0x0048U;
                __Vdly__t__DOT__b = 0x1fU;
                if (VL_UNLIKELY(((6U != (vlSelfRef.t__DOT__cyc 
                                         << 1U))))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 135, "");
                }
                if (VL_UNLIKELY(((9U != ((IData)(3U) 
                                         * vlSelfRef.t__DOT__cyc))))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 136, "");
                }
                if (VL_UNLIKELY(((0x0000000cU != (vlSelfRef.t__DOT__cyc 
                                                  << 2U))))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 137, "");
                }
                if (VL_UNLIKELY(((3U != vlSelfRef.t__DOT__cyc)))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 138, "");
                }
                if (VL_UNLIKELY(((6U != VL_MULS_III(32, (IData)(2U), vlSelfRef.t__DOT__cyc))))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 139, "");
                }
                if (VL_UNLIKELY(((9U != VL_MULS_III(32, (IData)(3U), vlSelfRef.t__DOT__cyc))))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 140, "");
                }
            } else if ((4U == vlSelfRef.t__DOT__cyc)) {
                vlSelfRef.t__DOT__a = 0x4154U;
                __Vdly__t__DOT__b = 2U;
            } else if ((5U == vlSelfRef.t__DOT__cyc)) {
                vlSelfRef.t__DOT__a = 0xc3e8U;
                __Vdly__t__DOT__b = 0x12U;
            } else if (VL_LIKELY(((6U == vlSelfRef.t__DOT__cyc)))) {
                vlSelfRef.t__DOT__a = 0x488bU;
                __Vdly__t__DOT__b = 2U;
            } else {
                vlCtx.writef(0U, 0U);
                VL_FINISH_MT("/root/reference/test_regress/t/t_math_signed.v", 156, "");
            }
        }
        if (((((((((0U == vlSelfRef.t__DOT__cyc) | 
                   (1U == vlSelfRef.t__DOT__cyc)) | 
                  (2U == vlSelfRef.t__DOT__cyc)) | 
                 (3U == vlSelfRef.t__DOT__cyc)) | (4U 
                                                   == vlSelfRef.t__DOT__cyc)) 
               | (5U == vlSelfRef.t__DOT__cyc)) | (6U 
                                                   == vlSelfRef.t__DOT__cyc)) 
             | (7U == vlSelfRef.t__DOT__cyc))) {
            if ((0U != vlSelfRef.t__DOT__cyc)) {
                if ((1U == vlSelfRef.t__DOT__cyc)) {
                    if (VL_UNLIKELY(((0U != (((0xffff0001U 
                                               ^ (((IData)(vlSelfRef.t__DOT__b_s) 
                                                   << 0x00000010U) 
                                                  | (1U 
                                                     & ((IData)(vlSelfRef.t__DOT__b) 
                                                        >> 4U)))) 
                                              | (IData)(
                                                        (((QData)((IData)(
                                                                          (((IData)(vlSelfRef.t__DOT__sr) 
                                                                            << 0x00000010U) 
                                                                           | (IData)(vlSelfRef.t__DOT__srs)))) 
                                                          << 0x00000020U) 
                                                         | (QData)((IData)(
                                                                           (((IData)(vlSelfRef.t__DOT__sl) 
                                                                             << 0x00000010U) 
                                                                            | (IData)(vlSelfRef.t__DOT__sl))))))) 
                                             | (0x0000ffffU 
                                                ^ (IData)(
                                                          ((((QData)((IData)(
                                                                             (((IData)(vlSelfRef.t__DOT__sr) 
                                                                               << 0x00000010U) 
                                                                              | (IData)(vlSelfRef.t__DOT__srs)))) 
                                                             << 0x00000020U) 
                                                            | (QData)((IData)(
                                                                              (((IData)(vlSelfRef.t__DOT__sl) 
                                                                                << 0x00000010U) 
                                                                               | (IData)(vlSelfRef.t__DOT__sl))))) 
                                                           >> 0x00000020U)))))))) {
                        VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 162, "");
                    }
                } else if ((2U == vlSelfRef.t__DOT__cyc)) {
                    if (VL_UNLIKELY(((0U != (((((IData)(vlSelfRef.t__DOT__b_s) 
                                                << 0x00000010U) 
                                               | (1U 
                                                  & ((IData)(vlSelfRef.t__DOT__b) 
                                                     >> 4U))) 
                                              | (0x16361636U 
                                                 ^ (IData)(
                                                           (((QData)((IData)(
                                                                             (((IData)(vlSelfRef.t__DOT__sr) 
                                                                               << 0x00000010U) 
                                                                              | (IData)(vlSelfRef.t__DOT__srs)))) 
                                                             << 0x00000020U) 
                                                            | (QData)((IData)(
                                                                              (((IData)(vlSelfRef.t__DOT__sl) 
                                                                                << 0x00000010U) 
                                                                               | (IData)(vlSelfRef.t__DOT__sl)))))))) 
                                             | (0x458dc58dU 
                                                ^ (IData)(
                                                          ((((QData)((IData)(
                                                                             (((IData)(vlSelfRef.t__DOT__sr) 
                                                                               << 0x00000010U) 
                                                                              | (IData)(vlSelfRef.t__DOT__srs)))) 
                                                             << 0x00000020U) 
                                                            | (QData)((IData)(
                                                                              (((IData)(vlSelfRef.t__DOT__sl) 
                                                                                << 0x00000010U) 
                                                                               | (IData)(vlSelfRef.t__DOT__sl))))) 
                                                           >> 0x00000020U)))))))) {
                        VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 163, "");
                    }
                } else if ((3U == vlSelfRef.t__DOT__cyc)) {
                    if (VL_UNLIKELY(((0U != (((0xffff0001U 
                                               ^ (((IData)(vlSelfRef.t__DOT__b_s) 
                                                   << 0x00000010U) 
                                                  | (1U 
                                                     & ((IData)(vlSelfRef.t__DOT__b) 
                                                        >> 4U)))) 
                                              | (IData)(
                                                        (((QData)((IData)(
                                                                          (((IData)(vlSelfRef.t__DOT__sr) 
                                                                            << 0x00000010U) 
                                                                           | (IData)(vlSelfRef.t__DOT__srs)))) 
                                                          << 0x00000020U) 
                                                         | (QData)((IData)(
                                                                           (((IData)(vlSelfRef.t__DOT__sl) 
                                                                             << 0x00000010U) 
                                                                            | (IData)(vlSelfRef.t__DOT__sl))))))) 
                                             | (0x0000ffffU 
                                                ^ (IData)(
                                                          ((((QData)((IData)(
                                                                             (((IData)(vlSelfRef.t__DOT__sr) 
                                                                               << 0x00000010U) 
                                                                              | (IData)(vlSelfRef.t__DOT__srs)))) 
                                                             << 0x00000020U) 
                                                            | (QData)((IData)(
                                                                              (((IData)(vlSelfRef.t__DOT__sl) 
                                                                                << 0x00000010U) 
                                                                               | (IData)(vlSelfRef.t__DOT__sl))))) 
                                                           >> 0x00000020U)))))))) {
                        VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 164, "");
                    }
                } else if ((4U == vlSelfRef.t__DOT__cyc)) {
                    if (VL_UNLIKELY(((0U != (((0xffff0001U 
                                               ^ (((IData)(vlSelfRef.t__DOT__b_s) 
                                                   << 0x00000010U) 
                                                  | (1U 
                                                     & ((IData)(vlSelfRef.t__DOT__b) 
                                                        >> 4U)))) 
                                              | (IData)(
                                                        (((QData)((IData)(
                                                                          (((IData)(vlSelfRef.t__DOT__sr) 
                                                                            << 0x00000010U) 
                                                                           | (IData)(vlSelfRef.t__DOT__srs)))) 
                                                          << 0x00000020U) 
                                                         | (QData)((IData)(
                                                                           (((IData)(vlSelfRef.t__DOT__sl) 
                                                                             << 0x00000010U) 
                                                                            | (IData)(vlSelfRef.t__DOT__sl))))))) 
                                             | (IData)(
                                                       ((((QData)((IData)(
                                                                          (((IData)(vlSelfRef.t__DOT__sr) 
                                                                            << 0x00000010U) 
                                                                           | (IData)(vlSelfRef.t__DOT__srs)))) 
                                                          << 0x00000020U) 
                                                         | (QData)((IData)(
                                                                           (((IData)(vlSelfRef.t__DOT__sl) 
                                                                             << 0x00000010U) 
                                                                            | (IData)(vlSelfRef.t__DOT__sl))))) 
                                                        >> 0x00000020U))))))) {
                        VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 165, "");
                    }
                } else if ((5U == vlSelfRef.t__DOT__cyc)) {
                    if (VL_UNLIKELY(((0U != (((((IData)(vlSelfRef.t__DOT__b_s) 
                                                << 0x00000010U) 
                                               | (1U 
                                                  & ((IData)(vlSelfRef.t__DOT__b) 
                                                     >> 4U))) 
                                              | (0x05500550U 
                                                 ^ (IData)(
                                                           (((QData)((IData)(
                                                                             (((IData)(vlSelfRef.t__DOT__sr) 
                                                                               << 0x00000010U) 
                                                                              | (IData)(vlSelfRef.t__DOT__srs)))) 
                                                             << 0x00000020U) 
                                                            | (QData)((IData)(
                                                                              (((IData)(vlSelfRef.t__DOT__sl) 
                                                                                << 0x00000010U) 
                                                                               | (IData)(vlSelfRef.t__DOT__sl)))))))) 
                                             | (0x10551055U 
                                                ^ (IData)(
                                                          ((((QData)((IData)(
                                                                             (((IData)(vlSelfRef.t__DOT__sr) 
                                                                               << 0x00000010U) 
                                                                              | (IData)(vlSelfRef.t__DOT__srs)))) 
                                                             << 0x00000020U) 
                                                            | (QData)((IData)(
                                                                              (((IData)(vlSelfRef.t__DOT__sl) 
                                                                                << 0x00000010U) 
                                                                               | (IData)(vlSelfRef.t__DOT__sl))))) 
                                                           >> 0x00000020U)))))))) {
                        VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 166, "");
                    }
                } else if ((6U == vlSelfRef.t__DOT__cyc)) {
                    if (VL_UNLIKELY(((0U != (((0xffff0001U 
                                               ^ (((IData)(vlSelfRef.t__DOT__b_s) 
                                                   << 0x00000010U) 
                                                  | (1U 
                                                     & ((IData)(vlSelfRef.t__DOT__b) 
                                                        >> 4U)))) 
                                              | (IData)(
                                                        (((QData)((IData)(
                                                                          (((IData)(vlSelfRef.t__DOT__sr) 
                                                                            << 0x00000010U) 
                                                                           | (IData)(vlSelfRef.t__DOT__srs)))) 
                                                          << 0x00000020U) 
                                                         | (QData)((IData)(
                                                                           (((IData)(vlSelfRef.t__DOT__sl) 
                                                                             << 0x00000010U) 
                                                                            | (IData)(vlSelfRef.t__DOT__sl))))))) 
                                             | (0x0000ffffU 
                                                ^ (IData)(
                                                          ((((QData)((IData)(
                                                                             (((IData)(vlSelfRef.t__DOT__sr) 
                                                                               << 0x00000010U) 
                                                                              | (IData)(vlSelfRef.t__DOT__srs)))) 
                                                             << 0x00000020U) 
                                                            | (QData)((IData)(
                                                                              (((IData)(vlSelfRef.t__DOT__sl) 
                                                                                << 0x00000010U) 
                                                                               | (IData)(vlSelfRef.t__DOT__sl))))) 
                                                           >> 0x00000020U)))))))) {
                        VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 167, "");
                    }
                } else if (VL_UNLIKELY(((0U != (((((IData)(vlSelfRef.t__DOT__b_s) 
                                                   << 0x00000010U) 
                                                  | (1U 
                                                     & ((IData)(vlSelfRef.t__DOT__b) 
                                                        >> 4U))) 
                                                 | (0x222c222cU 
                                                    ^ (IData)(
                                                              (((QData)((IData)(
                                                                                (((IData)(vlSelfRef.t__DOT__sr) 
                                                                                << 0x00000010U) 
                                                                                | (IData)(vlSelfRef.t__DOT__srs)))) 
                                                                << 0x00000020U) 
                                                               | (QData)((IData)(
                                                                                (((IData)(vlSelfRef.t__DOT__sl) 
                                                                                << 0x00000010U) 
                                                                                | (IData)(vlSelfRef.t__DOT__sl)))))))) 
                                                | (0x12221222U 
                                                   ^ (IData)(
                                                             ((((QData)((IData)(
                                                                                (((IData)(vlSelfRef.t__DOT__sr) 
                                                                                << 0x00000010U) 
                                                                                | (IData)(vlSelfRef.t__DOT__srs)))) 
                                                                << 0x00000020U) 
                                                               | (QData)((IData)(
                                                                                (((IData)(vlSelfRef.t__DOT__sl) 
                                                                                << 0x00000010U) 
                                                                                | (IData)(vlSelfRef.t__DOT__sl))))) 
                                                              >> 0x00000020U)))))))) {
                    VL_STOP_MT("/root/reference/test_regress/t/t_math_signed.v", 168, "");
                }
            }
        }
        vlSelfRef.t__DOT__w1__DOT__cyc = __Vdly__t__DOT__w1__DOT__cyc;
        vlSelfRef.t__DOT__w1__DOT__i1 = __Vdly__t__DOT__w1__DOT__i1;
        vlSelfRef.t__DOT__w1__DOT__i63 = __Vdly__t__DOT__w1__DOT__i63;
        vlSelfRef.t__DOT__w1__DOT__i65[0U] = __Vdly__t__DOT__w1__DOT__i65[0U];
        vlSelfRef.t__DOT__w1__DOT__i65[1U] = __Vdly__t__DOT__w1__DOT__i65[1U];
        vlSelfRef.t__DOT__w1__DOT__i65[2U] = __Vdly__t__DOT__w1__DOT__i65[2U];
        vlSelfRef.t__DOT__w32__DOT__cyc = __Vdly__t__DOT__w32__DOT__cyc;
        vlSelfRef.t__DOT__w32__DOT__i63 = __Vdly__t__DOT__w32__DOT__i63;
        vlSelfRef.t__DOT__w32__DOT__i65[0U] = __Vdly__t__DOT__w32__DOT__i65[0U];
        vlSelfRef.t__DOT__w32__DOT__i65[1U] = __Vdly__t__DOT__w32__DOT__i65[1U];
        vlSelfRef.t__DOT__w32__DOT__i65[2U] = __Vdly__t__DOT__w32__DOT__i65[2U];
        vlSelfRef.t__DOT__w31__DOT__cyc = __Vdly__t__DOT__w31__DOT__cyc;
        vlSelfRef.t__DOT__w31__DOT__i63 = __Vdly__t__DOT__w31__DOT__i63;
        vlSelfRef.t__DOT__w31__DOT__i65[0U] = __Vdly__t__DOT__w31__DOT__i65[0U];
        vlSelfRef.t__DOT__w31__DOT__i65[1U] = __Vdly__t__DOT__w31__DOT__i65[1U];
        vlSelfRef.t__DOT__w31__DOT__i65[2U] = __Vdly__t__DOT__w31__DOT__i65[2U];
        vlSelfRef.t__DOT__w65__DOT__cyc = __Vdly__t__DOT__w65__DOT__cyc;
        vlSelfRef.t__DOT__w65__DOT__i65[0U] = __Vdly__t__DOT__w65__DOT__i65[0U];
        vlSelfRef.t__DOT__w65__DOT__i65[1U] = __Vdly__t__DOT__w65__DOT__i65[1U];
        vlSelfRef.t__DOT__w65__DOT__i65[2U] = __Vdly__t__DOT__w65__DOT__i65[2U];
        vlSelfRef.t__DOT__w63__DOT__cyc = __Vdly__t__DOT__w63__DOT__cyc;
        vlSelfRef.t__DOT__w63__DOT__i63 = __Vdly__t__DOT__w63__DOT__i63;
        vlSelfRef.t__DOT__w33__DOT__cyc = __Vdly__t__DOT__w33__DOT__cyc;
        vlSelfRef.t__DOT__w33__DOT__i63 = __Vdly__t__DOT__w33__DOT__i63;
        vlSelfRef.t__DOT__w97__DOT__cyc = __Vdly__t__DOT__w97__DOT__cyc;
        vlSelfRef.t__DOT__w96__DOT__cyc = __Vdly__t__DOT__w96__DOT__cyc;
        vlSelfRef.t__DOT__w95__DOT__cyc = __Vdly__t__DOT__w95__DOT__cyc;
        vlSelfRef.t__DOT__w64__DOT__cyc = __Vdly__t__DOT__w64__DOT__cyc;
        vlSelfRef.t__DOT__cyc = __Vdly__t__DOT__cyc;
        vlSelfRef.t__DOT__b = __Vdly__t__DOT__b;
        vlSelfRef.t__DOT__w32__DOT__i1extp = VL_EXTENDS_II(32,1, (IData)(vlSelfRef.t__DOT__w32__DOT__i1));
        vlSelfRef.t__DOT__w31__DOT__i1extp = (0x7fffffffU 
                                              & VL_EXTENDS_II(31,1, (IData)(vlSelfRef.t__DOT__w31__DOT__i1)));
        VL_EXTENDS_WI(65,1, __Vtemp_68, (IData)(vlSelfRef.t__DOT__w65__DOT__i1));
        vlSelfRef.t__DOT__w65__DOT__i1extp[0U] = __Vtemp_68[0U];
        vlSelfRef.t__DOT__w65__DOT__i1extp[1U] = __Vtemp_68[1U];
        vlSelfRef.t__DOT__w65__DOT__i1extp[2U] = (1U 
                                                  & __Vtemp_68[2U]);
        VL_EXTENDS_WQ(65,63, __Vtemp_70, vlSelfRef.t__DOT__w65__DOT__i63);
        vlSelfRef.t__DOT__w65__DOT__i63ext[0U] = __Vtemp_70[0U];
        vlSelfRef.t__DOT__w65__DOT__i63ext[1U] = __Vtemp_70[1U];
        vlSelfRef.t__DOT__w65__DOT__i63ext[2U] = (1U 
                                                  & __Vtemp_70[2U]);
        vlSelfRef.t__DOT__w63__DOT__i1extp = (0x7fffffffffffffffULL 
                                              & VL_EXTENDS_QI(63,1, (IData)(vlSelfRef.t__DOT__w63__DOT__i1)));
        vlSelfRef.t__DOT__w63__DOT__i65ext = (0x7fffffffffffffffULL 
                                              & (((QData)((IData)(vlSelfRef.t__DOT__w63__DOT__i65[1U])) 
                                                  << 0x00000020U) 
                                                 | (QData)((IData)(vlSelfRef.t__DOT__w63__DOT__i65[0U]))));
        vlSelfRef.t__DOT__w33__DOT__i1extp = (0x00000001ffffffffULL 
                                              & VL_EXTENDS_QI(33,1, (IData)(vlSelfRef.t__DOT__w33__DOT__i1)));
        vlSelfRef.t__DOT__w33__DOT__i65ext = (0x00000001ffffffffULL 
                                              & (((QData)((IData)(vlSelfRef.t__DOT__w33__DOT__i65[1U])) 
                                                  << 0x00000020U) 
                                                 | (QData)((IData)(vlSelfRef.t__DOT__w33__DOT__i65[0U]))));
        VL_EXTENDS_WI(97,1, __Vtemp_72, (IData)(vlSelfRef.t__DOT__w97__DOT__i1));
        vlSelfRef.t__DOT__w97__DOT__i1extp[0U] = __Vtemp_72[0U];
        vlSelfRef.t__DOT__w97__DOT__i1extp[1U] = __Vtemp_72[1U];
        vlSelfRef.t__DOT__w97__DOT__i1extp[2U] = __Vtemp_72[2U];
        vlSelfRef.t__DOT__w97__DOT__i1extp[3U] = (1U 
                                                  & __Vtemp_72[3U]);
        VL_EXTENDS_WQ(97,63, __Vtemp_74, vlSelfRef.t__DOT__w97__DOT__i63);
        vlSelfRef.t__DOT__w97__DOT__i63ext[0U] = __Vtemp_74[0U];
        vlSelfRef.t__DOT__w97__DOT__i63ext[1U] = __Vtemp_74[1U];
        vlSelfRef.t__DOT__w97__DOT__i63ext[2U] = __Vtemp_74[2U];
        vlSelfRef.t__DOT__w97__DOT__i63ext[3U] = (1U 
                                                  & __Vtemp_74[3U]);
        VL_EXTENDS_WW(97,65, __Vtemp_76, vlSelfRef.t__DOT__w97__DOT__i65);
        vlSelfRef.t__DOT__w97__DOT__i65ext[0U] = __Vtemp_76[0U];
        vlSelfRef.t__DOT__w97__DOT__i65ext[1U] = __Vtemp_76[1U];
        vlSelfRef.t__DOT__w97__DOT__i65ext[2U] = __Vtemp_76[2U];
        vlSelfRef.t__DOT__w97__DOT__i65ext[3U] = (1U 
                                                  & __Vtemp_76[3U]);
        VL_EXTENDS_WI(96,1, vlSelfRef.t__DOT__w96__DOT__i1extp, (IData)(vlSelfRef.t__DOT__w96__DOT__i1));
        VL_EXTENDS_WQ(96,63, vlSelfRef.t__DOT__w96__DOT__i63ext, vlSelfRef.t__DOT__w96__DOT__i63);
        VL_EXTENDS_WW(96,65, vlSelfRef.t__DOT__w96__DOT__i65ext, vlSelfRef.t__DOT__w96__DOT__i65);
        VL_EXTENDS_WI(95,1, __Vtemp_78, (IData)(vlSelfRef.t__DOT__w95__DOT__i1));
        vlSelfRef.t__DOT__w95__DOT__i1extp[0U] = __Vtemp_78[0U];
        vlSelfRef.t__DOT__w95__DOT__i1extp[1U] = __Vtemp_78[1U];
        vlSelfRef.t__DOT__w95__DOT__i1extp[2U] = (0x7fffffffU 
                                                  & __Vtemp_78[2U]);
        VL_EXTENDS_WQ(95,63, __Vtemp_80, vlSelfRef.t__DOT__w95__DOT__i63);
        vlSelfRef.t__DOT__w95__DOT__i63ext[0U] = __Vtemp_80[0U];
        vlSelfRef.t__DOT__w95__DOT__i63ext[1U] = __Vtemp_80[1U];
        vlSelfRef.t__DOT__w95__DOT__i63ext[2U] = (0x7fffffffU 
                                                  & __Vtemp_80[2U]);
        VL_EXTENDS_WW(95,65, __Vtemp_82, vlSelfRef.t__DOT__w95__DOT__i65);
        vlSelfRef.t__DOT__w95__DOT__i65ext[0U] = __Vtemp_82[0U];
        vlSelfRef.t__DOT__w95__DOT__i65ext[1U] = __Vtemp_82[1U];
        vlSelfRef.t__DOT__w95__DOT__i65ext[2U] = (0x7fffffffU 
                                                  & __Vtemp_82[2U]);
        vlSelfRef.t__DOT__w64__DOT__i1extp = VL_EXTENDS_QI(64,1, (IData)(vlSelfRef.t__DOT__w64__DOT__i1));
        vlSelfRef.t__DOT__w64__DOT__i63ext = VL_EXTENDS_QQ(64,63, vlSelfRef.t__DOT__w64__DOT__i63);
        vlSelfRef.t__DOT__w64__DOT__i65ext = (((QData)((IData)(vlSelfRef.t__DOT__w64__DOT__i65[1U])) 
                                               << 0x00000020U) 
                                              | (QData)((IData)(vlSelfRef.t__DOT__w64__DOT__i65[0U])));
        vlSelfRef.t__DOT__b_s = (0x0000ffffU & VL_SHIFTRS_III(16,16,32, 
                                                              (0x0000ffffU 
                                                               & VL_EXTENDS_II(16,5, (IData)(vlSelfRef.t__DOT__b))), 4U));
        vlSelfRef.t__DOT__sr = ((IData)(vlSelfRef.t__DOT__a) 
                                >> (IData)(vlSelfRef.t__DOT__b));
        vlSelfRef.t__DOT__srs = (0x0000ffffU & VL_SHIFTRS_III(16,16,5, (IData)(vlSelfRef.t__DOT__a), (IData)(vlSelfRef.t__DOT__b)));
        vlSelfRef.t__DOT__sl = (0x0000ffffU & ((IData)(vlSelfRef.t__DOT__a) 
                                               << (IData)(vlSelfRef.t__DOT__b)));
    }

    VLCU_D static void Vt_math_signed___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt_math_signed___024root___eval_phase__act(Vt_math_signed___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt_math_signed___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vt_math_signed___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vt_math_signed___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vt_math_signed___024root___eval_phase__nba(Vt_math_signed___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vt_math_signed___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    Vt_math_signed___024root___nba_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
            Vt_math_signed___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vt_math_signed___024root___eval(Vt_math_signed___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt_math_signed___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("/root/reference/test_regress/t/t_math_signed.v", 7, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vt_math_signed___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("/root/reference/test_regress/t/t_math_signed.v", 7, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vt_math_signed___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vt_math_signed___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vt_math_signed___024root___ctor_var_reset(Vt_math_signed___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        vlSelf->t__DOT__a = VL_SCOPED_RAND_RESET_I(16, __VscopeHash, 568261790826559771ull);
        vlSelf->t__DOT__b = VL_SCOPED_RAND_RESET_I(5, __VscopeHash, 11816130287683444364ull);
        vlSelf->t__DOT__sr = VL_SCOPED_RAND_RESET_I(16, __VscopeHash, 2750434385867559212ull);
        vlSelf->t__DOT__srs = VL_SCOPED_RAND_RESET_I(16, __VscopeHash, 15390884037904855503ull);
        vlSelf->t__DOT__sl = VL_SCOPED_RAND_RESET_I(16, __VscopeHash, 12576617150736452623ull);
        vlSelf->t__DOT__b_s = VL_SCOPED_RAND_RESET_I(16, __VscopeHash, 6037124860917247449ull);
        vlSelf->t__DOT__cyc = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 15508623621340971111ull);
        vlSelf->t__DOT__w97__DOT__i1 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 928132118708518522ull);
        vlSelf->t__DOT__w97__DOT__i63 = VL_SCOPED_RAND_RESET_Q(63, __VscopeHash, 10728607501746702105ull);
        VL_SCOPED_RAND_RESET_W(65, vlSelf->t__DOT__w97__DOT__i65, __VscopeHash, 1991379704812757442ull);
        VL_SCOPED_RAND_RESET_W(97, vlSelf->t__DOT__w97__DOT__i1extp, __VscopeHash, 14139856529600468064ull);
        VL_SCOPED_RAND_RESET_W(97, vlSelf->t__DOT__w97__DOT__i63ext, __VscopeHash, 3992087731905400127ull);
        VL_SCOPED_RAND_RESET_W(97, vlSelf->t__DOT__w97__DOT__i65ext, __VscopeHash, 2713039251910081070ull);
        vlSelf->t__DOT__w97__DOT__cyc = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 14877792376444498807ull);
        vlSelf->t__DOT__w96__DOT__i1 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 4671433720129338194ull);
        vlSelf->t__DOT__w96__DOT__i63 = VL_SCOPED_RAND_RESET_Q(63, __VscopeHash, 12469660357121288776ull);
        VL_SCOPED_RAND_RESET_W(65, vlSelf->t__DOT__w96__DOT__i65, __VscopeHash, 17909506344301382243ull);
        VL_SCOPED_RAND_RESET_W(96, vlSelf->t__DOT__w96__DOT__i1extp, __VscopeHash, 14765162764492131632ull);
        VL_SCOPED_RAND_RESET_W(96, vlSelf->t__DOT__w96__DOT__i63ext, __VscopeHash, 4349675223643063758ull);
        VL_SCOPED_RAND_RESET_W(96, vlSelf->t__DOT__w96__DOT__i65ext, __VscopeHash, 17280920404760602862ull);
        vlSelf->t__DOT__w96__DOT__cyc = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 2470780131329577180ull);
        vlSelf->t__DOT__w95__DOT__i1 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 9633990323095769851ull);
        vlSelf->t__DOT__w95__DOT__i63 = VL_SCOPED_RAND_RESET_Q(63, __VscopeHash, 10106303152791304131ull);
        VL_SCOPED_RAND_RESET_W(65, vlSelf->t__DOT__w95__DOT__i65, __VscopeHash, 1523862408162255112ull);
        VL_SCOPED_RAND_RESET_W(95, vlSelf->t__DOT__w95__DOT__i1extp, __VscopeHash, 17640346159714024600ull);
        VL_SCOPED_RAND_RESET_W(95, vlSelf->t__DOT__w95__DOT__i63ext, __VscopeHash, 13917068782124885402ull);
        VL_SCOPED_RAND_RESET_W(95, vlSelf->t__DOT__w95__DOT__i65ext, __VscopeHash, 11839134844225795275ull);
        vlSelf->t__DOT__w95__DOT__cyc = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 18049102714926638590ull);
        vlSelf->t__DOT__w65__DOT__i1 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 11356633277252247504ull);
        vlSelf->t__DOT__w65__DOT__i63 = VL_SCOPED_RAND_RESET_Q(63, __VscopeHash, 16282141810105593127ull);
        VL_SCOPED_RAND_RESET_W(65, vlSelf->t__DOT__w65__DOT__i65, __VscopeHash, 3941226083518680104ull);
        VL_SCOPED_RAND_RESET_W(65, vlSelf->t__DOT__w65__DOT__i1extp, __VscopeHash, 5609611401802257271ull);
        VL_SCOPED_RAND_RESET_W(65, vlSelf->t__DOT__w65__DOT__i63ext, __VscopeHash, 13245392145410526347ull);
        vlSelf->t__DOT__w65__DOT__cyc = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 14977582148508349490ull);
        vlSelf->t__DOT__w64__DOT__i1 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 1351449351727902784ull);
        vlSelf->t__DOT__w64__DOT__i63 = VL_SCOPED_RAND_RESET_Q(63, __VscopeHash, 15514771027609797764ull);
        VL_SCOPED_RAND_RESET_W(65, vlSelf->t__DOT__w64__DOT__i65, __VscopeHash, 14512245650090879358ull);
        vlSelf->t__DOT__w64__DOT__i1extp = VL_SCOPED_RAND_RESET_Q(64, __VscopeHash, 12096248981822954216ull);
        vlSelf->t__DOT__w64__DOT__i63ext = VL_SCOPED_RAND_RESET_Q(64, __VscopeHash, 3170110798577787612ull);
        vlSelf->t__DOT__w64__DOT__i65ext = VL_SCOPED_RAND_RESET_Q(64, __VscopeHash, 6516804838129056391ull);
        vlSelf->t__DOT__w64__DOT__cyc = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 10742687480892236343ull);
        vlSelf->t__DOT__w63__DOT__i1 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 13189972654478479732ull);
        vlSelf->t__DOT__w63__DOT__i63 = VL_SCOPED_RAND_RESET_Q(63, __VscopeHash, 10378480398353179771ull);
        VL_SCOPED_RAND_RESET_W(65, vlSelf->t__DOT__w63__DOT__i65, __VscopeHash, 5886723681734179274ull);
        vlSelf->t__DOT__w63__DOT__i1extp = VL_SCOPED_RAND_RESET_Q(63, __VscopeHash, 14702652998672930521ull);
        vlSelf->t__DOT__w63__DOT__i65ext = VL_SCOPED_RAND_RESET_Q(63, __VscopeHash, 7418387297949876486ull);
        vlSelf->t__DOT__w63__DOT__cyc = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 879351119192842322ull);
        vlSelf->t__DOT__w33__DOT__i1 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 10186506911670961590ull);
        vlSelf->t__DOT__w33__DOT__i63 = VL_SCOPED_RAND_RESET_Q(63, __VscopeHash, 4379730796371361724ull);
        VL_SCOPED_RAND_RESET_W(65, vlSelf->t__DOT__w33__DOT__i65, __VscopeHash, 8171093368965423468ull);
        vlSelf->t__DOT__w33__DOT__i1extp = VL_SCOPED_RAND_RESET_Q(33, __VscopeHash, 6352154937968211117ull);
        vlSelf->t__DOT__w33__DOT__i65ext = VL_SCOPED_RAND_RESET_Q(33, __VscopeHash, 21266667878512104ull);
        vlSelf->t__DOT__w33__DOT__cyc = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 9306998173495372047ull);
        vlSelf->t__DOT__w32__DOT__i1 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 8511542180099395367ull);
        vlSelf->t__DOT__w32__DOT__i63 = VL_SCOPED_RAND_RESET_Q(63, __VscopeHash, 13483170450574494576ull);
        VL_SCOPED_RAND_RESET_W(65, vlSelf->t__DOT__w32__DOT__i65, __VscopeHash, 6621871242206721270ull);
        vlSelf->t__DOT__w32__DOT__i1extp = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 2384304530595863633ull);
        vlSelf->t__DOT__w32__DOT__cyc = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 5143352985657504868ull);
        vlSelf->t__DOT__w31__DOT__i1 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 9146567382562866363ull);
        vlSelf->t__DOT__w31__DOT__i63 = VL_SCOPED_RAND_RESET_Q(63, __VscopeHash, 3100057028929605908ull);
        VL_SCOPED_RAND_RESET_W(65, vlSelf->t__DOT__w31__DOT__i65, __VscopeHash, 221576554945826189ull);
        vlSelf->t__DOT__w31__DOT__i1extp = VL_SCOPED_RAND_RESET_I(31, __VscopeHash, 17811965241784396099ull);
        vlSelf->t__DOT__w31__DOT__cyc = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 11923240209730056722ull);
        vlSelf->t__DOT__w1__DOT__i1 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 14190428321251324004ull);
        vlSelf->t__DOT__w1__DOT__i63 = VL_SCOPED_RAND_RESET_Q(63, __VscopeHash, 11069287561971881218ull);
        VL_SCOPED_RAND_RESET_W(65, vlSelf->t__DOT__w1__DOT__i65, __VscopeHash, 7555192903547261145ull);
        vlSelf->t__DOT__w1__DOT__cyc = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 17199264867717782774ull);
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
        Vt_math_signed___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_signed___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_signed___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_signed___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_signed___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_signed___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt_math_signed__cuda::info() {
    typedef Vt_math_signed__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"t__DOT__b", "t.b", VLCU_MEMBER_STATE, 1, 0, 1, 5, 4, 0, T::OFF_t__DOT__b},
        {"t__DOT__w97__DOT__i1", "t.w97.i1", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__w97__DOT__i1},
        {"t__DOT__w96__DOT__i1", "t.w96.i1", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__w96__DOT__i1},
        {"t__DOT__w95__DOT__i1", "t.w95.i1", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__w95__DOT__i1},
        {"t__DOT__w65__DOT__i1", "t.w65.i1", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__w65__DOT__i1},
        {"t__DOT__w64__DOT__i1", "t.w64.i1", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__w64__DOT__i1},
        {"t__DOT__w63__DOT__i1", "t.w63.i1", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__w63__DOT__i1},
        {"t__DOT__w33__DOT__i1", "t.w33.i1", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__w33__DOT__i1},
        {"t__DOT__w32__DOT__i1", "t.w32.i1", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__w32__DOT__i1},
        {"t__DOT__w31__DOT__i1", "t.w31.i1", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__w31__DOT__i1},
        {"t__DOT__w1__DOT__i1", "t.w1.i1", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__w1__DOT__i1},
        {"__VstlFirstIteration", "__VstlFirstIteration", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlFirstIteration},
        {"__VstlPhaseResult", "__VstlPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"t__DOT__a", "t.a", VLCU_MEMBER_STATE, 2, 0, 1, 16, 15, 0, T::OFF_t__DOT__a},
        {"t__DOT__sr", "t.sr", VLCU_MEMBER_STATE, 2, 0, 1, 16, 15, 0, T::OFF_t__DOT__sr},
        {"t__DOT__srs", "t.srs", VLCU_MEMBER_STATE, 2, 0, 1, 16, 15, 0, T::OFF_t__DOT__srs},
        {"t__DOT__sl", "t.sl", VLCU_MEMBER_STATE, 2, 0, 1, 16, 15, 0, T::OFF_t__DOT__sl},
        {"t__DOT__b_s", "t.b_s", VLCU_MEMBER_STATE, 2, 0, 1, 16, 15, 0, T::OFF_t__DOT__b_s},
        {"t__DOT__cyc", "t.cyc", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__cyc},
        {"t__DOT__w97__DOT__i65", "t.w97.i65", VLCU_MEMBER_STATE, 4, 0, 3, 65, 64, 0, T::OFF_t__DOT__w97__DOT__i65},
        {"t__DOT__w97__DOT__i1extp", "t.w97.i1extp", VLCU_MEMBER_STATE, 4, 0, 4, 97, 96, 0, T::OFF_t__DOT__w97__DOT__i1extp},
        {"t__DOT__w97__DOT__i63ext", "t.w97.i63ext", VLCU_MEMBER_STATE, 4, 0, 4, 97, 96, 0, T::OFF_t__DOT__w97__DOT__i63ext},
        {"t__DOT__w97__DOT__i65ext", "t.w97.i65ext", VLCU_MEMBER_STATE, 4, 0, 4, 97, 96, 0, T::OFF_t__DOT__w97__DOT__i65ext},
        {"t__DOT__w97__DOT__cyc", "t.w97.cyc", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__w97__DOT__cyc},
        {"t__DOT__w96__DOT__i65", "t.w96.i65", VLCU_MEMBER_STATE, 4, 0, 3, 65, 64, 0, T::OFF_t__DOT__w96__DOT__i65},
        {"t__DOT__w96__DOT__i1extp", "t.w96.i1extp", VLCU_MEMBER_STATE, 4, 0, 3, 96, 95, 0, T::OFF_t__DOT__w96__DOT__i1extp},
        {"t__DOT__w96__DOT__i63ext", "t.w96.i63ext", VLCU_MEMBER_STATE, 4, 0, 3, 96, 95, 0, T::OFF_t__DOT__w96__DOT__i63ext},
        {"t__DOT__w96__DOT__i65ext", "t.w96.i65ext", VLCU_MEMBER_STATE, 4, 0, 3, 96, 95, 0, T::OFF_t__DOT__w96__DOT__i65ext},
        {"t__DOT__w96__DOT__cyc", "t.w96.cyc", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__w96__DOT__cyc},
        {"t__DOT__w95__DOT__i65", "t.w95.i65", VLCU_MEMBER_STATE, 4, 0, 3, 65, 64, 0, T::OFF_t__DOT__w95__DOT__i65},
        {"t__DOT__w95__DOT__i1extp", "t.w95.i1extp", VLCU_MEMBER_STATE, 4, 0, 3, 95, 94, 0, T::OFF_t__DOT__w95__DOT__i1extp},
        {"t__DOT__w95__DOT__i63ext", "t.w95.i63ext", VLCU_MEMBER_STATE, 4, 0, 3, 95, 94, 0, T::OFF_t__DOT__w95__DOT__i63ext},
        {"t__DOT__w95__DOT__i65ext", "t.w95.i65ext", VLCU_MEMBER_STATE, 4, 0, 3, 95, 94, 0, T::OFF_t__DOT__w95__DOT__i65ext},
        {"t__DOT__w95__DOT__cyc", "t.w95.cyc", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__w95__DOT__cyc},
        {"t__DOT__w65__DOT__i65", "t.w65.i65", VLCU_MEMBER_STATE, 4, 0, 3, 65, 64, 0, T::OFF_t__DOT__w65__DOT__i65},
        {"t__DOT__w65__DOT__i1extp", "t.w65.i1extp", VLCU_MEMBER_STATE, 4, 0, 3, 65, 64, 0, T::OFF_t__DOT__w65__DOT__i1extp},
        {"t__DOT__w65__DOT__i63ext", "t.w65.i63ext", VLCU_MEMBER_STATE, 4, 0, 3, 65, 64, 0, T::OFF_t__DOT__w65__DOT__i63ext},
        {"t__DOT__w65__DOT__cyc", "t.w65.cyc", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__w65__DOT__cyc},
        {"t__DOT__w64__DOT__i65", "t.w64.i65", VLCU_MEMBER_STATE, 4, 0, 3, 65, 64, 0, T::OFF_t__DOT__w64__DOT__i65},
        {"t__DOT__w64__DOT__cyc", "t.w64.cyc", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__w64__DOT__cyc},
        {"t__DOT__w63__DOT__i65", "t.w63.i65", VLCU_MEMBER_STATE, 4, 0, 3, 65, 64, 0, T::OFF_t__DOT__w63__DOT__i65},
        {"t__DOT__w63__DOT__cyc", "t.w63.cyc", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__w63__DOT__cyc},
        {"t__DOT__w33__DOT__i65", "t.w33.i65", VLCU_MEMBER_STATE, 4, 0, 3, 65, 64, 0, T::OFF_t__DOT__w33__DOT__i65},
        {"t__DOT__w33__DOT__cyc", "t.w33.cyc", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__w33__DOT__cyc},
        {"t__DOT__w32__DOT__i65", "t.w32.i65", VLCU_MEMBER_STATE, 4, 0, 3, 65, 64, 0, T::OFF_t__DOT__w32__DOT__i65},
        {"t__DOT__w32__DOT__i1extp", "t.w32.i1extp", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__w32__DOT__i1extp},
        {"t__DOT__w32__DOT__cyc", "t.w32.cyc", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__w32__DOT__cyc},
        {"t__DOT__w31__DOT__i65", "t.w31.i65", VLCU_MEMBER_STATE, 4, 0, 3, 65, 64, 0, T::OFF_t__DOT__w31__DOT__i65},
        {"t__DOT__w31__DOT__i1extp", "t.w31.i1extp", VLCU_MEMBER_STATE, 4, 0, 1, 31, 30, 0, T::OFF_t__DOT__w31__DOT__i1extp},
        {"t__DOT__w31__DOT__cyc", "t.w31.cyc", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__w31__DOT__cyc},
        {"t__DOT__w1__DOT__i65", "t.w1.i65", VLCU_MEMBER_STATE, 4, 0, 3, 65, 64, 0, T::OFF_t__DOT__w1__DOT__i65},
        {"t__DOT__w1__DOT__cyc", "t.w1.cyc", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__w1__DOT__cyc},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"t__DOT__w97__DOT__i63", "t.w97.i63", VLCU_MEMBER_STATE, 8, 0, 1, 63, 62, 0, T::OFF_t__DOT__w97__DOT__i63},
        {"t__DOT__w96__DOT__i63", "t.w96.i63", VLCU_MEMBER_STATE, 8, 0, 1, 63, 62, 0, T::OFF_t__DOT__w96__DOT__i63},
        {"t__DOT__w95__DOT__i63", "t.w95.i63", VLCU_MEMBER_STATE, 8, 0, 1, 63, 62, 0, T::OFF_t__DOT__w95__DOT__i63},
        {"t__DOT__w65__DOT__i63", "t.w65.i63", VLCU_MEMBER_STATE, 8, 0, 1, 63, 62, 0, T::OFF_t__DOT__w65__DOT__i63},
        {"t__DOT__w64__DOT__i63", "t.w64.i63", VLCU_MEMBER_STATE, 8, 0, 1, 63, 62, 0, T::OFF_t__DOT__w64__DOT__i63},
        {"t__DOT__w64__DOT__i1extp", "t.w64.i1extp", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF_t__DOT__w64__DOT__i1extp},
        {"t__DOT__w64__DOT__i63ext", "t.w64.i63ext", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF_t__DOT__w64__DOT__i63ext},
        {"t__DOT__w64__DOT__i65ext", "t.w64.i65ext", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF_t__DOT__w64__DOT__i65ext},
        {"t__DOT__w63__DOT__i63", "t.w63.i63", VLCU_MEMBER_STATE, 8, 0, 1, 63, 62, 0, T::OFF_t__DOT__w63__DOT__i63},
        {"t__DOT__w63__DOT__i1extp", "t.w63.i1extp", VLCU_MEMBER_STATE, 8, 0, 1, 63, 62, 0, T::OFF_t__DOT__w63__DOT__i1extp},
        {"t__DOT__w63__DOT__i65ext", "t.w63.i65ext", VLCU_MEMBER_STATE, 8, 0, 1, 63, 62, 0, T::OFF_t__DOT__w63__DOT__i65ext},
        {"t__DOT__w33__DOT__i63", "t.w33.i63", VLCU_MEMBER_STATE, 8, 0, 1, 63, 62, 0, T::OFF_t__DOT__w33__DOT__i63},
        {"t__DOT__w33__DOT__i1extp", "t.w33.i1extp", VLCU_MEMBER_STATE, 8, 0, 1, 33, 32, 0, T::OFF_t__DOT__w33__DOT__i1extp},
        {"t__DOT__w33__DOT__i65ext", "t.w33.i65ext", VLCU_MEMBER_STATE, 8, 0, 1, 33, 32, 0, T::OFF_t__DOT__w33__DOT__i65ext},
        {"t__DOT__w32__DOT__i63", "t.w32.i63", VLCU_MEMBER_STATE, 8, 0, 1, 63, 62, 0, T::OFF_t__DOT__w32__DOT__i63},
        {"t__DOT__w31__DOT__i63", "t.w31.i63", VLCU_MEMBER_STATE, 8, 0, 1, 63, 62, 0, T::OFF_t__DOT__w31__DOT__i63},
        {"t__DOT__w1__DOT__i63", "t.w1.i63", VLCU_MEMBER_STATE, 8, 0, 1, 63, 62, 0, T::OFF_t__DOT__w1__DOT__i63},
        {"__VstlTriggered", "__VstlTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VstlTriggered},
        {"__VactTriggered", "__VactTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VactTriggered},
        {"__VnbaTriggered", "__VnbaTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VnbaTriggered},
    };
    static const char* const messages[] = {"*-* All Finished *-*\n", nullptr};
    static const char* const plusargs[] = {nullptr};
    static const VlcuTraceDecl traceDecls[] = {
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0}};  // (terminator, not counted)
    static const uint32_t traceOrder[] = {0};  // codes in dump order (+ terminator)
    static const uint32_t saveOrder[] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 36, 37, 38, 39, 40, 41, 42, 43, 44, 45, 46, 47, 48, 49, 50, 51, 52, 53, 54, 55, 56, 57, 58, 59, 60, 61, 62, 63, 64, 65, 66, 67, 68, 69, 70, 71, 72, 73, 74, 75, 76, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "t", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 1, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 77, 0, 0x99599e39f227d7e3ULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt_math_signed__cuda VLCU_TESTBENCHES);
