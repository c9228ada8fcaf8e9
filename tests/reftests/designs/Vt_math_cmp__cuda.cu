// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt_math_cmp"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt_math_cmp___024root final {
    CData/*0:0*/ clk;
    CData/*2:0*/ t__DOT__index_a;
    CData/*2:0*/ t__DOT__index_b;
    CData/*7:0*/ t__DOT__p128__DOT__exp;
    CData/*7:0*/ t__DOT__p128__DOT__got;
    CData/*7:0*/ t__DOT__p126__DOT__exp;
    CData/*7:0*/ t__DOT__p126__DOT__got;
    CData/*7:0*/ t__DOT__p72__DOT__exp;
    CData/*7:0*/ t__DOT__p72__DOT__got;
    CData/*7:0*/ t__DOT__p64__DOT__exp;
    CData/*7:0*/ t__DOT__p64__DOT__got;
    CData/*7:0*/ t__DOT__p63__DOT__exp;
    CData/*7:0*/ t__DOT__p63__DOT__got;
    CData/*7:0*/ t__DOT__p32__DOT__exp;
    CData/*7:0*/ t__DOT__p32__DOT__got;
    CData/*3:0*/ t__DOT__p4__DOT__as;
    CData/*3:0*/ t__DOT__p4__DOT__bs;
    CData/*7:0*/ t__DOT__p4__DOT__exp;
    CData/*7:0*/ t__DOT__p4__DOT__got;
    CData/*0:0*/ __VstlFirstIteration;
    CData/*0:0*/ __VstlPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    IData/*31:0*/ t__DOT__cyc;
    VlWide<4>/*127:0*/ t__DOT__p128__DOT__as;
    VlWide<4>/*127:0*/ t__DOT__p128__DOT__bs;
    IData/*31:0*/ t__DOT__p128__DOT__cyc;
    VlWide<4>/*125:0*/ t__DOT__p126__DOT__as;
    VlWide<4>/*125:0*/ t__DOT__p126__DOT__bs;
    IData/*31:0*/ t__DOT__p126__DOT__cyc;
    VlWide<3>/*71:0*/ t__DOT__p72__DOT__as;
    VlWide<3>/*71:0*/ t__DOT__p72__DOT__bs;
    IData/*31:0*/ t__DOT__p72__DOT__cyc;
    IData/*31:0*/ t__DOT__p64__DOT__cyc;
    IData/*31:0*/ t__DOT__p63__DOT__cyc;
    IData/*31:0*/ t__DOT__p32__DOT__as;
    IData/*31:0*/ t__DOT__p32__DOT__bs;
    IData/*31:0*/ t__DOT__p32__DOT__cyc;
    IData/*31:0*/ t__DOT__p4__DOT__cyc;
    IData/*31:0*/ __VactIterCount;
    QData/*63:0*/ t__DOT__p64__DOT__as;
    QData/*63:0*/ t__DOT__p64__DOT__bs;
    QData/*62:0*/ t__DOT__p63__DOT__as;
    QData/*62:0*/ t__DOT__p63__DOT__bs;
    VlcuMdArray<VlcuSharedArray<CData/*7:0*/, 25>, 5, 5> t__DOT__p128__DOT__results;  // run-time indexed: shared memory [index][lane]
    VlcuMdArray<VlcuSharedArray<CData/*7:0*/, 25>, 5, 5> t__DOT__p126__DOT__results;  // run-time indexed: shared memory [index][lane]
    VlcuMdArray<VlcuLocalArray<CData/*7:0*/, 25>, 5, 5> t__DOT__p72__DOT__results;  // thread-local memory
    VlcuMdArray<VlcuLocalArray<CData/*7:0*/, 25>, 5, 5> t__DOT__p64__DOT__results;  // thread-local memory
    VlcuMdArray<VlcuLocalArray<CData/*7:0*/, 25>, 5, 5> t__DOT__p63__DOT__results;  // thread-local memory
    VlcuMdArray<VlcuLocalArray<CData/*7:0*/, 25>, 5, 5> t__DOT__p32__DOT__results;  // thread-local memory
    VlcuMdArray<VlcuLocalArray<CData/*7:0*/, 25>, 5, 5> t__DOT__p4__DOT__results;  // thread-local memory
    VlUnpacked<QData/*63:0*/, 1> __VstlTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vt_math_cmp__cuda final {
    typedef Vt_math_cmp___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 188,
        OFF_t__DOT__index_a = 189,
        OFF_t__DOT__index_b = 190,
        OFF_t__DOT__p128__DOT__exp = 191,
        OFF_t__DOT__p128__DOT__got = 192,
        OFF_t__DOT__p126__DOT__exp = 193,
        OFF_t__DOT__p126__DOT__got = 194,
        OFF_t__DOT__p72__DOT__exp = 195,
        OFF_t__DOT__p72__DOT__got = 196,
        OFF_t__DOT__p64__DOT__exp = 197,
        OFF_t__DOT__p64__DOT__got = 198,
        OFF_t__DOT__p63__DOT__exp = 199,
        OFF_t__DOT__p63__DOT__got = 200,
        OFF_t__DOT__p32__DOT__exp = 201,
        OFF_t__DOT__p32__DOT__got = 202,
        OFF_t__DOT__p4__DOT__as = 203,
        OFF_t__DOT__p4__DOT__bs = 204,
        OFF_t__DOT__p4__DOT__exp = 205,
        OFF_t__DOT__p4__DOT__got = 206,
        OFF___VstlFirstIteration = 207,
        OFF___VstlPhaseResult = 208,
        OFF___Vtrigprevexpr___TOP__clk__0 = 209,
        OFF___VactPhaseResult = 210,
        OFF___VnbaPhaseResult = 211,
        OFF_t__DOT__cyc = 56,
        OFF_t__DOT__p128__DOT__as = 60,
        OFF_t__DOT__p128__DOT__bs = 76,
        OFF_t__DOT__p128__DOT__cyc = 92,
        OFF_t__DOT__p126__DOT__as = 96,
        OFF_t__DOT__p126__DOT__bs = 112,
        OFF_t__DOT__p126__DOT__cyc = 128,
        OFF_t__DOT__p72__DOT__as = 132,
        OFF_t__DOT__p72__DOT__bs = 144,
        OFF_t__DOT__p72__DOT__cyc = 156,
        OFF_t__DOT__p64__DOT__cyc = 160,
        OFF_t__DOT__p63__DOT__cyc = 164,
        OFF_t__DOT__p32__DOT__as = 168,
        OFF_t__DOT__p32__DOT__bs = 172,
        OFF_t__DOT__p32__DOT__cyc = 176,
        OFF_t__DOT__p4__DOT__cyc = 180,
        OFF___VactIterCount = 184,
        OFF_t__DOT__p64__DOT__as = 0,
        OFF_t__DOT__p64__DOT__bs = 8,
        OFF_t__DOT__p63__DOT__as = 16,
        OFF_t__DOT__p63__DOT__bs = 24,
        OFF_t__DOT__p128__DOT__results = 212,
        OFF_t__DOT__p126__DOT__results = 237,
        OFF_t__DOT__p72__DOT__results = 262,
        OFF_t__DOT__p64__DOT__results = 287,
        OFF_t__DOT__p63__DOT__results = 312,
        OFF_t__DOT__p32__DOT__results = 337,
        OFF_t__DOT__p4__DOT__results = 362,
        OFF___VstlTriggered = 32,
        OFF___VactTriggered = 40,
        OFF___VnbaTriggered = 48,
        kStateBytes = 392,  // S_persist
        kTouchedReadBytes = 209,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 201,  // S_state_wr
    };
    struct Locals final {
        CData/*7:0*/ t__DOT__p72__DOT__results[25];
        CData/*7:0*/ t__DOT__p64__DOT__results[25];
        CData/*7:0*/ t__DOT__p63__DOT__results[25];
        CData/*7:0*/ t__DOT__p32__DOT__results[25];
        CData/*7:0*/ t__DOT__p4__DOT__results[25];
    };
    struct Shared final {
        CData/*7:0*/ t__DOT__p128__DOT__results[25][VLCU_BLOCK];
        CData/*7:0*/ t__DOT__p126__DOT__results[25][VLCU_BLOCK];
    };
    VLCU_D static void bind(Root& s, Locals& l, Shared& sh, const VlcuState& st, uint64_t i) {
        (void)l; (void)sh; (void)st; (void)i;
        s.vlNamep = "TOP";
        s.t__DOT__p128__DOT__results.bind(sh.t__DOT__p128__DOT__results, threadIdx.x);
        s.t__DOT__p126__DOT__results.bind(sh.t__DOT__p126__DOT__results, threadIdx.x);
        s.t__DOT__p72__DOT__results.bind(l.t__DOT__p72__DOT__results);
        s.t__DOT__p64__DOT__results.bind(l.t__DOT__p64__DOT__results);
        s.t__DOT__p63__DOT__results.bind(l.t__DOT__p63__DOT__results);
        s.t__DOT__p32__DOT__results.bind(l.t__DOT__p32__DOT__results);
        s.t__DOT__p4__DOT__results.bind(l.t__DOT__p4__DOT__results);
    }
    VLCU_D static void load(Root& s, const VlcuState& st, uint64_t i) {
        s.clk = st.ld<VlcuMem8>(OFF_clk, i);
        s.t__DOT__index_a = st.ld<VlcuMem8>(OFF_t__DOT__index_a, i);
        s.t__DOT__index_b = st.ld<VlcuMem8>(OFF_t__DOT__index_b, i);
        s.t__DOT__p128__DOT__exp = st.ld<VlcuMem8>(OFF_t__DOT__p128__DOT__exp, i);
        s.t__DOT__p128__DOT__got = st.ld<VlcuMem8>(OFF_t__DOT__p128__DOT__got, i);
        s.t__DOT__p126__DOT__exp = st.ld<VlcuMem8>(OFF_t__DOT__p126__DOT__exp, i);
        s.t__DOT__p126__DOT__got = st.ld<VlcuMem8>(OFF_t__DOT__p126__DOT__got, i);
        s.t__DOT__p72__DOT__exp = st.ld<VlcuMem8>(OFF_t__DOT__p72__DOT__exp, i);
        s.t__DOT__p72__DOT__got = st.ld<VlcuMem8>(OFF_t__DOT__p72__DOT__got, i);
        s.t__DOT__p64__DOT__exp = st.ld<VlcuMem8>(OFF_t__DOT__p64__DOT__exp, i);
        s.t__DOT__p64__DOT__got = st.ld<VlcuMem8>(OFF_t__DOT__p64__DOT__got, i);
        s.t__DOT__p63__DOT__exp = st.ld<VlcuMem8>(OFF_t__DOT__p63__DOT__exp, i);
        s.t__DOT__p63__DOT__got = st.ld<VlcuMem8>(OFF_t__DOT__p63__DOT__got, i);
        s.t__DOT__p32__DOT__exp = st.ld<VlcuMem8>(OFF_t__DOT__p32__DOT__exp, i);
        s.t__DOT__p32__DOT__got = st.ld<VlcuMem8>(OFF_t__DOT__p32__DOT__got, i);
        s.t__DOT__p4__DOT__as = st.ld<VlcuMem8>(OFF_t__DOT__p4__DOT__as, i);
        s.t__DOT__p4__DOT__bs = st.ld<VlcuMem8>(OFF_t__DOT__p4__DOT__bs, i);
        s.t__DOT__p4__DOT__exp = st.ld<VlcuMem8>(OFF_t__DOT__p4__DOT__exp, i);
        s.t__DOT__p4__DOT__got = st.ld<VlcuMem8>(OFF_t__DOT__p4__DOT__got, i);
        s.__VstlFirstIteration = st.ld<VlcuMem8>(OFF___VstlFirstIteration, i);
        s.__VstlPhaseResult = st.ld<VlcuMem8>(OFF___VstlPhaseResult, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        s.t__DOT__cyc = st.ld<IData>(OFF_t__DOT__cyc, i);
        st.ldW(OFF_t__DOT__p128__DOT__as, i, s.t__DOT__p128__DOT__as);
        st.ldW(OFF_t__DOT__p128__DOT__bs, i, s.t__DOT__p128__DOT__bs);
        s.t__DOT__p128__DOT__cyc = st.ld<IData>(OFF_t__DOT__p128__DOT__cyc, i);
        st.ldW(OFF_t__DOT__p126__DOT__as, i, s.t__DOT__p126__DOT__as);
        st.ldW(OFF_t__DOT__p126__DOT__bs, i, s.t__DOT__p126__DOT__bs);
        s.t__DOT__p126__DOT__cyc = st.ld<IData>(OFF_t__DOT__p126__DOT__cyc, i);
        st.ldW(OFF_t__DOT__p72__DOT__as, i, s.t__DOT__p72__DOT__as);
        st.ldW(OFF_t__DOT__p72__DOT__bs, i, s.t__DOT__p72__DOT__bs);
        s.t__DOT__p72__DOT__cyc = st.ld<IData>(OFF_t__DOT__p72__DOT__cyc, i);
        s.t__DOT__p64__DOT__cyc = st.ld<IData>(OFF_t__DOT__p64__DOT__cyc, i);
        s.t__DOT__p63__DOT__cyc = st.ld<IData>(OFF_t__DOT__p63__DOT__cyc, i);
        s.t__DOT__p32__DOT__as = st.ld<IData>(OFF_t__DOT__p32__DOT__as, i);
        s.t__DOT__p32__DOT__bs = st.ld<IData>(OFF_t__DOT__p32__DOT__bs, i);
        s.t__DOT__p32__DOT__cyc = st.ld<IData>(OFF_t__DOT__p32__DOT__cyc, i);
        s.t__DOT__p4__DOT__cyc = st.ld<IData>(OFF_t__DOT__p4__DOT__cyc, i);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        s.t__DOT__p64__DOT__as = st.ld<QData>(OFF_t__DOT__p64__DOT__as, i);
        s.t__DOT__p64__DOT__bs = st.ld<QData>(OFF_t__DOT__p64__DOT__bs, i);
        s.t__DOT__p63__DOT__as = st.ld<QData>(OFF_t__DOT__p63__DOT__as, i);
        s.t__DOT__p63__DOT__bs = st.ld<QData>(OFF_t__DOT__p63__DOT__bs, i);
        for (uint32_t e = 0; e < 25; ++e) s.t__DOT__p128__DOT__results.flat()[e] = st.ldE<VlcuMem8>(OFF_t__DOT__p128__DOT__results, e, i);
        for (uint32_t e = 0; e < 25; ++e) s.t__DOT__p126__DOT__results.flat()[e] = st.ldE<VlcuMem8>(OFF_t__DOT__p126__DOT__results, e, i);
        for (uint32_t e = 0; e < 25; ++e) s.t__DOT__p72__DOT__results.flat()[e] = st.ldE<VlcuMem8>(OFF_t__DOT__p72__DOT__results, e, i);
        for (uint32_t e = 0; e < 25; ++e) s.t__DOT__p64__DOT__results.flat()[e] = st.ldE<VlcuMem8>(OFF_t__DOT__p64__DOT__results, e, i);
        for (uint32_t e = 0; e < 25; ++e) s.t__DOT__p63__DOT__results.flat()[e] = st.ldE<VlcuMem8>(OFF_t__DOT__p63__DOT__results, e, i);
        for (uint32_t e = 0; e < 25; ++e) s.t__DOT__p32__DOT__results.flat()[e] = st.ldE<VlcuMem8>(OFF_t__DOT__p32__DOT__results, e, i);
        for (uint32_t e = 0; e < 25; ++e) s.t__DOT__p4__DOT__results.flat()[e] = st.ldE<VlcuMem8>(OFF_t__DOT__p4__DOT__results, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VstlTriggered[e] = st.ldE<QData>(OFF___VstlTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.st<VlcuMem8>(OFF_t__DOT__index_a, i, s.t__DOT__index_a);
        st.st<VlcuMem8>(OFF_t__DOT__index_b, i, s.t__DOT__index_b);
        st.st<VlcuMem8>(OFF_t__DOT__p128__DOT__exp, i, s.t__DOT__p128__DOT__exp);
        st.st<VlcuMem8>(OFF_t__DOT__p128__DOT__got, i, s.t__DOT__p128__DOT__got);
        st.st<VlcuMem8>(OFF_t__DOT__p126__DOT__exp, i, s.t__DOT__p126__DOT__exp);
        st.st<VlcuMem8>(OFF_t__DOT__p126__DOT__got, i, s.t__DOT__p126__DOT__got);
        st.st<VlcuMem8>(OFF_t__DOT__p72__DOT__exp, i, s.t__DOT__p72__DOT__exp);
        st.st<VlcuMem8>(OFF_t__DOT__p72__DOT__got, i, s.t__DOT__p72__DOT__got);
        st.st<VlcuMem8>(OFF_t__DOT__p64__DOT__exp, i, s.t__DOT__p64__DOT__exp);
        st.st<VlcuMem8>(OFF_t__DOT__p64__DOT__got, i, s.t__DOT__p64__DOT__got);
        st.st<VlcuMem8>(OFF_t__DOT__p63__DOT__exp, i, s.t__DOT__p63__DOT__exp);
        st.st<VlcuMem8>(OFF_t__DOT__p63__DOT__got, i, s.t__DOT__p63__DOT__got);
        st.st<VlcuMem8>(OFF_t__DOT__p32__DOT__exp, i, s.t__DOT__p32__DOT__exp);
        st.st<VlcuMem8>(OFF_t__DOT__p32__DOT__got, i, s.t__DOT__p32__DOT__got);
        st.st<VlcuMem8>(OFF_t__DOT__p4__DOT__as, i, s.t__DOT__p4__DOT__as);
        st.st<VlcuMem8>(OFF_t__DOT__p4__DOT__bs, i, s.t__DOT__p4__DOT__bs);
        st.st<VlcuMem8>(OFF_t__DOT__p4__DOT__exp, i, s.t__DOT__p4__DOT__exp);
        st.st<VlcuMem8>(OFF_t__DOT__p4__DOT__got, i, s.t__DOT__p4__DOT__got);
        st.st<VlcuMem8>(OFF___VstlFirstIteration, i, s.__VstlFirstIteration);
        st.st<VlcuMem8>(OFF___VstlPhaseResult, i, s.__VstlPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.st<IData>(OFF_t__DOT__cyc, i, s.t__DOT__cyc);
        st.stW(OFF_t__DOT__p128__DOT__as, i, s.t__DOT__p128__DOT__as);
        st.stW(OFF_t__DOT__p128__DOT__bs, i, s.t__DOT__p128__DOT__bs);
        st.st<IData>(OFF_t__DOT__p128__DOT__cyc, i, s.t__DOT__p128__DOT__cyc);
        st.stW(OFF_t__DOT__p126__DOT__as, i, s.t__DOT__p126__DOT__as);
        st.stW(OFF_t__DOT__p126__DOT__bs, i, s.t__DOT__p126__DOT__bs);
        st.st<IData>(OFF_t__DOT__p126__DOT__cyc, i, s.t__DOT__p126__DOT__cyc);
        st.stW(OFF_t__DOT__p72__DOT__as, i, s.t__DOT__p72__DOT__as);
        st.stW(OFF_t__DOT__p72__DOT__bs, i, s.t__DOT__p72__DOT__bs);
        st.st<IData>(OFF_t__DOT__p72__DOT__cyc, i, s.t__DOT__p72__DOT__cyc);
        st.st<IData>(OFF_t__DOT__p64__DOT__cyc, i, s.t__DOT__p64__DOT__cyc);
        st.st<IData>(OFF_t__DOT__p63__DOT__cyc, i, s.t__DOT__p63__DOT__cyc);
        st.st<IData>(OFF_t__DOT__p32__DOT__as, i, s.t__DOT__p32__DOT__as);
        st.st<IData>(OFF_t__DOT__p32__DOT__bs, i, s.t__DOT__p32__DOT__bs);
        st.st<IData>(OFF_t__DOT__p32__DOT__cyc, i, s.t__DOT__p32__DOT__cyc);
        st.st<IData>(OFF_t__DOT__p4__DOT__cyc, i, s.t__DOT__p4__DOT__cyc);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
        st.st<QData>(OFF_t__DOT__p64__DOT__as, i, s.t__DOT__p64__DOT__as);
        st.st<QData>(OFF_t__DOT__p64__DOT__bs, i, s.t__DOT__p64__DOT__bs);
        st.st<QData>(OFF_t__DOT__p63__DOT__as, i, s.t__DOT__p63__DOT__as);
        st.st<QData>(OFF_t__DOT__p63__DOT__bs, i, s.t__DOT__p63__DOT__bs);
        for (uint32_t e = 0; e < 25; ++e) st.stE<VlcuMem8>(OFF_t__DOT__p128__DOT__results, e, i, s.t__DOT__p128__DOT__results.flat()[e]);
        for (uint32_t e = 0; e < 25; ++e) st.stE<VlcuMem8>(OFF_t__DOT__p126__DOT__results, e, i, s.t__DOT__p126__DOT__results.flat()[e]);
        for (uint32_t e = 0; e < 25; ++e) st.stE<VlcuMem8>(OFF_t__DOT__p72__DOT__results, e, i, s.t__DOT__p72__DOT__results.flat()[e]);
        for (uint32_t e = 0; e < 25; ++e) st.stE<VlcuMem8>(OFF_t__DOT__p64__DOT__results, e, i, s.t__DOT__p64__DOT__results.flat()[e]);
        for (uint32_t e = 0; e < 25; ++e) st.stE<VlcuMem8>(OFF_t__DOT__p63__DOT__results, e, i, s.t__DOT__p63__DOT__results.flat()[e]);
        for (uint32_t e = 0; e < 25; ++e) st.stE<VlcuMem8>(OFF_t__DOT__p32__DOT__results, e, i, s.t__DOT__p32__DOT__results.flat()[e]);
        for (uint32_t e = 0; e < 25; ++e) st.stE<VlcuMem8>(OFF_t__DOT__p4__DOT__results, e, i, s.t__DOT__p4__DOT__results.flat()[e]);
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

    VLCU_D static void Vt_math_cmp___024root___eval_static(Vt_math_cmp___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        {
            // Inlined CFunc: _eval_static__TOP
            vlSelfRef.t__DOT__p128__DOT__cyc = 0U;
            vlSelfRef.t__DOT__p126__DOT__cyc = 0U;
            vlSelfRef.t__DOT__p72__DOT__cyc = 0U;
            vlSelfRef.t__DOT__p64__DOT__cyc = 0U;
            vlSelfRef.t__DOT__p63__DOT__cyc = 0U;
            vlSelfRef.t__DOT__p32__DOT__cyc = 0U;
            vlSelfRef.t__DOT__p4__DOT__cyc = 0U;
        }
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
    }

    VLCU_D static void Vt_math_cmp___024root___eval_initial(Vt_math_cmp___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        Vt_math_cmp___024root___eval_initial__TOP(vlSelf, vlCtx);
    }

    VLCU_D static void Vt_math_cmp___024root___eval_initial__TOP(Vt_math_cmp___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlSelfRef.t__DOT__cyc = 0U;
        vlSelfRef.t__DOT__index_a = 0U;
        vlSelfRef.t__DOT__index_b = 0U;
        vlSelfRef.t__DOT__p128__DOT__results[0U][0U] = 0x33U;
        vlSelfRef.t__DOT__p128__DOT__results[0U][1U] = 0x0fU;
        vlSelfRef.t__DOT__p128__DOT__results[0U][2U] = 0x33U;
        vlSelfRef.t__DOT__p128__DOT__results[0U][3U] = 0x5aU;
        vlSelfRef.t__DOT__p128__DOT__results[0U][4U] = 0x5aU;
        vlSelfRef.t__DOT__p128__DOT__results[1U][0U] = 0xf0U;
        vlSelfRef.t__DOT__p128__DOT__results[1U][1U] = 0x33U;
        vlSelfRef.t__DOT__p128__DOT__results[1U][2U] = 0xf0U;
        vlSelfRef.t__DOT__p128__DOT__results[1U][3U] = 0x5aU;
        vlSelfRef.t__DOT__p128__DOT__results[1U][4U] = 0x5aU;
        vlSelfRef.t__DOT__p128__DOT__results[2U][0U] = 0x33U;
        vlSelfRef.t__DOT__p128__DOT__results[2U][1U] = 0x0fU;
        vlSelfRef.t__DOT__p128__DOT__results[2U][2U] = 0x33U;
        vlSelfRef.t__DOT__p128__DOT__results[2U][3U] = 0x5aU;
        vlSelfRef.t__DOT__p128__DOT__results[2U][4U] = 0x5aU;
        vlSelfRef.t__DOT__p128__DOT__results[3U][0U] = 0xa5U;
        vlSelfRef.t__DOT__p128__DOT__results[3U][1U] = 0xa5U;
        vlSelfRef.t__DOT__p128__DOT__results[3U][2U] = 0xa5U;
        vlSelfRef.t__DOT__p128__DOT__results[3U][3U] = 0x33U;
        vlSelfRef.t__DOT__p128__DOT__results[3U][4U] = 0xf0U;
        vlSelfRef.t__DOT__p128__DOT__results[4U][0U] = 0xa5U;
        vlSelfRef.t__DOT__p128__DOT__results[4U][1U] = 0xa5U;
        vlSelfRef.t__DOT__p128__DOT__results[4U][2U] = 0xa5U;
        vlSelfRef.t__DOT__p128__DOT__results[4U][3U] = 0x0fU;
        vlSelfRef.t__DOT__p128__DOT__results[4U][4U] = 0x33U;
        vlSelfRef.t__DOT__p126__DOT__results[0U][0U] = 0x33U;
        vlSelfRef.t__DOT__p126__DOT__results[0U][1U] = 0x0fU;
        vlSelfRef.t__DOT__p126__DOT__results[0U][2U] = 0x33U;
        vlSelfRef.t__DOT__p126__DOT__results[0U][3U] = 0x5aU;
        vlSelfRef.t__DOT__p126__DOT__results[0U][4U] = 0x5aU;
        vlSelfRef.t__DOT__p126__DOT__results[1U][0U] = 0xf0U;
        vlSelfRef.t__DOT__p126__DOT__results[1U][1U] = 0x33U;
        vlSelfRef.t__DOT__p126__DOT__results[1U][2U] = 0xf0U;
        vlSelfRef.t__DOT__p126__DOT__results[1U][3U] = 0x5aU;
        vlSelfRef.t__DOT__p126__DOT__results[1U][4U] = 0x5aU;
        vlSelfRef.t__DOT__p126__DOT__results[2U][0U] = 0x33U;
        vlSelfRef.t__DOT__p126__DOT__results[2U][1U] = 0x0fU;
        vlSelfRef.t__DOT__p126__DOT__results[2U][2U] = 0x33U;
        vlSelfRef.t__DOT__p126__DOT__results[2U][3U] = 0x5aU;
        vlSelfRef.t__DOT__p126__DOT__results[2U][4U] = 0x5aU;
        vlSelfRef.t__DOT__p126__DOT__results[3U][0U] = 0xa5U;
        vlSelfRef.t__DOT__p126__DOT__results[3U][1U] = 0xa5U;
        vlSelfRef.t__DOT__p126__DOT__results[3U][2U] = 0xa5U;
        vlSelfRef.t__DOT__p126__DOT__results[3U][3U] = 0x33U;
        vlSelfRef.t__DOT__p126__DOT__results[3U][4U] = 0xf0U;
        vlSelfRef.t__DOT__p126__DOT__results[4U][0U] = 0xa5U;
        vlSelfRef.t__DOT__p126__DOT__results[4U][1U] = 0xa5U;
        vlSelfRef.t__DOT__p126__DOT__results[4U][2U] = 0xa5U;
        vlSelfRef.t__DOT__p126__DOT__results[4U][3U] = 0x0fU;
        vlSelfRef.t__DOT__p126__DOT__results[4U][4U] = 0x33U;
        vlSelfRef.t__DOT__p72__DOT__results[0U][0U] = 0x33U;
        vlSelfRef.t__DOT__p72__DOT__results[0U][1U] = 0x0fU;
        vlSelfRef.t__DOT__p72__DOT__results[0U][2U] = 0x33U;
        vlSelfRef.t__DOT__p72__DOT__results[0U][3U] = 0x5aU;
        vlSelfRef.t__DOT__p72__DOT__results[0U][4U] = 0x5aU;
        vlSelfRef.t__DOT__p72__DOT__results[1U][0U] = 0xf0U;
        vlSelfRef.t__DOT__p72__DOT__results[1U][1U] = 0x33U;
        vlSelfRef.t__DOT__p72__DOT__results[1U][2U] = 0xf0U;
        vlSelfRef.t__DOT__p72__DOT__results[1U][3U] = 0x5aU;
        vlSelfRef.t__DOT__p72__DOT__results[1U][4U] = 0x5aU;
        vlSelfRef.t__DOT__p72__DOT__results[2U][0U] = 0x33U;
        vlSelfRef.t__DOT__p72__DOT__results[2U][1U] = 0x0fU;
        vlSelfRef.t__DOT__p72__DOT__results[2U][2U] = 0x33U;
        vlSelfRef.t__DOT__p72__DOT__results[2U][3U] = 0x5aU;
        vlSelfRef.t__DOT__p72__DOT__results[2U][4U] = 0x5aU;
        vlSelfRef.t__DOT__p72__DOT__results[3U][0U] = 0xa5U;
        vlSelfRef.t__DOT__p72__DOT__results[3U][1U] = 0xa5U;
        vlSelfRef.t__DOT__p72__DOT__results[3U][2U] = 0xa5U;
        vlSelfRef.t__DOT__p72__DOT__results[3U][3U] = 0x33U;
        vlSelfRef.t__DOT__p72__DOT__results[3U][4U] = 0xf0U;
        vlSelfRef.t__DOT__p72__DOT__results[4U][0U] = 0xa5U;
        vlSelfRef.t__DOT__p72__DOT__results[4U][1U] = 0xa5U;
        vlSelfRef.t__DOT__p72__DOT__results[4U][2U] = 0xa5U;
        vlSelfRef.t__DOT__p72__DOT__results[4U][3U] = 0x0fU;
        vlSelfRef.t__DOT__p72__DOT__results[4U][4U] = 0x33U;
        vlSelfRef.t__DOT__p64__DOT__results[0U][0U] = 0x33U;
        vlSelfRef.t__DOT__p64__DOT__results[0U][1U] = 0x0fU;
        vlSelfRef.t__DOT__p64__DOT__results[0U][2U] = 0x33U;
        vlSelfRef.t__DOT__p64__DOT__results[0U][3U] = 0x5aU;
        vlSelfRef.t__DOT__p64__DOT__results[0U][4U] = 0x5aU;
        vlSelfRef.t__DOT__p64__DOT__results[1U][0U] = 0xf0U;
        vlSelfRef.t__DOT__p64__DOT__results[1U][1U] = 0x33U;
        vlSelfRef.t__DOT__p64__DOT__results[1U][2U] = 0xf0U;
        vlSelfRef.t__DOT__p64__DOT__results[1U][3U] = 0x5aU;
        vlSelfRef.t__DOT__p64__DOT__results[1U][4U] = 0x5aU;
        vlSelfRef.t__DOT__p64__DOT__results[2U][0U] = 0x33U;
        vlSelfRef.t__DOT__p64__DOT__results[2U][1U] = 0x0fU;
        vlSelfRef.t__DOT__p64__DOT__results[2U][2U] = 0x33U;
        vlSelfRef.t__DOT__p64__DOT__results[2U][3U] = 0x5aU;
        vlSelfRef.t__DOT__p64__DOT__results[2U][4U] = 0x5aU;
        vlSelfRef.t__DOT__p64__DOT__results[3U][0U] = 0xa5U;
        vlSelfRef.t__DOT__p64__DOT__results[3U][1U] = 0xa5U;
        vlSelfRef.t__DOT__p64__DOT__results[3U][2U] = 0xa5U;
        vlSelfRef.t__DOT__p64__DOT__results[3U][3U] = 0x33U;
        vlSelfRef.t__DOT__p64__DOT__results[3U][4U] = 0xf0U;
        vlSelfRef.t__DOT__p64__DOT__results[4U][0U] = 0xa5U;
        vlSelfRef.t__DOT__p64__DOT__results[4U][1U] = 0xa5U;
        vlSelfRef.t__DOT__p64__DOT__results[4U][2U] = 0xa5U;
        vlSelfRef.t__DOT__p64__DOT__results[4U][3U] = 0x0fU;
        vlSelfRef.t__DOT__p64__DOT__results[4U][4U] = 0x33U;
        vlSelfRef.t__DOT__p63__DOT__results[0U][0U] = 0x33U;
        vlSelfRef.t__DOT__p63__DOT__results[0U][1U] = 0x0fU;
        vlSelfRef.t__DOT__p63__DOT__results[0U][2U] = 0x33U;
        vlSelfRef.t__DOT__p63__DOT__results[0U][3U] = 0x5aU;
        vlSelfRef.t__DOT__p63__DOT__results[0U][4U] = 0x5aU;
        vlSelfRef.t__DOT__p63__DOT__results[1U][0U] = 0xf0U;
        vlSelfRef.t__DOT__p63__DOT__results[1U][1U] = 0x33U;
        vlSelfRef.t__DOT__p63__DOT__results[1U][2U] = 0xf0U;
        vlSelfRef.t__DOT__p63__DOT__results[1U][3U] = 0x5aU;
        vlSelfRef.t__DOT__p63__DOT__results[1U][4U] = 0x5aU;
        vlSelfRef.t__DOT__p63__DOT__results[2U][0U] = 0x33U;
        vlSelfRef.t__DOT__p63__DOT__results[2U][1U] = 0x0fU;
        vlSelfRef.t__DOT__p63__DOT__results[2U][2U] = 0x33U;
        vlSelfRef.t__DOT__p63__DOT__results[2U][3U] = 0x5aU;
        vlSelfRef.t__DOT__p63__DOT__results[2U][4U] = 0x5aU;
        vlSelfRef.t__DOT__p63__DOT__results[3U][0U] = 0xa5U;
        vlSelfRef.t__DOT__p63__DOT__results[3U][1U] = 0xa5U;
        vlSelfRef.t__DOT__p63__DOT__results[3U][2U] = 0xa5U;
        vlSelfRef.t__DOT__p63__DOT__results[3U][3U] = 0x33U;
        vlSelfRef.t__DOT__p63__DOT__results[3U][4U] = 0xf0U;
        vlSelfRef.t__DOT__p63__DOT__results[4U][0U] = 0xa5U;
        vlSelfRef.t__DOT__p63__DOT__results[4U][1U] = 0xa5U;
        vlSelfRef.t__DOT__p63__DOT__results[4U][2U] = 0xa5U;
        vlSelfRef.t__DOT__p63__DOT__results[4U][3U] = 0x0fU;
        vlSelfRef.t__DOT__p63__DOT__results[4U][4U] = 0x33U;
        vlSelfRef.t__DOT__p32__DOT__results[0U][0U] = 0x33U;
        vlSelfRef.t__DOT__p32__DOT__results[0U][1U] = 0x0fU;
        vlSelfRef.t__DOT__p32__DOT__results[0U][2U] = 0x33U;
        vlSelfRef.t__DOT__p32__DOT__results[0U][3U] = 0x5aU;
        vlSelfRef.t__DOT__p32__DOT__results[0U][4U] = 0x5aU;
        vlSelfRef.t__DOT__p32__DOT__results[1U][0U] = 0xf0U;
        vlSelfRef.t__DOT__p32__DOT__results[1U][1U] = 0x33U;
        vlSelfRef.t__DOT__p32__DOT__results[1U][2U] = 0xf0U;
        vlSelfRef.t__DOT__p32__DOT__results[1U][3U] = 0x5aU;
        vlSelfRef.t__DOT__p32__DOT__results[1U][4U] = 0x5aU;
        vlSelfRef.t__DOT__p32__DOT__results[2U][0U] = 0x33U;
        vlSelfRef.t__DOT__p32__DOT__results[2U][1U] = 0x0fU;
        vlSelfRef.t__DOT__p32__DOT__results[2U][2U] = 0x33U;
        vlSelfRef.t__DOT__p32__DOT__results[2U][3U] = 0x5aU;
        vlSelfRef.t__DOT__p32__DOT__results[2U][4U] = 0x5aU;
        vlSelfRef.t__DOT__p32__DOT__results[3U][0U] = 0xa5U;
        vlSelfRef.t__DOT__p32__DOT__results[3U][1U] = 0xa5U;
        vlSelfRef.t__DOT__p32__DOT__results[3U][2U] = 0xa5U;
        vlSelfRef.t__DOT__p32__DOT__results[3U][3U] = 0x33U;
        vlSelfRef.t__DOT__p32__DOT__results[3U][4U] = 0xf0U;
        vlSelfRef.t__DOT__p32__DOT__results[4U][0U] = 0xa5U;
        vlSelfRef.t__DOT__p32__DOT__results[4U][1U] = 0xa5U;
        vlSelfRef.t__DOT__p32__DOT__results[4U][2U] = 0xa5U;
        vlSelfRef.t__DOT__p32__DOT__results[4U][3U] = 0x0fU;
        vlSelfRef.t__DOT__p32__DOT__results[4U][4U] = 0x33U;
        vlSelfRef.t__DOT__p4__DOT__results[0U][0U] = 0x33U;
        vlSelfRef.t__DOT__p4__DOT__results[0U][1U] = 0x0fU;
        vlSelfRef.t__DOT__p4__DOT__results[0U][2U] = 0x33U;
        vlSelfRef.t__DOT__p4__DOT__results[0U][3U] = 0x5aU;
        vlSelfRef.t__DOT__p4__DOT__results[0U][4U] = 0x5aU;
        vlSelfRef.t__DOT__p4__DOT__results[1U][0U] = 0xf0U;
        vlSelfRef.t__DOT__p4__DOT__results[1U][1U] = 0x33U;
        vlSelfRef.t__DOT__p4__DOT__results[1U][2U] = 0xf0U;
        vlSelfRef.t__DOT__p4__DOT__results[1U][3U] = 0x5aU;
        vlSelfRef.t__DOT__p4__DOT__results[1U][4U] = 0x5aU;
        vlSelfRef.t__DOT__p4__DOT__results[2U][0U] = 0x33U;
        vlSelfRef.t__DOT__p4__DOT__results[2U][1U] = 0x0fU;
        vlSelfRef.t__DOT__p4__DOT__results[2U][2U] = 0x33U;
        vlSelfRef.t__DOT__p4__DOT__results[2U][3U] = 0x5aU;
        vlSelfRef.t__DOT__p4__DOT__results[2U][4U] = 0x5aU;
        vlSelfRef.t__DOT__p4__DOT__results[3U][0U] = 0xa5U;
        vlSelfRef.t__DOT__p4__DOT__results[3U][1U] = 0xa5U;
        vlSelfRef.t__DOT__p4__DOT__results[3U][2U] = 0xa5U;
        vlSelfRef.t__DOT__p4__DOT__results[3U][3U] = 0x33U;
        vlSelfRef.t__DOT__p4__DOT__results[3U][4U] = 0xf0U;
        vlSelfRef.t__DOT__p4__DOT__results[4U][0U] = 0xa5U;
        vlSelfRef.t__DOT__p4__DOT__results[4U][1U] = 0xa5U;
        vlSelfRef.t__DOT__p4__DOT__results[4U][2U] = 0xa5U;
        vlSelfRef.t__DOT__p4__DOT__results[4U][3U] = 0x0fU;
        vlSelfRef.t__DOT__p4__DOT__results[4U][4U] = 0x33U;
    }

    VLCU_D static void Vt_math_cmp___024root___eval_final(Vt_math_cmp___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_math_cmp___024root___eval_settle(Vt_math_cmp___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt_math_cmp___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
#endif
                VL_FATAL_MT("/root/reference/test_regress/t/t_math_cmp.v", 7, "", "DIDNOTCONVERGE: Settle region did not converge after '--converge-limit' of 10000 tries");
            }
            __VstlIterCount = ((IData)(1U) + __VstlIterCount);
            vlSelfRef.__VstlPhaseResult = Vt_math_cmp___024root___eval_phase__stl(vlSelf, vlCtx);
            vlSelfRef.__VstlFirstIteration = 0U;
        } while (vlSelfRef.__VstlPhaseResult);
    }

    VLCU_D static bool Vt_math_cmp___024root___trigger_anySet__stl(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt_math_cmp___024root___stl_sequent__TOP__0(Vt_math_cmp___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlSelfRef.t__DOT__index_a = (7U & vlSelfRef.t__DOT__cyc);
        if ((4U < (IData)(vlSelfRef.t__DOT__index_a))) {
            vlSelfRef.t__DOT__index_a = 4U;
        }
        vlSelfRef.t__DOT__index_b = (7U & (vlSelfRef.t__DOT__cyc 
                                           >> 3U));
        if ((4U < (IData)(vlSelfRef.t__DOT__index_b))) {
            vlSelfRef.t__DOT__index_b = 4U;
        }
        if ((4U & (IData)(vlSelfRef.t__DOT__index_a))) {
            if (VL_UNLIKELY(((2U & (IData)(vlSelfRef.t__DOT__index_a))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 101, "");
            } else if (VL_UNLIKELY(((1U & (IData)(vlSelfRef.t__DOT__index_a))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 101, "");
            } else {
                vlSelfRef.t__DOT__p128__DOT__as[0U] = 0U;
                vlSelfRef.t__DOT__p128__DOT__as[1U] = 0U;
                vlSelfRef.t__DOT__p128__DOT__as[2U] = 0U;
                vlSelfRef.t__DOT__p128__DOT__as[3U] = 0x80000000U;
            }
        } else if ((2U & (IData)(vlSelfRef.t__DOT__index_a))) {
            if ((1U & (IData)(vlSelfRef.t__DOT__index_a))) {
                vlSelfRef.t__DOT__p128__DOT__as[0U] = 0xffffffffU;
                vlSelfRef.t__DOT__p128__DOT__as[1U] = 0xffffffffU;
                vlSelfRef.t__DOT__p128__DOT__as[2U] = 0xffffffffU;
                vlSelfRef.t__DOT__p128__DOT__as[3U] = 0xffffffffU;
            } else {
                vlSelfRef.t__DOT__p128__DOT__as[0U] = 0U;
                vlSelfRef.t__DOT__p128__DOT__as[1U] = 0U;
                vlSelfRef.t__DOT__p128__DOT__as[2U] = 0U;
                vlSelfRef.t__DOT__p128__DOT__as[3U] = 0U;
            }
        } else {
            vlSelfRef.t__DOT__p128__DOT__as[0U] = (
                                                   (1U 
                                                    & (IData)(vlSelfRef.t__DOT__index_a))
                                                    ? 1U
                                                    : 0U);
            vlSelfRef.t__DOT__p128__DOT__as[1U] = 0U;
            vlSelfRef.t__DOT__p128__DOT__as[2U] = 0U;
            vlSelfRef.t__DOT__p128__DOT__as[3U] = 0U;
        }
        if ((4U & (IData)(vlSelfRef.t__DOT__index_b))) {
            if (VL_UNLIKELY(((2U & (IData)(vlSelfRef.t__DOT__index_b))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 109, "");
            } else if (VL_UNLIKELY(((1U & (IData)(vlSelfRef.t__DOT__index_b))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 109, "");
            } else {
                vlSelfRef.t__DOT__p128__DOT__bs[0U] = 0U;
                vlSelfRef.t__DOT__p128__DOT__bs[1U] = 0U;
                vlSelfRef.t__DOT__p128__DOT__bs[2U] = 0U;
                vlSelfRef.t__DOT__p128__DOT__bs[3U] = 0x80000000U;
            }
        } else if ((2U & (IData)(vlSelfRef.t__DOT__index_b))) {
            if ((1U & (IData)(vlSelfRef.t__DOT__index_b))) {
                vlSelfRef.t__DOT__p128__DOT__bs[0U] = 0xffffffffU;
                vlSelfRef.t__DOT__p128__DOT__bs[1U] = 0xffffffffU;
                vlSelfRef.t__DOT__p128__DOT__bs[2U] = 0xffffffffU;
                vlSelfRef.t__DOT__p128__DOT__bs[3U] = 0xffffffffU;
            } else {
                vlSelfRef.t__DOT__p128__DOT__bs[0U] = 0U;
                vlSelfRef.t__DOT__p128__DOT__bs[1U] = 0U;
                vlSelfRef.t__DOT__p128__DOT__bs[2U] = 0U;
                vlSelfRef.t__DOT__p128__DOT__bs[3U] = 0U;
            }
        } else {
            vlSelfRef.t__DOT__p128__DOT__bs[0U] = (
                                                   (1U 
                                                    & (IData)(vlSelfRef.t__DOT__index_b))
                                                    ? 1U
                                                    : 0U);
            vlSelfRef.t__DOT__p128__DOT__bs[1U] = 0U;
            vlSelfRef.t__DOT__p128__DOT__bs[2U] = 0U;
            vlSelfRef.t__DOT__p128__DOT__bs[3U] = 0U;
        }
        if ((4U & (IData)(vlSelfRef.t__DOT__index_a))) {
            if (VL_UNLIKELY(((2U & (IData)(vlSelfRef.t__DOT__index_a))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 101, "");
            } else if (VL_UNLIKELY(((1U & (IData)(vlSelfRef.t__DOT__index_a))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 101, "");
            } else {
                vlSelfRef.t__DOT__p126__DOT__as[0U] = 0U;
                vlSelfRef.t__DOT__p126__DOT__as[1U] = 0U;
                vlSelfRef.t__DOT__p126__DOT__as[2U] = 0U;
                vlSelfRef.t__DOT__p126__DOT__as[3U] = 0x20000000U;
            }
        } else if ((2U & (IData)(vlSelfRef.t__DOT__index_a))) {
            if ((1U & (IData)(vlSelfRef.t__DOT__index_a))) {
                vlSelfRef.t__DOT__p126__DOT__as[0U] = 0xffffffffU;
                vlSelfRef.t__DOT__p126__DOT__as[1U] = 0xffffffffU;
                vlSelfRef.t__DOT__p126__DOT__as[2U] = 0xffffffffU;
                vlSelfRef.t__DOT__p126__DOT__as[3U] = 0x3fffffffU;
            } else {
                vlSelfRef.t__DOT__p126__DOT__as[0U] = 0U;
                vlSelfRef.t__DOT__p126__DOT__as[1U] = 0U;
                vlSelfRef.t__DOT__p126__DOT__as[2U] = 0U;
                vlSelfRef.t__DOT__p126__DOT__as[3U] = 0U;
            }
        } else {
            vlSelfRef.t__DOT__p126__DOT__as[0U] = (
                                                   (1U 
                                                    & (IData)(vlSelfRef.t__DOT__index_a))
                                                    ? 1U
                                                    : 0U);
            vlSelfRef.t__DOT__p126__DOT__as[1U] = 0U;
            vlSelfRef.t__DOT__p126__DOT__as[2U] = 0U;
            vlSelfRef.t__DOT__p126__DOT__as[3U] = 0U;
        }
        if ((4U & (IData)(vlSelfRef.t__DOT__index_b))) {
            if (VL_UNLIKELY(((2U & (IData)(vlSelfRef.t__DOT__index_b))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 109, "");
            } else if (VL_UNLIKELY(((1U & (IData)(vlSelfRef.t__DOT__index_b))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 109, "");
            } else {
                vlSelfRef.t__DOT__p126__DOT__bs[0U] = 0U;
                vlSelfRef.t__DOT__p126__DOT__bs[1U] = 0U;
                vlSelfRef.t__DOT__p126__DOT__bs[2U] = 0U;
                vlSelfRef.t__DOT__p126__DOT__bs[3U] = 0x20000000U;
            }
        } else if ((2U & (IData)(vlSelfRef.t__DOT__index_b))) {
            if ((1U & (IData)(vlSelfRef.t__DOT__index_b))) {
                vlSelfRef.t__DOT__p126__DOT__bs[0U] = 0xffffffffU;
                vlSelfRef.t__DOT__p126__DOT__bs[1U] = 0xffffffffU;
                vlSelfRef.t__DOT__p126__DOT__bs[2U] = 0xffffffffU;
                vlSelfRef.t__DOT__p126__DOT__bs[3U] = 0x3fffffffU;
            } else {
                vlSelfRef.t__DOT__p126__DOT__bs[0U] = 0U;
                vlSelfRef.t__DOT__p126__DOT__bs[1U] = 0U;
                vlSelfRef.t__DOT__p126__DOT__bs[2U] = 0U;
                vlSelfRef.t__DOT__p126__DOT__bs[3U] = 0U;
            }
        } else {
            vlSelfRef.t__DOT__p126__DOT__bs[0U] = (
                                                   (1U 
                                                    & (IData)(vlSelfRef.t__DOT__index_b))
                                                    ? 1U
                                                    : 0U);
            vlSelfRef.t__DOT__p126__DOT__bs[1U] = 0U;
            vlSelfRef.t__DOT__p126__DOT__bs[2U] = 0U;
            vlSelfRef.t__DOT__p126__DOT__bs[3U] = 0U;
        }
        if ((4U & (IData)(vlSelfRef.t__DOT__index_a))) {
            if (VL_UNLIKELY(((2U & (IData)(vlSelfRef.t__DOT__index_a))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 101, "");
            } else if (VL_UNLIKELY(((1U & (IData)(vlSelfRef.t__DOT__index_a))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 101, "");
            } else {
                vlSelfRef.t__DOT__p72__DOT__as[0U] = 0U;
                vlSelfRef.t__DOT__p72__DOT__as[1U] = 0U;
                vlSelfRef.t__DOT__p72__DOT__as[2U] = 0x00000080U;
            }
        } else if ((2U & (IData)(vlSelfRef.t__DOT__index_a))) {
            if ((1U & (IData)(vlSelfRef.t__DOT__index_a))) {
                vlSelfRef.t__DOT__p72__DOT__as[0U] = 0xffffffffU;
                vlSelfRef.t__DOT__p72__DOT__as[1U] = 0xffffffffU;
                vlSelfRef.t__DOT__p72__DOT__as[2U] = 0x000000ffU;
            } else {
                vlSelfRef.t__DOT__p72__DOT__as[0U] = 0U;
                vlSelfRef.t__DOT__p72__DOT__as[1U] = 0U;
                vlSelfRef.t__DOT__p72__DOT__as[2U] = 0U;
            }
        } else {
            vlSelfRef.t__DOT__p72__DOT__as[0U] = ((1U 
                                                   & (IData)(vlSelfRef.t__DOT__index_a))
                                                   ? 1U
                                                   : 0U);
            vlSelfRef.t__DOT__p72__DOT__as[1U] = 0U;
            vlSelfRef.t__DOT__p72__DOT__as[2U] = 0U;
        }
        if ((4U & (IData)(vlSelfRef.t__DOT__index_b))) {
            if (VL_UNLIKELY(((2U & (IData)(vlSelfRef.t__DOT__index_b))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 109, "");
            } else if (VL_UNLIKELY(((1U & (IData)(vlSelfRef.t__DOT__index_b))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 109, "");
            } else {
                vlSelfRef.t__DOT__p72__DOT__bs[0U] = 0U;
                vlSelfRef.t__DOT__p72__DOT__bs[1U] = 0U;
                vlSelfRef.t__DOT__p72__DOT__bs[2U] = 0x00000080U;
            }
        } else if ((2U & (IData)(vlSelfRef.t__DOT__index_b))) {
            if ((1U & (IData)(vlSelfRef.t__DOT__index_b))) {
                vlSelfRef.t__DOT__p72__DOT__bs[0U] = 0xffffffffU;
                vlSelfRef.t__DOT__p72__DOT__bs[1U] = 0xffffffffU;
                vlSelfRef.t__DOT__p72__DOT__bs[2U] = 0x000000ffU;
            } else {
                vlSelfRef.t__DOT__p72__DOT__bs[0U] = 0U;
                vlSelfRef.t__DOT__p72__DOT__bs[1U] = 0U;
                vlSelfRef.t__DOT__p72__DOT__bs[2U] = 0U;
            }
        } else {
            vlSelfRef.t__DOT__p72__DOT__bs[0U] = ((1U 
                                                   & (IData)(vlSelfRef.t__DOT__index_b))
                                                   ? 1U
                                                   : 0U);
            vlSelfRef.t__DOT__p72__DOT__bs[1U] = 0U;
            vlSelfRef.t__DOT__p72__DOT__bs[2U] = 0U;
        }
        if ((4U & (IData)(vlSelfRef.t__DOT__index_a))) {
            if (VL_UNLIKELY(((2U & (IData)(vlSelfRef.t__DOT__index_a))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 101, "");
            } else if (VL_UNLIKELY(((1U & (IData)(vlSelfRef.t__DOT__index_a))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 101, "");
            } else {
                vlSelfRef.t__DOT__p64__DOT__as = 0x8000000000000000ULL;
            }
        } else {
            vlSelfRef.t__DOT__p64__DOT__as = ((2U & (IData)(vlSelfRef.t__DOT__index_a))
                                               ? ((1U 
                                                   & (IData)(vlSelfRef.t__DOT__index_a))
                                                   ? 0xffffffffffffffffULL
                                                   : 0ULL)
                                               : ((1U 
                                                   & (IData)(vlSelfRef.t__DOT__index_a))
                                                   ? 1ULL
                                                   : 0ULL));
        }
        if ((4U & (IData)(vlSelfRef.t__DOT__index_b))) {
            if (VL_UNLIKELY(((2U & (IData)(vlSelfRef.t__DOT__index_b))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 109, "");
            } else if (VL_UNLIKELY(((1U & (IData)(vlSelfRef.t__DOT__index_b))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 109, "");
            } else {
                vlSelfRef.t__DOT__p64__DOT__bs = 0x8000000000000000ULL;
            }
        } else {
            vlSelfRef.t__DOT__p64__DOT__bs = ((2U & (IData)(vlSelfRef.t__DOT__index_b))
                                               ? ((1U 
                                                   & (IData)(vlSelfRef.t__DOT__index_b))
                                                   ? 0xffffffffffffffffULL
                                                   : 0ULL)
                                               : ((1U 
                                                   & (IData)(vlSelfRef.t__DOT__index_b))
                                                   ? 1ULL
                                                   : 0ULL));
        }
        if ((4U & (IData)(vlSelfRef.t__DOT__index_a))) {
            if (VL_UNLIKELY(((2U & (IData)(vlSelfRef.t__DOT__index_a))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 101, "");
            } else if (VL_UNLIKELY(((1U & (IData)(vlSelfRef.t__DOT__index_a))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 101, "");
            } else {
                vlSelfRef.t__DOT__p63__DOT__as = 0x4000000000000000ULL;
            }
        } else {
            vlSelfRef.t__DOT__p63__DOT__as = ((2U & (IData)(vlSelfRef.t__DOT__index_a))
                                               ? ((1U 
                                                   & (IData)(vlSelfRef.t__DOT__index_a))
                                                   ? 0x7fffffffffffffffULL
                                                   : 0ULL)
                                               : ((1U 
                                                   & (IData)(vlSelfRef.t__DOT__index_a))
                                                   ? 1ULL
                                                   : 0ULL));
        }
        if ((4U & (IData)(vlSelfRef.t__DOT__index_b))) {
            if (VL_UNLIKELY(((2U & (IData)(vlSelfRef.t__DOT__index_b))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 109, "");
            } else if (VL_UNLIKELY(((1U & (IData)(vlSelfRef.t__DOT__index_b))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 109, "");
            } else {
                vlSelfRef.t__DOT__p63__DOT__bs = 0x4000000000000000ULL;
            }
        } else {
            vlSelfRef.t__DOT__p63__DOT__bs = ((2U & (IData)(vlSelfRef.t__DOT__index_b))
                                               ? ((1U 
                                                   & (IData)(vlSelfRef.t__DOT__index_b))
                                                   ? 0x7fffffffffffffffULL
                                                   : 0ULL)
                                               : ((1U 
                                                   & (IData)(vlSelfRef.t__DOT__index_b))
                                                   ? 1ULL
                                                   : 0ULL));
        }
        if ((4U & (IData)(vlSelfRef.t__DOT__index_a))) {
            if (VL_UNLIKELY(((2U & (IData)(vlSelfRef.t__DOT__index_a))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 101, "");
            } else if (VL_UNLIKELY(((1U & (IData)(vlSelfRef.t__DOT__index_a))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 101, "");
            } else {
                vlSelfRef.t__DOT__p32__DOT__as = 0x80000000U;
            }
        } else {
            vlSelfRef.t__DOT__p32__DOT__as = ((2U & (IData)(vlSelfRef.t__DOT__index_a))
                                               ? ((1U 
                                                   & (IData)(vlSelfRef.t__DOT__index_a))
                                                   ? 0xffffffffU
                                                   : 0U)
                                               : ((1U 
                                                   & (IData)(vlSelfRef.t__DOT__index_a))
                                                   ? 1U
                                                   : 0U));
        }
        if ((4U & (IData)(vlSelfRef.t__DOT__index_b))) {
            if (VL_UNLIKELY(((2U & (IData)(vlSelfRef.t__DOT__index_b))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 109, "");
            } else if (VL_UNLIKELY(((1U & (IData)(vlSelfRef.t__DOT__index_b))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 109, "");
            } else {
                vlSelfRef.t__DOT__p32__DOT__bs = 0x80000000U;
            }
        } else {
            vlSelfRef.t__DOT__p32__DOT__bs = ((2U & (IData)(vlSelfRef.t__DOT__index_b))
                                               ? ((1U 
                                                   & (IData)(vlSelfRef.t__DOT__index_b))
                                                   ? 0xffffffffU
                                                   : 0U)
                                               : ((1U 
                                                   & (IData)(vlSelfRef.t__DOT__index_b))
                                                   ? 1U
                                                   : 0U));
        }
        if ((4U & (IData)(vlSelfRef.t__DOT__index_a))) {
            if (VL_UNLIKELY(((2U & (IData)(vlSelfRef.t__DOT__index_a))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 101, "");
            } else if (VL_UNLIKELY(((1U & (IData)(vlSelfRef.t__DOT__index_a))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 101, "");
            } else {
                vlSelfRef.t__DOT__p4__DOT__as = 8U;
            }
        } else {
            vlSelfRef.t__DOT__p4__DOT__as = ((2U & (IData)(vlSelfRef.t__DOT__index_a))
                                              ? ((1U 
                                                  & (IData)(vlSelfRef.t__DOT__index_a))
                                                  ? 0x0fU
                                                  : 0U)
                                              : ((1U 
                                                  & (IData)(vlSelfRef.t__DOT__index_a))
                                                  ? 1U
                                                  : 0U));
        }
        if ((4U & (IData)(vlSelfRef.t__DOT__index_b))) {
            if (VL_UNLIKELY(((2U & (IData)(vlSelfRef.t__DOT__index_b))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 109, "");
            } else if (VL_UNLIKELY(((1U & (IData)(vlSelfRef.t__DOT__index_b))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 109, "");
            } else {
                vlSelfRef.t__DOT__p4__DOT__bs = 8U;
            }
        } else {
            vlSelfRef.t__DOT__p4__DOT__bs = ((2U & (IData)(vlSelfRef.t__DOT__index_b))
                                              ? ((1U 
                                                  & (IData)(vlSelfRef.t__DOT__index_b))
                                                  ? 0x0fU
                                                  : 0U)
                                              : ((1U 
                                                  & (IData)(vlSelfRef.t__DOT__index_b))
                                                  ? 1U
                                                  : 0U));
        }
    }

    VLCU_D static bool Vt_math_cmp___024root___eval_phase__stl(Vt_math_cmp___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt_math_cmp___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
        }
#endif
        __VstlExecute = Vt_math_cmp___024root___trigger_anySet__stl(vlSelfRef.__VstlTriggered);
        if (__VstlExecute) {
            {
                // Inlined CFunc: _eval_stl
                if ((1ULL & vlSelfRef.__VstlTriggered[0U])) {
                    Vt_math_cmp___024root___stl_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
        }
        return (__VstlExecute);
    }

    VLCU_D static bool Vt_math_cmp___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt_math_cmp___024root___nba_sequent__TOP__0(Vt_math_cmp___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        IData/*31:0*/ __Vdly__t__DOT__cyc;
        __Vdly__t__DOT__cyc = 0;
        IData/*31:0*/ __Vdly__t__DOT__p128__DOT__cyc;
        __Vdly__t__DOT__p128__DOT__cyc = 0;
        IData/*31:0*/ __Vdly__t__DOT__p126__DOT__cyc;
        __Vdly__t__DOT__p126__DOT__cyc = 0;
        IData/*31:0*/ __Vdly__t__DOT__p72__DOT__cyc;
        __Vdly__t__DOT__p72__DOT__cyc = 0;
        IData/*31:0*/ __Vdly__t__DOT__p64__DOT__cyc;
        __Vdly__t__DOT__p64__DOT__cyc = 0;
        IData/*31:0*/ __Vdly__t__DOT__p63__DOT__cyc;
        __Vdly__t__DOT__p63__DOT__cyc = 0;
        IData/*31:0*/ __Vdly__t__DOT__p32__DOT__cyc;
        __Vdly__t__DOT__p32__DOT__cyc = 0;
        IData/*31:0*/ __Vdly__t__DOT__p4__DOT__cyc;
        __Vdly__t__DOT__p4__DOT__cyc = 0;
        // Body
        __Vdly__t__DOT__p128__DOT__cyc = vlSelfRef.t__DOT__p128__DOT__cyc;
        __Vdly__t__DOT__p126__DOT__cyc = vlSelfRef.t__DOT__p126__DOT__cyc;
        __Vdly__t__DOT__p72__DOT__cyc = vlSelfRef.t__DOT__p72__DOT__cyc;
        __Vdly__t__DOT__p64__DOT__cyc = vlSelfRef.t__DOT__p64__DOT__cyc;
        __Vdly__t__DOT__p63__DOT__cyc = vlSelfRef.t__DOT__p63__DOT__cyc;
        __Vdly__t__DOT__p32__DOT__cyc = vlSelfRef.t__DOT__p32__DOT__cyc;
        __Vdly__t__DOT__p4__DOT__cyc = vlSelfRef.t__DOT__p4__DOT__cyc;
        __Vdly__t__DOT__cyc = vlSelfRef.t__DOT__cyc;
        __Vdly__t__DOT__p128__DOT__cyc = ((IData)(1U) 
                                          + vlSelfRef.t__DOT__p128__DOT__cyc);
        if (VL_LTS_III(32, 2U, vlSelfRef.t__DOT__p128__DOT__cyc)) {
            vlSelfRef.t__DOT__p128__DOT__exp = ((4U 
                                                 >= (IData)(vlSelfRef.t__DOT__index_b))
                                                 ? vlSelfRef.t__DOT__p128__DOT__results
                                                [((4U 
                                                   >= (IData)(vlSelfRef.t__DOT__index_a))
                                                   ? (IData)(vlSelfRef.t__DOT__index_a)
                                                   : 0U)]
                                                [vlSelfRef.t__DOT__index_b]
                                                 : 0U);
            vlSelfRef.t__DOT__p128__DOT__got = ((((
                                                   (VL_GT_W(4, vlSelfRef.t__DOT__p128__DOT__as, vlSelfRef.t__DOT__p128__DOT__bs) 
                                                    << 3U) 
                                                   | (VL_GTS_IWW(128, vlSelfRef.t__DOT__p128__DOT__as, vlSelfRef.t__DOT__p128__DOT__bs) 
                                                      << 2U)) 
                                                  | ((VL_GTE_W(4, vlSelfRef.t__DOT__p128__DOT__as, vlSelfRef.t__DOT__p128__DOT__bs) 
                                                      << 1U) 
                                                     | VL_GTES_IWW(128, vlSelfRef.t__DOT__p128__DOT__as, vlSelfRef.t__DOT__p128__DOT__bs))) 
                                                 << 4U) 
                                                | (((VL_LT_W(4, vlSelfRef.t__DOT__p128__DOT__as, vlSelfRef.t__DOT__p128__DOT__bs) 
                                                     << 3U) 
                                                    | (VL_LTS_IWW(128, vlSelfRef.t__DOT__p128__DOT__as, vlSelfRef.t__DOT__p128__DOT__bs) 
                                                       << 2U)) 
                                                   | ((VL_LTE_W(4, vlSelfRef.t__DOT__p128__DOT__as, vlSelfRef.t__DOT__p128__DOT__bs) 
                                                       << 1U) 
                                                      | VL_LTES_IWW(128, vlSelfRef.t__DOT__p128__DOT__as, vlSelfRef.t__DOT__p128__DOT__bs))));
            if (VL_UNLIKELY((((IData)(vlSelfRef.t__DOT__p128__DOT__exp) 
                              != (IData)(vlSelfRef.t__DOT__p128__DOT__got))))) {
                vlCtx.writef(0U, 4U, vlcu_msg_arg(vlSelfRef.t__DOT__index_a), vlcu_msg_arg((IData)(vlSelfRef.t__DOT__index_b)), vlcu_msg_arg(vlSelfRef.t__DOT__p128__DOT__got), vlcu_msg_arg((IData)(vlSelfRef.t__DOT__p128__DOT__exp)));
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 141, "");
            }
        }
        __Vdly__t__DOT__p126__DOT__cyc = ((IData)(1U) 
                                          + vlSelfRef.t__DOT__p126__DOT__cyc);
        if (VL_LTS_III(32, 2U, vlSelfRef.t__DOT__p126__DOT__cyc)) {
            vlSelfRef.t__DOT__p126__DOT__exp = ((4U 
                                                 >= (IData)(vlSelfRef.t__DOT__index_b))
                                                 ? vlSelfRef.t__DOT__p126__DOT__results
                                                [((4U 
                                                   >= (IData)(vlSelfRef.t__DOT__index_a))
                                                   ? (IData)(vlSelfRef.t__DOT__index_a)
                                                   : 0U)]
                                                [vlSelfRef.t__DOT__index_b]
                                                 : 0U);
            vlSelfRef.t__DOT__p126__DOT__got = ((((
                                                   (VL_GT_W(4, vlSelfRef.t__DOT__p126__DOT__as, vlSelfRef.t__DOT__p126__DOT__bs) 
                                                    << 3U) 
                                                   | (VL_GTS_IWW(126, vlSelfRef.t__DOT__p126__DOT__as, vlSelfRef.t__DOT__p126__DOT__bs) 
                                                      << 2U)) 
                                                  | ((VL_GTE_W(4, vlSelfRef.t__DOT__p126__DOT__as, vlSelfRef.t__DOT__p126__DOT__bs) 
                                                      << 1U) 
                                                     | VL_GTES_IWW(126, vlSelfRef.t__DOT__p126__DOT__as, vlSelfRef.t__DOT__p126__DOT__bs))) 
                                                 << 4U) 
                                                | (((VL_LT_W(4, vlSelfRef.t__DOT__p126__DOT__as, vlSelfRef.t__DOT__p126__DOT__bs) 
                                                     << 3U) 
                                                    | (VL_LTS_IWW(126, vlSelfRef.t__DOT__p126__DOT__as, vlSelfRef.t__DOT__p126__DOT__bs) 
                                                       << 2U)) 
                                                   | ((VL_LTE_W(4, vlSelfRef.t__DOT__p126__DOT__as, vlSelfRef.t__DOT__p126__DOT__bs) 
                                                       << 1U) 
                                                      | VL_LTES_IWW(126, vlSelfRef.t__DOT__p126__DOT__as, vlSelfRef.t__DOT__p126__DOT__bs))));
            if (VL_UNLIKELY((((IData)(vlSelfRef.t__DOT__p126__DOT__exp) 
                              != (IData)(vlSelfRef.t__DOT__p126__DOT__got))))) {
                vlCtx.writef(1U, 4U, vlcu_msg_arg(vlSelfRef.t__DOT__index_a), vlcu_msg_arg((IData)(vlSelfRef.t__DOT__index_b)), vlcu_msg_arg(vlSelfRef.t__DOT__p126__DOT__got), vlcu_msg_arg((IData)(vlSelfRef.t__DOT__p126__DOT__exp)));
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 141, "");
            }
        }
        __Vdly__t__DOT__p72__DOT__cyc = ((IData)(1U) 
                                         + vlSelfRef.t__DOT__p72__DOT__cyc);
        if (VL_LTS_III(32, 2U, vlSelfRef.t__DOT__p72__DOT__cyc)) {
            vlSelfRef.t__DOT__p72__DOT__exp = ((4U 
                                                >= (IData)(vlSelfRef.t__DOT__index_b))
                                                ? vlSelfRef.t__DOT__p72__DOT__results
                                               [((4U 
                                                  >= (IData)(vlSelfRef.t__DOT__index_a))
                                                  ? (IData)(vlSelfRef.t__DOT__index_a)
                                                  : 0U)]
                                               [vlSelfRef.t__DOT__index_b]
                                                : 0U);
            vlSelfRef.t__DOT__p72__DOT__got = (((((
                                                   VL_GT_W(3, vlSelfRef.t__DOT__p72__DOT__as, vlSelfRef.t__DOT__p72__DOT__bs) 
                                                   << 3U) 
                                                  | (VL_GTS_IWW(72, vlSelfRef.t__DOT__p72__DOT__as, vlSelfRef.t__DOT__p72__DOT__bs) 
                                                     << 2U)) 
                                                 | ((VL_GTE_W(3, vlSelfRef.t__DOT__p72__DOT__as, vlSelfRef.t__DOT__p72__DOT__bs) 
                                                     << 1U) 
                                                    | VL_GTES_IWW(72, vlSelfRef.t__DOT__p72__DOT__as, vlSelfRef.t__DOT__p72__DOT__bs))) 
                                                << 4U) 
                                               | (((VL_LT_W(3, vlSelfRef.t__DOT__p72__DOT__as, vlSelfRef.t__DOT__p72__DOT__bs) 
                                                    << 3U) 
                                                   | (VL_LTS_IWW(72, vlSelfRef.t__DOT__p72__DOT__as, vlSelfRef.t__DOT__p72__DOT__bs) 
                                                      << 2U)) 
                                                  | ((VL_LTE_W(3, vlSelfRef.t__DOT__p72__DOT__as, vlSelfRef.t__DOT__p72__DOT__bs) 
                                                      << 1U) 
                                                     | VL_LTES_IWW(72, vlSelfRef.t__DOT__p72__DOT__as, vlSelfRef.t__DOT__p72__DOT__bs))));
            if (VL_UNLIKELY((((IData)(vlSelfRef.t__DOT__p72__DOT__exp) 
                              != (IData)(vlSelfRef.t__DOT__p72__DOT__got))))) {
                vlCtx.writef(2U, 4U, vlcu_msg_arg(vlSelfRef.t__DOT__index_a), vlcu_msg_arg((IData)(vlSelfRef.t__DOT__index_b)), vlcu_msg_arg(vlSelfRef.t__DOT__p72__DOT__got), vlcu_msg_arg((IData)(vlSelfRef.t__DOT__p72__DOT__exp)));
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 141, "");
            }
        }
        __Vdly__t__DOT__p64__DOT__cyc = ((IData)(1U) 
                                         + vlSelfRef.t__DOT__p64__DOT__cyc);
        if (VL_LTS_III(32, 2U, vlSelfRef.t__DOT__p64__DOT__cyc)) {
            vlSelfRef.t__DOT__p64__DOT__exp = ((4U 
                                                >= (IData)(vlSelfRef.t__DOT__index_b))
                                                ? vlSelfRef.t__DOT__p64__DOT__results
                                               [((4U 
                                                  >= (IData)(vlSelfRef.t__DOT__index_a))
                                                  ? (IData)(vlSelfRef.t__DOT__index_a)
                                                  : 0U)]
                                               [vlSelfRef.t__DOT__index_b]
                                                : 0U);
            vlSelfRef.t__DOT__p64__DOT__got = (((((
                                                   (vlSelfRef.t__DOT__p64__DOT__as 
                                                    > vlSelfRef.t__DOT__p64__DOT__bs) 
                                                   << 3U) 
                                                  | (VL_GTS_IQQ(64, vlSelfRef.t__DOT__p64__DOT__as, vlSelfRef.t__DOT__p64__DOT__bs) 
                                                     << 2U)) 
                                                 | (((vlSelfRef.t__DOT__p64__DOT__as 
                                                      >= vlSelfRef.t__DOT__p64__DOT__bs) 
                                                     << 1U) 
                                                    | VL_GTES_IQQ(64, vlSelfRef.t__DOT__p64__DOT__as, vlSelfRef.t__DOT__p64__DOT__bs))) 
                                                << 4U) 
                                               | ((((vlSelfRef.t__DOT__p64__DOT__as 
                                                     < vlSelfRef.t__DOT__p64__DOT__bs) 
                                                    << 3U) 
                                                   | (VL_LTS_IQQ(64, vlSelfRef.t__DOT__p64__DOT__as, vlSelfRef.t__DOT__p64__DOT__bs) 
                                                      << 2U)) 
                                                  | (((vlSelfRef.t__DOT__p64__DOT__as 
                                                       <= vlSelfRef.t__DOT__p64__DOT__bs) 
                                                      << 1U) 
                                                     | VL_LTES_IQQ(64, vlSelfRef.t__DOT__p64__DOT__as, vlSelfRef.t__DOT__p64__DOT__bs))));
            if (VL_UNLIKELY((((IData)(vlSelfRef.t__DOT__p64__DOT__exp) 
                              != (IData)(vlSelfRef.t__DOT__p64__DOT__got))))) {
                vlCtx.writef(3U, 4U, vlcu_msg_arg(vlSelfRef.t__DOT__index_a), vlcu_msg_arg((IData)(vlSelfRef.t__DOT__index_b)), vlcu_msg_arg(vlSelfRef.t__DOT__p64__DOT__got), vlcu_msg_arg((IData)(vlSelfRef.t__DOT__p64__DOT__exp)));
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 141, "");
            }
        }
        __Vdly__t__DOT__p63__DOT__cyc = ((IData)(1U) 
                                         + vlSelfRef.t__DOT__p63__DOT__cyc);
        if (VL_LTS_III(32, 2U, vlSelfRef.t__DOT__p63__DOT__cyc)) {
            vlSelfRef.t__DOT__p63__DOT__exp = ((4U 
                                                >= (IData)(vlSelfRef.t__DOT__index_b))
                                                ? vlSelfRef.t__DOT__p63__DOT__results
                                               [((4U 
                                                  >= (IData)(vlSelfRef.t__DOT__index_a))
                                                  ? (IData)(vlSelfRef.t__DOT__index_a)
                                                  : 0U)]
                                               [vlSelfRef.t__DOT__index_b]
                                                : 0U);
            vlSelfRef.t__DOT__p63__DOT__got = (((((
                                                   (vlSelfRef.t__DOT__p63__DOT__as 
                                                    > vlSelfRef.t__DOT__p63__DOT__bs) 
                                                   << 3U) 
                                                  | (VL_GTS_IQQ(63, vlSelfRef.t__DOT__p63__DOT__as, vlSelfRef.t__DOT__p63__DOT__bs) 
                                                     << 2U)) 
                                                 | (((vlSelfRef.t__DOT__p63__DOT__as 
                                                      >= vlSelfRef.t__DOT__p63__DOT__bs) 
                                                     << 1U) 
                                                    | VL_GTES_IQQ(63, vlSelfRef.t__DOT__p63__DOT__as, vlSelfRef.t__DOT__p63__DOT__bs))) 
                                                << 4U) 
                                               | ((((vlSelfRef.t__DOT__p63__DOT__as 
                                                     < vlSelfRef.t__DOT__p63__DOT__bs) 
                                                    << 3U) 
                                                   | (VL_LTS_IQQ(63, vlSelfRef.t__DOT__p63__DOT__as, vlSelfRef.t__DOT__p63__DOT__bs) 
                                                      << 2U)) 
                                                  | (((vlSelfRef.t__DOT__p63__DOT__as 
                                                       <= vlSelfRef.t__DOT__p63__DOT__bs) 
                                                      << 1U) 
                                                     | VL_LTES_IQQ(63, vlSelfRef.t__DOT__p63__DOT__as, vlSelfRef.t__DOT__p63__DOT__bs))));
            if (VL_UNLIKELY((((IData)(vlSelfRef.t__DOT__p63__DOT__exp) 
                              != (IData)(vlSelfRef.t__DOT__p63__DOT__got))))) {
                vlCtx.writef(4U, 4U, vlcu_msg_arg(vlSelfRef.t__DOT__index_a), vlcu_msg_arg((IData)(vlSelfRef.t__DOT__index_b)), vlcu_msg_arg(vlSelfRef.t__DOT__p63__DOT__got), vlcu_msg_arg((IData)(vlSelfRef.t__DOT__p63__DOT__exp)));
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 141, "");
            }
        }
        __Vdly__t__DOT__p32__DOT__cyc = ((IData)(1U) 
                                         + vlSelfRef.t__DOT__p32__DOT__cyc);
        if (VL_LTS_III(32, 2U, vlSelfRef.t__DOT__p32__DOT__cyc)) {
            vlSelfRef.t__DOT__p32__DOT__exp = ((4U 
                                                >= (IData)(vlSelfRef.t__DOT__index_b))
                                                ? vlSelfRef.t__DOT__p32__DOT__results
                                               [((4U 
                                                  >= (IData)(vlSelfRef.t__DOT__index_a))
                                                  ? (IData)(vlSelfRef.t__DOT__index_a)
                                                  : 0U)]
                                               [vlSelfRef.t__DOT__index_b]
                                                : 0U);
            vlSelfRef.t__DOT__p32__DOT__got = (((((
                                                   (vlSelfRef.t__DOT__p32__DOT__as 
                                                    > vlSelfRef.t__DOT__p32__DOT__bs) 
                                                   << 3U) 
                                                  | (VL_GTS_III(32, vlSelfRef.t__DOT__p32__DOT__as, vlSelfRef.t__DOT__p32__DOT__bs) 
                                                     << 2U)) 
                                                 | (((vlSelfRef.t__DOT__p32__DOT__as 
                                                      >= vlSelfRef.t__DOT__p32__DOT__bs) 
                                                     << 1U) 
                                                    | VL_GTES_III(32, vlSelfRef.t__DOT__p32__DOT__as, vlSelfRef.t__DOT__p32__DOT__bs))) 
                                                << 4U) 
                                               | ((((vlSelfRef.t__DOT__p32__DOT__as 
                                                     < vlSelfRef.t__DOT__p32__DOT__bs) 
                                                    << 3U) 
                                                   | (VL_LTS_III(32, vlSelfRef.t__DOT__p32__DOT__as, vlSelfRef.t__DOT__p32__DOT__bs) 
                                                      << 2U)) 
                                                  | (((vlSelfRef.t__DOT__p32__DOT__as 
                                                       <= vlSelfRef.t__DOT__p32__DOT__bs) 
                                                      << 1U) 
                                                     | VL_LTES_III(32, vlSelfRef.t__DOT__p32__DOT__as, vlSelfRef.t__DOT__p32__DOT__bs))));
            if (VL_UNLIKELY((((IData)(vlSelfRef.t__DOT__p32__DOT__exp) 
                              != (IData)(vlSelfRef.t__DOT__p32__DOT__got))))) {
                vlCtx.writef(5U, 4U, vlcu_msg_arg(vlSelfRef.t__DOT__index_a), vlcu_msg_arg((IData)(vlSelfRef.t__DOT__index_b)), vlcu_msg_arg(vlSelfRef.t__DOT__p32__DOT__got), vlcu_msg_arg((IData)(vlSelfRef.t__DOT__p32__DOT__exp)));
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 141, "");
            }
        }
        __Vdly__t__DOT__p4__DOT__cyc = ((IData)(1U) 
                                        + vlSelfRef.t__DOT__p4__DOT__cyc);
        if (VL_LTS_III(32, 2U, vlSelfRef.t__DOT__p4__DOT__cyc)) {
            vlSelfRef.t__DOT__p4__DOT__exp = ((4U >= (IData)(vlSelfRef.t__DOT__index_b))
                                               ? vlSelfRef.t__DOT__p4__DOT__results
                                              [((4U 
                                                 >= (IData)(vlSelfRef.t__DOT__index_a))
                                                 ? (IData)(vlSelfRef.t__DOT__index_a)
                                                 : 0U)]
                                              [vlSelfRef.t__DOT__index_b]
                                               : 0U);
            vlSelfRef.t__DOT__p4__DOT__got = (((((((IData)(vlSelfRef.t__DOT__p4__DOT__as) 
                                                   > (IData)(vlSelfRef.t__DOT__p4__DOT__bs)) 
                                                  << 3U) 
                                                 | (VL_GTS_III(4, (IData)(vlSelfRef.t__DOT__p4__DOT__as), (IData)(vlSelfRef.t__DOT__p4__DOT__bs)) 
                                                    << 2U)) 
                                                | ((((IData)(vlSelfRef.t__DOT__p4__DOT__as) 
                                                     >= (IData)(vlSelfRef.t__DOT__p4__DOT__bs)) 
                                                    << 1U) 
                                                   | VL_GTES_III(4, (IData)(vlSelfRef.t__DOT__p4__DOT__as), (IData)(vlSelfRef.t__DOT__p4__DOT__bs)))) 
                                               << 4U) 
                                              | (((((IData)(vlSelfRef.t__DOT__p4__DOT__as) 
                                                    < (IData)(vlSelfRef.t__DOT__p4__DOT__bs)) 
                                                   << 3U) 
                                                  | (VL_LTS_III(4, (IData)(vlSelfRef.t__DOT__p4__DOT__as), (IData)(vlSelfRef.t__DOT__p4__DOT__bs)) 
                                                     << 2U)) 
                                                 | ((((IData)(vlSelfRef.t__DOT__p4__DOT__as) 
                                                      <= (IData)(vlSelfRef.t__DOT__p4__DOT__bs)) 
                                                     << 1U) 
                                                    | VL_LTES_III(4, (IData)(vlSelfRef.t__DOT__p4__DOT__as), (IData)(vlSelfRef.t__DOT__p4__DOT__bs)))));
            if (VL_UNLIKELY((((IData)(vlSelfRef.t__DOT__p4__DOT__exp) 
                              != (IData)(vlSelfRef.t__DOT__p4__DOT__got))))) {
                vlCtx.writef(6U, 4U, vlcu_msg_arg(vlSelfRef.t__DOT__index_a), vlcu_msg_arg((IData)(vlSelfRef.t__DOT__index_b)), vlcu_msg_arg(vlSelfRef.t__DOT__p4__DOT__got), vlcu_msg_arg((IData)(vlSelfRef.t__DOT__p4__DOT__exp)));
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 141, "");
            }
        }
        __Vdly__t__DOT__cyc = ((IData)(1U) + vlSelfRef.t__DOT__cyc);
        if (VL_UNLIKELY(((0x00000063U == vlSelfRef.t__DOT__cyc)))) {
            vlCtx.writef(7U, 0U);
            VL_FINISH_MT("/root/reference/test_regress/t/t_math_cmp.v", 73, "");
        }
        vlSelfRef.t__DOT__p128__DOT__cyc = __Vdly__t__DOT__p128__DOT__cyc;
        vlSelfRef.t__DOT__p126__DOT__cyc = __Vdly__t__DOT__p126__DOT__cyc;
        vlSelfRef.t__DOT__p72__DOT__cyc = __Vdly__t__DOT__p72__DOT__cyc;
        vlSelfRef.t__DOT__p64__DOT__cyc = __Vdly__t__DOT__p64__DOT__cyc;
        vlSelfRef.t__DOT__p63__DOT__cyc = __Vdly__t__DOT__p63__DOT__cyc;
        vlSelfRef.t__DOT__p32__DOT__cyc = __Vdly__t__DOT__p32__DOT__cyc;
        vlSelfRef.t__DOT__p4__DOT__cyc = __Vdly__t__DOT__p4__DOT__cyc;
        vlSelfRef.t__DOT__cyc = __Vdly__t__DOT__cyc;
        vlSelfRef.t__DOT__index_a = (7U & vlSelfRef.t__DOT__cyc);
        if ((4U < (IData)(vlSelfRef.t__DOT__index_a))) {
            vlSelfRef.t__DOT__index_a = 4U;
        }
        vlSelfRef.t__DOT__index_b = (7U & (vlSelfRef.t__DOT__cyc 
                                           >> 3U));
        if ((4U < (IData)(vlSelfRef.t__DOT__index_b))) {
            vlSelfRef.t__DOT__index_b = 4U;
        }
        if ((4U & (IData)(vlSelfRef.t__DOT__index_a))) {
            if (VL_UNLIKELY(((2U & (IData)(vlSelfRef.t__DOT__index_a))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 101, "");
            } else if (VL_UNLIKELY(((1U & (IData)(vlSelfRef.t__DOT__index_a))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 101, "");
            } else {
                vlSelfRef.t__DOT__p128__DOT__as[0U] = 0U;
                vlSelfRef.t__DOT__p128__DOT__as[1U] = 0U;
                vlSelfRef.t__DOT__p128__DOT__as[2U] = 0U;
                vlSelfRef.t__DOT__p128__DOT__as[3U] = 0x80000000U;
            }
        } else if ((2U & (IData)(vlSelfRef.t__DOT__index_a))) {
            if ((1U & (IData)(vlSelfRef.t__DOT__index_a))) {
                vlSelfRef.t__DOT__p128__DOT__as[0U] = 0xffffffffU;
                vlSelfRef.t__DOT__p128__DOT__as[1U] = 0xffffffffU;
                vlSelfRef.t__DOT__p128__DOT__as[2U] = 0xffffffffU;
                vlSelfRef.t__DOT__p128__DOT__as[3U] = 0xffffffffU;
            } else {
                vlSelfRef.t__DOT__p128__DOT__as[0U] = 0U;
                vlSelfRef.t__DOT__p128__DOT__as[1U] = 0U;
                vlSelfRef.t__DOT__p128__DOT__as[2U] = 0U;
                vlSelfRef.t__DOT__p128__DOT__as[3U] = 0U;
            }
        } else {
            vlSelfRef.t__DOT__p128__DOT__as[0U] = (
                                                   (1U 
                                                    & (IData)(vlSelfRef.t__DOT__index_a))
                                                    ? 1U
                                                    : 0U);
            vlSelfRef.t__DOT__p128__DOT__as[1U] = 0U;
            vlSelfRef.t__DOT__p128__DOT__as[2U] = 0U;
            vlSelfRef.t__DOT__p128__DOT__as[3U] = 0U;
        }
        if ((4U & (IData)(vlSelfRef.t__DOT__index_b))) {
            if (VL_UNLIKELY(((2U & (IData)(vlSelfRef.t__DOT__index_b))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 109, "");
            } else if (VL_UNLIKELY(((1U & (IData)(vlSelfRef.t__DOT__index_b))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 109, "");
            } else {
                vlSelfRef.t__DOT__p128__DOT__bs[0U] = 0U;
                vlSelfRef.t__DOT__p128__DOT__bs[1U] = 0U;
                vlSelfRef.t__DOT__p128__DOT__bs[2U] = 0U;
                vlSelfRef.t__DOT__p128__DOT__bs[3U] = 0x80000000U;
            }
        } else if ((2U & (IData)(vlSelfRef.t__DOT__index_b))) {
            if ((1U & (IData)(vlSelfRef.t__DOT__index_b))) {
                vlSelfRef.t__DOT__p128__DOT__bs[0U] = 0xffffffffU;
                vlSelfRef.t__DOT__p128__DOT__bs[1U] = 0xffffffffU;
                vlSelfRef.t__DOT__p128__DOT__bs[2U] = 0xffffffffU;
                vlSelfRef.t__DOT__p128__DOT__bs[3U] = 0xffffffffU;
            } else {
                vlSelfRef.t__DOT__p128__DOT__bs[0U] = 0U;
                vlSelfRef.t__DOT__p128__DOT__bs[1U] = 0U;
                vlSelfRef.t__DOT__p128__DOT__bs[2U] = 0U;
                vlSelfRef.t__DOT__p128__DOT__bs[3U] = 0U;
            }
        } else {
            vlSelfRef.t__DOT__p128__DOT__bs[0U] = (
                                                   (1U 
                                                    & (IData)(vlSelfRef.t__DOT__index_b))
                                                    ? 1U
                                                    : 0U);
            vlSelfRef.t__DOT__p128__DOT__bs[1U] = 0U;
            vlSelfRef.t__DOT__p128__DOT__bs[2U] = 0U;
            vlSelfRef.t__DOT__p128__DOT__bs[3U] = 0U;
        }
        if ((4U & (IData)(vlSelfRef.t__DOT__index_a))) {
            if (VL_UNLIKELY(((2U & (IData)(vlSelfRef.t__DOT__index_a))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 101, "");
            } else if (VL_UNLIKELY(((1U & (IData)(vlSelfRef.t__DOT__index_a))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 101, "");
            } else {
                vlSelfRef.t__DOT__p126__DOT__as[0U] = 0U;
                vlSelfRef.t__DOT__p126__DOT__as[1U] = 0U;
                vlSelfRef.t__DOT__p126__DOT__as[2U] = 0U;
                vlSelfRef.t__DOT__p126__DOT__as[3U] = 0x20000000U;
            }
        } else if ((2U & (IData)(vlSelfRef.t__DOT__index_a))) {
            if ((1U & (IData)(vlSelfRef.t__DOT__index_a))) {
                vlSelfRef.t__DOT__p126__DOT__as[0U] = 0xffffffffU;
                vlSelfRef.t__DOT__p126__DOT__as[1U] = 0xffffffffU;
                vlSelfRef.t__DOT__p126__DOT__as[2U] = 0xffffffffU;
                vlSelfRef.t__DOT__p126__DOT__as[3U] = 0x3fffffffU;
            } else {
                vlSelfRef.t__DOT__p126__DOT__as[0U] = 0U;
                vlSelfRef.t__DOT__p126__DOT__as[1U] = 0U;
                vlSelfRef.t__DOT__p126__DOT__as[2U] = 0U;
                vlSelfRef.t__DOT__p126__DOT__as[3U] = 0U;
            }
        } else {
            vlSelfRef.t__DOT__p126__DOT__as[0U] = (
                                                   (1U 
                                                    & (IData)(vlSelfRef.t__DOT__index_a))
                                                    ? 1U
                                                    : 0U);
            vlSelfRef.t__DOT__p126__DOT__as[1U] = 0U;
            vlSelfRef.t__DOT__p126__DOT__as[2U] = 0U;
            vlSelfRef.t__DOT__p126__DOT__as[3U] = 0U;
        }
        if ((4U & (IData)(vlSelfRef.t__DOT__index_b))) {
            if (VL_UNLIKELY(((2U & (IData)(vlSelfRef.t__DOT__index_b))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 109, "");
            } else if (VL_UNLIKELY(((1U & (IData)(vlSelfRef.t__DOT__index_b))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 109, "");
            } else {
                vlSelfRef.t__DOT__p126__DOT__bs[0U] = 0U;
                vlSelfRef.t__DOT__p126__DOT__bs[1U] = 0U;
                vlSelfRef.t__DOT__p126__DOT__bs[2U] = 0U;
                vlSelfRef.t__DOT__p126__DOT__bs[3U] = 0x20000000U;
            }
        } else if ((2U & (IData)(vlSelfRef.t__DOT__index_b))) {
            if ((1U & (IData)(vlSelfRef.t__DOT__index_b))) {
                vlSelfRef.t__DOT__p126__DOT__bs[0U] = 0xffffffffU;
                vlSelfRef.t__DOT__p126__DOT__bs[1U] = 0xffffffffU;
                vlSelfRef.t__DOT__p126__DOT__bs[2U] = 0xffffffffU;
                vlSelfRef.t__DOT__p126__DOT__bs[3U] = 0x3fffffffU;
            } else {
                vlSelfRef.t__DOT__p126__DOT__bs[0U] = 0U;
                vlSelfRef.t__DOT__p126__DOT__bs[1U] = 0U;
                vlSelfRef.t__DOT__p126__DOT__bs[2U] = 0U;
                vlSelfRef.t__DOT__p126__DOT__bs[3U] = 0U;
            }
        } else {
            vlSelfRef.t__DOT__p126__DOT__bs[0U] = (
                                                   (1U 
                                                    & (IData)(vlSelfRef.t__DOT__index_b))
                                                    ? 1U
                                                    : 0U);
            vlSelfRef.t__DOT__p126__DOT__bs[1U] = 0U;
            vlSelfRef.t__DOT__p126__DOT__bs[2U] = 0U;
            vlSelfRef.t__DOT__p126__DOT__bs[3U] = 0U;
        }
        if ((4U & (IData)(vlSelfRef.t__DOT__index_a))) {
            if (VL_UNLIKELY(((2U & (IData)(vlSelfRef.t__DOT__index_a))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 101, "");
            } else if (VL_UNLIKELY(((1U & (IData)(vlSelfRef.t__DOT__index_a))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 101, "");
            } else {
                vlSelfRef.t__DOT__p72__DOT__as[0U] = 0U;
                vlSelfRef.t__DOT__p72__DOT__as[1U] = 0U;
                vlSelfRef.t__DOT__p72__DOT__as[2U] = 0x00000080U;
            }
        } else if ((2U & (IData)(vlSelfRef.t__DOT__index_a))) {
            if ((1U & (IData)(vlSelfRef.t__DOT__index_a))) {
                vlSelfRef.t__DOT__p72__DOT__as[0U] = 0xffffffffU;
                vlSelfRef.t__DOT__p72__DOT__as[1U] = 0xffffffffU;
                vlSelfRef.t__DOT__p72__DOT__as[2U] = 0x000000ffU;
            } else {
                vlSelfRef.t__DOT__p72__DOT__as[0U] = 0U;
                vlSelfRef.t__DOT__p72__DOT__as[1U] = 0U;
                vlSelfRef.t__DOT__p72__DOT__as[2U] = 0U;
            }
        } else {
            vlSelfRef.t__DOT__p72__DOT__as[0U] = ((1U 
                                                   & (IData)(vlSelfRef.t__DOT__index_a))
                                                   ? 1U
                                                   : 0U);
            vlSelfRef.t__DOT__p72__DOT__as[1U] = 0U;
            vlSelfRef.t__DOT__p72__DOT__as[2U] = 0U;
        }
        if ((4U & (IData)(vlSelfRef.t__DOT__index_b))) {
            if (VL_UNLIKELY(((2U & (IData)(vlSelfRef.t__DOT__index_b))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 109, "");
            } else if (VL_UNLIKELY(((1U & (IData)(vlSelfRef.t__DOT__index_b))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 109, "");
            } else {
                vlSelfRef.t__DOT__p72__DOT__bs[0U] = 0U;
                vlSelfRef.t__DOT__p72__DOT__bs[1U] = 0U;
                vlSelfRef.t__DOT__p72__DOT__bs[2U] = 0x00000080U;
            }
        } else if ((2U & (IData)(vlSelfRef.t__DOT__index_b))) {
            if ((1U & (IData)(vlSelfRef.t__DOT__index_b))) {
                vlSelfRef.t__DOT__p72__DOT__bs[0U] = 0xffffffffU;
                vlSelfRef.t__DOT__p72__DOT__bs[1U] = 0xffffffffU;
                vlSelfRef.t__DOT__p72__DOT__bs[2U] = 0x000000ffU;
            } else {
                vlSelfRef.t__DOT__p72__DOT__bs[0U] = 0U;
                vlSelfRef.t__DOT__p72__DOT__bs[1U] = 0U;
                vlSelfRef.t__DOT__p72__DOT__bs[2U] = 0U;
            }
        } else {
            vlSelfRef.t__DOT__p72__DOT__bs[0U] = ((1U 
                                                   & (IData)(vlSelfRef.t__DOT__index_b))
                                                   ? 1U
                                                   : 0U);
            vlSelfRef.t__DOT__p72__DOT__bs[1U] = 0U;
            vlSelfRef.t__DOT__p72__DOT__bs[2U] = 0U;
        }
        if ((4U & (IData)(vlSelfRef.t__DOT__index_a))) {
            if (VL_UNLIKELY(((2U & (IData)(vlSelfRef.t__DOT__index_a))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 101, "");
            } else if (VL_UNLIKELY(((1U & (IData)(vlSelfRef.t__DOT__index_a))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 101, "");
            } else {
                vlSelfRef.t__DOT__p64__DOT__as = 0x8000000000000000ULL;
            }
        } else {
            vlSelfRef.t__DOT__p64__DOT__as = ((2U & (IData)(vlSelfRef.t__DOT__index_a))
                                               ? ((1U 
                                                   & (IData)(vlSelfRef.t__DOT__index_a))
                                                   ? 0xffffffffffffffffULL
                                                   : 0ULL)
                                               : ((1U 
                                                   & (IData)(vlSelfRef.t__DOT__index_a))
                                                   ? 1ULL
                                                   : 0ULL));
        }
        if ((4U & (IData)(vlSelfRef.t__DOT__index_b))) {
            if (VL_UNLIKELY(((2U & (IData)(vlSelfRef.t__DOT__index_b))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 109, "");
            } else if (VL_UNLIKELY(((1U & (IData)(vlSelfRef.t__DOT__index_b))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 109, "");
            } else {
                vlSelfRef.t__DOT__p64__DOT__bs = 0x8000000000000000ULL;
            }
        } else {
            vlSelfRef.t__DOT__p64__DOT__bs = ((2U & (IData)(vlSelfRef.t__DOT__index_b))
                                               ? ((1U 
                                                   & (IData)(vlSelfRef.t__DOT__index_b))
                                                   ? 0xffffffffffffffffULL
                                                   : 0ULL)
                                               : ((1U 
                                                   & (IData)(vlSelfRef.t__DOT__index_b))
                                                   ? 1ULL
                                                   : 0ULL));
        }
        if ((4U & (IData)(vlSelfRef.t__DOT__index_a))) {
            if (VL_UNLIKELY(((2U & (IData)(vlSelfRef.t__DOT__index_a))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 101, "");
            } else if (VL_UNLIKELY(((1U & (IData)(vlSelfRef.t__DOT__index_a))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 101, "");
            } else {
                vlSelfRef.t__DOT__p63__DOT__as = 0x4000000000000000ULL;
            }
        } else {
            vlSelfRef.t__DOT__p63__DOT__as = ((2U & (IData)(vlSelfRef.t__DOT__index_a))
                                               ? ((1U 
                                                   & (IData)(vlSelfRef.t__DOT__index_a))
                                                   ? 0x7fffffffffffffffULL
                                                   : 0ULL)
                                               : ((1U 
                                                   & (IData)(vlSelfRef.t__DOT__index_a))
                                                   ? 1ULL
                                                   : 0ULL));
        }
        if ((4U & (IData)(vlSelfRef.t__DOT__index_b))) {
            if (VL_UNLIKELY(((2U & (IData)(vlSelfRef.t__DOT__index_b))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 109, "");
            } else if (VL_UNLIKELY(((1U & (IData)(vlSelfRef.t__DOT__index_b))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 109, "");
            } else {
                vlSelfRef.t__DOT__p63__DOT__bs = 0x4000000000000000ULL;
            }
        } else {
            vlSelfRef.t__DOT__p63__DOT__bs = ((2U & (IData)(vlSelfRef.t__DOT__index_b))
                                               ? ((1U 
                                                   & (IData)(vlSelfRef.t__DOT__index_b))
                                                   ? 0x7fffffffffffffffULL
                                                   : 0ULL)
                                               : ((1U 
                                                   & (IData)(vlSelfRef.t__DOT__index_b))
                                                   ? 1ULL
                                                   : 0ULL));
        }
        if ((4U & (IData)(vlSelfRef.t__DOT__index_a))) {
            if (VL_UNLIKELY(((2U & (IData)(vlSelfRef.t__DOT__index_a))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 101, "");
            } else if (VL_UNLIKELY(((1U & (IData)(vlSelfRef.t__DOT__index_a))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 101, "");
            } else {
                vlSelfRef.t__DOT__p32__DOT__as = 0x80000000U;
            }
        } else {
            vlSelfRef.t__DOT__p32__DOT__as = ((2U & (IData)(vlSelfRef.t__DOT__index_a))
                                               ? ((1U 
                                                   & (IData)(vlSelfRef.t__DOT__index_a))
                                                   ? 0xffffffffU
                                                   : 0U)
                                               : ((1U 
                                                   & (IData)(vlSelfRef.t__DOT__index_a))
                                                   ? 1U
                                                   : 0U));
        }
        if ((4U & (IData)(vlSelfRef.t__DOT__index_b))) {
            if (VL_UNLIKELY(((2U & (IData)(vlSelfRef.t__DOT__index_b))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 109, "");
            } else if (VL_UNLIKELY(((1U & (IData)(vlSelfRef.t__DOT__index_b))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 109, "");
            } else {
                vlSelfRef.t__DOT__p32__DOT__bs = 0x80000000U;
            }
        } else {
            vlSelfRef.t__DOT__p32__DOT__bs = ((2U & (IData)(vlSelfRef.t__DOT__index_b))
                                               ? ((1U 
                                                   & (IData)(vlSelfRef.t__DOT__index_b))
                                                   ? 0xffffffffU
                                                   : 0U)
                                               : ((1U 
                                                   & (IData)(vlSelfRef.t__DOT__index_b))
                                                   ? 1U
                                                   : 0U));
        }
        if ((4U & (IData)(vlSelfRef.t__DOT__index_a))) {
            if (VL_UNLIKELY(((2U & (IData)(vlSelfRef.t__DOT__index_a))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 101, "");
            } else if (VL_UNLIKELY(((1U & (IData)(vlSelfRef.t__DOT__index_a))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 101, "");
            } else {
                vlSelfRef.t__DOT__p4__DOT__as = 8U;
            }
        } else {
            vlSelfRef.t__DOT__p4__DOT__as = ((2U & (IData)(vlSelfRef.t__DOT__index_a))
                                              ? ((1U 
                                                  & (IData)(vlSelfRef.t__DOT__index_a))
                                                  ? 0x0fU
                                                  : 0U)
                                              : ((1U 
                                                  & (IData)(vlSelfRef.t__DOT__index_a))
                                                  ? 1U
                                                  : 0U));
        }
        if ((4U & (IData)(vlSelfRef.t__DOT__index_b))) {
            if (VL_UNLIKELY(((2U & (IData)(vlSelfRef.t__DOT__index_b))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 109, "");
            } else if (VL_UNLIKELY(((1U & (IData)(vlSelfRef.t__DOT__index_b))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_math_cmp.v", 109, "");
            } else {
                vlSelfRef.t__DOT__p4__DOT__bs = 8U;
            }
        } else {
            vlSelfRef.t__DOT__p4__DOT__bs = ((2U & (IData)(vlSelfRef.t__DOT__index_b))
                                              ? ((1U 
                                                  & (IData)(vlSelfRef.t__DOT__index_b))
                                                  ? 0x0fU
                                                  : 0U)
                                              : ((1U 
                                                  & (IData)(vlSelfRef.t__DOT__index_b))
                                                  ? 1U
                                                  : 0U));
        }
    }

    VLCU_D static void Vt_math_cmp___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt_math_cmp___024root___eval_phase__act(Vt_math_cmp___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt_math_cmp___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vt_math_cmp___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vt_math_cmp___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vt_math_cmp___024root___eval_phase__nba(Vt_math_cmp___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vt_math_cmp___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    Vt_math_cmp___024root___nba_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
            Vt_math_cmp___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vt_math_cmp___024root___eval(Vt_math_cmp___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt_math_cmp___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("/root/reference/test_regress/t/t_math_cmp.v", 7, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vt_math_cmp___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("/root/reference/test_regress/t/t_math_cmp.v", 7, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vt_math_cmp___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vt_math_cmp___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vt_math_cmp___024root___ctor_var_reset(Vt_math_cmp___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        vlSelf->t__DOT__index_a = VL_SCOPED_RAND_RESET_I(3, __VscopeHash, 2407613259116646721ull);
        vlSelf->t__DOT__index_b = VL_SCOPED_RAND_RESET_I(3, __VscopeHash, 7073485808186199881ull);
        vlSelf->t__DOT__cyc = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 15508623621340971111ull);
        VL_SCOPED_RAND_RESET_W(128, vlSelf->t__DOT__p128__DOT__as, __VscopeHash, 12559802784005411829ull);
        VL_SCOPED_RAND_RESET_W(128, vlSelf->t__DOT__p128__DOT__bs, __VscopeHash, 12711901933494776699ull);
        for (int __Vi0 = 0; __Vi0 < 5; ++__Vi0) {
            for (int __Vi1 = 0; __Vi1 < 5; ++__Vi1) {
                vlSelf->t__DOT__p128__DOT__results[__Vi0][__Vi1] = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 11416079230318097719ull);
            }
        }
        vlSelf->t__DOT__p128__DOT__exp = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 15433416984162184792ull);
        vlSelf->t__DOT__p128__DOT__got = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 6605795900602597642ull);
        VL_SCOPED_RAND_RESET_W(126, vlSelf->t__DOT__p126__DOT__as, __VscopeHash, 12031359733370202625ull);
        VL_SCOPED_RAND_RESET_W(126, vlSelf->t__DOT__p126__DOT__bs, __VscopeHash, 2971452536138407149ull);
        for (int __Vi0 = 0; __Vi0 < 5; ++__Vi0) {
            for (int __Vi1 = 0; __Vi1 < 5; ++__Vi1) {
                vlSelf->t__DOT__p126__DOT__results[__Vi0][__Vi1] = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 16367855628312935357ull);
            }
        }
        vlSelf->t__DOT__p126__DOT__exp = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 3337957579618654617ull);
        vlSelf->t__DOT__p126__DOT__got = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 1588037863241722290ull);
        VL_SCOPED_RAND_RESET_W(72, vlSelf->t__DOT__p72__DOT__as, __VscopeHash, 2678539708784888823ull);
        VL_SCOPED_RAND_RESET_W(72, vlSelf->t__DOT__p72__DOT__bs, __VscopeHash, 32202185481449692ull);
        for (int __Vi0 = 0; __Vi0 < 5; ++__Vi0) {
            for (int __Vi1 = 0; __Vi1 < 5; ++__Vi1) {
                vlSelf->t__DOT__p72__DOT__results[__Vi0][__Vi1] = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 9195434246681870443ull);
            }
        }
        vlSelf->t__DOT__p72__DOT__exp = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 10861102423663155349ull);
        vlSelf->t__DOT__p72__DOT__got = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 13653854856463253212ull);
        vlSelf->t__DOT__p64__DOT__as = VL_SCOPED_RAND_RESET_Q(64, __VscopeHash, 16429950364848318021ull);
        vlSelf->t__DOT__p64__DOT__bs = VL_SCOPED_RAND_RESET_Q(64, __VscopeHash, 5273408322168909733ull);
        for (int __Vi0 = 0; __Vi0 < 5; ++__Vi0) {
            for (int __Vi1 = 0; __Vi1 < 5; ++__Vi1) {
                vlSelf->t__DOT__p64__DOT__results[__Vi0][__Vi1] = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 14334878451025236705ull);
            }
        }
        vlSelf->t__DOT__p64__DOT__exp = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 17398844373069228584ull);
        vlSelf->t__DOT__p64__DOT__got = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 4581546083434798168ull);
        vlSelf->t__DOT__p63__DOT__as = VL_SCOPED_RAND_RESET_Q(63, __VscopeHash, 12971991024947615782ull);
        vlSelf->t__DOT__p63__DOT__bs = VL_SCOPED_RAND_RESET_Q(63, __VscopeHash, 3193064635847948993ull);
        for (int __Vi0 = 0; __Vi0 < 5; ++__Vi0) {
            for (int __Vi1 = 0; __Vi1 < 5; ++__Vi1) {
                vlSelf->t__DOT__p63__DOT__results[__Vi0][__Vi1] = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 12593047627579649419ull);
            }
        }
        vlSelf->t__DOT__p63__DOT__exp = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 11361925576897692608ull);
        vlSelf->t__DOT__p63__DOT__got = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 17943093305537313229ull);
        vlSelf->t__DOT__p32__DOT__as = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 3530640014806543574ull);
        vlSelf->t__DOT__p32__DOT__bs = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 7012730046392079685ull);
        for (int __Vi0 = 0; __Vi0 < 5; ++__Vi0) {
            for (int __Vi1 = 0; __Vi1 < 5; ++__Vi1) {
                vlSelf->t__DOT__p32__DOT__results[__Vi0][__Vi1] = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 15908650382972479457ull);
            }
        }
        vlSelf->t__DOT__p32__DOT__exp = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 2138853865905342227ull);
        vlSelf->t__DOT__p32__DOT__got = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 4449562615181693406ull);
        vlSelf->t__DOT__p4__DOT__as = VL_SCOPED_RAND_RESET_I(4, __VscopeHash, 1810306453155136599ull);
        vlSelf->t__DOT__p4__DOT__bs = VL_SCOPED_RAND_RESET_I(4, __VscopeHash, 3661765191834074140ull);
        for (int __Vi0 = 0; __Vi0 < 5; ++__Vi0) {
            for (int __Vi1 = 0; __Vi1 < 5; ++__Vi1) {
                vlSelf->t__DOT__p4__DOT__results[__Vi0][__Vi1] = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 13302715575302566511ull);
            }
        }
        vlSelf->t__DOT__p4__DOT__exp = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 11274953760226789757ull);
        vlSelf->t__DOT__p4__DOT__got = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 12320067829263375001ull);
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
        Vt_math_cmp___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_cmp___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_cmp___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_cmp___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_cmp___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_cmp___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt_math_cmp__cuda::info() {
    typedef Vt_math_cmp__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"t__DOT__index_a", "t.index_a", VLCU_MEMBER_STATE, 1, 0, 1, 3, 2, 0, T::OFF_t__DOT__index_a},
        {"t__DOT__index_b", "t.index_b", VLCU_MEMBER_STATE, 1, 0, 1, 3, 2, 0, T::OFF_t__DOT__index_b},
        {"t__DOT__p128__DOT__exp", "t.p128.exp", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF_t__DOT__p128__DOT__exp},
        {"t__DOT__p128__DOT__got", "t.p128.got", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF_t__DOT__p128__DOT__got},
        {"t__DOT__p126__DOT__exp", "t.p126.exp", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF_t__DOT__p126__DOT__exp},
        {"t__DOT__p126__DOT__got", "t.p126.got", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF_t__DOT__p126__DOT__got},
        {"t__DOT__p72__DOT__exp", "t.p72.exp", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF_t__DOT__p72__DOT__exp},
        {"t__DOT__p72__DOT__got", "t.p72.got", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF_t__DOT__p72__DOT__got},
        {"t__DOT__p64__DOT__exp", "t.p64.exp", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF_t__DOT__p64__DOT__exp},
        {"t__DOT__p64__DOT__got", "t.p64.got", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF_t__DOT__p64__DOT__got},
        {"t__DOT__p63__DOT__exp", "t.p63.exp", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF_t__DOT__p63__DOT__exp},
        {"t__DOT__p63__DOT__got", "t.p63.got", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF_t__DOT__p63__DOT__got},
        {"t__DOT__p32__DOT__exp", "t.p32.exp", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF_t__DOT__p32__DOT__exp},
        {"t__DOT__p32__DOT__got", "t.p32.got", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF_t__DOT__p32__DOT__got},
        {"t__DOT__p4__DOT__as", "t.p4.as", VLCU_MEMBER_STATE, 1, 0, 1, 4, 3, 0, T::OFF_t__DOT__p4__DOT__as},
        {"t__DOT__p4__DOT__bs", "t.p4.bs", VLCU_MEMBER_STATE, 1, 0, 1, 4, 3, 0, T::OFF_t__DOT__p4__DOT__bs},
        {"t__DOT__p4__DOT__exp", "t.p4.exp", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF_t__DOT__p4__DOT__exp},
        {"t__DOT__p4__DOT__got", "t.p4.got", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF_t__DOT__p4__DOT__got},
        {"__VstlFirstIteration", "__VstlFirstIteration", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlFirstIteration},
        {"__VstlPhaseResult", "__VstlPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"t__DOT__cyc", "t.cyc", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__cyc},
        {"t__DOT__p128__DOT__as", "t.p128.as", VLCU_MEMBER_STATE, 4, 0, 4, 128, 127, 0, T::OFF_t__DOT__p128__DOT__as},
        {"t__DOT__p128__DOT__bs", "t.p128.bs", VLCU_MEMBER_STATE, 4, 0, 4, 128, 127, 0, T::OFF_t__DOT__p128__DOT__bs},
        {"t__DOT__p128__DOT__cyc", "t.p128.cyc", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__p128__DOT__cyc},
        {"t__DOT__p126__DOT__as", "t.p126.as", VLCU_MEMBER_STATE, 4, 0, 4, 126, 125, 0, T::OFF_t__DOT__p126__DOT__as},
        {"t__DOT__p126__DOT__bs", "t.p126.bs", VLCU_MEMBER_STATE, 4, 0, 4, 126, 125, 0, T::OFF_t__DOT__p126__DOT__bs},
        {"t__DOT__p126__DOT__cyc", "t.p126.cyc", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__p126__DOT__cyc},
        {"t__DOT__p72__DOT__as", "t.p72.as", VLCU_MEMBER_STATE, 4, 0, 3, 72, 71, 0, T::OFF_t__DOT__p72__DOT__as},
        {"t__DOT__p72__DOT__bs", "t.p72.bs", VLCU_MEMBER_STATE, 4, 0, 3, 72, 71, 0, T::OFF_t__DOT__p72__DOT__bs},
        {"t__DOT__p72__DOT__cyc", "t.p72.cyc", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__p72__DOT__cyc},
        {"t__DOT__p64__DOT__cyc", "t.p64.cyc", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__p64__DOT__cyc},
        {"t__DOT__p63__DOT__cyc", "t.p63.cyc", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__p63__DOT__cyc},
        {"t__DOT__p32__DOT__as", "t.p32.as", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__p32__DOT__as},
        {"t__DOT__p32__DOT__bs", "t.p32.bs", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__p32__DOT__bs},
        {"t__DOT__p32__DOT__cyc", "t.p32.cyc", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__p32__DOT__cyc},
        {"t__DOT__p4__DOT__cyc", "t.p4.cyc", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__p4__DOT__cyc},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"t__DOT__p64__DOT__as", "t.p64.as", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF_t__DOT__p64__DOT__as},
        {"t__DOT__p64__DOT__bs", "t.p64.bs", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF_t__DOT__p64__DOT__bs},
        {"t__DOT__p63__DOT__as", "t.p63.as", VLCU_MEMBER_STATE, 8, 0, 1, 63, 62, 0, T::OFF_t__DOT__p63__DOT__as},
        {"t__DOT__p63__DOT__bs", "t.p63.bs", VLCU_MEMBER_STATE, 8, 0, 1, 63, 62, 0, T::OFF_t__DOT__p63__DOT__bs},
        {"t__DOT__p128__DOT__results", "t.p128.results", VLCU_MEMBER_STATE, 1, 0, 25, 8, 7, 0, T::OFF_t__DOT__p128__DOT__results},
        {"t__DOT__p126__DOT__results", "t.p126.results", VLCU_MEMBER_STATE, 1, 0, 25, 8, 7, 0, T::OFF_t__DOT__p126__DOT__results},
        {"t__DOT__p72__DOT__results", "t.p72.results", VLCU_MEMBER_STATE, 1, 0, 25, 8, 7, 0, T::OFF_t__DOT__p72__DOT__results},
        {"t__DOT__p64__DOT__results", "t.p64.results", VLCU_MEMBER_STATE, 1, 0, 25, 8, 7, 0, T::OFF_t__DOT__p64__DOT__results},
        {"t__DOT__p63__DOT__results", "t.p63.results", VLCU_MEMBER_STATE, 1, 0, 25, 8, 7, 0, T::OFF_t__DOT__p63__DOT__results},
        {"t__DOT__p32__DOT__results", "t.p32.results", VLCU_MEMBER_STATE, 1, 0, 25, 8, 7, 0, T::OFF_t__DOT__p32__DOT__results},
        {"t__DOT__p4__DOT__results", "t.p4.results", VLCU_MEMBER_STATE, 1, 0, 25, 8, 7, 0, T::OFF_t__DOT__p4__DOT__results},
        {"__VstlTriggered", "__VstlTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VstlTriggered},
        {"__VactTriggered", "__VactTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VactTriggered},
        {"__VnbaTriggered", "__VnbaTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VnbaTriggered},
    };
    static const char* const messages[] = {"%%Error: bad comparison width=128: %d/%d got=%b exp=%b\n", "%%Error: bad comparison width=126: %d/%d got=%b exp=%b\n", "%%Error: bad comparison width=72: %d/%d got=%b exp=%b\n", "%%Error: bad comparison width=64: %d/%d got=%b exp=%b\n", "%%Error: bad comparison width=63: %d/%d got=%b exp=%b\n", "%%Error: bad comparison width=32: %d/%d got=%b exp=%b\n", "%%Error: bad comparison width=4: %d/%d got=%b exp=%b\n", "*-* All Finished *-*\n", nullptr};
    static const char* const plusargs[] = {nullptr};
    static const VlcuTraceDecl traceDecls[] = {
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0}};  // (terminator, not counted)
    static const uint32_t traceOrder[] = {0};  // codes in dump order (+ terminator)
    static const uint32_t saveOrder[] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 36, 37, 38, 39, 40, 41, 42, 43, 44, 45, 46, 47, 48, 49, 50, 51, 52, 53, 54, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "t", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 8, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 55, 0, 0x4abfb8a786cb87a9ULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt_math_cmp__cuda VLCU_TESTBENCHES);
