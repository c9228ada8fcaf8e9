// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt_opt_const_red"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt_opt_const_red___024root final {
    CData/*0:0*/ clk;
    CData/*0:0*/ t__DOT__a1;
    CData/*0:0*/ t__DOT__a10;
    CData/*0:0*/ t__DOT__a11;
    CData/*0:0*/ t__DOT__a2;
    CData/*0:0*/ t__DOT__a3;
    CData/*0:0*/ t__DOT__a4;
    CData/*0:0*/ t__DOT__a5;
    CData/*0:0*/ t__DOT__a6;
    CData/*0:0*/ t__DOT__a7;
    CData/*0:0*/ t__DOT__a8;
    CData/*0:0*/ t__DOT__a9;
    CData/*0:0*/ t__DOT__o1;
    CData/*0:0*/ t__DOT__o10;
    CData/*0:0*/ t__DOT__o11;
    CData/*0:0*/ t__DOT__o2;
    CData/*0:0*/ t__DOT__o3;
    CData/*0:0*/ t__DOT__o4;
    CData/*0:0*/ t__DOT__o5;
    CData/*0:0*/ t__DOT__o6;
    CData/*0:0*/ t__DOT__o7;
    CData/*0:0*/ t__DOT__o8;
    CData/*0:0*/ t__DOT__o9;
    CData/*0:0*/ t__DOT__x1;
    CData/*0:0*/ t__DOT__x2;
    CData/*0:0*/ t__DOT__x3;
    CData/*0:0*/ t__DOT__x4;
    CData/*0:0*/ t__DOT__x5;
    CData/*0:0*/ t__DOT__x6;
    CData/*0:0*/ t__DOT__x7;
    CData/*0:0*/ t__DOT__x8;
    CData/*0:0*/ t__DOT__x9;
    CData/*0:0*/ t__DOT__z1;
    CData/*0:0*/ t__DOT__z2;
    CData/*0:0*/ t__DOT__z3;
    CData/*0:0*/ t__DOT__z4;
    CData/*0:0*/ t__DOT__z5;
    CData/*0:0*/ t__DOT__z6;
    CData/*0:0*/ t__DOT__z7;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    IData/*31:0*/ t__DOT__cyc;
    VlWide<4>/*127:0*/ t__DOT__test__DOT__d;
    IData/*17:0*/ t__DOT__test__DOT__e;
    IData/*31:0*/ __VactIterCount;
    QData/*63:0*/ t__DOT__crc;
    QData/*63:0*/ t__DOT__sum;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vt_opt_const_red__cuda final {
    typedef Vt_opt_const_red___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 60,
        OFF_t__DOT__a1 = 61,
        OFF_t__DOT__a10 = 62,
        OFF_t__DOT__a11 = 63,
        OFF_t__DOT__a2 = 64,
        OFF_t__DOT__a3 = 65,
        OFF_t__DOT__a4 = 66,
        OFF_t__DOT__a5 = 67,
        OFF_t__DOT__a6 = 68,
        OFF_t__DOT__a7 = 69,
        OFF_t__DOT__a8 = 70,
        OFF_t__DOT__a9 = 71,
        OFF_t__DOT__o1 = 72,
        OFF_t__DOT__o10 = 73,
        OFF_t__DOT__o11 = 74,
        OFF_t__DOT__o2 = 75,
        OFF_t__DOT__o3 = 76,
        OFF_t__DOT__o4 = 77,
        OFF_t__DOT__o5 = 78,
        OFF_t__DOT__o6 = 79,
        OFF_t__DOT__o7 = 80,
        OFF_t__DOT__o8 = 81,
        OFF_t__DOT__o9 = 82,
        OFF_t__DOT__x1 = 83,
        OFF_t__DOT__x2 = 84,
        OFF_t__DOT__x3 = 85,
        OFF_t__DOT__x4 = 86,
        OFF_t__DOT__x5 = 87,
        OFF_t__DOT__x6 = 88,
        OFF_t__DOT__x7 = 89,
        OFF_t__DOT__x8 = 90,
        OFF_t__DOT__x9 = 91,
        OFF_t__DOT__z1 = 92,
        OFF_t__DOT__z2 = 93,
        OFF_t__DOT__z3 = 94,
        OFF_t__DOT__z4 = 95,
        OFF_t__DOT__z5 = 96,
        OFF_t__DOT__z6 = 97,
        OFF_t__DOT__z7 = 98,
        OFF___Vtrigprevexpr___TOP__clk__0 = 99,
        OFF___VactPhaseResult = 100,
        OFF___VnbaPhaseResult = 101,
        OFF_t__DOT__cyc = 32,
        OFF_t__DOT__test__DOT__d = 36,
        OFF_t__DOT__test__DOT__e = 52,
        OFF___VactIterCount = 56,
        OFF_t__DOT__crc = 0,
        OFF_t__DOT__sum = 8,
        OFF___VactTriggered = 16,
        OFF___VnbaTriggered = 24,
        kStateBytes = 104,  // S_persist
        kTouchedReadBytes = 102,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 101,  // S_state_wr
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
        s.t__DOT__a1 = st.ld<VlcuMem8>(OFF_t__DOT__a1, i);
        s.t__DOT__a10 = st.ld<VlcuMem8>(OFF_t__DOT__a10, i);
        s.t__DOT__a11 = st.ld<VlcuMem8>(OFF_t__DOT__a11, i);
        s.t__DOT__a2 = st.ld<VlcuMem8>(OFF_t__DOT__a2, i);
        s.t__DOT__a3 = st.ld<VlcuMem8>(OFF_t__DOT__a3, i);
        s.t__DOT__a4 = st.ld<VlcuMem8>(OFF_t__DOT__a4, i);
        s.t__DOT__a5 = st.ld<VlcuMem8>(OFF_t__DOT__a5, i);
        s.t__DOT__a6 = st.ld<VlcuMem8>(OFF_t__DOT__a6, i);
        s.t__DOT__a7 = st.ld<VlcuMem8>(OFF_t__DOT__a7, i);
        s.t__DOT__a8 = st.ld<VlcuMem8>(OFF_t__DOT__a8, i);
        s.t__DOT__a9 = st.ld<VlcuMem8>(OFF_t__DOT__a9, i);
        s.t__DOT__o1 = st.ld<VlcuMem8>(OFF_t__DOT__o1, i);
        s.t__DOT__o10 = st.ld<VlcuMem8>(OFF_t__DOT__o10, i);
        s.t__DOT__o11 = st.ld<VlcuMem8>(OFF_t__DOT__o11, i);
        s.t__DOT__o2 = st.ld<VlcuMem8>(OFF_t__DOT__o2, i);
        s.t__DOT__o3 = st.ld<VlcuMem8>(OFF_t__DOT__o3, i);
        s.t__DOT__o4 = st.ld<VlcuMem8>(OFF_t__DOT__o4, i);
        s.t__DOT__o5 = st.ld<VlcuMem8>(OFF_t__DOT__o5, i);
        s.t__DOT__o6 = st.ld<VlcuMem8>(OFF_t__DOT__o6, i);
        s.t__DOT__o7 = st.ld<VlcuMem8>(OFF_t__DOT__o7, i);
        s.t__DOT__o8 = st.ld<VlcuMem8>(OFF_t__DOT__o8, i);
        s.t__DOT__o9 = st.ld<VlcuMem8>(OFF_t__DOT__o9, i);
        s.t__DOT__x1 = st.ld<VlcuMem8>(OFF_t__DOT__x1, i);
        s.t__DOT__x2 = st.ld<VlcuMem8>(OFF_t__DOT__x2, i);
        s.t__DOT__x3 = st.ld<VlcuMem8>(OFF_t__DOT__x3, i);
        s.t__DOT__x4 = st.ld<VlcuMem8>(OFF_t__DOT__x4, i);
        s.t__DOT__x5 = st.ld<VlcuMem8>(OFF_t__DOT__x5, i);
        s.t__DOT__x6 = st.ld<VlcuMem8>(OFF_t__DOT__x6, i);
        s.t__DOT__x7 = st.ld<VlcuMem8>(OFF_t__DOT__x7, i);
        s.t__DOT__x8 = st.ld<VlcuMem8>(OFF_t__DOT__x8, i);
        s.t__DOT__x9 = st.ld<VlcuMem8>(OFF_t__DOT__x9, i);
        s.t__DOT__z1 = st.ld<VlcuMem8>(OFF_t__DOT__z1, i);
        s.t__DOT__z2 = st.ld<VlcuMem8>(OFF_t__DOT__z2, i);
        s.t__DOT__z3 = st.ld<VlcuMem8>(OFF_t__DOT__z3, i);
        s.t__DOT__z4 = st.ld<VlcuMem8>(OFF_t__DOT__z4, i);
        s.t__DOT__z5 = st.ld<VlcuMem8>(OFF_t__DOT__z5, i);
        s.t__DOT__z6 = st.ld<VlcuMem8>(OFF_t__DOT__z6, i);
        s.t__DOT__z7 = st.ld<VlcuMem8>(OFF_t__DOT__z7, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        s.t__DOT__cyc = st.ld<IData>(OFF_t__DOT__cyc, i);
        st.ldW(OFF_t__DOT__test__DOT__d, i, s.t__DOT__test__DOT__d);
        s.t__DOT__test__DOT__e = st.ld<IData>(OFF_t__DOT__test__DOT__e, i);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        s.t__DOT__crc = st.ld<QData>(OFF_t__DOT__crc, i);
        s.t__DOT__sum = st.ld<QData>(OFF_t__DOT__sum, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.st<VlcuMem8>(OFF_t__DOT__a1, i, s.t__DOT__a1);
        st.st<VlcuMem8>(OFF_t__DOT__a10, i, s.t__DOT__a10);
        st.st<VlcuMem8>(OFF_t__DOT__a11, i, s.t__DOT__a11);
        st.st<VlcuMem8>(OFF_t__DOT__a2, i, s.t__DOT__a2);
        st.st<VlcuMem8>(OFF_t__DOT__a3, i, s.t__DOT__a3);
        st.st<VlcuMem8>(OFF_t__DOT__a4, i, s.t__DOT__a4);
        st.st<VlcuMem8>(OFF_t__DOT__a5, i, s.t__DOT__a5);
        st.st<VlcuMem8>(OFF_t__DOT__a6, i, s.t__DOT__a6);
        st.st<VlcuMem8>(OFF_t__DOT__a7, i, s.t__DOT__a7);
        st.st<VlcuMem8>(OFF_t__DOT__a8, i, s.t__DOT__a8);
        st.st<VlcuMem8>(OFF_t__DOT__a9, i, s.t__DOT__a9);
        st.st<VlcuMem8>(OFF_t__DOT__o1, i, s.t__DOT__o1);
        st.st<VlcuMem8>(OFF_t__DOT__o10, i, s.t__DOT__o10);
        st.st<VlcuMem8>(OFF_t__DOT__o11, i, s.t__DOT__o11);
        st.st<VlcuMem8>(OFF_t__DOT__o2, i, s.t__DOT__o2);
        st.st<VlcuMem8>(OFF_t__DOT__o3, i, s.t__DOT__o3);
        st.st<VlcuMem8>(OFF_t__DOT__o4, i, s.t__DOT__o4);
        st.st<VlcuMem8>(OFF_t__DOT__o5, i, s.t__DOT__o5);
        st.st<VlcuMem8>(OFF_t__DOT__o6, i, s.t__DOT__o6);
        st.st<VlcuMem8>(OFF_t__DOT__o7, i, s.t__DOT__o7);
        st.st<VlcuMem8>(OFF_t__DOT__o8, i, s.t__DOT__o8);
        st.st<VlcuMem8>(OFF_t__DOT__o9, i, s.t__DOT__o9);
        st.st<VlcuMem8>(OFF_t__DOT__x1, i, s.t__DOT__x1);
        st.st<VlcuMem8>(OFF_t__DOT__x2, i, s.t__DOT__x2);
        st.st<VlcuMem8>(OFF_t__DOT__x3, i, s.t__DOT__x3);
        st.st<VlcuMem8>(OFF_t__DOT__x4, i, s.t__DOT__x4);
        st.st<VlcuMem8>(OFF_t__DOT__x5, i, s.t__DOT__x5);
        st.st<VlcuMem8>(OFF_t__DOT__x6, i, s.t__DOT__x6);
        st.st<VlcuMem8>(OFF_t__DOT__x7, i, s.t__DOT__x7);
        st.st<VlcuMem8>(OFF_t__DOT__x8, i, s.t__DOT__x8);
        st.st<VlcuMem8>(OFF_t__DOT__x9, i, s.t__DOT__x9);
        st.st<VlcuMem8>(OFF_t__DOT__z1, i, s.t__DOT__z1);
        st.st<VlcuMem8>(OFF_t__DOT__z2, i, s.t__DOT__z2);
        st.st<VlcuMem8>(OFF_t__DOT__z3, i, s.t__DOT__z3);
        st.st<VlcuMem8>(OFF_t__DOT__z4, i, s.t__DOT__z4);
        st.st<VlcuMem8>(OFF_t__DOT__z5, i, s.t__DOT__z5);
        st.st<VlcuMem8>(OFF_t__DOT__z6, i, s.t__DOT__z6);
        st.st<VlcuMem8>(OFF_t__DOT__z7, i, s.t__DOT__z7);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.st<IData>(OFF_t__DOT__cyc, i, s.t__DOT__cyc);
        st.stW(OFF_t__DOT__test__DOT__d, i, s.t__DOT__test__DOT__d);
        st.st<IData>(OFF_t__DOT__test__DOT__e, i, s.t__DOT__test__DOT__e);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
        st.st<QData>(OFF_t__DOT__crc, i, s.t__DOT__crc);
        st.st<QData>(OFF_t__DOT__sum, i, s.t__DOT__sum);
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

    VLCU_D static void Vt_opt_const_red___024root___eval_static(Vt_opt_const_red___024root* vlSelf, VlcuContext& vlCtx) {
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

    VLCU_D static void Vt_opt_const_red___024root___eval_initial(Vt_opt_const_red___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_opt_const_red___024root___eval_final(Vt_opt_const_red___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_opt_const_red___024root___eval_settle(Vt_opt_const_red___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static bool Vt_opt_const_red___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt_opt_const_red___024root___nba_sequent__TOP__0(Vt_opt_const_red___024root* vlSelf, VlcuContext& vlCtx) {
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
        // Body
        __Vdly__t__DOT__cyc = vlSelfRef.t__DOT__cyc;
        __Vdly__t__DOT__crc = vlSelfRef.t__DOT__crc;
        __Vdly__t__DOT__sum = vlSelfRef.t__DOT__sum;
        __Vdly__t__DOT__cyc = ((IData)(1U) + vlSelfRef.t__DOT__cyc);
        __Vdly__t__DOT__crc = ((vlSelfRef.t__DOT__crc 
                                << 1U) | (QData)((IData)(
                                                         (1U 
                                                          & VL_REDXOR_64(
                                                                         (0x8000000000000005ULL 
                                                                          & vlSelfRef.t__DOT__crc))))));
        __Vdly__t__DOT__sum = ((QData)((IData)(((((
                                                   (((((IData)(vlSelfRef.t__DOT__a1) 
                                                       << 3U) 
                                                      | ((IData)(vlSelfRef.t__DOT__a2) 
                                                         << 2U)) 
                                                     | (((IData)(vlSelfRef.t__DOT__a3) 
                                                         << 1U) 
                                                        | (IData)(vlSelfRef.t__DOT__a4))) 
                                                    << 0x0000000cU) 
                                                   | (((((IData)(vlSelfRef.t__DOT__a5) 
                                                         << 3U) 
                                                        | ((IData)(vlSelfRef.t__DOT__a6) 
                                                           << 2U)) 
                                                       | (((IData)(vlSelfRef.t__DOT__a7) 
                                                           << 1U) 
                                                          | (IData)(vlSelfRef.t__DOT__a8))) 
                                                      << 8U)) 
                                                  | ((((((IData)(vlSelfRef.t__DOT__a9) 
                                                         << 3U) 
                                                        | ((IData)(vlSelfRef.t__DOT__a10) 
                                                           << 2U)) 
                                                       | (((IData)(vlSelfRef.t__DOT__a11) 
                                                           << 1U) 
                                                          | (IData)(vlSelfRef.t__DOT__o1))) 
                                                      << 4U) 
                                                     | ((((IData)(vlSelfRef.t__DOT__o2) 
                                                          << 3U) 
                                                         | ((IData)(vlSelfRef.t__DOT__o3) 
                                                            << 2U)) 
                                                        | (((IData)(vlSelfRef.t__DOT__o4) 
                                                            << 1U) 
                                                           | (IData)(vlSelfRef.t__DOT__o5))))) 
                                                 << 0x0000000fU) 
                                                | (((((((IData)(vlSelfRef.t__DOT__o6) 
                                                        << 3U) 
                                                       | ((IData)(vlSelfRef.t__DOT__o7) 
                                                          << 2U)) 
                                                      | (((IData)(vlSelfRef.t__DOT__o8) 
                                                          << 1U) 
                                                         | (IData)(vlSelfRef.t__DOT__o9))) 
                                                     << 0x0000000bU) 
                                                    | (((((IData)(vlSelfRef.t__DOT__o10) 
                                                          << 3U) 
                                                         | ((IData)(vlSelfRef.t__DOT__o11) 
                                                            << 2U)) 
                                                        | (((IData)(vlSelfRef.t__DOT__x1) 
                                                            << 1U) 
                                                           | (IData)(vlSelfRef.t__DOT__x2))) 
                                                       << 7U)) 
                                                   | ((((((IData)(vlSelfRef.t__DOT__x3) 
                                                          << 3U) 
                                                         | ((IData)(vlSelfRef.t__DOT__x4) 
                                                            << 2U)) 
                                                        | (((IData)(vlSelfRef.t__DOT__x5) 
                                                            << 1U) 
                                                           | (IData)(vlSelfRef.t__DOT__x6))) 
                                                       << 3U) 
                                                      | (((IData)(vlSelfRef.t__DOT__x7) 
                                                          << 2U) 
                                                         | (((IData)(vlSelfRef.t__DOT__x8) 
                                                             << 1U) 
                                                            | (IData)(vlSelfRef.t__DOT__x9)))))))) 
                               ^ ((vlSelfRef.t__DOT__sum 
                                   << 1U) | (QData)((IData)(
                                                            (1U 
                                                             & VL_REDXOR_64(
                                                                            (0x8000000000000005ULL 
                                                                             & vlSelfRef.t__DOT__sum)))))));
        if ((0U == vlSelfRef.t__DOT__cyc)) {
            __Vdly__t__DOT__crc = 0x5aef0c8dd70a4497ULL;
            __Vdly__t__DOT__sum = 0ULL;
        } else if (VL_GTS_III(32, 0x0000000aU, vlSelfRef.t__DOT__cyc)) {
            __Vdly__t__DOT__sum = 0ULL;
        } else if (VL_LIKELY((VL_GTS_III(32, 0x00000063U, vlSelfRef.t__DOT__cyc)))) {
            if (VL_UNLIKELY((((IData)(vlSelfRef.t__DOT__a1) 
                              != (IData)(vlSelfRef.t__DOT__a2))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_opt_const_red.v", 151, "");
            }
            if (VL_UNLIKELY((((IData)(vlSelfRef.t__DOT__a1) 
                              != (IData)(vlSelfRef.t__DOT__a3))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_opt_const_red.v", 152, "");
            }
            if (VL_UNLIKELY((((IData)(vlSelfRef.t__DOT__a1) 
                              != (IData)(vlSelfRef.t__DOT__a4))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_opt_const_red.v", 153, "");
            }
            if (VL_UNLIKELY((((IData)(vlSelfRef.t__DOT__a1) 
                              != (IData)(vlSelfRef.t__DOT__a5))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_opt_const_red.v", 154, "");
            }
            if (VL_UNLIKELY((((IData)(vlSelfRef.t__DOT__a6) 
                              != (IData)(vlSelfRef.t__DOT__a7))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_opt_const_red.v", 155, "");
            }
            if (VL_UNLIKELY((((IData)(vlSelfRef.t__DOT__a8) 
                              != (IData)(vlSelfRef.t__DOT__a9))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_opt_const_red.v", 156, "");
            }
            if (VL_UNLIKELY((((IData)(vlSelfRef.t__DOT__o1) 
                              != (IData)(vlSelfRef.t__DOT__o2))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_opt_const_red.v", 157, "");
            }
            if (VL_UNLIKELY((((IData)(vlSelfRef.t__DOT__o1) 
                              != (IData)(vlSelfRef.t__DOT__o3))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_opt_const_red.v", 158, "");
            }
            if (VL_UNLIKELY((((IData)(vlSelfRef.t__DOT__o1) 
                              != (IData)(vlSelfRef.t__DOT__o4))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_opt_const_red.v", 159, "");
            }
            if (VL_UNLIKELY((((IData)(vlSelfRef.t__DOT__o1) 
                              != (IData)(vlSelfRef.t__DOT__o5))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_opt_const_red.v", 160, "");
            }
            if (VL_UNLIKELY((((IData)(vlSelfRef.t__DOT__o6) 
                              != (IData)(vlSelfRef.t__DOT__o7))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_opt_const_red.v", 161, "");
            }
            if (VL_UNLIKELY((((IData)(vlSelfRef.t__DOT__o8) 
                              != (IData)(vlSelfRef.t__DOT__o9))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_opt_const_red.v", 162, "");
            }
            if (VL_UNLIKELY((((IData)(vlSelfRef.t__DOT__x1) 
                              != (IData)(vlSelfRef.t__DOT__x2))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_opt_const_red.v", 163, "");
            }
            if (VL_UNLIKELY((((IData)(vlSelfRef.t__DOT__x1) 
                              != (IData)(vlSelfRef.t__DOT__x3))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_opt_const_red.v", 164, "");
            }
            if (VL_UNLIKELY((((IData)(vlSelfRef.t__DOT__x1) 
                              != (IData)(vlSelfRef.t__DOT__x4))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_opt_const_red.v", 165, "");
            }
            if (VL_UNLIKELY((((IData)(vlSelfRef.t__DOT__x1) 
                              != (IData)(vlSelfRef.t__DOT__x5))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_opt_const_red.v", 166, "");
            }
            if (VL_UNLIKELY((((IData)(vlSelfRef.t__DOT__x1) 
                              != (IData)(vlSelfRef.t__DOT__x6))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_opt_const_red.v", 167, "");
            }
            if (VL_UNLIKELY((((IData)(vlSelfRef.t__DOT__x1) 
                              != (IData)(vlSelfRef.t__DOT__x7))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_opt_const_red.v", 168, "");
            }
            if (VL_UNLIKELY((vlSelfRef.t__DOT__z1))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_opt_const_red.v", 169, "");
            }
            if (VL_UNLIKELY(((1U & (~ (IData)(vlSelfRef.t__DOT__z2)))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_opt_const_red.v", 170, "");
            }
            if (VL_UNLIKELY((vlSelfRef.t__DOT__z3))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_opt_const_red.v", 171, "");
            }
            if (VL_UNLIKELY((vlSelfRef.t__DOT__z4))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_opt_const_red.v", 172, "");
            }
            if (VL_UNLIKELY(((1U & (~ (IData)(vlSelfRef.t__DOT__z5)))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_opt_const_red.v", 173, "");
            }
            if (VL_UNLIKELY(((1U & (~ (IData)(vlSelfRef.t__DOT__z6)))))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_opt_const_red.v", 174, "");
            }
            if (VL_UNLIKELY((vlSelfRef.t__DOT__z7))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_opt_const_red.v", 175, "");
            }
        } else {
            vlCtx.writef(0U, 4U, vlcu_msg_arg(VL_TIME_UNITED_Q(1)), vlcu_msg_arg(vlSelfRef.t__DOT__cyc), vlcu_msg_arg(vlSelfRef.t__DOT__crc), vlcu_msg_arg(vlSelfRef.t__DOT__sum));
            if (VL_UNLIKELY(((0xc77bb9b3784ea091ULL 
                              != vlSelfRef.t__DOT__crc)))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_opt_const_red.v", 183, "");
            }
            if (VL_UNLIKELY(((0x727fb78d09c1981eULL 
                              != vlSelfRef.t__DOT__sum)))) {
                VL_STOP_MT("/root/reference/test_regress/t/t_opt_const_red.v", 186, "");
            }
            vlCtx.writef(1U, 0U);
            VL_FINISH_MT("/root/reference/test_regress/t/t_opt_const_red.v", 188, "");
        }
        vlSelfRef.t__DOT__cyc = __Vdly__t__DOT__cyc;
        vlSelfRef.t__DOT__sum = __Vdly__t__DOT__sum;
        vlSelfRef.t__DOT__a1 = (IData)((0x0000000000000022ULL 
                                        == (0x000000000000002aULL 
                                            & vlSelfRef.t__DOT__crc)));
        vlSelfRef.t__DOT__a2 = (IData)((0x0000000000000022ULL 
                                        == (0x000000000000002aULL 
                                            & vlSelfRef.t__DOT__crc)));
        vlSelfRef.t__DOT__a3 = (IData)((0x0000000000000022ULL 
                                        == (0x000000000000002aULL 
                                            & vlSelfRef.t__DOT__crc)));
        vlSelfRef.t__DOT__a4 = (0x00000022U == (0x0000002aU 
                                                & (IData)(vlSelfRef.t__DOT__crc)));
        vlSelfRef.t__DOT__a5 = (IData)((0x0000000000000022ULL 
                                        == (0x000000000000002aULL 
                                            & vlSelfRef.t__DOT__crc)));
        vlSelfRef.t__DOT__a6 = (7U == (7U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 3U))));
        vlSelfRef.t__DOT__a7 = (IData)((0x0000000000000038ULL 
                                        == (0x0000000000000038ULL 
                                            & vlSelfRef.t__DOT__crc)));
        vlSelfRef.t__DOT__a8 = (7U == (7U & (~ (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 3U)))));
        vlSelfRef.t__DOT__a9 = (IData)((0ULL == (0x0000000000000038ULL 
                                                 & vlSelfRef.t__DOT__crc)));
        vlSelfRef.t__DOT__o1 = (IData)((0x0000000000000022ULL 
                                        != (0x000000000000002aULL 
                                            & vlSelfRef.t__DOT__crc)));
        vlSelfRef.t__DOT__o2 = (IData)((0x0000000000000022ULL 
                                        != (0x000000000000002aULL 
                                            & vlSelfRef.t__DOT__crc)));
        vlSelfRef.t__DOT__o3 = (IData)((0x0000000000000022ULL 
                                        != (0x000000000000002aULL 
                                            & vlSelfRef.t__DOT__crc)));
        vlSelfRef.t__DOT__o4 = (0x00000022U != (0x0000002aU 
                                                & (IData)(vlSelfRef.t__DOT__crc)));
        vlSelfRef.t__DOT__o5 = (IData)((0x0000000000000022ULL 
                                        != (0x000000000000002aULL 
                                            & vlSelfRef.t__DOT__crc)));
        vlSelfRef.t__DOT__o6 = (0U != (7U & (IData)(
                                                    (vlSelfRef.t__DOT__crc 
                                                     >> 3U))));
        vlSelfRef.t__DOT__o7 = (IData)((0ULL != (0x0000000000000038ULL 
                                                 & vlSelfRef.t__DOT__crc)));
        vlSelfRef.t__DOT__o8 = (0U != (7U & (~ (IData)(
                                                       (vlSelfRef.t__DOT__crc 
                                                        >> 3U)))));
        vlSelfRef.t__DOT__o9 = (IData)((0x0000000000000038ULL 
                                        != (0x0000000000000038ULL 
                                            & vlSelfRef.t__DOT__crc)));
        vlSelfRef.t__DOT__x1 = (1U & (~ VL_REDXOR_8(
                                                    (0x000000000000002aULL 
                                                     & vlSelfRef.t__DOT__crc))));
        vlSelfRef.t__DOT__x2 = (1U & (~ VL_REDXOR_8(
                                                    (0x000000000000002aULL 
                                                     & vlSelfRef.t__DOT__crc))));
        vlSelfRef.t__DOT__x3 = (1U & (~ VL_REDXOR_8(
                                                    (0x000000000000002aULL 
                                                     & vlSelfRef.t__DOT__crc))));
        vlSelfRef.t__DOT__x4 = (1U & (~ VL_REDXOR_32(
                                                     (0x0000002aU 
                                                      & (IData)(vlSelfRef.t__DOT__crc)))));
        vlSelfRef.t__DOT__x5 = (1U & (~ VL_REDXOR_8(
                                                    (0x000000000000002aULL 
                                                     & vlSelfRef.t__DOT__crc))));
        vlSelfRef.t__DOT__x6 = (1U & (~ VL_REDXOR_8(
                                                    (0x000000000000002aULL 
                                                     & vlSelfRef.t__DOT__crc))));
        vlSelfRef.t__DOT__x7 = (1U & (~ VL_REDXOR_8(
                                                    (0x000000000000002aULL 
                                                     & vlSelfRef.t__DOT__crc))));
        vlSelfRef.t__DOT__z1 = 0U;
        vlSelfRef.t__DOT__z2 = 1U;
        vlSelfRef.t__DOT__z3 = 0U;
        vlSelfRef.t__DOT__z4 = 0U;
        vlSelfRef.t__DOT__z5 = 1U;
        vlSelfRef.t__DOT__z6 = 1U;
        vlSelfRef.t__DOT__z7 = 0U;
        vlSelfRef.t__DOT__a10 = (1U & ((~ ((IData)(
                                                   (vlSelfRef.t__DOT__crc 
                                                    >> 5U)) 
                                           & (~ (vlSelfRef.t__DOT__test__DOT__d[0U] 
                                                 >> 3U)))) 
                                       & ((~ (IData)(
                                                     (vlSelfRef.t__DOT__crc 
                                                      >> 5U))) 
                                          & (vlSelfRef.t__DOT__test__DOT__d[0U] 
                                             >> 1U))));
        vlSelfRef.t__DOT__o10 = (1U & ((~ ((~ (IData)(
                                                      (vlSelfRef.t__DOT__crc 
                                                       >> 5U))) 
                                           | (vlSelfRef.t__DOT__test__DOT__d[0U] 
                                              >> 3U))) 
                                       | ((IData)((vlSelfRef.t__DOT__crc 
                                                   >> 5U)) 
                                          | (~ (vlSelfRef.t__DOT__test__DOT__d[0U] 
                                                >> 1U)))));
        vlSelfRef.t__DOT__x8 = (1U & (~ VL_REDXOR_4(
                                                    (0x0000000aU 
                                                     & vlSelfRef.t__DOT__test__DOT__d[0U]))));
        vlSelfRef.t__DOT__a11 = (IData)((((((1U == 
                                             (0x80000001U 
                                              & vlSelfRef.t__DOT__test__DOT__d[0U])) 
                                            & (2U == 
                                               (0x40000002U 
                                                & vlSelfRef.t__DOT__test__DOT__d[1U]))) 
                                           & (4U == 
                                              (0x20000004U 
                                               & vlSelfRef.t__DOT__test__DOT__d[2U]))) 
                                          & (8U == 
                                             (0x10000008U 
                                              & vlSelfRef.t__DOT__test__DOT__d[3U]))) 
                                         & (5U == (7U 
                                                   & vlSelfRef.t__DOT__test__DOT__e))));
        vlSelfRef.t__DOT__o11 = (1U & (IData)((((((0x80000000U 
                                                   != 
                                                   (0x80000001U 
                                                    & vlSelfRef.t__DOT__test__DOT__d[0U])) 
                                                  | (0x40000000U 
                                                     != 
                                                     (0x40000002U 
                                                      & vlSelfRef.t__DOT__test__DOT__d[1U]))) 
                                                 | (0x20000000U 
                                                    != 
                                                    (0x20000004U 
                                                     & vlSelfRef.t__DOT__test__DOT__d[2U]))) 
                                                | (0x10000000U 
                                                   != 
                                                   (0x10000008U 
                                                    & vlSelfRef.t__DOT__test__DOT__d[3U]))) 
                                               | (2U 
                                                  != 
                                                  (7U 
                                                   & vlSelfRef.t__DOT__test__DOT__e)))));
        vlSelfRef.t__DOT__x9 = (1U & (~ ((((VL_REDXOR_32(
                                                         (0x80000001U 
                                                          & vlSelfRef.t__DOT__test__DOT__d[0U])) 
                                            ^ VL_REDXOR_32(
                                                           (0x40000002U 
                                                            & vlSelfRef.t__DOT__test__DOT__d[1U]))) 
                                           ^ VL_REDXOR_32(
                                                          (0x20000004U 
                                                           & vlSelfRef.t__DOT__test__DOT__d[2U]))) 
                                          ^ VL_REDXOR_32(
                                                         (0x10000008U 
                                                          & vlSelfRef.t__DOT__test__DOT__d[3U]))) 
                                         ^ VL_REDXOR_4(
                                                       (7U 
                                                        & vlSelfRef.t__DOT__test__DOT__e)))));
        vlSelfRef.t__DOT__test__DOT__e = (0x0003ffffU 
                                          & (IData)(vlSelfRef.t__DOT__crc));
        vlSelfRef.t__DOT__test__DOT__d[0U] = (IData)(
                                                     (((QData)((IData)(
                                                                       (~ (IData)(vlSelfRef.t__DOT__crc)))) 
                                                       << 0x00000020U) 
                                                      | (QData)((IData)(vlSelfRef.t__DOT__crc))));
        vlSelfRef.t__DOT__test__DOT__d[1U] = (IData)(
                                                     ((((QData)((IData)(
                                                                        (~ (IData)(vlSelfRef.t__DOT__crc)))) 
                                                        << 0x00000020U) 
                                                       | (QData)((IData)(vlSelfRef.t__DOT__crc))) 
                                                      >> 0x00000020U));
        vlSelfRef.t__DOT__test__DOT__d[2U] = (IData)(
                                                     (((QData)((IData)(vlSelfRef.t__DOT__crc)) 
                                                       << 0x00000020U) 
                                                      | (QData)((IData)(
                                                                        (~ (IData)(vlSelfRef.t__DOT__crc))))));
        vlSelfRef.t__DOT__test__DOT__d[3U] = (IData)(
                                                     ((((QData)((IData)(vlSelfRef.t__DOT__crc)) 
                                                        << 0x00000020U) 
                                                       | (QData)((IData)(
                                                                         (~ (IData)(vlSelfRef.t__DOT__crc))))) 
                                                      >> 0x00000020U));
        vlSelfRef.t__DOT__crc = __Vdly__t__DOT__crc;
    }

    VLCU_D static void Vt_opt_const_red___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt_opt_const_red___024root___eval_phase__act(Vt_opt_const_red___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt_opt_const_red___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vt_opt_const_red___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vt_opt_const_red___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vt_opt_const_red___024root___eval_phase__nba(Vt_opt_const_red___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vt_opt_const_red___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    Vt_opt_const_red___024root___nba_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
            Vt_opt_const_red___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vt_opt_const_red___024root___eval(Vt_opt_const_red___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt_opt_const_red___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("/root/reference/test_regress/t/t_opt_const_red.v", 7, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vt_opt_const_red___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("/root/reference/test_regress/t/t_opt_const_red.v", 7, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vt_opt_const_red___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vt_opt_const_red___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vt_opt_const_red___024root___ctor_var_reset(Vt_opt_const_red___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        vlSelf->t__DOT__crc = VL_SCOPED_RAND_RESET_Q(64, __VscopeHash, 17680824324711171631ull);
        vlSelf->t__DOT__sum = VL_SCOPED_RAND_RESET_Q(64, __VscopeHash, 1514072735762163239ull);
        vlSelf->t__DOT__a1 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 4366861477568935566ull);
        vlSelf->t__DOT__a10 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 15050927298126456118ull);
        vlSelf->t__DOT__a11 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16320715397657832256ull);
        vlSelf->t__DOT__a2 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 1828157105111444767ull);
        vlSelf->t__DOT__a3 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 5922719172576119868ull);
        vlSelf->t__DOT__a4 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 911307258566730252ull);
        vlSelf->t__DOT__a5 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 10113598769988735898ull);
        vlSelf->t__DOT__a6 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 4826927137749829205ull);
        vlSelf->t__DOT__a7 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 18021243644292094052ull);
        vlSelf->t__DOT__a8 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 7477144232297018548ull);
        vlSelf->t__DOT__a9 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 805445553162773218ull);
        vlSelf->t__DOT__o1 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 9010135485891048095ull);
        vlSelf->t__DOT__o10 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 14407099543014902190ull);
        vlSelf->t__DOT__o11 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 8841585019015039461ull);
        vlSelf->t__DOT__o2 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 5222920395842680255ull);
        vlSelf->t__DOT__o3 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 7088017066039472267ull);
        vlSelf->t__DOT__o4 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 10586860685099884992ull);
        vlSelf->t__DOT__o5 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 13403120952154058583ull);
        vlSelf->t__DOT__o6 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 13128126042036107613ull);
        vlSelf->t__DOT__o7 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 10465803787554107993ull);
        vlSelf->t__DOT__o8 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 105646776429392716ull);
        vlSelf->t__DOT__o9 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 8333203150570829696ull);
        vlSelf->t__DOT__x1 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 2481947800948510401ull);
        vlSelf->t__DOT__x2 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 3077931057464295972ull);
        vlSelf->t__DOT__x3 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 415608802982304187ull);
        vlSelf->t__DOT__x4 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 4729841474307995368ull);
        vlSelf->t__DOT__x5 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 2066035081906823631ull);
        vlSelf->t__DOT__x6 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 13392550882030870371ull);
        vlSelf->t__DOT__x7 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 9554199659865895570ull);
        vlSelf->t__DOT__x8 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 10790817954505986597ull);
        vlSelf->t__DOT__x9 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 14894568408700153673ull);
        vlSelf->t__DOT__z1 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 2221872139190500375ull);
        vlSelf->t__DOT__z2 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 2008685885837387513ull);
        vlSelf->t__DOT__z3 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16617078737382059286ull);
        vlSelf->t__DOT__z4 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 3798597338399216679ull);
        vlSelf->t__DOT__z5 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 900737188443428294ull);
        vlSelf->t__DOT__z6 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 2765833858640189290ull);
        vlSelf->t__DOT__z7 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 9284935611378647587ull);
        VL_SCOPED_RAND_RESET_W(128, vlSelf->t__DOT__test__DOT__d, __VscopeHash, 1135660206730878013ull);
        vlSelf->t__DOT__test__DOT__e = VL_SCOPED_RAND_RESET_I(18, __VscopeHash, 16066567358933958484ull);
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
        Vt_opt_const_red___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt_opt_const_red___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt_opt_const_red___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt_opt_const_red___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vt_opt_const_red___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt_opt_const_red___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt_opt_const_red__cuda::info() {
    typedef Vt_opt_const_red__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"t__DOT__a1", "t.a1", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__a1},
        {"t__DOT__a10", "t.a10", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__a10},
        {"t__DOT__a11", "t.a11", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__a11},
        {"t__DOT__a2", "t.a2", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__a2},
        {"t__DOT__a3", "t.a3", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__a3},
        {"t__DOT__a4", "t.a4", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__a4},
        {"t__DOT__a5", "t.a5", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__a5},
        {"t__DOT__a6", "t.a6", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__a6},
        {"t__DOT__a7", "t.a7", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__a7},
        {"t__DOT__a8", "t.a8", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__a8},
        {"t__DOT__a9", "t.a9", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__a9},
        {"t__DOT__o1", "t.o1", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__o1},
        {"t__DOT__o10", "t.o10", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__o10},
        {"t__DOT__o11", "t.o11", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__o11},
        {"t__DOT__o2", "t.o2", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__o2},
        {"t__DOT__o3", "t.o3", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__o3},
        {"t__DOT__o4", "t.o4", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__o4},
        {"t__DOT__o5", "t.o5", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__o5},
        {"t__DOT__o6", "t.o6", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__o6},
        {"t__DOT__o7", "t.o7", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__o7},
        {"t__DOT__o8", "t.o8", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__o8},
        {"t__DOT__o9", "t.o9", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__o9},
        {"t__DOT__x1", "t.x1", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__x1},
        {"t__DOT__x2", "t.x2", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__x2},
        {"t__DOT__x3", "t.x3", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__x3},
        {"t__DOT__x4", "t.x4", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__x4},
        {"t__DOT__x5", "t.x5", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__x5},
        {"t__DOT__x6", "t.x6", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__x6},
        {"t__DOT__x7", "t.x7", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__x7},
        {"t__DOT__x8", "t.x8", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__x8},
        {"t__DOT__x9", "t.x9", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__x9},
        {"t__DOT__z1", "t.z1", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__z1},
        {"t__DOT__z2", "t.z2", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__z2},
        {"t__DOT__z3", "t.z3", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__z3},
        {"t__DOT__z4", "t.z4", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__z4},
        {"t__DOT__z5", "t.z5", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__z5},
        {"t__DOT__z6", "t.z6", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__z6},
        {"t__DOT__z7", "t.z7", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t__DOT__z7},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"t__DOT__cyc", "t.cyc", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__cyc},
        {"t__DOT__test__DOT__d", "t.test.d", VLCU_MEMBER_STATE, 4, 0, 4, 128, 127, 0, T::OFF_t__DOT__test__DOT__d},
        {"t__DOT__test__DOT__e", "t.test.e", VLCU_MEMBER_STATE, 4, 0, 1, 18, 17, 0, T::OFF_t__DOT__test__DOT__e},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"t__DOT__crc", "t.crc", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF_t__DOT__crc},
        {"t__DOT__sum", "t.sum", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF_t__DOT__sum},
        {"__VactTriggered", "__VactTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VactTriggered},
        {"__VnbaTriggered", "__VnbaTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VnbaTriggered},
    };
    static const char* const messages[] = {"[%0t] cyc==%0d crc=%x sum=%x\n", "*-* All Finished *-*\n", nullptr};
    static const char* const plusargs[] = {nullptr};
    static const VlcuTraceDecl traceDecls[] = {
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0}};  // (terminator, not counted)
    static const uint32_t traceOrder[] = {0};  // codes in dump order (+ terminator)
    static const uint32_t saveOrder[] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 36, 37, 38, 39, 40, 41, 42, 43, 44, 45, 46, 47, 48, 49, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "t", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 2, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 50, 0, 0xdbd1eb0e99358474ULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt_opt_const_red__cuda VLCU_TESTBENCHES);
