// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

// 40 bytes of register-resident state: latency-bound, compile for full occupancy
#ifndef VLCU_MIN_BLOCKS
#define VLCU_MIN_BLOCKS 32
#endif

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt_math_signed2"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt_math_signed2___024root final {
    CData/*0:0*/ clk;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    SData/*10:0*/ t__DOT__in;
    IData/*31:0*/ __VactIterCount;
    VlcuLocalArray<CData/*7:0*/, 32> t__DOT__delay_minmax;  // thread-local memory
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vt_math_signed2__cuda final {
    typedef Vt_math_signed2___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 22,
        OFF___Vtrigprevexpr___TOP__clk__0 = 23,
        OFF___VactPhaseResult = 24,
        OFF___VnbaPhaseResult = 25,
        OFF_t__DOT__in = 20,
        OFF___VactIterCount = 16,
        OFF_t__DOT__delay_minmax = 26,
        OFF___VactTriggered = 0,
        OFF___VnbaTriggered = 8,
        kStateBytes = 64,  // S_persist
        kTouchedReadBytes = 58,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 57,  // S_state_wr
    };
    struct Locals final {
        CData/*7:0*/ t__DOT__delay_minmax[32];
    };
    struct Shared final {
    };
    VLCU_D static void bind(Root& s, Locals& l, Shared& sh, const VlcuState& st, uint64_t i) {
        (void)l; (void)sh; (void)st; (void)i;
        s.vlNamep = "TOP";
        s.t__DOT__delay_minmax.bind(l.t__DOT__delay_minmax);
    }
    VLCU_D static void load(Root& s, const VlcuState& st, uint64_t i) {
        s.clk = st.ld<VlcuMem8>(OFF_clk, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        s.t__DOT__in = st.ld<VlcuMem16>(OFF_t__DOT__in, i);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        for (uint32_t e = 0; e < 32; ++e) s.t__DOT__delay_minmax[e] = st.ldE<VlcuMem8>(OFF_t__DOT__delay_minmax, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.st<VlcuMem16>(OFF_t__DOT__in, i, s.t__DOT__in);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
        for (uint32_t e = 0; e < 32; ++e) st.stE<VlcuMem8>(OFF_t__DOT__delay_minmax, e, i, s.t__DOT__delay_minmax[e]);
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

    VLCU_D static void Vt_math_signed2___024root___eval_static(Vt_math_signed2___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
    }

    VLCU_D static void Vt_math_signed2___024root___eval_initial(Vt_math_signed2___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        {
            // Inlined CFunc: _eval_initial__TOP
            vlSelfRef.t__DOT__in = 0x0408U;
            vlSelfRef.t__DOT__delay_minmax[0U] = 0U;
            vlSelfRef.t__DOT__delay_minmax[1U] = 0U;
            vlSelfRef.t__DOT__delay_minmax[2U] = 0U;
            vlSelfRef.t__DOT__delay_minmax[3U] = 0U;
            vlSelfRef.t__DOT__delay_minmax[4U] = 0U;
            vlSelfRef.t__DOT__delay_minmax[5U] = 0U;
            vlSelfRef.t__DOT__delay_minmax[6U] = 0U;
            vlSelfRef.t__DOT__delay_minmax[7U] = 0U;
            vlSelfRef.t__DOT__delay_minmax[8U] = 0U;
            vlSelfRef.t__DOT__delay_minmax[9U] = 0U;
            vlSelfRef.t__DOT__delay_minmax[10U] = 0U;
            vlSelfRef.t__DOT__delay_minmax[11U] = 0U;
            vlSelfRef.t__DOT__delay_minmax[12U] = 0U;
            vlSelfRef.t__DOT__delay_minmax[13U] = 0U;
            vlSelfRef.t__DOT__delay_minmax[14U] = 0U;
            vlSelfRef.t__DOT__delay_minmax[15U] = 0U;
            vlSelfRef.t__DOT__delay_minmax[16U] = 0U;
            vlSelfRef.t__DOT__delay_minmax[17U] = 0U;
            vlSelfRef.t__DOT__delay_minmax[18U] = 0U;
            vlSelfRef.t__DOT__delay_minmax[19U] = 0U;
            vlSelfRef.t__DOT__delay_minmax[20U] = 0U;
            vlSelfRef.t__DOT__delay_minmax[21U] = 0U;
            vlSelfRef.t__DOT__delay_minmax[22U] = 0U;
            vlSelfRef.t__DOT__delay_minmax[23U] = 0U;
            vlSelfRef.t__DOT__delay_minmax[24U] = 0U;
            vlSelfRef.t__DOT__delay_minmax[25U] = 0U;
            vlSelfRef.t__DOT__delay_minmax[26U] = 0U;
            vlSelfRef.t__DOT__delay_minmax[27U] = 0U;
            vlSelfRef.t__DOT__delay_minmax[28U] = 0U;
            vlSelfRef.t__DOT__delay_minmax[29U] = 0U;
            vlSelfRef.t__DOT__delay_minmax[30U] = 0U;
            vlSelfRef.t__DOT__delay_minmax[31U] = 0U;
        }
    }

    VLCU_D static void Vt_math_signed2___024root___eval_final(Vt_math_signed2___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_math_signed2___024root___eval_settle(Vt_math_signed2___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static bool Vt_math_signed2___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt_math_signed2___024root___nba_sequent__TOP__0(Vt_math_signed2___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*7:0*/ t__DOT__min;
        t__DOT__min = 0;
        CData/*7:0*/ t__DOT__max;
        t__DOT__max = 0;
        // Body
        vlSelfRef.t__DOT__in = (0x000007ffU & ((IData)(8U) 
                                               + (IData)(vlSelfRef.t__DOT__in)));
        vlSelfRef.t__DOT__delay_minmax[31U] = vlSelfRef.t__DOT__delay_minmax[30U];
        vlSelfRef.t__DOT__delay_minmax[30U] = vlSelfRef.t__DOT__delay_minmax[29U];
        vlSelfRef.t__DOT__delay_minmax[29U] = vlSelfRef.t__DOT__delay_minmax[28U];
        vlSelfRef.t__DOT__delay_minmax[28U] = vlSelfRef.t__DOT__delay_minmax[27U];
        vlSelfRef.t__DOT__delay_minmax[27U] = vlSelfRef.t__DOT__delay_minmax[26U];
        vlSelfRef.t__DOT__delay_minmax[26U] = vlSelfRef.t__DOT__delay_minmax[25U];
        vlSelfRef.t__DOT__delay_minmax[25U] = vlSelfRef.t__DOT__delay_minmax[24U];
        vlSelfRef.t__DOT__delay_minmax[24U] = vlSelfRef.t__DOT__delay_minmax[23U];
        vlSelfRef.t__DOT__delay_minmax[23U] = vlSelfRef.t__DOT__delay_minmax[22U];
        vlSelfRef.t__DOT__delay_minmax[22U] = vlSelfRef.t__DOT__delay_minmax[21U];
        vlSelfRef.t__DOT__delay_minmax[21U] = vlSelfRef.t__DOT__delay_minmax[20U];
        vlSelfRef.t__DOT__delay_minmax[20U] = vlSelfRef.t__DOT__delay_minmax[19U];
        vlSelfRef.t__DOT__delay_minmax[19U] = vlSelfRef.t__DOT__delay_minmax[18U];
        vlSelfRef.t__DOT__delay_minmax[18U] = vlSelfRef.t__DOT__delay_minmax[17U];
        vlSelfRef.t__DOT__delay_minmax[17U] = vlSelfRef.t__DOT__delay_minmax[16U];
        vlSelfRef.t__DOT__delay_minmax[16U] = vlSelfRef.t__DOT__delay_minmax[15U];
        vlSelfRef.t__DOT__delay_minmax[15U] = vlSelfRef.t__DOT__delay_minmax[14U];
        vlSelfRef.t__DOT__delay_minmax[14U] = vlSelfRef.t__DOT__delay_minmax[13U];
        vlSelfRef.t__DOT__delay_minmax[13U] = vlSelfRef.t__DOT__delay_minmax[12U];
        vlSelfRef.t__DOT__delay_minmax[12U] = vlSelfRef.t__DOT__delay_minmax[11U];
        vlSelfRef.t__DOT__delay_minmax[11U] = vlSelfRef.t__DOT__delay_minmax[10U];
        vlSelfRef.t__DOT__delay_minmax[10U] = vlSelfRef.t__DOT__delay_minmax[9U];
        vlSelfRef.t__DOT__delay_minmax[9U] = vlSelfRef.t__DOT__delay_minmax[8U];
        vlSelfRef.t__DOT__delay_minmax[8U] = vlSelfRef.t__DOT__delay_minmax[7U];
        vlSelfRef.t__DOT__delay_minmax[7U] = vlSelfRef.t__DOT__delay_minmax[6U];
        vlSelfRef.t__DOT__delay_minmax[6U] = vlSelfRef.t__DOT__delay_minmax[5U];
        vlSelfRef.t__DOT__delay_minmax[5U] = vlSelfRef.t__DOT__delay_minmax[4U];
        vlSelfRef.t__DOT__delay_minmax[4U] = vlSelfRef.t__DOT__delay_minmax[3U];
        vlSelfRef.t__DOT__delay_minmax[3U] = vlSelfRef.t__DOT__delay_minmax[2U];
        vlSelfRef.t__DOT__delay_minmax[2U] = vlSelfRef.t__DOT__delay_minmax[1U];
        vlSelfRef.t__DOT__delay_minmax[1U] = vlSelfRef.t__DOT__delay_minmax[0U];
        vlSelfRef.t__DOT__delay_minmax[0U] = (0x000000ffU 
                                              & ((IData)(vlSelfRef.t__DOT__in) 
                                                 >> 3U));
        t__DOT__min = 0x7fU;
        t__DOT__max = 0x80U;
        if (VL_LTS_III(8, 0x80U, vlSelfRef.t__DOT__delay_minmax[0U])) {
            t__DOT__max = vlSelfRef.t__DOT__delay_minmax[0U];
        }
        if (VL_GTS_III(8, 0x7fU, vlSelfRef.t__DOT__delay_minmax[0U])) {
            t__DOT__min = vlSelfRef.t__DOT__delay_minmax[0U];
        }
        if (VL_GTS_III(8, vlSelfRef.t__DOT__delay_minmax[1U], (IData)(t__DOT__max))) {
            t__DOT__max = vlSelfRef.t__DOT__delay_minmax[1U];
        }
        if (VL_LTS_III(8, vlSelfRef.t__DOT__delay_minmax[1U], (IData)(t__DOT__min))) {
            t__DOT__min = vlSelfRef.t__DOT__delay_minmax[1U];
        }
        if (VL_GTS_III(8, vlSelfRef.t__DOT__delay_minmax[2U], (IData)(t__DOT__max))) {
            t__DOT__max = vlSelfRef.t__DOT__delay_minmax[2U];
        }
        if (VL_LTS_III(8, vlSelfRef.t__DOT__delay_minmax[2U], (IData)(t__DOT__min))) {
            t__DOT__min = vlSelfRef.t__DOT__delay_minmax[2U];
        }
        if (VL_GTS_III(8, vlSelfRef.t__DOT__delay_minmax[3U], (IData)(t__DOT__max))) {
            t__DOT__max = vlSelfRef.t__DOT__delay_minmax[3U];
        }
        if (VL_LTS_III(8, vlSelfRef.t__DOT__delay_minmax[3U], (IData)(t__DOT__min))) {
            t__DOT__min = vlSelfRef.t__DOT__delay_minmax[3U];
        }
        if (VL_GTS_III(8, vlSelfRef.t__DOT__delay_minmax[4U], (IData)(t__DOT__max))) {
            t__DOT__max = vlSelfRef.t__DOT__delay_minmax[4U];
        }
        if (VL_LTS_III(8, vlSelfRef.t__DOT__delay_minmax[4U], (IData)(t__DOT__min))) {
            t__DOT__min = vlSelfRef.t__DOT__delay_minmax[4U];
        }
        if (VL_GTS_III(8, vlSelfRef.t__DOT__delay_minmax[5U], (IData)(t__DOT__max))) {
            t__DOT__max = vlSelfRef.t__DOT__delay_minmax[5U];
        }
        if (VL_LTS_III(8, vlSelfRef.t__DOT__delay_minmax[5U], (IData)(t__DOT__min))) {
            t__DOT__min = vlSelfRef.t__DOT__delay_minmax[5U];
        }
        if (VL_GTS_III(8, vlSelfRef.t__DOT__delay_minmax[6U], (IData)(t__DOT__max))) {
            t__DOT__max = vlSelfRef.t__DOT__delay_minmax[6U];
        }
        if (VL_LTS_III(8, vlSelfRef.t__DOT__delay_minmax[6U], (IData)(t__DOT__min))) {
            t__DOT__min = vlSelfRef.t__DOT__delay_minmax[6U];
        }
        if (VL_GTS_III(8, vlSelfRef.t__DOT__delay_minmax[7U], (IData)(t__DOT__max))) {
            t__DOT__max = vlSelfRef.t__DOT__delay_minmax[7U];
        }
        if (VL_LTS_III(8, vlSelfRef.t__DOT__delay_minmax[7U], (IData)(t__DOT__min))) {
            t__DOT__min = vlSelfRef.t__DOT__delay_minmax[7U];
        }
        if (VL_GTS_III(8, vlSelfRef.t__DOT__delay_minmax[8U], (IData)(t__DOT__max))) {
            t__DOT__max = vlSelfRef.t__DOT__delay_minmax[8U];
        }
        if (VL_LTS_III(8, vlSelfRef.t__DOT__delay_minmax[8U], (IData)(t__DOT__min))) {
            t__DOT__min = vlSelfRef.t__DOT__delay_minmax[8U];
        }
        if (VL_GTS_III(8, vlSelfRef.t__DOT__delay_minmax[9U], (IData)(t__DOT__max))) {
            t__DOT__max = vlSelfRef.t__DOT__delay_minmax[9U];
        }
        if (VL_LTS_III(8, vlSelfRef.t__DOT__delay_minmax[9U], (IData)(t__DOT__min))) {
            t__DOT__min = vlSelfRef.t__DOT__delay_minmax[9U];
        }
        if (VL_GTS_III(8, vlSelfRef.t__DOT__delay_minmax[10U], (IData)(t__DOT__max))) {
            t__DOT__max = vlSelfRef.t__DOT__delay_minmax[10U];
        }
        if (VL_LTS_III(8, vlSelfRef.t__DOT__delay_minmax[10U], (IData)(t__DOT__min))) {
            t__DOT__min = vlSelfRef.t__DOT__delay_minmax[10U];
        }
        if (VL_GTS_III(8, vlSelfRef.t__DOT__delay_minmax[11U], (IData)(t__DOT__max))) {
            t__DOT__max = vlSelfRef.t__DOT__delay_minmax[11U];
        }
        if (VL_LTS_III(8, vlSelfRef.t__DOT__delay_minmax[11U], (IData)(t__DOT__min))) {
            t__DOT__min = vlSelfRef.t__DOT__delay_minmax[11U];
        }
        if (VL_GTS_III(8, vlSelfRef.t__DOT__delay_minmax[12U], (IData)(t__DOT__max))) {
            t__DOT__max = vlSelfRef.t__DOT__delay_minmax[12U];
        }
        if (VL_LTS_III(8, vlSelfRef.t__DOT__delay_minmax[12U], (IData)(t__DOT__min))) {
            t__DOT__min = vlSelfRef.t__DOT__delay_minmax[12U];
        }
        if (VL_GTS_III(8, vlSelfRef.t__DOT__delay_minmax[13U], (IData)(t__DOT__max))) {
            t__DOT__max = vlSelfRef.t__DOT__delay_minmax[13U];
        }
        if (VL_LTS_III(8, vlSelfRef.t__DOT__delay_minmax[13U], (IData)(t__DOT__min))) {
            t__DOT__min = vlSelfRef.t__DOT__delay_minmax[13U];
        }
        if (VL_GTS_III(8, vlSelfRef.t__DOT__delay_minmax[14U], (IData)(t__DOT__max))) {
            t__DOT__max = vlSelfRef.t__DOT__delay_minmax[14U];
        }
        if (VL_LTS_III(8, vlSelfRef.t__DOT__delay_minmax[14U], (IData)(t__DOT__min))) {
            t__DOT__min = vlSelfRef.t__DOT__delay_minmax[14U];
        }
        if (VL_GTS_III(8, vlSelfRef.t__DOT__delay_minmax[15U], (IData)(t__DOT__max))) {
            t__DOT__max = vlSelfRef.t__DOT__delay_minmax[15U];
        }
        if (VL_LTS_III(8, vlSelfRef.t__DOT__delay_minmax[15U], (IData)(t__DOT__min))) {
            t__DOT__min = vlSelfRef.t__DOT__delay_minmax[15U];
        }
        if (VL_GTS_III(8, vlSelfRef.t__DOT__delay_minmax[16U], (IData)(t__DOT__max))) {
            t__DOT__max = vlSelfRef.t__DOT__delay_minmax[16U];
        }
        if (VL_LTS_III(8, vlSelfRef.t__DOT__delay_minmax[16U], (IData)(t__DOT__min))) {
            t__DOT__min = vlSelfRef.t__DOT__delay_minmax[16U];
        }
        if (VL_GTS_III(8, vlSelfRef.t__DOT__delay_minmax[17U], (IData)(t__DOT__max))) {
            t__DOT__max = vlSelfRef.t__DOT__delay_minmax[17U];
        }
        if (VL_LTS_III(8, vlSelfRef.t__DOT__delay_minmax[17U], (IData)(t__DOT__min))) {
            t__DOT__min = vlSelfRef.t__DOT__delay_minmax[17U];
        }
        if (VL_GTS_III(8, vlSelfRef.t__DOT__delay_minmax[18U], (IData)(t__DOT__max))) {
            t__DOT__max = vlSelfRef.t__DOT__delay_minmax[18U];
        }
        if (VL_LTS_III(8, vlSelfRef.t__DOT__delay_minmax[18U], (IData)(t__DOT__min))) {
            t__DOT__min = vlSelfRef.t__DOT__delay_minmax[18U];
        }
        if (VL_GTS_III(8, vlSelfRef.t__DOT__delay_minmax[19U], (IData)(t__DOT__max))) {
            t__DOT__max = vlSelfRef.t__DOT__delay_minmax[19U];
        }
        if (VL_LTS_III(8, vlSelfRef.t__DOT__delay_minmax[19U], (IData)(t__DOT__min))) {
            t__DOT__min = vlSelfRef.t__DOT__delay_minmax[19U];
        }
        if (VL_GTS_III(8, vlSelfRef.t__DOT__delay_minmax[20U], (IData)(t__DOT__max))) {
            t__DOT__max = vlSelfRef.t__DOT__delay_minmax[20U];
        }
        if (VL_LTS_III(8, vlSelfRef.t__DOT__delay_minmax[20U], (IData)(t__DOT__min))) {
            t__DOT__min = vlSelfRef.t__DOT__delay_minmax[20U];
        }
        if (VL_GTS_III(8, vlSelfRef.t__DOT__delay_minmax[21U], (IData)(t__DOT__max))) {
            t__DOT__max = vlSelfRef.t__DOT__delay_minmax[21U];
        }
        if (VL_LTS_III(8, vlSelfRef.t__DOT__delay_minmax[21U], (IData)(t__DOT__min))) {
            t__DOT__min = vlSelfRef.t__DOT__delay_minmax[21U];
        }
        if (VL_GTS_III(8, vlSelfRef.t__DOT__delay_minmax[22U], (IData)(t__DOT__max))) {
            t__DOT__max = vlSelfRef.t__DOT__delay_minmax[22U];
        }
        if (VL_LTS_III(8, vlSelfRef.t__DOT__delay_minmax[22U], (IData)(t__DOT__min))) {
            t__DOT__min = vlSelfRef.t__DOT__delay_minmax[22U];
        }
        if (VL_GTS_III(8, vlSelfRef.t__DOT__delay_minmax[23U], (IData)(t__DOT__max))) {
            t__DOT__max = vlSelfRef.t__DOT__delay_minmax[23U];
        }
        if (VL_LTS_III(8, vlSelfRef.t__DOT__delay_minmax[23U], (IData)(t__DOT__min))) {
            t__DOT__min = vlSelfRef.t__DOT__delay_minmax[23U];
        }
        if (VL_GTS_III(8, vlSelfRef.t__DOT__delay_minmax[24U], (IData)(t__DOT__max))) {
            t__DOT__max = vlSelfRef.t__DOT__delay_minmax[24U];
        }
        if (VL_LTS_III(8, vlSelfRef.t__DOT__delay_minmax[24U], (IData)(t__DOT__min))) {
            t__DOT__min = vlSelfRef.t__DOT__delay_minmax[24U];
        }
        if (VL_GTS_III(8, vlSelfRef.t__DOT__delay_minmax[25U], (IData)(t__DOT__max))) {
            t__DOT__max = vlSelfRef.t__DOT__delay_minmax[25U];
        }
        if (VL_LTS_III(8, vlSelfRef.t__DOT__delay_minmax[25U], (IData)(t__DOT__min))) {
            t__DOT__min = vlSelfRef.t__DOT__delay_minmax[25U];
        }
        if (VL_GTS_III(8, vlSelfRef.t__DOT__delay_minmax[26U], (IData)(t__DOT__max))) {
            t__DOT__max = vlSelfRef.t__DOT__delay_minmax[26U];
        }
        if (VL_LTS_III(8, vlSelfRef.t__DOT__delay_minmax[26U], (IData)(t__DOT__min))) {
            t__DOT__min = vlSelfRef.t__DOT__delay_minmax[26U];
        }
        if (VL_GTS_III(8, vlSelfRef.t__DOT__delay_minmax[27U], (IData)(t__DOT__max))) {
            t__DOT__max = vlSelfRef.t__DOT__delay_minmax[27U];
        }
        if (VL_LTS_III(8, vlSelfRef.t__DOT__delay_minmax[27U], (IData)(t__DOT__min))) {
            t__DOT__min = vlSelfRef.t__DOT__delay_minmax[27U];
        }
        if (VL_GTS_III(8, vlSelfRef.t__DOT__delay_minmax[28U], (IData)(t__DOT__max))) {
            t__DOT__max = vlSelfRef.t__DOT__delay_minmax[28U];
        }
        if (VL_LTS_III(8, vlSelfRef.t__DOT__delay_minmax[28U], (IData)(t__DOT__min))) {
            t__DOT__min = vlSelfRef.t__DOT__delay_minmax[28U];
        }
        if (VL_GTS_III(8, vlSelfRef.t__DOT__delay_minmax[29U], (IData)(t__DOT__max))) {
            t__DOT__max = vlSelfRef.t__DOT__delay_minmax[29U];
        }
        if (VL_LTS_III(8, vlSelfRef.t__DOT__delay_minmax[29U], (IData)(t__DOT__min))) {
            t__DOT__min = vlSelfRef.t__DOT__delay_minmax[29U];
        }
        if (VL_GTS_III(8, vlSelfRef.t__DOT__delay_minmax[30U], (IData)(t__DOT__max))) {
            t__DOT__max = vlSelfRef.t__DOT__delay_minmax[30U];
        }
        if (VL_LTS_III(8, vlSelfRef.t__DOT__delay_minmax[30U], (IData)(t__DOT__min))) {
            t__DOT__min = vlSelfRef.t__DOT__delay_minmax[30U];
        }
        if (VL_GTS_III(8, vlSelfRef.t__DOT__delay_minmax[31U], (IData)(t__DOT__max))) {
            t__DOT__max = vlSelfRef.t__DOT__delay_minmax[31U];
        }
        if (VL_LTS_III(8, vlSelfRef.t__DOT__delay_minmax[31U], (IData)(t__DOT__min))) {
            t__DOT__min = vlSelfRef.t__DOT__delay_minmax[31U];
        }
        if (VL_UNLIKELY(((0x0000007fU == VL_EXTENDS_II(32,8, (IData)(t__DOT__min)))))) {
            VL_STOP_MT("/root/reference/test_regress/t/t_math_signed2.v", 60, "");
        } else if (VL_UNLIKELY((VL_LTES_III(32, 0xffffffc3U, 
                                            VL_EXTENDS_II(32,8, 
                                                          (0x000000ffU 
                                                           & ((IData)(vlSelfRef.t__DOT__in) 
                                                              >> 3U))))))) {
            vlCtx.writef(0U, 0U);
            VL_FINISH_MT("/root/reference/test_regress/t/t_math_signed2.v", 64, "");
        }
    }

    VLCU_D static void Vt_math_signed2___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt_math_signed2___024root___eval_phase__act(Vt_math_signed2___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt_math_signed2___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vt_math_signed2___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vt_math_signed2___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vt_math_signed2___024root___eval_phase__nba(Vt_math_signed2___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vt_math_signed2___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    Vt_math_signed2___024root___nba_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
            Vt_math_signed2___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vt_math_signed2___024root___eval(Vt_math_signed2___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt_math_signed2___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("/root/reference/test_regress/t/t_math_signed2.v", 7, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vt_math_signed2___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("/root/reference/test_regress/t/t_math_signed2.v", 7, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vt_math_signed2___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vt_math_signed2___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vt_math_signed2___024root___ctor_var_reset(Vt_math_signed2___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        vlSelf->t__DOT__in = VL_SCOPED_RAND_RESET_I(11, __VscopeHash, 18129128471828695690ull);
        for (int __Vi0 = 0; __Vi0 < 32; ++__Vi0) {
            vlSelf->t__DOT__delay_minmax[__Vi0] = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 4334265529568846339ull);
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
        Vt_math_signed2___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_signed2___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_signed2___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_signed2___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_signed2___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_signed2___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt_math_signed2__cuda::info() {
    typedef Vt_math_signed2__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"t__DOT__in", "t.in", VLCU_MEMBER_STATE, 2, 0, 1, 11, 10, 0, T::OFF_t__DOT__in},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"t__DOT__delay_minmax", "t.delay_minmax", VLCU_MEMBER_STATE, 1, 0, 32, 8, 7, 0, T::OFF_t__DOT__delay_minmax},
        {"__VactTriggered", "__VactTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VactTriggered},
        {"__VnbaTriggered", "__VnbaTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VnbaTriggered},
    };
    static const char* const messages[] = {"*-* All Finished *-*\n", nullptr};
    static const char* const plusargs[] = {nullptr};
    static const VlcuTraceDecl traceDecls[] = {
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0}};  // (terminator, not counted)
    static const uint32_t traceOrder[] = {0};  // codes in dump order (+ terminator)
    static const uint32_t saveOrder[] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "t", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 1, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 9, 0, 0xbccfa4fb21748de3ULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt_math_signed2__cuda VLCU_TESTBENCHES);
