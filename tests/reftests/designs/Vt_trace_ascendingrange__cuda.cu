// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt_trace_ascendingrange"
#endif

// Constant pool
static __device__ const VlWide<9>/*287:0*/ Vt_trace_ascendingrange__ConstPool__CONST_h3dd8fb2c_0 = VlWide<9>{{
        0x00000000, 0x00000000, 0x00000000, 0x00000000,
        0x00000000, 0x00000000, 0x00000000, 0x00000000,
        0x00000000
}};
static __device__ const VlWide<8>/*255:0*/ Vt_trace_ascendingrange__ConstPool__CONST_h5b979007_0 = VlWide<8>{{
        0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff,
        0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff
}};
static __device__ const VlWide<8>/*255:0*/ Vt_trace_ascendingrange__ConstPool__CONST_h6ad97645_0 = VlWide<8>{{
        0x00000000, 0x00000000, 0x00000000, 0x00000000,
        0x00000000, 0x00000000, 0x00000000, 0x00000000
}};
static __device__ const VlWide<16>/*511:0*/ Vt_trace_ascendingrange__ConstPool__CONST_h6d27141e_0 = VlWide<16>{{
        0x00000000, 0x00000000, 0x00000000, 0x00000000,
        0x00000000, 0x00000000, 0x00000000, 0x00000000,
        0x00000000, 0x00000000, 0x00000000, 0x00000000,
        0x00000000, 0x00000000, 0x00000000, 0x00000000
}};
static __device__ const VlWide<9>/*287:0*/ Vt_trace_ascendingrange__ConstPool__CONST_h7d18c5a8_0 = VlWide<9>{{
        0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff,
        0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff,
        0x00000001
}};
static __device__ const VlWide<16>/*511:0*/ Vt_trace_ascendingrange__ConstPool__CONST_h8b2d9f06_0 = VlWide<16>{{
        0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff,
        0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff,
        0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff,
        0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff
}};
static __device__ const VlWide<16>/*511:0*/ Vt_trace_ascendingrange__ConstPool__CONST_h936cb6fc_0 = VlWide<16>{{
        0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff,
        0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff,
        0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff,
        0xffffffff, 0xffffffff, 0xffffffff, 0x7fffffff
}};
static __device__ const VlWide<16>/*511:0*/ Vt_trace_ascendingrange__ConstPool__CONST_h93e1b771_0 = VlWide<16>{{
        0x00000000, 0x00000000, 0x00000000, 0x00000000,
        0x00000000, 0x00000000, 0x00000000, 0x00000000,
        0x00000000, 0x00000000, 0x00000000, 0x00000000,
        0x00000000, 0x00000000, 0x00000000, 0x00000000
}};
static __device__ const VlWide<8>/*255:0*/ Vt_trace_ascendingrange__ConstPool__CONST_h9e67c271_0 = VlWide<8>{{
        0x00000000, 0x00000000, 0x00000000, 0x00000000,
        0x00000000, 0x00000000, 0x00000000, 0x00000000
}};
static __device__ const VlWide<8>/*255:0*/ Vt_trace_ascendingrange__ConstPool__CONST_ha837507f_0 = VlWide<8>{{
        0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff,
        0xffffffff, 0xffffffff, 0xffffffff, 0x7fffffff
}};

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt_trace_ascendingrange___024root final {
    CData/*0:0*/ clk;
    CData/*1:0*/ t__DOT__v_b;
    CData/*7:0*/ t__DOT__v_c;
    CData/*2:0*/ t__DOT__v_p;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    SData/*8:0*/ t__DOT__v_d;
    SData/*15:0*/ t__DOT__v_e;
    SData/*14:0*/ t__DOT__v_q;
    IData/*31:0*/ t__DOT__cyc;
    IData/*16:0*/ t__DOT__v_f;
    IData/*31:0*/ t__DOT__v_g;
    VlWide<3>/*64:0*/ t__DOT__v_j;
    VlWide<4>/*127:0*/ t__DOT__v_k;
    VlWide<5>/*128:0*/ t__DOT__v_l;
    VlWide<8>/*255:0*/ t__DOT__v_m;
    VlWide<9>/*256:0*/ t__DOT__v_n;
    VlWide<16>/*511:0*/ t__DOT__v_o;
    IData/*30:0*/ t__DOT__v_r;
    VlWide<4>/*126:0*/ t__DOT__v_t;
    VlWide<8>/*254:0*/ t__DOT__v_u;
    VlWide<16>/*510:0*/ t__DOT__v_v;
    IData/*31:0*/ __VactIterCount;
    QData/*32:0*/ t__DOT__v_h;
    QData/*63:0*/ t__DOT__v_i;
    QData/*62:0*/ t__DOT__v_s;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vt_trace_ascendingrange__cuda final {
    typedef Vt_trace_ascendingrange___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 358,
        OFF_t__DOT__v_b = 359,
        OFF_t__DOT__v_c = 360,
        OFF_t__DOT__v_p = 361,
        OFF___Vtrigprevexpr___TOP__clk__0 = 362,
        OFF___VactPhaseResult = 363,
        OFF___VnbaPhaseResult = 364,
        OFF_t__DOT__v_d = 352,
        OFF_t__DOT__v_e = 354,
        OFF_t__DOT__v_q = 356,
        OFF_t__DOT__cyc = 40,
        OFF_t__DOT__v_f = 44,
        OFF_t__DOT__v_g = 48,
        OFF_t__DOT__v_j = 52,
        OFF_t__DOT__v_k = 64,
        OFF_t__DOT__v_l = 80,
        OFF_t__DOT__v_m = 100,
        OFF_t__DOT__v_n = 132,
        OFF_t__DOT__v_o = 168,
        OFF_t__DOT__v_r = 232,
        OFF_t__DOT__v_t = 236,
        OFF_t__DOT__v_u = 252,
        OFF_t__DOT__v_v = 284,
        OFF___VactIterCount = 348,
        OFF_t__DOT__v_h = 0,
        OFF_t__DOT__v_i = 8,
        OFF_t__DOT__v_s = 16,
        OFF___VactTriggered = 24,
        OFF___VnbaTriggered = 32,
        kStateBytes = 368,  // S_persist
        kTouchedReadBytes = 365,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 364,  // S_state_wr
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
        s.t__DOT__v_b = st.ld<VlcuMem8>(OFF_t__DOT__v_b, i);
        s.t__DOT__v_c = st.ld<VlcuMem8>(OFF_t__DOT__v_c, i);
        s.t__DOT__v_p = st.ld<VlcuMem8>(OFF_t__DOT__v_p, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        s.t__DOT__v_d = st.ld<VlcuMem16>(OFF_t__DOT__v_d, i);
        s.t__DOT__v_e = st.ld<VlcuMem16>(OFF_t__DOT__v_e, i);
        s.t__DOT__v_q = st.ld<VlcuMem16>(OFF_t__DOT__v_q, i);
        s.t__DOT__cyc = st.ld<IData>(OFF_t__DOT__cyc, i);
        s.t__DOT__v_f = st.ld<IData>(OFF_t__DOT__v_f, i);
        s.t__DOT__v_g = st.ld<IData>(OFF_t__DOT__v_g, i);
        st.ldW(OFF_t__DOT__v_j, i, s.t__DOT__v_j);
        st.ldW(OFF_t__DOT__v_k, i, s.t__DOT__v_k);
        st.ldW(OFF_t__DOT__v_l, i, s.t__DOT__v_l);
        st.ldW(OFF_t__DOT__v_m, i, s.t__DOT__v_m);
        st.ldW(OFF_t__DOT__v_n, i, s.t__DOT__v_n);
        st.ldW(OFF_t__DOT__v_o, i, s.t__DOT__v_o);
        s.t__DOT__v_r = st.ld<IData>(OFF_t__DOT__v_r, i);
        st.ldW(OFF_t__DOT__v_t, i, s.t__DOT__v_t);
        st.ldW(OFF_t__DOT__v_u, i, s.t__DOT__v_u);
        st.ldW(OFF_t__DOT__v_v, i, s.t__DOT__v_v);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        s.t__DOT__v_h = st.ld<QData>(OFF_t__DOT__v_h, i);
        s.t__DOT__v_i = st.ld<QData>(OFF_t__DOT__v_i, i);
        s.t__DOT__v_s = st.ld<QData>(OFF_t__DOT__v_s, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.st<VlcuMem8>(OFF_t__DOT__v_b, i, s.t__DOT__v_b);
        st.st<VlcuMem8>(OFF_t__DOT__v_c, i, s.t__DOT__v_c);
        st.st<VlcuMem8>(OFF_t__DOT__v_p, i, s.t__DOT__v_p);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.st<VlcuMem16>(OFF_t__DOT__v_d, i, s.t__DOT__v_d);
        st.st<VlcuMem16>(OFF_t__DOT__v_e, i, s.t__DOT__v_e);
        st.st<VlcuMem16>(OFF_t__DOT__v_q, i, s.t__DOT__v_q);
        st.st<IData>(OFF_t__DOT__cyc, i, s.t__DOT__cyc);
        st.st<IData>(OFF_t__DOT__v_f, i, s.t__DOT__v_f);
        st.st<IData>(OFF_t__DOT__v_g, i, s.t__DOT__v_g);
        st.stW(OFF_t__DOT__v_j, i, s.t__DOT__v_j);
        st.stW(OFF_t__DOT__v_k, i, s.t__DOT__v_k);
        st.stW(OFF_t__DOT__v_l, i, s.t__DOT__v_l);
        st.stW(OFF_t__DOT__v_m, i, s.t__DOT__v_m);
        st.stW(OFF_t__DOT__v_n, i, s.t__DOT__v_n);
        st.stW(OFF_t__DOT__v_o, i, s.t__DOT__v_o);
        st.st<IData>(OFF_t__DOT__v_r, i, s.t__DOT__v_r);
        st.stW(OFF_t__DOT__v_t, i, s.t__DOT__v_t);
        st.stW(OFF_t__DOT__v_u, i, s.t__DOT__v_u);
        st.stW(OFF_t__DOT__v_v, i, s.t__DOT__v_v);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
        st.st<QData>(OFF_t__DOT__v_h, i, s.t__DOT__v_h);
        st.st<QData>(OFF_t__DOT__v_i, i, s.t__DOT__v_i);
        st.st<QData>(OFF_t__DOT__v_s, i, s.t__DOT__v_s);
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

    VLCU_D static void Vt_trace_ascendingrange___024root___eval_static(Vt_trace_ascendingrange___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        {
            // Inlined CFunc: _eval_static__TOP
            vlSelfRef.t__DOT__cyc = 0U;
            vlSelfRef.t__DOT__v_b = 0U;
            vlSelfRef.t__DOT__v_c = 0U;
            vlSelfRef.t__DOT__v_d = 0U;
            vlSelfRef.t__DOT__v_e = 0U;
            vlSelfRef.t__DOT__v_f = 0U;
            vlSelfRef.t__DOT__v_g = 0U;
            vlSelfRef.t__DOT__v_h = 0ULL;
            vlSelfRef.t__DOT__v_i = 0ULL;
            vlSelfRef.t__DOT__v_j[0U] = 0U;
            vlSelfRef.t__DOT__v_j[1U] = 0U;
            vlSelfRef.t__DOT__v_j[2U] = 0U;
            vlSelfRef.t__DOT__v_k[0U] = 0U;
            vlSelfRef.t__DOT__v_k[1U] = 0U;
            vlSelfRef.t__DOT__v_k[2U] = 0U;
            vlSelfRef.t__DOT__v_k[3U] = 0U;
            vlSelfRef.t__DOT__v_l[0U] = 0U;
            vlSelfRef.t__DOT__v_l[1U] = 0U;
            vlSelfRef.t__DOT__v_l[2U] = 0U;
            vlSelfRef.t__DOT__v_l[3U] = 0U;
            vlSelfRef.t__DOT__v_l[4U] = 0U;
            VL_ASSIGN_W(256, vlSelfRef.t__DOT__v_m, Vt_trace_ascendingrange__ConstPool__CONST_h9e67c271_0);
            VL_ASSIGN_W(257, vlSelfRef.t__DOT__v_n, Vt_trace_ascendingrange__ConstPool__CONST_h3dd8fb2c_0);
            VL_ASSIGN_W(512, vlSelfRef.t__DOT__v_o, Vt_trace_ascendingrange__ConstPool__CONST_h93e1b771_0);
            vlSelfRef.t__DOT__v_p = 0U;
            vlSelfRef.t__DOT__v_q = 0U;
            vlSelfRef.t__DOT__v_r = 0U;
            vlSelfRef.t__DOT__v_s = 0ULL;
            vlSelfRef.t__DOT__v_t[0U] = 0U;
            vlSelfRef.t__DOT__v_t[1U] = 0U;
            vlSelfRef.t__DOT__v_t[2U] = 0U;
            vlSelfRef.t__DOT__v_t[3U] = 0U;
            VL_ASSIGN_W(255, vlSelfRef.t__DOT__v_u, Vt_trace_ascendingrange__ConstPool__CONST_h6ad97645_0);
            VL_ASSIGN_W(511, vlSelfRef.t__DOT__v_v, Vt_trace_ascendingrange__ConstPool__CONST_h6d27141e_0);
        }
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
    }

    VLCU_D static void Vt_trace_ascendingrange___024root___eval_initial(Vt_trace_ascendingrange___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_trace_ascendingrange___024root___eval_final(Vt_trace_ascendingrange___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_trace_ascendingrange___024root___eval_settle(Vt_trace_ascendingrange___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static bool Vt_trace_ascendingrange___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt_trace_ascendingrange___024root___nba_sequent__TOP__0(Vt_trace_ascendingrange___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*1:0*/ __Vdly__t__DOT__v_b;
        __Vdly__t__DOT__v_b = 0;
        CData/*7:0*/ __Vdly__t__DOT__v_c;
        __Vdly__t__DOT__v_c = 0;
        SData/*8:0*/ __Vdly__t__DOT__v_d;
        __Vdly__t__DOT__v_d = 0;
        SData/*15:0*/ __Vdly__t__DOT__v_e;
        __Vdly__t__DOT__v_e = 0;
        IData/*16:0*/ __Vdly__t__DOT__v_f;
        __Vdly__t__DOT__v_f = 0;
        IData/*31:0*/ __Vdly__t__DOT__v_g;
        __Vdly__t__DOT__v_g = 0;
        QData/*32:0*/ __Vdly__t__DOT__v_h;
        __Vdly__t__DOT__v_h = 0;
        QData/*63:0*/ __Vdly__t__DOT__v_i;
        __Vdly__t__DOT__v_i = 0;
        VlWide<3>/*64:0*/ __Vdly__t__DOT__v_j;
        VL_ZERO_W(65, __Vdly__t__DOT__v_j);
        VlWide<4>/*127:0*/ __Vdly__t__DOT__v_k;
        VL_ZERO_W(128, __Vdly__t__DOT__v_k);
        VlWide<5>/*128:0*/ __Vdly__t__DOT__v_l;
        VL_ZERO_W(129, __Vdly__t__DOT__v_l);
        VlWide<8>/*255:0*/ __Vdly__t__DOT__v_m;
        VL_ZERO_W(256, __Vdly__t__DOT__v_m);
        VlWide<9>/*256:0*/ __Vdly__t__DOT__v_n;
        VL_ZERO_W(257, __Vdly__t__DOT__v_n);
        VlWide<16>/*511:0*/ __Vdly__t__DOT__v_o;
        VL_ZERO_W(512, __Vdly__t__DOT__v_o);
        CData/*2:0*/ __Vdly__t__DOT__v_p;
        __Vdly__t__DOT__v_p = 0;
        SData/*14:0*/ __Vdly__t__DOT__v_q;
        __Vdly__t__DOT__v_q = 0;
        IData/*30:0*/ __Vdly__t__DOT__v_r;
        __Vdly__t__DOT__v_r = 0;
        QData/*62:0*/ __Vdly__t__DOT__v_s;
        __Vdly__t__DOT__v_s = 0;
        VlWide<4>/*126:0*/ __Vdly__t__DOT__v_t;
        VL_ZERO_W(127, __Vdly__t__DOT__v_t);
        VlWide<8>/*254:0*/ __Vdly__t__DOT__v_u;
        VL_ZERO_W(255, __Vdly__t__DOT__v_u);
        VlWide<16>/*510:0*/ __Vdly__t__DOT__v_v;
        VL_ZERO_W(511, __Vdly__t__DOT__v_v);
        IData/*31:0*/ __Vdly__t__DOT__cyc;
        __Vdly__t__DOT__cyc = 0;
        VlWide<3>/*95:0*/ __Vtemp_2;
        VlWide<5>/*159:0*/ __Vtemp_4;
        VlWide<9>/*287:0*/ __Vtemp_5;
        VlWide<4>/*127:0*/ __Vtemp_7;
        VlWide<8>/*255:0*/ __Vtemp_8;
        VlWide<16>/*511:0*/ __Vtemp_9;
        // Body
        __Vdly__t__DOT__v_b = vlSelfRef.t__DOT__v_b;
        __Vdly__t__DOT__v_c = vlSelfRef.t__DOT__v_c;
        __Vdly__t__DOT__v_d = vlSelfRef.t__DOT__v_d;
        __Vdly__t__DOT__v_e = vlSelfRef.t__DOT__v_e;
        __Vdly__t__DOT__v_f = vlSelfRef.t__DOT__v_f;
        __Vdly__t__DOT__v_g = vlSelfRef.t__DOT__v_g;
        __Vdly__t__DOT__v_h = vlSelfRef.t__DOT__v_h;
        __Vdly__t__DOT__v_i = vlSelfRef.t__DOT__v_i;
        __Vdly__t__DOT__v_j[0U] = vlSelfRef.t__DOT__v_j[0U];
        __Vdly__t__DOT__v_j[1U] = vlSelfRef.t__DOT__v_j[1U];
        __Vdly__t__DOT__v_j[2U] = vlSelfRef.t__DOT__v_j[2U];
        __Vdly__t__DOT__v_k[0U] = vlSelfRef.t__DOT__v_k[0U];
        __Vdly__t__DOT__v_k[1U] = vlSelfRef.t__DOT__v_k[1U];
        __Vdly__t__DOT__v_k[2U] = vlSelfRef.t__DOT__v_k[2U];
        __Vdly__t__DOT__v_k[3U] = vlSelfRef.t__DOT__v_k[3U];
        __Vdly__t__DOT__v_l[0U] = vlSelfRef.t__DOT__v_l[0U];
        __Vdly__t__DOT__v_l[1U] = vlSelfRef.t__DOT__v_l[1U];
        __Vdly__t__DOT__v_l[2U] = vlSelfRef.t__DOT__v_l[2U];
        __Vdly__t__DOT__v_l[3U] = vlSelfRef.t__DOT__v_l[3U];
        __Vdly__t__DOT__v_l[4U] = vlSelfRef.t__DOT__v_l[4U];
        __Vdly__t__DOT__v_m[0U] = vlSelfRef.t__DOT__v_m[0U];
        __Vdly__t__DOT__v_m[1U] = vlSelfRef.t__DOT__v_m[1U];
        __Vdly__t__DOT__v_m[2U] = vlSelfRef.t__DOT__v_m[2U];
        __Vdly__t__DOT__v_m[3U] = vlSelfRef.t__DOT__v_m[3U];
        __Vdly__t__DOT__v_m[4U] = vlSelfRef.t__DOT__v_m[4U];
        __Vdly__t__DOT__v_m[5U] = vlSelfRef.t__DOT__v_m[5U];
        __Vdly__t__DOT__v_m[6U] = vlSelfRef.t__DOT__v_m[6U];
        __Vdly__t__DOT__v_m[7U] = vlSelfRef.t__DOT__v_m[7U];
        __Vdly__t__DOT__v_n[0U] = vlSelfRef.t__DOT__v_n[0U];
        __Vdly__t__DOT__v_n[1U] = vlSelfRef.t__DOT__v_n[1U];
        __Vdly__t__DOT__v_n[2U] = vlSelfRef.t__DOT__v_n[2U];
        __Vdly__t__DOT__v_n[3U] = vlSelfRef.t__DOT__v_n[3U];
        __Vdly__t__DOT__v_n[4U] = vlSelfRef.t__DOT__v_n[4U];
        __Vdly__t__DOT__v_n[5U] = vlSelfRef.t__DOT__v_n[5U];
        __Vdly__t__DOT__v_n[6U] = vlSelfRef.t__DOT__v_n[6U];
        __Vdly__t__DOT__v_n[7U] = vlSelfRef.t__DOT__v_n[7U];
        __Vdly__t__DOT__v_n[8U] = vlSelfRef.t__DOT__v_n[8U];
        __Vdly__t__DOT__v_o[0U] = vlSelfRef.t__DOT__v_o[0U];
        __Vdly__t__DOT__v_o[1U] = vlSelfRef.t__DOT__v_o[1U];
        __Vdly__t__DOT__v_o[2U] = vlSelfRef.t__DOT__v_o[2U];
        __Vdly__t__DOT__v_o[3U] = vlSelfRef.t__DOT__v_o[3U];
        __Vdly__t__DOT__v_o[4U] = vlSelfRef.t__DOT__v_o[4U];
        __Vdly__t__DOT__v_o[5U] = vlSelfRef.t__DOT__v_o[5U];
        __Vdly__t__DOT__v_o[6U] = vlSelfRef.t__DOT__v_o[6U];
        __Vdly__t__DOT__v_o[7U] = vlSelfRef.t__DOT__v_o[7U];
        __Vdly__t__DOT__v_o[8U] = vlSelfRef.t__DOT__v_o[8U];
        __Vdly__t__DOT__v_o[9U] = vlSelfRef.t__DOT__v_o[9U];
        __Vdly__t__DOT__v_o[10U] = vlSelfRef.t__DOT__v_o[10U];
        __Vdly__t__DOT__v_o[11U] = vlSelfRef.t__DOT__v_o[11U];
        __Vdly__t__DOT__v_o[12U] = vlSelfRef.t__DOT__v_o[12U];
        __Vdly__t__DOT__v_o[13U] = vlSelfRef.t__DOT__v_o[13U];
        __Vdly__t__DOT__v_o[14U] = vlSelfRef.t__DOT__v_o[14U];
        __Vdly__t__DOT__v_o[15U] = vlSelfRef.t__DOT__v_o[15U];
        __Vdly__t__DOT__v_p = vlSelfRef.t__DOT__v_p;
        __Vdly__t__DOT__v_q = vlSelfRef.t__DOT__v_q;
        __Vdly__t__DOT__v_r = vlSelfRef.t__DOT__v_r;
        __Vdly__t__DOT__v_s = vlSelfRef.t__DOT__v_s;
        __Vdly__t__DOT__v_t[0U] = vlSelfRef.t__DOT__v_t[0U];
        __Vdly__t__DOT__v_t[1U] = vlSelfRef.t__DOT__v_t[1U];
        __Vdly__t__DOT__v_t[2U] = vlSelfRef.t__DOT__v_t[2U];
        __Vdly__t__DOT__v_t[3U] = vlSelfRef.t__DOT__v_t[3U];
        __Vdly__t__DOT__v_u[0U] = vlSelfRef.t__DOT__v_u[0U];
        __Vdly__t__DOT__v_u[1U] = vlSelfRef.t__DOT__v_u[1U];
        __Vdly__t__DOT__v_u[2U] = vlSelfRef.t__DOT__v_u[2U];
        __Vdly__t__DOT__v_u[3U] = vlSelfRef.t__DOT__v_u[3U];
        __Vdly__t__DOT__v_u[4U] = vlSelfRef.t__DOT__v_u[4U];
        __Vdly__t__DOT__v_u[5U] = vlSelfRef.t__DOT__v_u[5U];
        __Vdly__t__DOT__v_u[6U] = vlSelfRef.t__DOT__v_u[6U];
        __Vdly__t__DOT__v_u[7U] = vlSelfRef.t__DOT__v_u[7U];
        __Vdly__t__DOT__v_v[0U] = vlSelfRef.t__DOT__v_v[0U];
        __Vdly__t__DOT__v_v[1U] = vlSelfRef.t__DOT__v_v[1U];
        __Vdly__t__DOT__v_v[2U] = vlSelfRef.t__DOT__v_v[2U];
        __Vdly__t__DOT__v_v[3U] = vlSelfRef.t__DOT__v_v[3U];
        __Vdly__t__DOT__v_v[4U] = vlSelfRef.t__DOT__v_v[4U];
        __Vdly__t__DOT__v_v[5U] = vlSelfRef.t__DOT__v_v[5U];
        __Vdly__t__DOT__v_v[6U] = vlSelfRef.t__DOT__v_v[6U];
        __Vdly__t__DOT__v_v[7U] = vlSelfRef.t__DOT__v_v[7U];
        __Vdly__t__DOT__v_v[8U] = vlSelfRef.t__DOT__v_v[8U];
        __Vdly__t__DOT__v_v[9U] = vlSelfRef.t__DOT__v_v[9U];
        __Vdly__t__DOT__v_v[10U] = vlSelfRef.t__DOT__v_v[10U];
        __Vdly__t__DOT__v_v[11U] = vlSelfRef.t__DOT__v_v[11U];
        __Vdly__t__DOT__v_v[12U] = vlSelfRef.t__DOT__v_v[12U];
        __Vdly__t__DOT__v_v[13U] = vlSelfRef.t__DOT__v_v[13U];
        __Vdly__t__DOT__v_v[14U] = vlSelfRef.t__DOT__v_v[14U];
        __Vdly__t__DOT__v_v[15U] = vlSelfRef.t__DOT__v_v[15U];
        __Vdly__t__DOT__cyc = vlSelfRef.t__DOT__cyc;
        if ((0U == vlSelfRef.t__DOT__cyc)) {
            __Vdly__t__DOT__v_b = 3U;
            __Vdly__t__DOT__v_c = 0xffU;
            __Vdly__t__DOT__v_d = 0x01ffU;
            __Vdly__t__DOT__v_e = 0xffffU;
            __Vdly__t__DOT__v_f = 0x0001ffffU;
            __Vdly__t__DOT__v_g = 0xffffffffU;
            __Vdly__t__DOT__v_h = 0x00000001ffffffffULL;
            __Vdly__t__DOT__v_i = 0xffffffffffffffffULL;
            __Vdly__t__DOT__v_j[0U] = 0xffffffffU;
            __Vdly__t__DOT__v_j[1U] = 0xffffffffU;
            __Vdly__t__DOT__v_j[2U] = 1U;
            __Vdly__t__DOT__v_k[0U] = 0xffffffffU;
            __Vdly__t__DOT__v_k[1U] = 0xffffffffU;
            __Vdly__t__DOT__v_k[2U] = 0xffffffffU;
            __Vdly__t__DOT__v_k[3U] = 0xffffffffU;
            __Vdly__t__DOT__v_l[0U] = 0xffffffffU;
            __Vdly__t__DOT__v_l[1U] = 0xffffffffU;
            __Vdly__t__DOT__v_l[2U] = 0xffffffffU;
            __Vdly__t__DOT__v_l[3U] = 0xffffffffU;
            __Vdly__t__DOT__v_l[4U] = 1U;
            VL_ASSIGN_W(256, __Vdly__t__DOT__v_m, Vt_trace_ascendingrange__ConstPool__CONST_h5b979007_0);
            VL_ASSIGN_W(257, __Vdly__t__DOT__v_n, Vt_trace_ascendingrange__ConstPool__CONST_h7d18c5a8_0);
            VL_ASSIGN_W(512, __Vdly__t__DOT__v_o, Vt_trace_ascendingrange__ConstPool__CONST_h8b2d9f06_0);
            __Vdly__t__DOT__v_p = 7U;
            __Vdly__t__DOT__v_q = 0x7fffU;
            __Vdly__t__DOT__v_r = 0x7fffffffU;
            __Vdly__t__DOT__v_s = 0x7fffffffffffffffULL;
            __Vdly__t__DOT__v_t[0U] = 0xffffffffU;
            __Vdly__t__DOT__v_t[1U] = 0xffffffffU;
            __Vdly__t__DOT__v_t[2U] = 0xffffffffU;
            __Vdly__t__DOT__v_t[3U] = 0x7fffffffU;
            VL_ASSIGN_W(255, __Vdly__t__DOT__v_u, Vt_trace_ascendingrange__ConstPool__CONST_ha837507f_0);
            VL_ASSIGN_W(511, __Vdly__t__DOT__v_v, Vt_trace_ascendingrange__ConstPool__CONST_h936cb6fc_0);
        } else {
            __Vdly__t__DOT__v_b = (3U & ((IData)(vlSelfRef.t__DOT__v_b) 
                                         << 1U));
            __Vdly__t__DOT__v_c = (0x000000ffU & ((IData)(vlSelfRef.t__DOT__v_c) 
                                                  << 1U));
            __Vdly__t__DOT__v_d = (0x000001ffU & ((IData)(vlSelfRef.t__DOT__v_d) 
                                                  << 1U));
            __Vdly__t__DOT__v_e = (0x0000ffffU & ((IData)(vlSelfRef.t__DOT__v_e) 
                                                  << 1U));
            __Vdly__t__DOT__v_f = (0x0001ffffU & (vlSelfRef.t__DOT__v_f 
                                                  << 1U));
            __Vdly__t__DOT__v_g = (vlSelfRef.t__DOT__v_g 
                                   << 1U);
            __Vdly__t__DOT__v_h = (0x00000001ffffffffULL 
                                   & (vlSelfRef.t__DOT__v_h 
                                      << 1U));
            __Vdly__t__DOT__v_i = (vlSelfRef.t__DOT__v_i 
                                   << 1U);
            VL_SHIFTL_WWI(65,65,32, __Vtemp_2, vlSelfRef.t__DOT__v_j, 1U);
            __Vdly__t__DOT__v_j[0U] = __Vtemp_2[0U];
            __Vdly__t__DOT__v_j[1U] = __Vtemp_2[1U];
            __Vdly__t__DOT__v_j[2U] = (1U & __Vtemp_2[2U]);
            VL_SHIFTL_WWI(128,128,32, __Vdly__t__DOT__v_k, vlSelfRef.t__DOT__v_k, 1U);
            VL_SHIFTL_WWI(129,129,32, __Vtemp_4, vlSelfRef.t__DOT__v_l, 1U);
            __Vdly__t__DOT__v_l[0U] = __Vtemp_4[0U];
            __Vdly__t__DOT__v_l[1U] = __Vtemp_4[1U];
            __Vdly__t__DOT__v_l[2U] = __Vtemp_4[2U];
            __Vdly__t__DOT__v_l[3U] = __Vtemp_4[3U];
            __Vdly__t__DOT__v_l[4U] = (1U & __Vtemp_4[4U]);
            VL_SHIFTL_WWI(256,256,32, __Vdly__t__DOT__v_m, vlSelfRef.t__DOT__v_m, 1U);
            VL_SHIFTL_WWI(257,257,32, __Vtemp_5, vlSelfRef.t__DOT__v_n, 1U);
            __Vdly__t__DOT__v_n[0U] = __Vtemp_5[0U];
            __Vdly__t__DOT__v_n[1U] = __Vtemp_5[1U];
            __Vdly__t__DOT__v_n[2U] = __Vtemp_5[2U];
            __Vdly__t__DOT__v_n[3U] = __Vtemp_5[3U];
            __Vdly__t__DOT__v_n[4U] = __Vtemp_5[4U];
            __Vdly__t__DOT__v_n[5U] = __Vtemp_5[5U];
            __Vdly__t__DOT__v_n[6U] = __Vtemp_5[6U];
            __Vdly__t__DOT__v_n[7U] = __Vtemp_5[7U];
            __Vdly__t__DOT__v_n[8U] = (1U & __Vtemp_5[8U]);
            VL_SHIFTL_WWI(512,512,32, __Vdly__t__DOT__v_o, vlSelfRef.t__DOT__v_o, 1U);
            __Vdly__t__DOT__v_p = (7U & ((IData)(vlSelfRef.t__DOT__v_p) 
                                         << 1U));
            __Vdly__t__DOT__v_q = (0x00007fffU & ((IData)(vlSelfRef.t__DOT__v_q) 
                                                  << 1U));
            __Vdly__t__DOT__v_r = (0x7fffffffU & (vlSelfRef.t__DOT__v_r 
                                                  << 1U));
            __Vdly__t__DOT__v_s = (0x7fffffffffffffffULL 
                                   & (vlSelfRef.t__DOT__v_s 
                                      << 1U));
            VL_SHIFTL_WWI(127,127,32, __Vtemp_7, vlSelfRef.t__DOT__v_t, 1U);
            __Vdly__t__DOT__v_t[0U] = __Vtemp_7[0U];
            __Vdly__t__DOT__v_t[1U] = __Vtemp_7[1U];
            __Vdly__t__DOT__v_t[2U] = __Vtemp_7[2U];
            __Vdly__t__DOT__v_t[3U] = (0x7fffffffU 
                                       & __Vtemp_7[3U]);
            VL_SHIFTL_WWI(255,255,32, __Vtemp_8, vlSelfRef.t__DOT__v_u, 1U);
            __Vdly__t__DOT__v_u[0U] = __Vtemp_8[0U];
            __Vdly__t__DOT__v_u[1U] = __Vtemp_8[1U];
            __Vdly__t__DOT__v_u[2U] = __Vtemp_8[2U];
            __Vdly__t__DOT__v_u[3U] = __Vtemp_8[3U];
            __Vdly__t__DOT__v_u[4U] = __Vtemp_8[4U];
            __Vdly__t__DOT__v_u[5U] = __Vtemp_8[5U];
            __Vdly__t__DOT__v_u[6U] = __Vtemp_8[6U];
            __Vdly__t__DOT__v_u[7U] = (0x7fffffffU 
                                       & __Vtemp_8[7U]);
            VL_SHIFTL_WWI(511,511,32, __Vtemp_9, vlSelfRef.t__DOT__v_v, 1U);
            __Vdly__t__DOT__v_v[0U] = __Vtemp_9[0U];
            __Vdly__t__DOT__v_v[1U] = __Vtemp_9[1U];
            __Vdly__t__DOT__v_v[2U] = __Vtemp_9[2U];
            __Vdly__t__DOT__v_v[3U] = __Vtemp_9[3U];
            __Vdly__t__DOT__v_v[4U] = __Vtemp_9[4U];
            __Vdly__t__DOT__v_v[5U] = __Vtemp_9[5U];
            __Vdly__t__DOT__v_v[6U] = __Vtemp_9[6U];
            __Vdly__t__DOT__v_v[7U] = __Vtemp_9[7U];
            __Vdly__t__DOT__v_v[8U] = __Vtemp_9[8U];
            __Vdly__t__DOT__v_v[9U] = __Vtemp_9[9U];
            __Vdly__t__DOT__v_v[10U] = __Vtemp_9[10U];
            __Vdly__t__DOT__v_v[11U] = __Vtemp_9[11U];
            __Vdly__t__DOT__v_v[12U] = __Vtemp_9[12U];
            __Vdly__t__DOT__v_v[13U] = __Vtemp_9[13U];
            __Vdly__t__DOT__v_v[14U] = __Vtemp_9[14U];
            __Vdly__t__DOT__v_v[15U] = (0x7fffffffU 
                                        & __Vtemp_9[15U]);
        }
        __Vdly__t__DOT__cyc = ((IData)(1U) + vlSelfRef.t__DOT__cyc);
        if (VL_UNLIKELY(((2U == vlSelfRef.t__DOT__cyc)))) {
            vlCtx.writef(0U, 0U);
            VL_FINISH_MT("/root/reference/test_regress/t/t_trace_ascendingrange.v", 93, "");
        }
        vlSelfRef.t__DOT__v_b = __Vdly__t__DOT__v_b;
        vlSelfRef.t__DOT__v_c = __Vdly__t__DOT__v_c;
        vlSelfRef.t__DOT__v_d = __Vdly__t__DOT__v_d;
        vlSelfRef.t__DOT__v_e = __Vdly__t__DOT__v_e;
        vlSelfRef.t__DOT__v_f = __Vdly__t__DOT__v_f;
        vlSelfRef.t__DOT__v_g = __Vdly__t__DOT__v_g;
        vlSelfRef.t__DOT__v_h = __Vdly__t__DOT__v_h;
        vlSelfRef.t__DOT__v_i = __Vdly__t__DOT__v_i;
        vlSelfRef.t__DOT__v_j[0U] = __Vdly__t__DOT__v_j[0U];
        vlSelfRef.t__DOT__v_j[1U] = __Vdly__t__DOT__v_j[1U];
        vlSelfRef.t__DOT__v_j[2U] = __Vdly__t__DOT__v_j[2U];
        vlSelfRef.t__DOT__v_k[0U] = __Vdly__t__DOT__v_k[0U];
        vlSelfRef.t__DOT__v_k[1U] = __Vdly__t__DOT__v_k[1U];
        vlSelfRef.t__DOT__v_k[2U] = __Vdly__t__DOT__v_k[2U];
        vlSelfRef.t__DOT__v_k[3U] = __Vdly__t__DOT__v_k[3U];
        vlSelfRef.t__DOT__v_l[0U] = __Vdly__t__DOT__v_l[0U];
        vlSelfRef.t__DOT__v_l[1U] = __Vdly__t__DOT__v_l[1U];
        vlSelfRef.t__DOT__v_l[2U] = __Vdly__t__DOT__v_l[2U];
        vlSelfRef.t__DOT__v_l[3U] = __Vdly__t__DOT__v_l[3U];
        vlSelfRef.t__DOT__v_l[4U] = __Vdly__t__DOT__v_l[4U];
        vlSelfRef.t__DOT__v_m[0U] = __Vdly__t__DOT__v_m[0U];
        vlSelfRef.t__DOT__v_m[1U] = __Vdly__t__DOT__v_m[1U];
        vlSelfRef.t__DOT__v_m[2U] = __Vdly__t__DOT__v_m[2U];
        vlSelfRef.t__DOT__v_m[3U] = __Vdly__t__DOT__v_m[3U];
        vlSelfRef.t__DOT__v_m[4U] = __Vdly__t__DOT__v_m[4U];
        vlSelfRef.t__DOT__v_m[5U] = __Vdly__t__DOT__v_m[5U];
        vlSelfRef.t__DOT__v_m[6U] = __Vdly__t__DOT__v_m[6U];
        vlSelfRef.t__DOT__v_m[7U] = __Vdly__t__DOT__v_m[7U];
        vlSelfRef.t__DOT__v_n[0U] = __Vdly__t__DOT__v_n[0U];
        vlSelfRef.t__DOT__v_n[1U] = __Vdly__t__DOT__v_n[1U];
        vlSelfRef.t__DOT__v_n[2U] = __Vdly__t__DOT__v_n[2U];
        vlSelfRef.t__DOT__v_n[3U] = __Vdly__t__DOT__v_n[3U];
        vlSelfRef.t__DOT__v_n[4U] = __Vdly__t__DOT__v_n[4U];
        vlSelfRef.t__DOT__v_n[5U] = __Vdly__t__DOT__v_n[5U];
        vlSelfRef.t__DOT__v_n[6U] = __Vdly__t__DOT__v_n[6U];
        vlSelfRef.t__DOT__v_n[7U] = __Vdly__t__DOT__v_n[7U];
        vlSelfRef.t__DOT__v_n[8U] = __Vdly__t__DOT__v_n[8U];
        vlSelfRef.t__DOT__v_o[0U] = __Vdly__t__DOT__v_o[0U];
        vlSelfRef.t__DOT__v_o[1U] = __Vdly__t__DOT__v_o[1U];
        vlSelfRef.t__DOT__v_o[2U] = __Vdly__t__DOT__v_o[2U];
        vlSelfRef.t__DOT__v_o[3U] = __Vdly__t__DOT__v_o[3U];
        vlSelfRef.t__DOT__v_o[4U] = __Vdly__t__DOT__v_o[4U];
        vlSelfRef.t__DOT__v_o[5U] = __Vdly__t__DOT__v_o[5U];
        vlSelfRef.t__DOT__v_o[6U] = __Vdly__t__DOT__v_o[6U];
        vlSelfRef.t__DOT__v_o[7U] = __Vdly__t__DOT__v_o[7U];
        vlSelfRef.t__DOT__v_o[8U] = __Vdly__t__DOT__v_o[8U];
        vlSelfRef.t__DOT__v_o[9U] = __Vdly__t__DOT__v_o[9U];
        vlSelfRef.t__DOT__v_o[10U] = __Vdly__t__DOT__v_o[10U];
        vlSelfRef.t__DOT__v_o[11U] = __Vdly__t__DOT__v_o[11U];
        vlSelfRef.t__DOT__v_o[12U] = __Vdly__t__DOT__v_o[12U];
        vlSelfRef.t__DOT__v_o[13U] = __Vdly__t__DOT__v_o[13U];
        vlSelfRef.t__DOT__v_o[14U] = __Vdly__t__DOT__v_o[14U];
        vlSelfRef.t__DOT__v_o[15U] = __Vdly__t__DOT__v_o[15U];
        vlSelfRef.t__DOT__v_p = __Vdly__t__DOT__v_p;
        vlSelfRef.t__DOT__v_q = __Vdly__t__DOT__v_q;
        vlSelfRef.t__DOT__v_r = __Vdly__t__DOT__v_r;
        vlSelfRef.t__DOT__v_s = __Vdly__t__DOT__v_s;
        vlSelfRef.t__DOT__v_t[0U] = __Vdly__t__DOT__v_t[0U];
        vlSelfRef.t__DOT__v_t[1U] = __Vdly__t__DOT__v_t[1U];
        vlSelfRef.t__DOT__v_t[2U] = __Vdly__t__DOT__v_t[2U];
        vlSelfRef.t__DOT__v_t[3U] = __Vdly__t__DOT__v_t[3U];
        vlSelfRef.t__DOT__v_u[0U] = __Vdly__t__DOT__v_u[0U];
        vlSelfRef.t__DOT__v_u[1U] = __Vdly__t__DOT__v_u[1U];
        vlSelfRef.t__DOT__v_u[2U] = __Vdly__t__DOT__v_u[2U];
        vlSelfRef.t__DOT__v_u[3U] = __Vdly__t__DOT__v_u[3U];
        vlSelfRef.t__DOT__v_u[4U] = __Vdly__t__DOT__v_u[4U];
        vlSelfRef.t__DOT__v_u[5U] = __Vdly__t__DOT__v_u[5U];
        vlSelfRef.t__DOT__v_u[6U] = __Vdly__t__DOT__v_u[6U];
        vlSelfRef.t__DOT__v_u[7U] = __Vdly__t__DOT__v_u[7U];
        vlSelfRef.t__DOT__v_v[0U] = __Vdly__t__DOT__v_v[0U];
        vlSelfRef.t__DOT__v_v[1U] = __Vdly__t__DOT__v_v[1U];
        vlSelfRef.t__DOT__v_v[2U] = __Vdly__t__DOT__v_v[2U];
        vlSelfRef.t__DOT__v_v[3U] = __Vdly__t__DOT__v_v[3U];
        vlSelfRef.t__DOT__v_v[4U] = __Vdly__t__DOT__v_v[4U];
        vlSelfRef.t__DOT__v_v[5U] = __Vdly__t__DOT__v_v[5U];
        vlSelfRef.t__DOT__v_v[6U] = __Vdly__t__DOT__v_v[6U];
        vlSelfRef.t__DOT__v_v[7U] = __Vdly__t__DOT__v_v[7U];
        vlSelfRef.t__DOT__v_v[8U] = __Vdly__t__DOT__v_v[8U];
        vlSelfRef.t__DOT__v_v[9U] = __Vdly__t__DOT__v_v[9U];
        vlSelfRef.t__DOT__v_v[10U] = __Vdly__t__DOT__v_v[10U];
        vlSelfRef.t__DOT__v_v[11U] = __Vdly__t__DOT__v_v[11U];
        vlSelfRef.t__DOT__v_v[12U] = __Vdly__t__DOT__v_v[12U];
        vlSelfRef.t__DOT__v_v[13U] = __Vdly__t__DOT__v_v[13U];
        vlSelfRef.t__DOT__v_v[14U] = __Vdly__t__DOT__v_v[14U];
        vlSelfRef.t__DOT__v_v[15U] = __Vdly__t__DOT__v_v[15U];
        vlSelfRef.t__DOT__cyc = __Vdly__t__DOT__cyc;
    }

    VLCU_D static void Vt_trace_ascendingrange___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt_trace_ascendingrange___024root___eval_phase__act(Vt_trace_ascendingrange___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt_trace_ascendingrange___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vt_trace_ascendingrange___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vt_trace_ascendingrange___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vt_trace_ascendingrange___024root___eval_phase__nba(Vt_trace_ascendingrange___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vt_trace_ascendingrange___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    Vt_trace_ascendingrange___024root___nba_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
            Vt_trace_ascendingrange___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vt_trace_ascendingrange___024root___eval(Vt_trace_ascendingrange___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt_trace_ascendingrange___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("/root/reference/test_regress/t/t_trace_ascendingrange.v", 7, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vt_trace_ascendingrange___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("/root/reference/test_regress/t/t_trace_ascendingrange.v", 7, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vt_trace_ascendingrange___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vt_trace_ascendingrange___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vt_trace_ascendingrange___024root___ctor_var_reset(Vt_trace_ascendingrange___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
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
        Vt_trace_ascendingrange___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt_trace_ascendingrange___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt_trace_ascendingrange___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt_trace_ascendingrange___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vt_trace_ascendingrange___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt_trace_ascendingrange___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt_trace_ascendingrange__cuda::info() {
    typedef Vt_trace_ascendingrange__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"t__DOT__v_b", "t.v_b", VLCU_MEMBER_STATE, 1, 0, 1, 2, 1, 0, T::OFF_t__DOT__v_b},
        {"t__DOT__v_c", "t.v_c", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF_t__DOT__v_c},
        {"t__DOT__v_p", "t.v_p", VLCU_MEMBER_STATE, 1, 0, 1, 3, 1, -1, T::OFF_t__DOT__v_p},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"t__DOT__v_d", "t.v_d", VLCU_MEMBER_STATE, 2, 0, 1, 9, 8, 0, T::OFF_t__DOT__v_d},
        {"t__DOT__v_e", "t.v_e", VLCU_MEMBER_STATE, 2, 0, 1, 16, 15, 0, T::OFF_t__DOT__v_e},
        {"t__DOT__v_q", "t.v_q", VLCU_MEMBER_STATE, 2, 0, 1, 15, 7, -7, T::OFF_t__DOT__v_q},
        {"t__DOT__cyc", "t.cyc", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__cyc},
        {"t__DOT__v_f", "t.v_f", VLCU_MEMBER_STATE, 4, 0, 1, 17, 16, 0, T::OFF_t__DOT__v_f},
        {"t__DOT__v_g", "t.v_g", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t__DOT__v_g},
        {"t__DOT__v_j", "t.v_j", VLCU_MEMBER_STATE, 4, 0, 3, 65, 64, 0, T::OFF_t__DOT__v_j},
        {"t__DOT__v_k", "t.v_k", VLCU_MEMBER_STATE, 4, 0, 4, 128, 127, 0, T::OFF_t__DOT__v_k},
        {"t__DOT__v_l", "t.v_l", VLCU_MEMBER_STATE, 4, 0, 5, 129, 128, 0, T::OFF_t__DOT__v_l},
        {"t__DOT__v_m", "t.v_m", VLCU_MEMBER_STATE, 4, 0, 8, 256, 255, 0, T::OFF_t__DOT__v_m},
        {"t__DOT__v_n", "t.v_n", VLCU_MEMBER_STATE, 4, 0, 9, 257, 256, 0, T::OFF_t__DOT__v_n},
        {"t__DOT__v_o", "t.v_o", VLCU_MEMBER_STATE, 4, 0, 16, 512, 511, 0, T::OFF_t__DOT__v_o},
        {"t__DOT__v_r", "t.v_r", VLCU_MEMBER_STATE, 4, 0, 1, 31, 15, -15, T::OFF_t__DOT__v_r},
        {"t__DOT__v_t", "t.v_t", VLCU_MEMBER_STATE, 4, 0, 4, 127, 63, -63, T::OFF_t__DOT__v_t},
        {"t__DOT__v_u", "t.v_u", VLCU_MEMBER_STATE, 4, 0, 8, 255, 127, -127, T::OFF_t__DOT__v_u},
        {"t__DOT__v_v", "t.v_v", VLCU_MEMBER_STATE, 4, 0, 16, 511, 255, -255, T::OFF_t__DOT__v_v},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"t__DOT__v_h", "t.v_h", VLCU_MEMBER_STATE, 8, 0, 1, 33, 32, 0, T::OFF_t__DOT__v_h},
        {"t__DOT__v_i", "t.v_i", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF_t__DOT__v_i},
        {"t__DOT__v_s", "t.v_s", VLCU_MEMBER_STATE, 8, 0, 1, 63, 31, -31, T::OFF_t__DOT__v_s},
        {"__VactTriggered", "__VactTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VactTriggered},
        {"__VnbaTriggered", "__VnbaTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VnbaTriggered},
    };
    static const char* const messages[] = {"*-* All Finished *-*\n", nullptr};
    static const char* const plusargs[] = {nullptr};
    static const VlcuTraceDecl traceDecls[] = {
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0}};  // (terminator, not counted)
    static const uint32_t traceOrder[] = {0};  // codes in dump order (+ terminator)
    static const uint32_t saveOrder[] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "t", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 1, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 29, 0, 0x8d500fcbce8c210bULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt_trace_ascendingrange__cuda VLCU_TESTBENCHES);
