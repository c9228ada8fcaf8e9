// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt_math_pow6"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt_math_pow6___024root final {
    CData/*3:0*/ a;
    QData/*39:0*/ a40;
    VlWide<3>/*69:0*/ a70;
    CData/*3:0*/ i65;
    CData/*3:0*/ j65;
    CData/*3:0*/ i33;
    CData/*3:0*/ j33;
    CData/*3:0*/ i30;
    CData/*3:0*/ j30;
    QData/*39:0*/ q65;
    QData/*39:0*/ r65;
    QData/*39:0*/ q33;
    QData/*39:0*/ r33;
    QData/*39:0*/ q30;
    QData/*39:0*/ r30;
    VlWide<3>/*69:0*/ w65;
    VlWide<3>/*69:0*/ x65;
    VlWide<3>/*69:0*/ w33;
    VlWide<3>/*69:0*/ x33;
    VlWide<3>/*69:0*/ w30;
    VlWide<3>/*69:0*/ x30;
    const char* vlNamep;
};

struct Vt_math_pow6__cuda final {
    typedef Vt_math_pow6___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_a = 146,
        OFF_a40 = 48,
        OFF_a70 = 128,
        OFF_i65 = 140,
        OFF_j65 = 141,
        OFF_i33 = 142,
        OFF_j33 = 143,
        OFF_i30 = 144,
        OFF_j30 = 145,
        OFF_q65 = 0,
        OFF_r65 = 8,
        OFF_q33 = 16,
        OFF_r33 = 24,
        OFF_q30 = 32,
        OFF_r30 = 40,
        OFF_w65 = 56,
        OFF_x65 = 68,
        OFF_w33 = 80,
        OFF_x33 = 92,
        OFF_w30 = 104,
        OFF_x30 = 116,
        kStateBytes = 152,  // S_persist
        kTouchedReadBytes = 0,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 0,  // S_state_wr
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
        s.a = st.ld<VlcuMem8>(OFF_a, i);
        s.a40 = st.ld<QData>(OFF_a40, i);
        st.ldW(OFF_a70, i, s.a70);
        s.i65 = st.ld<VlcuMem8>(OFF_i65, i);
        s.j65 = st.ld<VlcuMem8>(OFF_j65, i);
        s.i33 = st.ld<VlcuMem8>(OFF_i33, i);
        s.j33 = st.ld<VlcuMem8>(OFF_j33, i);
        s.i30 = st.ld<VlcuMem8>(OFF_i30, i);
        s.j30 = st.ld<VlcuMem8>(OFF_j30, i);
        s.q65 = st.ld<QData>(OFF_q65, i);
        s.r65 = st.ld<QData>(OFF_r65, i);
        s.q33 = st.ld<QData>(OFF_q33, i);
        s.r33 = st.ld<QData>(OFF_r33, i);
        s.q30 = st.ld<QData>(OFF_q30, i);
        s.r30 = st.ld<QData>(OFF_r30, i);
        st.ldW(OFF_w65, i, s.w65);
        st.ldW(OFF_x65, i, s.x65);
        st.ldW(OFF_w33, i, s.w33);
        st.ldW(OFF_x33, i, s.x33);
        st.ldW(OFF_w30, i, s.w30);
        st.ldW(OFF_x30, i, s.x30);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_a, i, s.a);
        st.st<QData>(OFF_a40, i, s.a40);
        st.stW(OFF_a70, i, s.a70);
        st.st<VlcuMem8>(OFF_i65, i, s.i65);
        st.st<VlcuMem8>(OFF_j65, i, s.j65);
        st.st<VlcuMem8>(OFF_i33, i, s.i33);
        st.st<VlcuMem8>(OFF_j33, i, s.j33);
        st.st<VlcuMem8>(OFF_i30, i, s.i30);
        st.st<VlcuMem8>(OFF_j30, i, s.j30);
        st.st<QData>(OFF_q65, i, s.q65);
        st.st<QData>(OFF_r65, i, s.r65);
        st.st<QData>(OFF_q33, i, s.q33);
        st.st<QData>(OFF_r33, i, s.r33);
        st.st<QData>(OFF_q30, i, s.q30);
        st.st<QData>(OFF_r30, i, s.r30);
        st.stW(OFF_w65, i, s.w65);
        st.stW(OFF_x65, i, s.x65);
        st.stW(OFF_w33, i, s.w33);
        st.stW(OFF_x33, i, s.x33);
        st.stW(OFF_w30, i, s.w30);
        st.stW(OFF_x30, i, s.x30);
    }
    VLCU_D static void load_inputs_row(Root& s, const uint32_t* rowp, uint64_t stride, uint64_t i) {
        s.a = static_cast<VlcuMem8>(rowp[0 * stride + i] & 0xfU);
        s.a40 = (static_cast<QData>(rowp[2 * stride + i] & 0xffU) << 32) | rowp[1 * stride + i];
        s.a70[0] = rowp[3 * stride + i] & 0xffffffffU;
        s.a70[1] = rowp[4 * stride + i] & 0xffffffffU;
        s.a70[2] = rowp[5 * stride + i] & 0x3fU;
    }
    VLCU_D static void store_outputs_row(const Root& s, uint32_t* rowp, uint64_t stride, uint64_t i) {
        rowp[0 * stride + i] = s.i65;
        rowp[1 * stride + i] = s.j65;
        rowp[2 * stride + i] = s.i33;
        rowp[3 * stride + i] = s.j33;
        rowp[4 * stride + i] = s.i30;
        rowp[5 * stride + i] = s.j30;
        rowp[6 * stride + i] = static_cast<uint32_t>(s.q65);
        rowp[7 * stride + i] = static_cast<uint32_t>(s.q65 >> 32);
        rowp[8 * stride + i] = static_cast<uint32_t>(s.r65);
        rowp[9 * stride + i] = static_cast<uint32_t>(s.r65 >> 32);
        rowp[10 * stride + i] = static_cast<uint32_t>(s.q33);
        rowp[11 * stride + i] = static_cast<uint32_t>(s.q33 >> 32);
        rowp[12 * stride + i] = static_cast<uint32_t>(s.r33);
        rowp[13 * stride + i] = static_cast<uint32_t>(s.r33 >> 32);
        rowp[14 * stride + i] = static_cast<uint32_t>(s.q30);
        rowp[15 * stride + i] = static_cast<uint32_t>(s.q30 >> 32);
        rowp[16 * stride + i] = static_cast<uint32_t>(s.r30);
        rowp[17 * stride + i] = static_cast<uint32_t>(s.r30 >> 32);
        rowp[18 * stride + i] = s.w65[0];
        rowp[19 * stride + i] = s.w65[1];
        rowp[20 * stride + i] = s.w65[2];
        rowp[21 * stride + i] = s.x65[0];
        rowp[22 * stride + i] = s.x65[1];
        rowp[23 * stride + i] = s.x65[2];
        rowp[24 * stride + i] = s.w33[0];
        rowp[25 * stride + i] = s.w33[1];
        rowp[26 * stride + i] = s.w33[2];
        rowp[27 * stride + i] = s.x33[0];
        rowp[28 * stride + i] = s.x33[1];
        rowp[29 * stride + i] = s.x33[2];
        rowp[30 * stride + i] = s.w30[0];
        rowp[31 * stride + i] = s.w30[1];
        rowp[32 * stride + i] = s.w30[2];
        rowp[33 * stride + i] = s.x30[0];
        rowp[34 * stride + i] = s.x30[1];
        rowp[35 * stride + i] = s.x30[2];
    }
    VLCU_D static uint64_t outputs_sig(const Root& s, uint64_t sig) {
        uint32_t acc = VLCU_SIG_ACC_INIT;
        acc = vlcu_sig_word(acc, s.i65);
        acc = vlcu_sig_word(acc, s.j65);
        acc = vlcu_sig_word(acc, s.i33);
        acc = vlcu_sig_word(acc, s.j33);
        acc = vlcu_sig_word(acc, s.i30);
        acc = vlcu_sig_word(acc, s.j30);
        acc = vlcu_sig_word(acc, static_cast<uint32_t>(s.q65));
        acc = vlcu_sig_word(acc, static_cast<uint32_t>(s.q65 >> 32));
        acc = vlcu_sig_word(acc, static_cast<uint32_t>(s.r65));
        acc = vlcu_sig_word(acc, static_cast<uint32_t>(s.r65 >> 32));
        acc = vlcu_sig_word(acc, static_cast<uint32_t>(s.q33));
        acc = vlcu_sig_word(acc, static_cast<uint32_t>(s.q33 >> 32));
        acc = vlcu_sig_word(acc, static_cast<uint32_t>(s.r33));
        acc = vlcu_sig_word(acc, static_cast<uint32_t>(s.r33 >> 32));
        acc = vlcu_sig_word(acc, static_cast<uint32_t>(s.q30));
        acc = vlcu_sig_word(acc, static_cast<uint32_t>(s.q30 >> 32));
        acc = vlcu_sig_word(acc, static_cast<uint32_t>(s.r30));
        acc = vlcu_sig_word(acc, static_cast<uint32_t>(s.r30 >> 32));
        acc = vlcu_sig_word(acc, s.w65[0]);
        acc = vlcu_sig_word(acc, s.w65[1]);
        acc = vlcu_sig_word(acc, s.w65[2]);
        acc = vlcu_sig_word(acc, s.x65[0]);
        acc = vlcu_sig_word(acc, s.x65[1]);
        acc = vlcu_sig_word(acc, s.x65[2]);
        acc = vlcu_sig_word(acc, s.w33[0]);
        acc = vlcu_sig_word(acc, s.w33[1]);
        acc = vlcu_sig_word(acc, s.w33[2]);
        acc = vlcu_sig_word(acc, s.x33[0]);
        acc = vlcu_sig_word(acc, s.x33[1]);
        acc = vlcu_sig_word(acc, s.x33[2]);
        acc = vlcu_sig_word(acc, s.w30[0]);
        acc = vlcu_sig_word(acc, s.w30[1]);
        acc = vlcu_sig_word(acc, s.w30[2]);
        acc = vlcu_sig_word(acc, s.x30[0]);
        acc = vlcu_sig_word(acc, s.x30[1]);
        acc = vlcu_sig_word(acc, s.x30[2]);
        return vlcu_sig_mix(sig, acc);
    }
    enum : uint32_t { kInWords = 6, kOutWords = 36, kInBytes = 21, kOutBytes = 126, kTraceWords = 0 };

    VLCU_D static void Vt_math_pow6___024root___eval_static(Vt_math_pow6___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        Vt_math_pow6___024root___eval_static__TOP(vlSelf, vlCtx);
    }

    VLCU_D static void Vt_math_pow6___024root___eval_static__TOP(Vt_math_pow6___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        VlWide<3>/*95:0*/ __Vtemp_2;
        VlWide<3>/*95:0*/ __Vtemp_3;
        VlWide<3>/*95:0*/ __Vtemp_5;
        VlWide<3>/*95:0*/ __Vtemp_7;
        VlWide<3>/*95:0*/ __Vtemp_8;
        VlWide<3>/*95:0*/ __Vtemp_10;
        VlWide<3>/*95:0*/ __Vtemp_12;
        VlWide<3>/*95:0*/ __Vtemp_13;
        VlWide<3>/*95:0*/ __Vtemp_15;
        VlWide<3>/*95:0*/ __Vtemp_16;
        VlWide<3>/*95:0*/ __Vtemp_18;
        VlWide<3>/*95:0*/ __Vtemp_19;
        VlWide<3>/*95:0*/ __Vtemp_21;
        VlWide<3>/*95:0*/ __Vtemp_23;
        VlWide<3>/*95:0*/ __Vtemp_24;
        VlWide<3>/*95:0*/ __Vtemp_26;
        // Body
        __Vtemp_2[0U] = 3U;
        __Vtemp_2[1U] = 0U;
        __Vtemp_2[2U] = 0U;
        VL_POW_WWI(65,65,4, __Vtemp_3, __Vtemp_2, (IData)(vlSelfRef.a));
        vlSelfRef.i65 = (0x0000000fU & __Vtemp_3[0U]);
        __Vtemp_5[0U] = 3U;
        __Vtemp_5[1U] = 0U;
        __Vtemp_5[2U] = 0U;
        vlSelfRef.j65 = (0x0000000fU & VL_POW_IIW(4,4,65, (IData)(vlSelfRef.a), __Vtemp_5));
        vlSelfRef.i33 = (0x0000000fU & (IData)((0x00000001ffffffffULL 
                                                & VL_POW_QQI(33,33,4, 3ULL, (IData)(vlSelfRef.a)))));
        vlSelfRef.j33 = (0x0000000fU & VL_POW_IIQ(4,4,33, (IData)(vlSelfRef.a), 3ULL));
        vlSelfRef.i30 = (0x0000000fU & VL_POW_III(30,30,4, (IData)(3U), (IData)(vlSelfRef.a)));
        vlSelfRef.j30 = (0x0000000fU & VL_POW_III(4,4,30, (IData)(vlSelfRef.a), 3U));
        __Vtemp_7[0U] = 3U;
        __Vtemp_7[1U] = 0U;
        __Vtemp_7[2U] = 0U;
        VL_POW_WWQ(65,65,40, __Vtemp_8, __Vtemp_7, vlSelfRef.a40);
        vlSelfRef.q65 = (0x000000ffffffffffULL & (((QData)((IData)(__Vtemp_8[1U])) 
                                                   << 0x00000020U) 
                                                  | (QData)((IData)(__Vtemp_8[0U]))));
        __Vtemp_10[0U] = 3U;
        __Vtemp_10[1U] = 0U;
        __Vtemp_10[2U] = 0U;
        vlSelfRef.r65 = (0x000000ffffffffffULL & VL_POW_QQW(40,40,65, vlSelfRef.a40, __Vtemp_10));
        vlSelfRef.q33 = (0x000000ffffffffffULL & VL_POW_QQQ(40,40,40, 3ULL, vlSelfRef.a40));
        vlSelfRef.r33 = (0x000000ffffffffffULL & VL_POW_QQQ(40,40,33, vlSelfRef.a40, 3ULL));
        vlSelfRef.q30 = (0x000000ffffffffffULL & VL_POW_QQQ(40,40,40, 3ULL, vlSelfRef.a40));
        vlSelfRef.r30 = (0x000000ffffffffffULL & VL_POW_QQI(40,40,30, vlSelfRef.a40, 3U));
        __Vtemp_12[0U] = 3U;
        __Vtemp_12[1U] = 0U;
        __Vtemp_12[2U] = 0U;
        VL_POW_WWW(70,70,70, __Vtemp_13, __Vtemp_12, vlSelfRef.a70);
        vlSelfRef.w65[0U] = __Vtemp_13[0U];
        vlSelfRef.w65[1U] = __Vtemp_13[1U];
        vlSelfRef.w65[2U] = (0x0000003fU & __Vtemp_13[2U]);
        __Vtemp_15[0U] = 3U;
        __Vtemp_15[1U] = 0U;
        __Vtemp_15[2U] = 0U;
        VL_POW_WWW(70,70,65, __Vtemp_16, vlSelfRef.a70, __Vtemp_15);
        vlSelfRef.x65[0U] = __Vtemp_16[0U];
        vlSelfRef.x65[1U] = __Vtemp_16[1U];
        vlSelfRef.x65[2U] = (0x0000003fU & __Vtemp_16[2U]);
        __Vtemp_18[0U] = 3U;
        __Vtemp_18[1U] = 0U;
        __Vtemp_18[2U] = 0U;
        VL_POW_WWW(70,70,70, __Vtemp_19, __Vtemp_18, vlSelfRef.a70);
        vlSelfRef.w33[0U] = __Vtemp_19[0U];
        vlSelfRef.w33[1U] = __Vtemp_19[1U];
        vlSelfRef.w33[2U] = (0x0000003fU & __Vtemp_19[2U]);
        VL_POW_WWQ(70,70,33, __Vtemp_21, vlSelfRef.a70, 3ULL);
        vlSelfRef.x33[0U] = __Vtemp_21[0U];
        vlSelfRef.x33[1U] = __Vtemp_21[1U];
        vlSelfRef.x33[2U] = (0x0000003fU & __Vtemp_21[2U]);
        __Vtemp_23[0U] = 3U;
        __Vtemp_23[1U] = 0U;
        __Vtemp_23[2U] = 0U;
        VL_POW_WWW(70,70,70, __Vtemp_24, __Vtemp_23, vlSelfRef.a70);
        vlSelfRef.w30[0U] = __Vtemp_24[0U];
        vlSelfRef.w30[1U] = __Vtemp_24[1U];
        vlSelfRef.w30[2U] = (0x0000003fU & __Vtemp_24[2U]);
        VL_POW_WWI(70,70,30, __Vtemp_26, vlSelfRef.a70, 3U);
        vlSelfRef.x30[0U] = __Vtemp_26[0U];
        vlSelfRef.x30[1U] = __Vtemp_26[1U];
        vlSelfRef.x30[2U] = (0x0000003fU & __Vtemp_26[2U]);
    }

    VLCU_D static void Vt_math_pow6___024root___eval_initial(Vt_math_pow6___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        {
            // Inlined CFunc: _eval_initial__TOP
            vlCtx.writef(0U, 0U);
            VL_FINISH_MT("/root/reference/test_regress/t/t_math_pow6.v", 45, "");
        }
    }

    VLCU_D static void Vt_math_pow6___024root___eval_final(Vt_math_pow6___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_math_pow6___024root___eval_settle(Vt_math_pow6___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_math_pow6___024root___eval(Vt_math_pow6___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt_math_pow6___024root___ctor_var_reset(Vt_math_pow6___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->a = VL_SCOPED_RAND_RESET_I(4, __VscopeHash, 510903276987443985ull);
        vlSelf->a40 = VL_SCOPED_RAND_RESET_Q(40, __VscopeHash, 12894432540807477949ull);
        VL_SCOPED_RAND_RESET_W(70, vlSelf->a70, __VscopeHash, 8629837237503696853ull);
    }

    // tfp->dump(): every traced value stored at its trace code.  Full functions, then the
    // constants, as VerilatedTrace::dump runs them the first time (verilated_trace_imp.h:348-358)
    VLCU_D static void trace_sample(Root* vlSelf, VlcuContext& vlCtx, uint32_t* valsp) {
        (void)vlSelf; (void)vlCtx; (void)valsp;
    }

    VLCU_D static void ctor_var_reset(Root* vlSelf, VlcuContext& vlCtx) {
        Vt_math_pow6___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_pow6___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_pow6___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_pow6___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_pow6___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt_math_pow6___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt_math_pow6__cuda::info() {
    typedef Vt_math_pow6__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"a", "a", VLCU_MEMBER_IN, 1, 0, 1, 4, 3, 0, T::OFF_a},
        {"a40", "a40", VLCU_MEMBER_IN, 8, 0, 1, 40, 39, 0, T::OFF_a40},
        {"a70", "a70", VLCU_MEMBER_IN, 4, 0, 3, 70, 69, 0, T::OFF_a70},
        {"i65", "i65", VLCU_MEMBER_OUT, 1, 0, 1, 4, 3, 0, T::OFF_i65},
        {"j65", "j65", VLCU_MEMBER_OUT, 1, 0, 1, 4, 3, 0, T::OFF_j65},
        {"i33", "i33", VLCU_MEMBER_OUT, 1, 0, 1, 4, 3, 0, T::OFF_i33},
        {"j33", "j33", VLCU_MEMBER_OUT, 1, 0, 1, 4, 3, 0, T::OFF_j33},
        {"i30", "i30", VLCU_MEMBER_OUT, 1, 0, 1, 4, 3, 0, T::OFF_i30},
        {"j30", "j30", VLCU_MEMBER_OUT, 1, 0, 1, 4, 3, 0, T::OFF_j30},
        {"q65", "q65", VLCU_MEMBER_OUT, 8, 0, 1, 40, 39, 0, T::OFF_q65},
        {"r65", "r65", VLCU_MEMBER_OUT, 8, 0, 1, 40, 39, 0, T::OFF_r65},
        {"q33", "q33", VLCU_MEMBER_OUT, 8, 0, 1, 40, 39, 0, T::OFF_q33},
        {"r33", "r33", VLCU_MEMBER_OUT, 8, 0, 1, 40, 39, 0, T::OFF_r33},
        {"q30", "q30", VLCU_MEMBER_OUT, 8, 0, 1, 40, 39, 0, T::OFF_q30},
        {"r30", "r30", VLCU_MEMBER_OUT, 8, 0, 1, 40, 39, 0, T::OFF_r30},
        {"w65", "w65", VLCU_MEMBER_OUT, 4, 0, 3, 70, 69, 0, T::OFF_w65},
        {"x65", "x65", VLCU_MEMBER_OUT, 4, 0, 3, 70, 69, 0, T::OFF_x65},
        {"w33", "w33", VLCU_MEMBER_OUT, 4, 0, 3, 70, 69, 0, T::OFF_w33},
        {"x33", "x33", VLCU_MEMBER_OUT, 4, 0, 3, 70, 69, 0, T::OFF_x33},
        {"w30", "w30", VLCU_MEMBER_OUT, 4, 0, 3, 70, 69, 0, T::OFF_w30},
        {"x30", "x30", VLCU_MEMBER_OUT, 4, 0, 3, 70, 69, 0, T::OFF_x30},
    };
    static const char* const messages[] = {"*-* All Finished *-*\n", nullptr};
    static const char* const plusargs[] = {nullptr};
    static const VlcuTraceDecl traceDecls[] = {
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0}};  // (terminator, not counted)
    static const uint32_t traceOrder[] = {0};  // codes in dump order (+ terminator)
    static const uint32_t saveOrder[] = {3, 4, 5, 6, 7, 8, 0, 15, 16, 17, 18, 19, 20, 2, 9, 10, 11, 12, 13, 14, 1, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "Vt_math_pow6", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 1, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 21, 0, 0x9c4ac0d2340936bULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt_math_pow6__cuda VLCU_TESTBENCHES);
