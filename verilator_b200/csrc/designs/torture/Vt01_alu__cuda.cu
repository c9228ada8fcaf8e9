// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt01_alu"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt01_alu___024root final {
    CData/*0:0*/ clk;
    CData/*3:0*/ op;
    CData/*7:0*/ a8;
    SData/*15:0*/ a16;
    IData/*31:0*/ a32;
    IData/*31:0*/ b32;
    QData/*63:0*/ a64;
    QData/*63:0*/ b64;
    CData/*7:0*/ q8;
    SData/*15:0*/ q16;
    IData/*31:0*/ q32;
    QData/*63:0*/ q64;
    CData/*6:0*/ flags;
    CData/*0:0*/ __VstlFirstIteration;
    CData/*0:0*/ __VstlPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*3:0*/ __Vtrigprevexpr___TOP__op__0;
    CData/*7:0*/ __Vtrigprevexpr___TOP__a8__0;
    CData/*0:0*/ __VicoDidInit;
    CData/*0:0*/ __VicoPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__1;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    SData/*15:0*/ __Vtrigprevexpr___TOP__a16__0;
    IData/*31:0*/ t01_alu__DOT__r32;
    IData/*31:0*/ __Vtrigprevexpr___TOP__a32__0;
    IData/*31:0*/ __Vtrigprevexpr___TOP__b32__0;
    IData/*31:0*/ __VactIterCount;
    QData/*63:0*/ t01_alu__DOT__r64;
    QData/*63:0*/ __Vtrigprevexpr___TOP__a64__0;
    QData/*63:0*/ __Vtrigprevexpr___TOP__b64__0;
    VlUnpacked<QData/*63:0*/, 1> __VstlTriggered;
    VlUnpacked<QData/*63:0*/, 2> __VicoTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vt01_alu__cuda final {
    typedef Vt01_alu___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 122,
        OFF_op = 123,
        OFF_a8 = 124,
        OFF_a16 = 116,
        OFF_a32 = 88,
        OFF_b32 = 92,
        OFF_a64 = 0,
        OFF_b64 = 8,
        OFF_q8 = 125,
        OFF_q16 = 118,
        OFF_q32 = 96,
        OFF_q64 = 16,
        OFF_flags = 126,
        OFF___VstlFirstIteration = 127,
        OFF___VstlPhaseResult = 128,
        OFF___Vtrigprevexpr___TOP__clk__0 = 129,
        OFF___Vtrigprevexpr___TOP__op__0 = 130,
        OFF___Vtrigprevexpr___TOP__a8__0 = 131,
        OFF___VicoDidInit = 132,
        OFF___VicoPhaseResult = 133,
        OFF___Vtrigprevexpr___TOP__clk__1 = 134,
        OFF___VactPhaseResult = 135,
        OFF___VnbaPhaseResult = 136,
        OFF___Vtrigprevexpr___TOP__a16__0 = 120,
        OFF_t01_alu__DOT__r32 = 100,
        OFF___Vtrigprevexpr___TOP__a32__0 = 104,
        OFF___Vtrigprevexpr___TOP__b32__0 = 108,
        OFF___VactIterCount = 112,
        OFF_t01_alu__DOT__r64 = 24,
        OFF___Vtrigprevexpr___TOP__a64__0 = 32,
        OFF___Vtrigprevexpr___TOP__b64__0 = 40,
        OFF___VstlTriggered = 48,
        OFF___VicoTriggered = 56,
        OFF___VactTriggered = 72,
        OFF___VnbaTriggered = 80,
        kStateBytes = 144,  // S_persist
        kTouchedReadBytes = 111,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 98,  // S_state_wr
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
        s.op = st.ld<VlcuMem8>(OFF_op, i);
        s.a8 = st.ld<VlcuMem8>(OFF_a8, i);
        s.a16 = st.ld<VlcuMem16>(OFF_a16, i);
        s.a32 = st.ld<IData>(OFF_a32, i);
        s.b32 = st.ld<IData>(OFF_b32, i);
        s.a64 = st.ld<QData>(OFF_a64, i);
        s.b64 = st.ld<QData>(OFF_b64, i);
        s.q8 = st.ld<VlcuMem8>(OFF_q8, i);
        s.q16 = st.ld<VlcuMem16>(OFF_q16, i);
        s.q32 = st.ld<IData>(OFF_q32, i);
        s.q64 = st.ld<QData>(OFF_q64, i);
        s.flags = st.ld<VlcuMem8>(OFF_flags, i);
        s.__VstlFirstIteration = st.ld<VlcuMem8>(OFF___VstlFirstIteration, i);
        s.__VstlPhaseResult = st.ld<VlcuMem8>(OFF___VstlPhaseResult, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__Vtrigprevexpr___TOP__op__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__op__0, i);
        s.__Vtrigprevexpr___TOP__a8__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__a8__0, i);
        s.__VicoDidInit = st.ld<VlcuMem8>(OFF___VicoDidInit, i);
        s.__VicoPhaseResult = st.ld<VlcuMem8>(OFF___VicoPhaseResult, i);
        s.__Vtrigprevexpr___TOP__clk__1 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__1, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        s.__Vtrigprevexpr___TOP__a16__0 = st.ld<VlcuMem16>(OFF___Vtrigprevexpr___TOP__a16__0, i);
        s.t01_alu__DOT__r32 = st.ld<IData>(OFF_t01_alu__DOT__r32, i);
        s.__Vtrigprevexpr___TOP__a32__0 = st.ld<IData>(OFF___Vtrigprevexpr___TOP__a32__0, i);
        s.__Vtrigprevexpr___TOP__b32__0 = st.ld<IData>(OFF___Vtrigprevexpr___TOP__b32__0, i);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        s.t01_alu__DOT__r64 = st.ld<QData>(OFF_t01_alu__DOT__r64, i);
        s.__Vtrigprevexpr___TOP__a64__0 = st.ld<QData>(OFF___Vtrigprevexpr___TOP__a64__0, i);
        s.__Vtrigprevexpr___TOP__b64__0 = st.ld<QData>(OFF___Vtrigprevexpr___TOP__b64__0, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VstlTriggered[e] = st.ldE<QData>(OFF___VstlTriggered, e, i);
        for (uint32_t e = 0; e < 2; ++e) s.__VicoTriggered[e] = st.ldE<QData>(OFF___VicoTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.st<VlcuMem8>(OFF_op, i, s.op);
        st.st<VlcuMem8>(OFF_a8, i, s.a8);
        st.st<VlcuMem16>(OFF_a16, i, s.a16);
        st.st<IData>(OFF_a32, i, s.a32);
        st.st<IData>(OFF_b32, i, s.b32);
        st.st<QData>(OFF_a64, i, s.a64);
        st.st<QData>(OFF_b64, i, s.b64);
        st.st<VlcuMem8>(OFF_q8, i, s.q8);
        st.st<VlcuMem16>(OFF_q16, i, s.q16);
        st.st<IData>(OFF_q32, i, s.q32);
        st.st<QData>(OFF_q64, i, s.q64);
        st.st<VlcuMem8>(OFF_flags, i, s.flags);
        st.st<VlcuMem8>(OFF___VstlFirstIteration, i, s.__VstlFirstIteration);
        st.st<VlcuMem8>(OFF___VstlPhaseResult, i, s.__VstlPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__op__0, i, s.__Vtrigprevexpr___TOP__op__0);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__a8__0, i, s.__Vtrigprevexpr___TOP__a8__0);
        st.st<VlcuMem8>(OFF___VicoDidInit, i, s.__VicoDidInit);
        st.st<VlcuMem8>(OFF___VicoPhaseResult, i, s.__VicoPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__1, i, s.__Vtrigprevexpr___TOP__clk__1);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.st<VlcuMem16>(OFF___Vtrigprevexpr___TOP__a16__0, i, s.__Vtrigprevexpr___TOP__a16__0);
        st.st<IData>(OFF_t01_alu__DOT__r32, i, s.t01_alu__DOT__r32);
        st.st<IData>(OFF___Vtrigprevexpr___TOP__a32__0, i, s.__Vtrigprevexpr___TOP__a32__0);
        st.st<IData>(OFF___Vtrigprevexpr___TOP__b32__0, i, s.__Vtrigprevexpr___TOP__b32__0);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
        st.st<QData>(OFF_t01_alu__DOT__r64, i, s.t01_alu__DOT__r64);
        st.st<QData>(OFF___Vtrigprevexpr___TOP__a64__0, i, s.__Vtrigprevexpr___TOP__a64__0);
        st.st<QData>(OFF___Vtrigprevexpr___TOP__b64__0, i, s.__Vtrigprevexpr___TOP__b64__0);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VstlTriggered, e, i, s.__VstlTriggered[e]);
        for (uint32_t e = 0; e < 2; ++e) st.stE<QData>(OFF___VicoTriggered, e, i, s.__VicoTriggered[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VactTriggered, e, i, s.__VactTriggered[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VnbaTriggered, e, i, s.__VnbaTriggered[e]);
    }
    VLCU_D static void load_inputs_row(Root& s, const uint32_t* rowp, uint64_t stride, uint64_t i) {
        s.clk = static_cast<VlcuMem8>(rowp[0 * stride + i] & 0x1U);
        s.op = static_cast<VlcuMem8>(rowp[1 * stride + i] & 0xfU);
        s.a8 = static_cast<VlcuMem8>(rowp[2 * stride + i] & 0xffU);
        s.a16 = static_cast<VlcuMem16>(rowp[3 * stride + i] & 0xffffU);
        s.a32 = static_cast<IData>(rowp[4 * stride + i] & 0xffffffffU);
        s.b32 = static_cast<IData>(rowp[5 * stride + i] & 0xffffffffU);
        s.a64 = (static_cast<QData>(rowp[7 * stride + i] & 0xffffffffU) << 32) | rowp[6 * stride + i];
        s.b64 = (static_cast<QData>(rowp[9 * stride + i] & 0xffffffffU) << 32) | rowp[8 * stride + i];
    }
    VLCU_D static void store_outputs_row(const Root& s, uint32_t* rowp, uint64_t stride, uint64_t i) {
        rowp[0 * stride + i] = s.q8;
        rowp[1 * stride + i] = s.q16;
        rowp[2 * stride + i] = s.q32;
        rowp[3 * stride + i] = static_cast<uint32_t>(s.q64);
        rowp[4 * stride + i] = static_cast<uint32_t>(s.q64 >> 32);
        rowp[5 * stride + i] = s.flags;
    }
    VLCU_D static uint64_t outputs_sig(const Root& s, uint64_t sig) {
        uint32_t acc = VLCU_SIG_ACC_INIT;
        acc = vlcu_sig_word(acc, s.q8);
        acc = vlcu_sig_word(acc, s.q16);
        acc = vlcu_sig_word(acc, s.q32);
        acc = vlcu_sig_word(acc, static_cast<uint32_t>(s.q64));
        acc = vlcu_sig_word(acc, static_cast<uint32_t>(s.q64 >> 32));
        acc = vlcu_sig_word(acc, s.flags);
        return vlcu_sig_mix(sig, acc);
    }
    enum : uint32_t { kInWords = 10, kOutWords = 6, kInBytes = 29, kOutBytes = 16, kTraceWords = 0 };

    VLCU_D static void Vt01_alu___024root___eval_static(Vt01_alu___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
        vlSelfRef.__Vtrigprevexpr___TOP__op__0 = vlSelfRef.op;
        vlSelfRef.__Vtrigprevexpr___TOP__a8__0 = vlSelfRef.a8;
        vlSelfRef.__Vtrigprevexpr___TOP__a16__0 = vlSelfRef.a16;
        vlSelfRef.__Vtrigprevexpr___TOP__a32__0 = vlSelfRef.a32;
        vlSelfRef.__Vtrigprevexpr___TOP__b32__0 = vlSelfRef.b32;
        vlSelfRef.__Vtrigprevexpr___TOP__a64__0 = vlSelfRef.a64;
        vlSelfRef.__Vtrigprevexpr___TOP__b64__0 = vlSelfRef.b64;
        vlSelfRef.__Vtrigprevexpr___TOP__clk__1 = vlSelfRef.clk;
    }

    VLCU_D static void Vt01_alu___024root___eval_initial(Vt01_alu___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt01_alu___024root___eval_final(Vt01_alu___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt01_alu___024root___eval_settle(Vt01_alu___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt01_alu___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t01_alu.v", 4, "", "DIDNOTCONVERGE: Settle region did not converge after '--converge-limit' of 10000 tries");
            }
            __VstlIterCount = ((IData)(1U) + __VstlIterCount);
            vlSelfRef.__VstlPhaseResult = Vt01_alu___024root___eval_phase__stl(vlSelf, vlCtx);
            vlSelfRef.__VstlFirstIteration = 0U;
        } while (vlSelfRef.__VstlPhaseResult);
    }

    VLCU_D static bool Vt01_alu___024root___trigger_anySet__stl(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt01_alu___024root___stl_sequent__TOP__0(Vt01_alu___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        if ((8U & (IData)(vlSelfRef.op))) {
            if ((4U & (IData)(vlSelfRef.op))) {
                if ((2U & (IData)(vlSelfRef.op))) {
                    if ((1U & (IData)(vlSelfRef.op))) {
                        vlSelfRef.t01_alu__DOT__r32 
                            = (- vlSelfRef.a32);
                        vlSelfRef.t01_alu__DOT__r64 
                            = ((~ vlSelfRef.a64) + (QData)((IData)(
                                                                   (1U 
                                                                    & (IData)(vlSelfRef.a8)))));
                    } else {
                        vlSelfRef.t01_alu__DOT__r32 
                            = ((1U & VL_REDXOR_32(vlSelfRef.a32)) 
                               + ((0U != (IData)(vlSelfRef.a16)) 
                                  + (0x000000ffU == (IData)(vlSelfRef.a8))));
                        vlSelfRef.t01_alu__DOT__r64 
                            = ((QData)((IData)((vlSelfRef.a64 
                                                >> 
                                                (0x0000001fU 
                                                 & vlSelfRef.b32)))) 
                               + (QData)((IData)((1U 
                                                  & (IData)(
                                                            (vlSelfRef.a64 
                                                             >> 
                                                             (0x0000003fU 
                                                              & vlSelfRef.b32)))))));
                    }
                } else if ((1U & (IData)(vlSelfRef.op))) {
                    vlSelfRef.t01_alu__DOT__r32 = (
                                                   (vlSelfRef.a32 
                                                    < vlSelfRef.b32)
                                                    ? vlSelfRef.a32
                                                    : 
                                                   (VL_LTS_III(32, vlSelfRef.a32, vlSelfRef.b32)
                                                     ? vlSelfRef.b32
                                                     : 0xdeadbeefU));
                    vlSelfRef.t01_alu__DOT__r64 = (
                                                   VL_GTES_IQQ(64, vlSelfRef.a64, vlSelfRef.b64)
                                                    ? 
                                                   (vlSelfRef.a64 
                                                    | vlSelfRef.b64)
                                                    : 
                                                   (vlSelfRef.a64 
                                                    & vlSelfRef.b64));
                } else {
                    vlSelfRef.t01_alu__DOT__r32 = (
                                                   (~ vlSelfRef.b32) 
                                                   & (((IData)(vlSelfRef.a8) 
                                                       << 0x00000018U) 
                                                      | (((IData)(vlSelfRef.a8) 
                                                          << 0x00000010U) 
                                                         | (((IData)(vlSelfRef.a8) 
                                                             << 8U) 
                                                            | (IData)(vlSelfRef.a8)))));
                    vlSelfRef.t01_alu__DOT__r64 = (
                                                   ((QData)((IData)(vlSelfRef.a64)) 
                                                    << 0x00000020U) 
                                                   | (QData)((IData)(
                                                                     (vlSelfRef.a64 
                                                                      >> 0x00000020U))));
                }
            } else if ((2U & (IData)(vlSelfRef.op))) {
                if ((1U & (IData)(vlSelfRef.op))) {
                    vlSelfRef.t01_alu__DOT__r32 = (
                                                   ((IData)(vlSelfRef.a16) 
                                                    << 0x00000010U) 
                                                   | (0x0000ffffU 
                                                      & (((IData)(vlSelfRef.a8) 
                                                          << 8U) 
                                                         | (IData)(vlSelfRef.a8))));
                    vlSelfRef.t01_alu__DOT__r64 = (
                                                   (((QData)((IData)(vlSelfRef.a32)) 
                                                     << 0x00000020U) 
                                                    | (QData)((IData)(vlSelfRef.a32))) 
                                                   ^ 
                                                   (((QData)((IData)(vlSelfRef.b32)) 
                                                     << 0x00000020U) 
                                                    | (QData)((IData)(vlSelfRef.a32))));
                } else {
                    vlSelfRef.t01_alu__DOT__r32 = VL_SHIFTRS_III(32,32,6, vlSelfRef.a32, 
                                                                 (0x0000003fU 
                                                                  & vlSelfRef.b32));
                    vlSelfRef.t01_alu__DOT__r64 = VL_SHIFTRS_QQI(64,64,6, vlSelfRef.a64, 
                                                                 (0x0000003fU 
                                                                  & vlSelfRef.b32));
                }
            } else if ((1U & (IData)(vlSelfRef.op))) {
                vlSelfRef.t01_alu__DOT__r32 = VL_SHIFTR_III(32,32,6, vlSelfRef.a32, 
                                                            (0x0000003fU 
                                                             & vlSelfRef.b32));
                vlSelfRef.t01_alu__DOT__r64 = (vlSelfRef.a64 
                                               >> (0x0000003fU 
                                                   & vlSelfRef.b32));
            } else {
                vlSelfRef.t01_alu__DOT__r32 = VL_SHIFTL_III(32,32,6, vlSelfRef.a32, 
                                                            (0x0000003fU 
                                                             & vlSelfRef.b32));
                vlSelfRef.t01_alu__DOT__r64 = (vlSelfRef.a64 
                                               << (0x0000003fU 
                                                   & vlSelfRef.b32));
            }
        } else if ((4U & (IData)(vlSelfRef.op))) {
            if ((2U & (IData)(vlSelfRef.op))) {
                if ((1U & (IData)(vlSelfRef.op))) {
                    vlSelfRef.t01_alu__DOT__r32 = VL_MULS_III(32, vlSelfRef.a32, vlSelfRef.b32);
                    vlSelfRef.t01_alu__DOT__r64 = VL_MULS_QQQ(64, vlSelfRef.a64, vlSelfRef.b64);
                } else {
                    vlSelfRef.t01_alu__DOT__r32 = VL_MODDIVS_III(32, vlSelfRef.a32, vlSelfRef.b32);
                    vlSelfRef.t01_alu__DOT__r64 = VL_MODDIVS_QQQ(64, vlSelfRef.a64, vlSelfRef.b64);
                }
            } else if ((1U & (IData)(vlSelfRef.op))) {
                vlSelfRef.t01_alu__DOT__r32 = VL_DIVS_III(32, vlSelfRef.a32, vlSelfRef.b32);
                vlSelfRef.t01_alu__DOT__r64 = VL_DIVS_QQQ(64, vlSelfRef.a64, vlSelfRef.b64);
            } else {
                vlSelfRef.t01_alu__DOT__r32 = VL_MODDIV_III(32, vlSelfRef.a32, vlSelfRef.b32);
                vlSelfRef.t01_alu__DOT__r64 = VL_MODDIV_QQQ(64, vlSelfRef.a64, vlSelfRef.b64);
            }
        } else if ((2U & (IData)(vlSelfRef.op))) {
            if ((1U & (IData)(vlSelfRef.op))) {
                vlSelfRef.t01_alu__DOT__r32 = VL_DIV_III(32, vlSelfRef.a32, vlSelfRef.b32);
                vlSelfRef.t01_alu__DOT__r64 = VL_DIV_QQQ(64, vlSelfRef.a64, vlSelfRef.b64);
            } else {
                vlSelfRef.t01_alu__DOT__r32 = (vlSelfRef.a32 
                                               * vlSelfRef.b32);
                vlSelfRef.t01_alu__DOT__r64 = (vlSelfRef.a64 
                                               * vlSelfRef.b64);
            }
        } else if ((1U & (IData)(vlSelfRef.op))) {
            vlSelfRef.t01_alu__DOT__r32 = (vlSelfRef.a32 
                                           - vlSelfRef.b32);
            vlSelfRef.t01_alu__DOT__r64 = (vlSelfRef.a64 
                                           - vlSelfRef.b64);
        } else {
            vlSelfRef.t01_alu__DOT__r32 = (vlSelfRef.a32 
                                           + vlSelfRef.b32);
            vlSelfRef.t01_alu__DOT__r64 = (vlSelfRef.a64 
                                           + vlSelfRef.b64);
        }
        vlSelfRef.flags = ((((((vlSelfRef.a32 == vlSelfRef.b32) 
                               << 3U) | ((vlSelfRef.a64 
                                          != vlSelfRef.b64) 
                                         << 2U)) | 
                             ((VL_LTES_III(32, vlSelfRef.a32, vlSelfRef.b32) 
                               << 1U) | VL_GTS_IQQ(64, vlSelfRef.a64, vlSelfRef.b64))) 
                            << 3U) | ((((IData)(vlSelfRef.a16) 
                                        > (0x0000ffffU 
                                           & vlSelfRef.b32)) 
                                       << 2U) | ((((IData)(vlSelfRef.a8) 
                                                   >= 
                                                   (0x000000ffU 
                                                    & vlSelfRef.b32)) 
                                                  << 1U) 
                                                 | (1U 
                                                    & VL_REDXOR_64(vlSelfRef.t01_alu__DOT__r64)))));
    }

    VLCU_D static bool Vt01_alu___024root___eval_phase__stl(Vt01_alu___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt01_alu___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
        }
#endif
        __VstlExecute = Vt01_alu___024root___trigger_anySet__stl(vlSelfRef.__VstlTriggered);
        if (__VstlExecute) {
            {
                // Inlined CFunc: _eval_stl
                if ((1ULL & vlSelfRef.__VstlTriggered[0U])) {
                    Vt01_alu___024root___stl_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
        }
        return (__VstlExecute);
    }

    VLCU_D static bool Vt01_alu___024root___trigger_anySet__ico(const VlUnpacked<QData/*63:0*/, 2> &in) {
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

    VLCU_D static void Vt01_alu___024root___eval_ico(Vt01_alu___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        if ((0x000000000000003eULL & vlSelfRef.__VicoTriggered[0U])) {
            {
                // Inlined CFunc: _ico_comb__TOP__0
                vlSelfRef.t01_alu__DOT__r32 = ((8U 
                                                & (IData)(vlSelfRef.op))
                                                ? (
                                                   (4U 
                                                    & (IData)(vlSelfRef.op))
                                                    ? 
                                                   ((2U 
                                                     & (IData)(vlSelfRef.op))
                                                     ? 
                                                    ((1U 
                                                      & (IData)(vlSelfRef.op))
                                                      ? 
                                                     (- vlSelfRef.a32)
                                                      : 
                                                     ((1U 
                                                       & VL_REDXOR_32(vlSelfRef.a32)) 
                                                      + 
                                                      ((0U 
                                                        != (IData)(vlSelfRef.a16)) 
                                                       + 
                                                       (0x000000ffU 
                                                        == (IData)(vlSelfRef.a8)))))
                                                     : 
                                                    ((1U 
                                                      & (IData)(vlSelfRef.op))
                                                      ? 
                                                     ((vlSelfRef.a32 
                                                       < vlSelfRef.b32)
                                                       ? vlSelfRef.a32
                                                       : 
                                                      (VL_LTS_III(32, vlSelfRef.a32, vlSelfRef.b32)
                                                        ? vlSelfRef.b32
                                                        : 0xdeadbeefU))
                                                      : 
                                                     ((~ vlSelfRef.b32) 
                                                      & (((IData)(vlSelfRef.a8) 
                                                          << 0x00000018U) 
                                                         | (((IData)(vlSelfRef.a8) 
                                                             << 0x00000010U) 
                                                            | (((IData)(vlSelfRef.a8) 
                                                                << 8U) 
                                                               | (IData)(vlSelfRef.a8)))))))
                                                    : 
                                                   ((2U 
                                                     & (IData)(vlSelfRef.op))
                                                     ? 
                                                    ((1U 
                                                      & (IData)(vlSelfRef.op))
                                                      ? 
                                                     (((IData)(vlSelfRef.a16) 
                                                       << 0x00000010U) 
                                                      | (0x0000ffffU 
                                                         & (((IData)(vlSelfRef.a8) 
                                                             << 8U) 
                                                            | (IData)(vlSelfRef.a8))))
                                                      : 
                                                     VL_SHIFTRS_III(32,32,6, vlSelfRef.a32, 
                                                                    (0x0000003fU 
                                                                     & vlSelfRef.b32)))
                                                     : 
                                                    ((1U 
                                                      & (IData)(vlSelfRef.op))
                                                      ? 
                                                     VL_SHIFTR_III(32,32,6, vlSelfRef.a32, 
                                                                   (0x0000003fU 
                                                                    & vlSelfRef.b32))
                                                      : 
                                                     VL_SHIFTL_III(32,32,6, vlSelfRef.a32, 
                                                                   (0x0000003fU 
                                                                    & vlSelfRef.b32)))))
                                                : (
                                                   (4U 
                                                    & (IData)(vlSelfRef.op))
                                                    ? 
                                                   ((2U 
                                                     & (IData)(vlSelfRef.op))
                                                     ? 
                                                    ((1U 
                                                      & (IData)(vlSelfRef.op))
                                                      ? 
                                                     VL_MULS_III(32, vlSelfRef.a32, vlSelfRef.b32)
                                                      : 
                                                     VL_MODDIVS_III(32, vlSelfRef.a32, vlSelfRef.b32))
                                                     : 
                                                    ((1U 
                                                      & (IData)(vlSelfRef.op))
                                                      ? 
                                                     VL_DIVS_III(32, vlSelfRef.a32, vlSelfRef.b32)
                                                      : 
                                                     VL_MODDIV_III(32, vlSelfRef.a32, vlSelfRef.b32)))
                                                    : 
                                                   ((2U 
                                                     & (IData)(vlSelfRef.op))
                                                     ? 
                                                    ((1U 
                                                      & (IData)(vlSelfRef.op))
                                                      ? 
                                                     VL_DIV_III(32, vlSelfRef.a32, vlSelfRef.b32)
                                                      : 
                                                     (vlSelfRef.a32 
                                                      * vlSelfRef.b32))
                                                     : 
                                                    ((1U 
                                                      & (IData)(vlSelfRef.op))
                                                      ? 
                                                     (vlSelfRef.a32 
                                                      - vlSelfRef.b32)
                                                      : 
                                                     (vlSelfRef.a32 
                                                      + vlSelfRef.b32)))));
            }
        }
        if ((0x00000000000000f6ULL & vlSelfRef.__VicoTriggered[0U])) {
            {
                // Inlined CFunc: _ico_comb__TOP__1
                vlSelfRef.t01_alu__DOT__r64 = ((8U 
                                                & (IData)(vlSelfRef.op))
                                                ? (
                                                   (4U 
                                                    & (IData)(vlSelfRef.op))
                                                    ? 
                                                   ((2U 
                                                     & (IData)(vlSelfRef.op))
                                                     ? 
                                                    ((1U 
                                                      & (IData)(vlSelfRef.op))
                                                      ? 
                                                     ((~ vlSelfRef.a64) 
                                                      + (QData)((IData)(
                                                                        (1U 
                                                                         & (IData)(vlSelfRef.a8)))))
                                                      : 
                                                     ((QData)((IData)(
                                                                      (vlSelfRef.a64 
                                                                       >> 
                                                                       (0x0000001fU 
                                                                        & vlSelfRef.b32)))) 
                                                      + (QData)((IData)(
                                                                        (1U 
                                                                         & (IData)(
                                                                                (vlSelfRef.a64 
                                                                                >> 
                                                                                (0x0000003fU 
                                                                                & vlSelfRef.b32))))))))
                                                     : 
                                                    ((1U 
                                                      & (IData)(vlSelfRef.op))
                                                      ? 
                                                     (VL_GTES_IQQ(64, vlSelfRef.a64, vlSelfRef.b64)
                                                       ? 
                                                      (vlSelfRef.a64 
                                                       | vlSelfRef.b64)
                                                       : 
                                                      (vlSelfRef.a64 
                                                       & vlSelfRef.b64))
                                                      : 
                                                     (((QData)((IData)(vlSelfRef.a64)) 
                                                       << 0x00000020U) 
                                                      | (QData)((IData)(
                                                                        (vlSelfRef.a64 
                                                                         >> 0x00000020U))))))
                                                    : 
                                                   ((2U 
                                                     & (IData)(vlSelfRef.op))
                                                     ? 
                                                    ((1U 
                                                      & (IData)(vlSelfRef.op))
                                                      ? 
                                                     ((((QData)((IData)(vlSelfRef.a32)) 
                                                        << 0x00000020U) 
                                                       | (QData)((IData)(vlSelfRef.a32))) 
                                                      ^ 
                                                      (((QData)((IData)(vlSelfRef.b32)) 
                                                        << 0x00000020U) 
                                                       | (QData)((IData)(vlSelfRef.a32))))
                                                      : 
                                                     VL_SHIFTRS_QQI(64,64,6, vlSelfRef.a64, 
                                                                    (0x0000003fU 
                                                                     & vlSelfRef.b32)))
                                                     : 
                                                    ((1U 
                                                      & (IData)(vlSelfRef.op))
                                                      ? 
                                                     (vlSelfRef.a64 
                                                      >> 
                                                      (0x0000003fU 
                                                       & vlSelfRef.b32))
                                                      : 
                                                     (vlSelfRef.a64 
                                                      << 
                                                      (0x0000003fU 
                                                       & vlSelfRef.b32)))))
                                                : (
                                                   (4U 
                                                    & (IData)(vlSelfRef.op))
                                                    ? 
                                                   ((2U 
                                                     & (IData)(vlSelfRef.op))
                                                     ? 
                                                    ((1U 
                                                      & (IData)(vlSelfRef.op))
                                                      ? 
                                                     VL_MULS_QQQ(64, vlSelfRef.a64, vlSelfRef.b64)
                                                      : 
                                                     VL_MODDIVS_QQQ(64, vlSelfRef.a64, vlSelfRef.b64))
                                                     : 
                                                    ((1U 
                                                      & (IData)(vlSelfRef.op))
                                                      ? 
                                                     VL_DIVS_QQQ(64, vlSelfRef.a64, vlSelfRef.b64)
                                                      : 
                                                     VL_MODDIV_QQQ(64, vlSelfRef.a64, vlSelfRef.b64)))
                                                    : 
                                                   ((2U 
                                                     & (IData)(vlSelfRef.op))
                                                     ? 
                                                    ((1U 
                                                      & (IData)(vlSelfRef.op))
                                                      ? 
                                                     VL_DIV_QQQ(64, vlSelfRef.a64, vlSelfRef.b64)
                                                      : 
                                                     (vlSelfRef.a64 
                                                      * vlSelfRef.b64))
                                                     : 
                                                    ((1U 
                                                      & (IData)(vlSelfRef.op))
                                                      ? 
                                                     (vlSelfRef.a64 
                                                      - vlSelfRef.b64)
                                                      : 
                                                     (vlSelfRef.a64 
                                                      + vlSelfRef.b64)))));
            }
        }
        if ((0x00000000000000feULL & vlSelfRef.__VicoTriggered[0U])) {
            {
                // Inlined CFunc: _ico_comb__TOP__2
                vlSelfRef.flags = ((((((vlSelfRef.a32 
                                        == vlSelfRef.b32) 
                                       << 3U) | ((vlSelfRef.a64 
                                                  != vlSelfRef.b64) 
                                                 << 2U)) 
                                     | ((VL_LTES_III(32, vlSelfRef.a32, vlSelfRef.b32) 
                                         << 1U) | VL_GTS_IQQ(64, vlSelfRef.a64, vlSelfRef.b64))) 
                                    << 3U) | ((((IData)(vlSelfRef.a16) 
                                                > (0x0000ffffU 
                                                   & vlSelfRef.b32)) 
                                               << 2U) 
                                              | ((((IData)(vlSelfRef.a8) 
                                                   >= 
                                                   (0x000000ffU 
                                                    & vlSelfRef.b32)) 
                                                  << 1U) 
                                                 | (1U 
                                                    & VL_REDXOR_64(vlSelfRef.t01_alu__DOT__r64)))));
            }
        }
    }

    VLCU_D static bool Vt01_alu___024root___eval_phase__ico(Vt01_alu___024root* vlSelf, VlcuContext& vlCtx) {
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
                                                            ((((((vlSelfRef.b64 
                                                                  != vlSelfRef.__Vtrigprevexpr___TOP__b64__0) 
                                                                 << 3U) 
                                                                | ((vlSelfRef.a64 
                                                                    != vlSelfRef.__Vtrigprevexpr___TOP__a64__0) 
                                                                   << 2U)) 
                                                               | (((vlSelfRef.b32 
                                                                    != vlSelfRef.__Vtrigprevexpr___TOP__b32__0) 
                                                                   << 1U) 
                                                                  | (vlSelfRef.a32 
                                                                     != vlSelfRef.__Vtrigprevexpr___TOP__a32__0))) 
                                                              << 4U) 
                                                             | (((((IData)(vlSelfRef.a16) 
                                                                   != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__a16__0)) 
                                                                  << 3U) 
                                                                 | (((IData)(vlSelfRef.a8) 
                                                                     != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__a8__0)) 
                                                                    << 2U)) 
                                                                | ((((IData)(vlSelfRef.op) 
                                                                     != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__op__0)) 
                                                                    << 1U) 
                                                                   | ((IData)(vlSelfRef.clk) 
                                                                      != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__clk__0)))))));
            vlSelfRef.__Vtrigprevexpr___TOP__clk__0 
                = vlSelfRef.clk;
            vlSelfRef.__Vtrigprevexpr___TOP__op__0 
                = vlSelfRef.op;
            vlSelfRef.__Vtrigprevexpr___TOP__a8__0 
                = vlSelfRef.a8;
            vlSelfRef.__Vtrigprevexpr___TOP__a16__0 
                = vlSelfRef.a16;
            vlSelfRef.__Vtrigprevexpr___TOP__a32__0 
                = vlSelfRef.a32;
            vlSelfRef.__Vtrigprevexpr___TOP__b32__0 
                = vlSelfRef.b32;
            vlSelfRef.__Vtrigprevexpr___TOP__a64__0 
                = vlSelfRef.a64;
            vlSelfRef.__Vtrigprevexpr___TOP__b64__0 
                = vlSelfRef.b64;
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
                vlSelfRef.__VicoTriggered[0U] = (0x0000000000000080ULL 
                                                 | vlSelfRef.__VicoTriggered[0U]);
            }
        }
#ifdef VL_DEBUG
        if (VL_UNLIKELY(vlSymsp->_vm_contextp__->debug())) {
            Vt01_alu___024root___dump_triggers__ico(vlSelfRef.__VicoTriggered, "ico"s);
        }
#endif
        __VicoExecute = Vt01_alu___024root___trigger_anySet__ico(vlSelfRef.__VicoTriggered);
        if (__VicoExecute) {
            Vt01_alu___024root___eval_ico(vlSelf, vlCtx);
        }
        return (__VicoExecute);
    }

    VLCU_D static bool Vt01_alu___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt01_alu___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt01_alu___024root___eval_phase__act(Vt01_alu___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt01_alu___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vt01_alu___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vt01_alu___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vt01_alu___024root___eval_phase__nba(Vt01_alu___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vt01_alu___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    {
                        // Inlined CFunc: _nba_sequent__TOP__0
                        vlSelfRef.q32 = vlSelfRef.t01_alu__DOT__r32;
                        vlSelfRef.q64 = vlSelfRef.t01_alu__DOT__r64;
                        vlSelfRef.q16 = (0x0000ffffU 
                                         & ((vlSelfRef.t01_alu__DOT__r32 
                                             ^ (IData)(
                                                       (vlSelfRef.t01_alu__DOT__r64 
                                                        >> 0x20U))) 
                                            ^ (IData)(vlSelfRef.a16)));
                        vlSelfRef.q8 = (0x000000ffU 
                                        & ((vlSelfRef.t01_alu__DOT__r32 
                                            + (IData)(
                                                      (vlSelfRef.t01_alu__DOT__r64 
                                                       >> 0x38U))) 
                                           + (IData)(vlSelfRef.a8)));
                    }
                }
            }
            Vt01_alu___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vt01_alu___024root___eval(Vt01_alu___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt01_alu___024root___dump_triggers__ico(vlSelfRef.__VicoTriggered, "ico"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t01_alu.v", 4, "", "DIDNOTCONVERGE: Input combinational region did not converge after '--converge-limit' of 10000 tries");
            }
            __VicoIterCount = ((IData)(1U) + __VicoIterCount);
            vlSelfRef.__VicoPhaseResult = Vt01_alu___024root___eval_phase__ico(vlSelf, vlCtx);
        } while (vlSelfRef.__VicoPhaseResult);
        __VnbaIterCount = 0U;
        do {
            if (VL_UNLIKELY(((0x00002710U < __VnbaIterCount)))) {
#ifdef VL_DEBUG
                Vt01_alu___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t01_alu.v", 4, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vt01_alu___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t01_alu.v", 4, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vt01_alu___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vt01_alu___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vt01_alu___024root___ctor_var_reset(Vt01_alu___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        vlSelf->op = VL_SCOPED_RAND_RESET_I(4, __VscopeHash, 3630531923276091163ull);
        vlSelf->a8 = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 1821884981173182555ull);
        vlSelf->a16 = VL_SCOPED_RAND_RESET_I(16, __VscopeHash, 8457099364513313400ull);
        vlSelf->a32 = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 8182655363289387084ull);
        vlSelf->b32 = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 512281595318465383ull);
        vlSelf->a64 = VL_SCOPED_RAND_RESET_Q(64, __VscopeHash, 10796914987831142057ull);
        vlSelf->b64 = VL_SCOPED_RAND_RESET_Q(64, __VscopeHash, 9340292202783580658ull);
        vlSelf->q8 = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 4048769987926657425ull);
        vlSelf->q16 = VL_SCOPED_RAND_RESET_I(16, __VscopeHash, 8973587424176523318ull);
        vlSelf->q32 = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 16465607520851010179ull);
        vlSelf->q64 = VL_SCOPED_RAND_RESET_Q(64, __VscopeHash, 2375150294754472520ull);
        vlSelf->flags = VL_SCOPED_RAND_RESET_I(7, __VscopeHash, 7921719909686688401ull);
        vlSelf->t01_alu__DOT__r32 = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 7651387025879530697ull);
        vlSelf->t01_alu__DOT__r64 = VL_SCOPED_RAND_RESET_Q(64, __VscopeHash, 11015893840876120400ull);
        for (int __Vi0 = 0; __Vi0 < 1; ++__Vi0) {
            vlSelf->__VstlTriggered[__Vi0] = 0;
        }
        for (int __Vi0 = 0; __Vi0 < 2; ++__Vi0) {
            vlSelf->__VicoTriggered[__Vi0] = 0;
        }
        vlSelf->__Vtrigprevexpr___TOP__clk__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__op__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__a8__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__a16__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__a32__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__b32__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__a64__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__b64__0 = 0;
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
        Vt01_alu___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt01_alu___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt01_alu___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt01_alu___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vt01_alu___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt01_alu___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt01_alu__cuda::info() {
    typedef Vt01_alu__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"op", "op", VLCU_MEMBER_IN, 1, 0, 1, 4, 3, 0, T::OFF_op},
        {"a8", "a8", VLCU_MEMBER_IN, 1, 0, 1, 8, 7, 0, T::OFF_a8},
        {"a16", "a16", VLCU_MEMBER_IN, 2, 0, 1, 16, 15, 0, T::OFF_a16},
        {"a32", "a32", VLCU_MEMBER_IN, 4, 0, 1, 32, 31, 0, T::OFF_a32},
        {"b32", "b32", VLCU_MEMBER_IN, 4, 0, 1, 32, 31, 0, T::OFF_b32},
        {"a64", "a64", VLCU_MEMBER_IN, 8, 0, 1, 64, 63, 0, T::OFF_a64},
        {"b64", "b64", VLCU_MEMBER_IN, 8, 0, 1, 64, 63, 0, T::OFF_b64},
        {"q8", "q8", VLCU_MEMBER_OUT, 1, 0, 1, 8, 7, 0, T::OFF_q8},
        {"q16", "q16", VLCU_MEMBER_OUT, 2, 0, 1, 16, 15, 0, T::OFF_q16},
        {"q32", "q32", VLCU_MEMBER_OUT, 4, 0, 1, 32, 31, 0, T::OFF_q32},
        {"q64", "q64", VLCU_MEMBER_OUT, 8, 0, 1, 64, 63, 0, T::OFF_q64},
        {"flags", "flags", VLCU_MEMBER_OUT, 1, 0, 1, 7, 6, 0, T::OFF_flags},
        {"__VstlFirstIteration", "__VstlFirstIteration", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlFirstIteration},
        {"__VstlPhaseResult", "__VstlPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__Vtrigprevexpr___TOP__op__0", "__Vtrigprevexpr___TOP__op__0", VLCU_MEMBER_STATE, 1, 0, 1, 4, 3, 0, T::OFF___Vtrigprevexpr___TOP__op__0},
        {"__Vtrigprevexpr___TOP__a8__0", "__Vtrigprevexpr___TOP__a8__0", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF___Vtrigprevexpr___TOP__a8__0},
        {"__VicoDidInit", "__VicoDidInit", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VicoDidInit},
        {"__VicoPhaseResult", "__VicoPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VicoPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__1", "__Vtrigprevexpr___TOP__clk__1", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__1},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"__Vtrigprevexpr___TOP__a16__0", "__Vtrigprevexpr___TOP__a16__0", VLCU_MEMBER_STATE, 2, 0, 1, 16, 15, 0, T::OFF___Vtrigprevexpr___TOP__a16__0},
        {"t01_alu__DOT__r32", "t01_alu.r32", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF_t01_alu__DOT__r32},
        {"__Vtrigprevexpr___TOP__a32__0", "__Vtrigprevexpr___TOP__a32__0", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___Vtrigprevexpr___TOP__a32__0},
        {"__Vtrigprevexpr___TOP__b32__0", "__Vtrigprevexpr___TOP__b32__0", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___Vtrigprevexpr___TOP__b32__0},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"t01_alu__DOT__r64", "t01_alu.r64", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF_t01_alu__DOT__r64},
        {"__Vtrigprevexpr___TOP__a64__0", "__Vtrigprevexpr___TOP__a64__0", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___Vtrigprevexpr___TOP__a64__0},
        {"__Vtrigprevexpr___TOP__b64__0", "__Vtrigprevexpr___TOP__b64__0", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___Vtrigprevexpr___TOP__b64__0},
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
    static const uint32_t saveOrder[] = {0, 1, 2, 8, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 3, 9, 23, 4, 5, 10, 24, 25, 26, 27, 6, 7, 11, 28, 29, 30, 31, 32, 33, 34, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "t01_alu", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 0, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 35, 0, 0x7c8d40b046c830b9ULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt01_alu__cuda VLCU_TESTBENCHES);
