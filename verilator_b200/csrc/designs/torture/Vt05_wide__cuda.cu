// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt05_wide"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt05_wide___024root final {
    CData/*0:0*/ clk;
    VlWide<3>/*95:0*/ a;
    VlWide<3>/*95:0*/ b;
    VlWide<5>/*129:0*/ c;
    CData/*1:0*/ sel;
    VlWide<3>/*95:0*/ q;
    VlWide<5>/*129:0*/ r;
    CData/*3:0*/ flags;
    VlWide<4>/*99:0*/ ctr;
    CData/*0:0*/ __VstlFirstIteration;
    CData/*0:0*/ __VstlPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*1:0*/ __Vtrigprevexpr___TOP__sel__0;
    CData/*0:0*/ __VicoDidInit;
    CData/*0:0*/ __VicoPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__1;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    VlWide<3>/*95:0*/ __Vtrigprevexpr___TOP__a__0;
    VlWide<3>/*95:0*/ __Vtrigprevexpr___TOP__b__0;
    VlWide<5>/*129:0*/ __Vtrigprevexpr___TOP__c__0;
    IData/*31:0*/ __VactIterCount;
    VlUnpacked<QData/*63:0*/, 1> __VstlTriggered;
    VlUnpacked<QData/*63:0*/, 2> __VicoTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vt05_wide__cuda final {
    typedef Vt05_wide___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 180,
        OFF_a = 40,
        OFF_b = 52,
        OFF_c = 64,
        OFF_sel = 181,
        OFF_q = 84,
        OFF_r = 96,
        OFF_flags = 182,
        OFF_ctr = 116,
        OFF___VstlFirstIteration = 183,
        OFF___VstlPhaseResult = 184,
        OFF___Vtrigprevexpr___TOP__clk__0 = 185,
        OFF___Vtrigprevexpr___TOP__sel__0 = 186,
        OFF___VicoDidInit = 187,
        OFF___VicoPhaseResult = 188,
        OFF___Vtrigprevexpr___TOP__clk__1 = 189,
        OFF___VactPhaseResult = 190,
        OFF___VnbaPhaseResult = 191,
        OFF___Vtrigprevexpr___TOP__a__0 = 132,
        OFF___Vtrigprevexpr___TOP__b__0 = 144,
        OFF___Vtrigprevexpr___TOP__c__0 = 156,
        OFF___VactIterCount = 176,
        OFF___VstlTriggered = 0,
        OFF___VicoTriggered = 8,
        OFF___VactTriggered = 24,
        OFF___VnbaTriggered = 32,
        kStateBytes = 192,  // S_persist
        kTouchedReadBytes = 161,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 136,  // S_state_wr
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
        st.ldW(OFF_a, i, s.a);
        st.ldW(OFF_b, i, s.b);
        st.ldW(OFF_c, i, s.c);
        s.sel = st.ld<VlcuMem8>(OFF_sel, i);
        st.ldW(OFF_q, i, s.q);
        st.ldW(OFF_r, i, s.r);
        s.flags = st.ld<VlcuMem8>(OFF_flags, i);
        st.ldW(OFF_ctr, i, s.ctr);
        s.__VstlFirstIteration = st.ld<VlcuMem8>(OFF___VstlFirstIteration, i);
        s.__VstlPhaseResult = st.ld<VlcuMem8>(OFF___VstlPhaseResult, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__Vtrigprevexpr___TOP__sel__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__sel__0, i);
        s.__VicoDidInit = st.ld<VlcuMem8>(OFF___VicoDidInit, i);
        s.__VicoPhaseResult = st.ld<VlcuMem8>(OFF___VicoPhaseResult, i);
        s.__Vtrigprevexpr___TOP__clk__1 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__1, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        st.ldW(OFF___Vtrigprevexpr___TOP__a__0, i, s.__Vtrigprevexpr___TOP__a__0);
        st.ldW(OFF___Vtrigprevexpr___TOP__b__0, i, s.__Vtrigprevexpr___TOP__b__0);
        st.ldW(OFF___Vtrigprevexpr___TOP__c__0, i, s.__Vtrigprevexpr___TOP__c__0);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VstlTriggered[e] = st.ldE<QData>(OFF___VstlTriggered, e, i);
        for (uint32_t e = 0; e < 2; ++e) s.__VicoTriggered[e] = st.ldE<QData>(OFF___VicoTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.stW(OFF_a, i, s.a);
        st.stW(OFF_b, i, s.b);
        st.stW(OFF_c, i, s.c);
        st.st<VlcuMem8>(OFF_sel, i, s.sel);
        st.stW(OFF_q, i, s.q);
        st.stW(OFF_r, i, s.r);
        st.st<VlcuMem8>(OFF_flags, i, s.flags);
        st.stW(OFF_ctr, i, s.ctr);
        st.st<VlcuMem8>(OFF___VstlFirstIteration, i, s.__VstlFirstIteration);
        st.st<VlcuMem8>(OFF___VstlPhaseResult, i, s.__VstlPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__sel__0, i, s.__Vtrigprevexpr___TOP__sel__0);
        st.st<VlcuMem8>(OFF___VicoDidInit, i, s.__VicoDidInit);
        st.st<VlcuMem8>(OFF___VicoPhaseResult, i, s.__VicoPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__1, i, s.__Vtrigprevexpr___TOP__clk__1);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.stW(OFF___Vtrigprevexpr___TOP__a__0, i, s.__Vtrigprevexpr___TOP__a__0);
        st.stW(OFF___Vtrigprevexpr___TOP__b__0, i, s.__Vtrigprevexpr___TOP__b__0);
        st.stW(OFF___Vtrigprevexpr___TOP__c__0, i, s.__Vtrigprevexpr___TOP__c__0);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VstlTriggered, e, i, s.__VstlTriggered[e]);
        for (uint32_t e = 0; e < 2; ++e) st.stE<QData>(OFF___VicoTriggered, e, i, s.__VicoTriggered[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VactTriggered, e, i, s.__VactTriggered[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VnbaTriggered, e, i, s.__VnbaTriggered[e]);
    }
    VLCU_D static void load_inputs_row(Root& s, const uint32_t* rowp, uint64_t stride, uint64_t i) {
        s.clk = static_cast<VlcuMem8>(rowp[0 * stride + i] & 0x1U);
        s.a[0] = rowp[1 * stride + i] & 0xffffffffU;
        s.a[1] = rowp[2 * stride + i] & 0xffffffffU;
        s.a[2] = rowp[3 * stride + i] & 0xffffffffU;
        s.b[0] = rowp[4 * stride + i] & 0xffffffffU;
        s.b[1] = rowp[5 * stride + i] & 0xffffffffU;
        s.b[2] = rowp[6 * stride + i] & 0xffffffffU;
        s.c[0] = rowp[7 * stride + i] & 0xffffffffU;
        s.c[1] = rowp[8 * stride + i] & 0xffffffffU;
        s.c[2] = rowp[9 * stride + i] & 0xffffffffU;
        s.c[3] = rowp[10 * stride + i] & 0xffffffffU;
        s.c[4] = rowp[11 * stride + i] & 0x3U;
        s.sel = static_cast<VlcuMem8>(rowp[12 * stride + i] & 0x3U);
    }
    VLCU_D static void store_outputs_row(const Root& s, uint32_t* rowp, uint64_t stride, uint64_t i) {
        rowp[0 * stride + i] = s.q[0];
        rowp[1 * stride + i] = s.q[1];
        rowp[2 * stride + i] = s.q[2];
        rowp[3 * stride + i] = s.r[0];
        rowp[4 * stride + i] = s.r[1];
        rowp[5 * stride + i] = s.r[2];
        rowp[6 * stride + i] = s.r[3];
        rowp[7 * stride + i] = s.r[4];
        rowp[8 * stride + i] = s.flags;
        rowp[9 * stride + i] = s.ctr[0];
        rowp[10 * stride + i] = s.ctr[1];
        rowp[11 * stride + i] = s.ctr[2];
        rowp[12 * stride + i] = s.ctr[3];
    }
    VLCU_D static uint64_t outputs_sig(const Root& s, uint64_t sig) {
        uint32_t acc = VLCU_SIG_ACC_INIT;
        acc = vlcu_sig_word(acc, s.q[0]);
        acc = vlcu_sig_word(acc, s.q[1]);
        acc = vlcu_sig_word(acc, s.q[2]);
        acc = vlcu_sig_word(acc, s.r[0]);
        acc = vlcu_sig_word(acc, s.r[1]);
        acc = vlcu_sig_word(acc, s.r[2]);
        acc = vlcu_sig_word(acc, s.r[3]);
        acc = vlcu_sig_word(acc, s.r[4]);
        acc = vlcu_sig_word(acc, s.flags);
        acc = vlcu_sig_word(acc, s.ctr[0]);
        acc = vlcu_sig_word(acc, s.ctr[1]);
        acc = vlcu_sig_word(acc, s.ctr[2]);
        acc = vlcu_sig_word(acc, s.ctr[3]);
        return vlcu_sig_mix(sig, acc);
    }
    enum : uint32_t { kInWords = 13, kOutWords = 13, kInBytes = 46, kOutBytes = 49, kTraceWords = 0 };

    VLCU_D static void Vt05_wide___024root___eval_static(Vt05_wide___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
        vlSelfRef.__Vtrigprevexpr___TOP__a__0[0U] = vlSelfRef.a[0U];
        vlSelfRef.__Vtrigprevexpr___TOP__a__0[1U] = vlSelfRef.a[1U];
        vlSelfRef.__Vtrigprevexpr___TOP__a__0[2U] = vlSelfRef.a[2U];
        vlSelfRef.__Vtrigprevexpr___TOP__b__0[0U] = vlSelfRef.b[0U];
        vlSelfRef.__Vtrigprevexpr___TOP__b__0[1U] = vlSelfRef.b[1U];
        vlSelfRef.__Vtrigprevexpr___TOP__b__0[2U] = vlSelfRef.b[2U];
        vlSelfRef.__Vtrigprevexpr___TOP__c__0[0U] = vlSelfRef.c[0U];
        vlSelfRef.__Vtrigprevexpr___TOP__c__0[1U] = vlSelfRef.c[1U];
        vlSelfRef.__Vtrigprevexpr___TOP__c__0[2U] = vlSelfRef.c[2U];
        vlSelfRef.__Vtrigprevexpr___TOP__c__0[3U] = vlSelfRef.c[3U];
        vlSelfRef.__Vtrigprevexpr___TOP__c__0[4U] = vlSelfRef.c[4U];
        vlSelfRef.__Vtrigprevexpr___TOP__sel__0 = vlSelfRef.sel;
        vlSelfRef.__Vtrigprevexpr___TOP__clk__1 = vlSelfRef.clk;
    }

    VLCU_D static void Vt05_wide___024root___eval_initial(Vt05_wide___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt05_wide___024root___eval_final(Vt05_wide___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt05_wide___024root___eval_settle(Vt05_wide___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt05_wide___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t05_wide.v", 4, "", "DIDNOTCONVERGE: Settle region did not converge after '--converge-limit' of 10000 tries");
            }
            __VstlIterCount = ((IData)(1U) + __VstlIterCount);
            vlSelfRef.__VstlPhaseResult = Vt05_wide___024root___eval_phase__stl(vlSelf, vlCtx);
            vlSelfRef.__VstlFirstIteration = 0U;
        } while (vlSelfRef.__VstlPhaseResult);
    }

    VLCU_D static bool Vt05_wide___024root___trigger_anySet__stl(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt05_wide___024root___eval_phase__stl(Vt05_wide___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt05_wide___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
        }
#endif
        __VstlExecute = Vt05_wide___024root___trigger_anySet__stl(vlSelfRef.__VstlTriggered);
        if (__VstlExecute) {
            {
                // Inlined CFunc: _eval_stl
                if ((1ULL & vlSelfRef.__VstlTriggered[0U])) {
                    {
                        // Inlined CFunc: _ico_comb__TOP__0
                        vlSelfRef.flags = ((((0U == 
                                              (((vlSelfRef.a[0U] 
                                                 ^ vlSelfRef.b[0U]) 
                                                | (vlSelfRef.a[1U] 
                                                   ^ vlSelfRef.b[1U])) 
                                               | (vlSelfRef.a[2U] 
                                                  ^ vlSelfRef.b[2U]))) 
                                             << 3U) 
                                            | ((0xffffffffU 
                                                == 
                                                ((((vlSelfRef.c[0U] 
                                                    & vlSelfRef.c[1U]) 
                                                   & vlSelfRef.c[2U]) 
                                                  & vlSelfRef.c[3U]) 
                                                 & (0xfffffffcU 
                                                    | vlSelfRef.c[4U]))) 
                                               << 2U)) 
                                           | ((2U & 
                                               (VL_REDXOR_32(
                                                             ((((vlSelfRef.c[0U] 
                                                                 ^ vlSelfRef.c[1U]) 
                                                                ^ vlSelfRef.c[2U]) 
                                                               ^ vlSelfRef.c[3U]) 
                                                              ^ vlSelfRef.c[4U])) 
                                                << 1U)) 
                                              | VL_LTS_IWW(96, vlSelfRef.a, vlSelfRef.b)));
                    }
                }
            }
        }
        return (__VstlExecute);
    }

    VLCU_D static void Vt05_wide___024root___eval_triggers_vec__ico(Vt05_wide___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlSelfRef.__VicoTriggered[0U] = (QData)((IData)(
                                                        ((((IData)(vlSelfRef.sel) 
                                                           != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__sel__0)) 
                                                          << 4U) 
                                                         | ((((0U 
                                                               != 
                                                               (((((vlSelfRef.c[0U] 
                                                                    ^ vlSelfRef.__Vtrigprevexpr___TOP__c__0[0U]) 
                                                                   | (vlSelfRef.c[1U] 
                                                                      ^ vlSelfRef.__Vtrigprevexpr___TOP__c__0[1U])) 
                                                                  | (vlSelfRef.c[2U] 
                                                                     ^ vlSelfRef.__Vtrigprevexpr___TOP__c__0[2U])) 
                                                                 | (vlSelfRef.c[3U] 
                                                                    ^ vlSelfRef.__Vtrigprevexpr___TOP__c__0[3U])) 
                                                                | (vlSelfRef.c[4U] 
                                                                   ^ vlSelfRef.__Vtrigprevexpr___TOP__c__0[4U]))) 
                                                              << 3U) 
                                                             | ((0U 
                                                                 != 
                                                                 (((vlSelfRef.b[0U] 
                                                                    ^ vlSelfRef.__Vtrigprevexpr___TOP__b__0[0U]) 
                                                                   | (vlSelfRef.b[1U] 
                                                                      ^ vlSelfRef.__Vtrigprevexpr___TOP__b__0[1U])) 
                                                                  | (vlSelfRef.b[2U] 
                                                                     ^ vlSelfRef.__Vtrigprevexpr___TOP__b__0[2U]))) 
                                                                << 2U)) 
                                                            | (((0U 
                                                                 != 
                                                                 (((vlSelfRef.a[0U] 
                                                                    ^ vlSelfRef.__Vtrigprevexpr___TOP__a__0[0U]) 
                                                                   | (vlSelfRef.a[1U] 
                                                                      ^ vlSelfRef.__Vtrigprevexpr___TOP__a__0[1U])) 
                                                                  | (vlSelfRef.a[2U] 
                                                                     ^ vlSelfRef.__Vtrigprevexpr___TOP__a__0[2U]))) 
                                                                << 1U) 
                                                               | ((IData)(vlSelfRef.clk) 
                                                                  != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__clk__0)))))));
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
        vlSelfRef.__Vtrigprevexpr___TOP__a__0[0U] = vlSelfRef.a[0U];
        vlSelfRef.__Vtrigprevexpr___TOP__a__0[1U] = vlSelfRef.a[1U];
        vlSelfRef.__Vtrigprevexpr___TOP__a__0[2U] = vlSelfRef.a[2U];
        vlSelfRef.__Vtrigprevexpr___TOP__b__0[0U] = vlSelfRef.b[0U];
        vlSelfRef.__Vtrigprevexpr___TOP__b__0[1U] = vlSelfRef.b[1U];
        vlSelfRef.__Vtrigprevexpr___TOP__b__0[2U] = vlSelfRef.b[2U];
        vlSelfRef.__Vtrigprevexpr___TOP__c__0[0U] = vlSelfRef.c[0U];
        vlSelfRef.__Vtrigprevexpr___TOP__c__0[1U] = vlSelfRef.c[1U];
        vlSelfRef.__Vtrigprevexpr___TOP__c__0[2U] = vlSelfRef.c[2U];
        vlSelfRef.__Vtrigprevexpr___TOP__c__0[3U] = vlSelfRef.c[3U];
        vlSelfRef.__Vtrigprevexpr___TOP__c__0[4U] = vlSelfRef.c[4U];
        vlSelfRef.__Vtrigprevexpr___TOP__sel__0 = vlSelfRef.sel;
        if (VL_UNLIKELY(((1U & (~ (IData)(vlSelfRef.__VicoDidInit)))))) {
            vlSelfRef.__VicoDidInit = 1U;
            vlSelfRef.__VicoTriggered[0U] = (1ULL | vlSelfRef.__VicoTriggered[0U]);
            vlSelfRef.__VicoTriggered[0U] = (2ULL | vlSelfRef.__VicoTriggered[0U]);
            vlSelfRef.__VicoTriggered[0U] = (4ULL | vlSelfRef.__VicoTriggered[0U]);
            vlSelfRef.__VicoTriggered[0U] = (8ULL | vlSelfRef.__VicoTriggered[0U]);
            vlSelfRef.__VicoTriggered[0U] = (0x0000000000000010ULL 
                                             | vlSelfRef.__VicoTriggered[0U]);
        }
    }

    VLCU_D static bool Vt05_wide___024root___trigger_anySet__ico(const VlUnpacked<QData/*63:0*/, 2> &in) {
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

    VLCU_D static bool Vt05_wide___024root___eval_phase__ico(Vt05_wide___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VicoExecute;
        // Body
        Vt05_wide___024root___eval_triggers_vec__ico(vlSelf, vlCtx);
#ifdef VL_DEBUG
        if (VL_UNLIKELY(vlSymsp->_vm_contextp__->debug())) {
            Vt05_wide___024root___dump_triggers__ico(vlSelfRef.__VicoTriggered, "ico"s);
        }
#endif
        __VicoExecute = Vt05_wide___024root___trigger_anySet__ico(vlSelfRef.__VicoTriggered);
        if (__VicoExecute) {
            {
                // Inlined CFunc: _eval_ico
                if ((0x000000000000000eULL & vlSelfRef.__VicoTriggered[0U])) {
                    {
                        // Inlined CFunc: _ico_comb__TOP__0
                        vlSelfRef.flags = ((((0U == 
                                              (((vlSelfRef.a[0U] 
                                                 ^ vlSelfRef.b[0U]) 
                                                | (vlSelfRef.a[1U] 
                                                   ^ vlSelfRef.b[1U])) 
                                               | (vlSelfRef.a[2U] 
                                                  ^ vlSelfRef.b[2U]))) 
                                             << 3U) 
                                            | ((0xffffffffU 
                                                == 
                                                ((((vlSelfRef.c[0U] 
                                                    & vlSelfRef.c[1U]) 
                                                   & vlSelfRef.c[2U]) 
                                                  & vlSelfRef.c[3U]) 
                                                 & (0xfffffffcU 
                                                    | vlSelfRef.c[4U]))) 
                                               << 2U)) 
                                           | ((2U & 
                                               (VL_REDXOR_32(
                                                             ((((vlSelfRef.c[0U] 
                                                                 ^ vlSelfRef.c[1U]) 
                                                                ^ vlSelfRef.c[2U]) 
                                                               ^ vlSelfRef.c[3U]) 
                                                              ^ vlSelfRef.c[4U])) 
                                                << 1U)) 
                                              | VL_LTS_IWW(96, vlSelfRef.a, vlSelfRef.b)));
                    }
                }
            }
        }
        return (__VicoExecute);
    }

    VLCU_D static bool Vt05_wide___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt05_wide___024root___nba_sequent__TOP__0(Vt05_wide___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        VlWide<5>/*159:0*/ __Vtemp_2;
        VlWide<5>/*159:0*/ __Vtemp_3;
        VlWide<5>/*159:0*/ __Vtemp_5;
        VlWide<5>/*159:0*/ __Vtemp_6;
        VlWide<4>/*127:0*/ __Vtemp_9;
        VlWide<4>/*127:0*/ __Vtemp_10;
        VlWide<4>/*127:0*/ __Vtemp_11;
        VlWide<4>/*127:0*/ __Vtemp_12;
        // Body
        if ((1U & (IData)(vlSelfRef.sel))) {
            VL_SHIFTL_WWI(130,130,8, __Vtemp_3, vlSelfRef.c, 
                          (0x000000ffU & vlSelfRef.b[0U]));
            __Vtemp_2[0U] = (__Vtemp_3[0U] ^ vlSelfRef.q[0U]);
            __Vtemp_2[1U] = (__Vtemp_3[1U] ^ vlSelfRef.q[1U]);
            __Vtemp_2[2U] = (__Vtemp_3[2U] ^ vlSelfRef.q[2U]);
            __Vtemp_2[3U] = __Vtemp_3[3U];
            __Vtemp_2[4U] = __Vtemp_3[4U];
        } else {
            VL_SHIFTR_WWI(130,130,8, __Vtemp_5, vlSelfRef.c, 
                          (vlSelfRef.a[2U] >> 0x00000018U));
            __Vtemp_6[0U] = vlSelfRef.ctr[0U];
            __Vtemp_6[1U] = vlSelfRef.ctr[1U];
            __Vtemp_6[2U] = vlSelfRef.ctr[2U];
            __Vtemp_6[3U] = vlSelfRef.ctr[3U];
            __Vtemp_6[4U] = 0U;
            VL_ADD_W(5, __Vtemp_2, __Vtemp_5, __Vtemp_6);
        }
        vlSelfRef.r[0U] = __Vtemp_2[0U];
        vlSelfRef.r[1U] = __Vtemp_2[1U];
        vlSelfRef.r[2U] = __Vtemp_2[2U];
        vlSelfRef.r[3U] = __Vtemp_2[3U];
        vlSelfRef.r[4U] = (3U & __Vtemp_2[4U]);
        __Vtemp_9[0U] = 1U;
        __Vtemp_9[1U] = 0U;
        __Vtemp_9[2U] = 0U;
        __Vtemp_9[3U] = 0U;
        __Vtemp_10[0U] = vlSelfRef.a[0U];
        __Vtemp_10[1U] = vlSelfRef.a[1U];
        __Vtemp_10[2U] = vlSelfRef.a[2U];
        __Vtemp_10[3U] = 0U;
        VL_ADD_W(4, __Vtemp_11, vlSelfRef.ctr, __Vtemp_10);
        VL_ADD_W(4, __Vtemp_12, __Vtemp_9, __Vtemp_11);
        vlSelfRef.ctr[0U] = __Vtemp_12[0U];
        vlSelfRef.ctr[1U] = __Vtemp_12[1U];
        vlSelfRef.ctr[2U] = __Vtemp_12[2U];
        vlSelfRef.ctr[3U] = (0x0000000fU & __Vtemp_12[3U]);
        if ((0U == (IData)(vlSelfRef.sel))) {
            VL_DIV_WWW(96, vlSelfRef.q, vlSelfRef.a, vlSelfRef.b);
        } else if ((1U == (IData)(vlSelfRef.sel))) {
            VL_MODDIV_WWW(96, vlSelfRef.q, vlSelfRef.a, vlSelfRef.b);
        } else if ((2U == (IData)(vlSelfRef.sel))) {
            VL_DIVS_WWW(96, vlSelfRef.q, vlSelfRef.a, vlSelfRef.b);
        } else {
            VL_MUL_W(3, vlSelfRef.q, vlSelfRef.a, vlSelfRef.b);
        }
    }

    VLCU_D static void Vt05_wide___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt05_wide___024root___eval_phase__act(Vt05_wide___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt05_wide___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vt05_wide___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vt05_wide___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vt05_wide___024root___eval_phase__nba(Vt05_wide___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vt05_wide___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    Vt05_wide___024root___nba_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
            Vt05_wide___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vt05_wide___024root___eval(Vt05_wide___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt05_wide___024root___dump_triggers__ico(vlSelfRef.__VicoTriggered, "ico"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t05_wide.v", 4, "", "DIDNOTCONVERGE: Input combinational region did not converge after '--converge-limit' of 10000 tries");
            }
            __VicoIterCount = ((IData)(1U) + __VicoIterCount);
            vlSelfRef.__VicoPhaseResult = Vt05_wide___024root___eval_phase__ico(vlSelf, vlCtx);
        } while (vlSelfRef.__VicoPhaseResult);
        __VnbaIterCount = 0U;
        do {
            if (VL_UNLIKELY(((0x00002710U < __VnbaIterCount)))) {
#ifdef VL_DEBUG
                Vt05_wide___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t05_wide.v", 4, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vt05_wide___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t05_wide.v", 4, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vt05_wide___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vt05_wide___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vt05_wide___024root___ctor_var_reset(Vt05_wide___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        VL_SCOPED_RAND_RESET_W(96, vlSelf->a, __VscopeHash, 510903276987443985ull);
        VL_SCOPED_RAND_RESET_W(96, vlSelf->b, __VscopeHash, 16900879642891266615ull);
        VL_SCOPED_RAND_RESET_W(130, vlSelf->c, __VscopeHash, 15598372446745583797ull);
        vlSelf->sel = VL_SCOPED_RAND_RESET_I(2, __VscopeHash, 14660815779230110548ull);
        VL_SCOPED_RAND_RESET_W(96, vlSelf->q, __VscopeHash, 8861071527689086543ull);
        VL_SCOPED_RAND_RESET_W(130, vlSelf->r, __VscopeHash, 16978132545290669629ull);
        vlSelf->flags = VL_SCOPED_RAND_RESET_I(4, __VscopeHash, 7921719909686688401ull);
        VL_SCOPED_RAND_RESET_W(100, vlSelf->ctr, __VscopeHash, 1299185730960472065ull);
        for (int __Vi0 = 0; __Vi0 < 1; ++__Vi0) {
            vlSelf->__VstlTriggered[__Vi0] = 0;
        }
        for (int __Vi0 = 0; __Vi0 < 2; ++__Vi0) {
            vlSelf->__VicoTriggered[__Vi0] = 0;
        }
        vlSelf->__Vtrigprevexpr___TOP__clk__0 = 0;
        VL_ZERO_RESET_W(96, vlSelf->__Vtrigprevexpr___TOP__a__0);
        VL_ZERO_RESET_W(96, vlSelf->__Vtrigprevexpr___TOP__b__0);
        VL_ZERO_RESET_W(130, vlSelf->__Vtrigprevexpr___TOP__c__0);
        vlSelf->__Vtrigprevexpr___TOP__sel__0 = 0;
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
        Vt05_wide___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt05_wide___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt05_wide___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt05_wide___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vt05_wide___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt05_wide___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt05_wide__cuda::info() {
    typedef Vt05_wide__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"a", "a", VLCU_MEMBER_IN, 4, 0, 3, 96, 95, 0, T::OFF_a},
        {"b", "b", VLCU_MEMBER_IN, 4, 0, 3, 96, 95, 0, T::OFF_b},
        {"c", "c", VLCU_MEMBER_IN, 4, 0, 5, 130, 129, 0, T::OFF_c},
        {"sel", "sel", VLCU_MEMBER_IN, 1, 0, 1, 2, 1, 0, T::OFF_sel},
        {"q", "q", VLCU_MEMBER_OUT, 4, 0, 3, 96, 95, 0, T::OFF_q},
        {"r", "r", VLCU_MEMBER_OUT, 4, 0, 5, 130, 129, 0, T::OFF_r},
        {"flags", "flags", VLCU_MEMBER_OUT, 1, 0, 1, 4, 3, 0, T::OFF_flags},
        {"ctr", "ctr", VLCU_MEMBER_OUT, 4, 0, 4, 100, 99, 0, T::OFF_ctr},
        {"__VstlFirstIteration", "__VstlFirstIteration", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlFirstIteration},
        {"__VstlPhaseResult", "__VstlPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__Vtrigprevexpr___TOP__sel__0", "__Vtrigprevexpr___TOP__sel__0", VLCU_MEMBER_STATE, 1, 0, 1, 2, 1, 0, T::OFF___Vtrigprevexpr___TOP__sel__0},
        {"__VicoDidInit", "__VicoDidInit", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VicoDidInit},
        {"__VicoPhaseResult", "__VicoPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VicoPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__1", "__Vtrigprevexpr___TOP__clk__1", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__1},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"__Vtrigprevexpr___TOP__a__0", "__Vtrigprevexpr___TOP__a__0", VLCU_MEMBER_STATE, 4, 0, 3, 96, 95, 0, T::OFF___Vtrigprevexpr___TOP__a__0},
        {"__Vtrigprevexpr___TOP__b__0", "__Vtrigprevexpr___TOP__b__0", VLCU_MEMBER_STATE, 4, 0, 3, 96, 95, 0, T::OFF___Vtrigprevexpr___TOP__b__0},
        {"__Vtrigprevexpr___TOP__c__0", "__Vtrigprevexpr___TOP__c__0", VLCU_MEMBER_STATE, 4, 0, 5, 130, 129, 0, T::OFF___Vtrigprevexpr___TOP__c__0},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
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
    static const uint32_t saveOrder[] = {0, 4, 7, 9, 10, 11, 12, 13, 14, 15, 16, 17, 1, 2, 3, 5, 6, 8, 18, 19, 20, 21, 22, 23, 24, 25, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "Vt05_wide", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 0, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 26, 0, 0x553002b1b73e24f0ULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt05_wide__cuda VLCU_TESTBENCHES);
