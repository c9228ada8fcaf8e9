// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt06_loops"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt06_loops___024root final {
    CData/*0:0*/ clk;
    CData/*0:0*/ en;
    IData/*31:0*/ din;
    IData/*31:0*/ rev;
    CData/*5:0*/ ones;
    IData/*31:0*/ tap;
    IData/*31:0*/ crc;
    CData/*0:0*/ __VstlFirstIteration;
    CData/*0:0*/ __VstlPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    IData/*31:0*/ __VactIterCount;
    VlUnpacked<IData/*31:0*/, 8> t06_loops__DOT__pipe;
    VlUnpacked<QData/*63:0*/, 1> __VstlTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vt06_loops__cuda final {
    typedef Vt06_loops___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 76,
        OFF_en = 77,
        OFF_din = 24,
        OFF_rev = 28,
        OFF_ones = 78,
        OFF_tap = 32,
        OFF_crc = 36,
        OFF___VstlFirstIteration = 79,
        OFF___VstlPhaseResult = 80,
        OFF___Vtrigprevexpr___TOP__clk__0 = 81,
        OFF___VactPhaseResult = 82,
        OFF___VnbaPhaseResult = 83,
        OFF___VactIterCount = 40,
        OFF_t06_loops__DOT__pipe = 44,
        OFF___VstlTriggered = 0,
        OFF___VactTriggered = 8,
        OFF___VnbaTriggered = 16,
        kStateBytes = 88,  // S_persist
        kTouchedReadBytes = 65,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 68,  // S_state_wr
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
        s.en = st.ld<VlcuMem8>(OFF_en, i);
        s.din = st.ld<IData>(OFF_din, i);
        s.rev = st.ld<IData>(OFF_rev, i);
        s.ones = st.ld<VlcuMem8>(OFF_ones, i);
        s.tap = st.ld<IData>(OFF_tap, i);
        s.crc = st.ld<IData>(OFF_crc, i);
        s.__VstlFirstIteration = st.ld<VlcuMem8>(OFF___VstlFirstIteration, i);
        s.__VstlPhaseResult = st.ld<VlcuMem8>(OFF___VstlPhaseResult, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        for (uint32_t e = 0; e < 8; ++e) s.t06_loops__DOT__pipe[e] = st.ldE<IData>(OFF_t06_loops__DOT__pipe, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VstlTriggered[e] = st.ldE<QData>(OFF___VstlTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.st<VlcuMem8>(OFF_en, i, s.en);
        st.st<IData>(OFF_din, i, s.din);
        st.st<IData>(OFF_rev, i, s.rev);
        st.st<VlcuMem8>(OFF_ones, i, s.ones);
        st.st<IData>(OFF_tap, i, s.tap);
        st.st<IData>(OFF_crc, i, s.crc);
        st.st<VlcuMem8>(OFF___VstlFirstIteration, i, s.__VstlFirstIteration);
        st.st<VlcuMem8>(OFF___VstlPhaseResult, i, s.__VstlPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
        for (uint32_t e = 0; e < 8; ++e) st.stE<IData>(OFF_t06_loops__DOT__pipe, e, i, s.t06_loops__DOT__pipe[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VstlTriggered, e, i, s.__VstlTriggered[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VactTriggered, e, i, s.__VactTriggered[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VnbaTriggered, e, i, s.__VnbaTriggered[e]);
    }
    VLCU_D static void load_inputs_row(Root& s, const uint32_t* rowp, uint64_t stride, uint64_t i) {
        s.clk = static_cast<VlcuMem8>(rowp[0 * stride + i] & 0x1U);
        s.en = static_cast<VlcuMem8>(rowp[1 * stride + i] & 0x1U);
        s.din = static_cast<IData>(rowp[2 * stride + i] & 0xffffffffU);
    }
    VLCU_D static void store_outputs_row(const Root& s, uint32_t* rowp, uint64_t stride, uint64_t i) {
        rowp[0 * stride + i] = s.rev;
        rowp[1 * stride + i] = s.ones;
        rowp[2 * stride + i] = s.tap;
        rowp[3 * stride + i] = s.crc;
    }
    VLCU_D static uint64_t outputs_sig(const Root& s, uint64_t sig) {
        uint32_t acc = VLCU_SIG_ACC_INIT;
        acc = vlcu_sig_word(acc, s.rev);
        acc = vlcu_sig_word(acc, s.ones);
        acc = vlcu_sig_word(acc, s.tap);
        acc = vlcu_sig_word(acc, s.crc);
        return vlcu_sig_mix(sig, acc);
    }
    enum : uint32_t { kInWords = 3, kOutWords = 4, kInBytes = 6, kOutBytes = 13, kTraceWords = 0 };

    VLCU_D static void Vt06_loops___024root___eval_static(Vt06_loops___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
    }

    VLCU_D static void Vt06_loops___024root___eval_initial(Vt06_loops___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt06_loops___024root___eval_final(Vt06_loops___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt06_loops___024root___eval_settle(Vt06_loops___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt06_loops___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t06_loops.v", 3, "", "DIDNOTCONVERGE: Settle region did not converge after '--converge-limit' of 10000 tries");
            }
            __VstlIterCount = ((IData)(1U) + __VstlIterCount);
            vlSelfRef.__VstlPhaseResult = Vt06_loops___024root___eval_phase__stl(vlSelf, vlCtx);
            vlSelfRef.__VstlFirstIteration = 0U;
        } while (vlSelfRef.__VstlPhaseResult);
    }

    VLCU_D static bool Vt06_loops___024root___trigger_anySet__stl(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt06_loops___024root___eval_phase__stl(Vt06_loops___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt06_loops___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
        }
#endif
        __VstlExecute = Vt06_loops___024root___trigger_anySet__stl(vlSelfRef.__VstlTriggered);
        if (__VstlExecute) {
            {
                // Inlined CFunc: _eval_stl
                if ((1ULL & vlSelfRef.__VstlTriggered[0U])) {
                    {
                        // Inlined CFunc: _stl_sequent__TOP__0
                        vlSelfRef.tap = (vlSelfRef.t06_loops__DOT__pipe[7U] 
                                         ^ vlSelfRef.t06_loops__DOT__pipe[3U]);
                    }
                }
            }
        }
        return (__VstlExecute);
    }

    VLCU_D static bool Vt06_loops___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt06_loops___024root___nba_sequent__TOP__0(Vt06_loops___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        IData/*31:0*/ t06_loops__DOT__t;
        t06_loops__DOT__t = 0;
        IData/*31:0*/ __VdlyVal__t06_loops__DOT__pipe__v0;
        __VdlyVal__t06_loops__DOT__pipe__v0 = 0;
        CData/*0:0*/ __VdlySet__t06_loops__DOT__pipe__v0;
        __VdlySet__t06_loops__DOT__pipe__v0 = 0;
        IData/*31:0*/ __VdlyVal__t06_loops__DOT__pipe__v1;
        __VdlyVal__t06_loops__DOT__pipe__v1 = 0;
        IData/*31:0*/ __VdlyVal__t06_loops__DOT__pipe__v2;
        __VdlyVal__t06_loops__DOT__pipe__v2 = 0;
        IData/*31:0*/ __VdlyVal__t06_loops__DOT__pipe__v3;
        __VdlyVal__t06_loops__DOT__pipe__v3 = 0;
        IData/*31:0*/ __VdlyVal__t06_loops__DOT__pipe__v4;
        __VdlyVal__t06_loops__DOT__pipe__v4 = 0;
        IData/*31:0*/ __VdlyVal__t06_loops__DOT__pipe__v5;
        __VdlyVal__t06_loops__DOT__pipe__v5 = 0;
        IData/*31:0*/ __VdlyVal__t06_loops__DOT__pipe__v6;
        __VdlyVal__t06_loops__DOT__pipe__v6 = 0;
        IData/*31:0*/ __VdlyVal__t06_loops__DOT__pipe__v7;
        __VdlyVal__t06_loops__DOT__pipe__v7 = 0;
        // Body
        __VdlySet__t06_loops__DOT__pipe__v0 = 0U;
        if (vlSelfRef.en) {
            __VdlyVal__t06_loops__DOT__pipe__v0 = vlSelfRef.din;
            __VdlySet__t06_loops__DOT__pipe__v0 = 1U;
            __VdlyVal__t06_loops__DOT__pipe__v1 = ((IData)(1U) 
                                                   + vlSelfRef.t06_loops__DOT__pipe[0U]);
            __VdlyVal__t06_loops__DOT__pipe__v2 = ((IData)(2U) 
                                                   + vlSelfRef.t06_loops__DOT__pipe[1U]);
            __VdlyVal__t06_loops__DOT__pipe__v3 = ((IData)(3U) 
                                                   + vlSelfRef.t06_loops__DOT__pipe[2U]);
            __VdlyVal__t06_loops__DOT__pipe__v4 = ((IData)(4U) 
                                                   + vlSelfRef.t06_loops__DOT__pipe[3U]);
            __VdlyVal__t06_loops__DOT__pipe__v5 = ((IData)(5U) 
                                                   + vlSelfRef.t06_loops__DOT__pipe[4U]);
            __VdlyVal__t06_loops__DOT__pipe__v6 = ((IData)(6U) 
                                                   + vlSelfRef.t06_loops__DOT__pipe[5U]);
            __VdlyVal__t06_loops__DOT__pipe__v7 = ((IData)(7U) 
                                                   + vlSelfRef.t06_loops__DOT__pipe[6U]);
        }
        t06_loops__DOT__t = 0U;
        t06_loops__DOT__t = ((0xfffffff8U & t06_loops__DOT__t) 
                             | ((4U & (vlSelfRef.din 
                                       >> 0x0000001bU)) 
                                | ((2U & (vlSelfRef.din 
                                          >> 0x0000001dU)) 
                                   | (vlSelfRef.din 
                                      >> 0x1fU))));
        t06_loops__DOT__t = ((0xffffffc7U & t06_loops__DOT__t) 
                             | (((4U & (vlSelfRef.din 
                                        >> 0x00000018U)) 
                                 | ((2U & (vlSelfRef.din 
                                           >> 0x0000001aU)) 
                                    | (1U & (vlSelfRef.din 
                                             >> 0x1cU)))) 
                                << 3U));
        t06_loops__DOT__t = ((0xfffffe3fU & t06_loops__DOT__t) 
                             | (((4U & (vlSelfRef.din 
                                        >> 0x00000015U)) 
                                 | ((2U & (vlSelfRef.din 
                                           >> 0x00000017U)) 
                                    | (1U & (vlSelfRef.din 
                                             >> 0x19U)))) 
                                << 6U));
        t06_loops__DOT__t = ((0xfffff1ffU & t06_loops__DOT__t) 
                             | (((4U & (vlSelfRef.din 
                                        >> 0x00000012U)) 
                                 | ((2U & (vlSelfRef.din 
                                           >> 0x00000014U)) 
                                    | (1U & (vlSelfRef.din 
                                             >> 0x16U)))) 
                                << 9U));
        t06_loops__DOT__t = ((0xffff8fffU & t06_loops__DOT__t) 
                             | (((4U & (vlSelfRef.din 
                                        >> 0x0000000fU)) 
                                 | ((2U & (vlSelfRef.din 
                                           >> 0x00000011U)) 
                                    | (1U & (vlSelfRef.din 
                                             >> 0x13U)))) 
                                << 0x0000000cU));
        t06_loops__DOT__t = ((0xfffc7fffU & t06_loops__DOT__t) 
                             | (((4U & (vlSelfRef.din 
                                        >> 0x0000000cU)) 
                                 | ((2U & (vlSelfRef.din 
                                           >> 0x0000000eU)) 
                                    | (1U & (vlSelfRef.din 
                                             >> 0x10U)))) 
                                << 0x0000000fU));
        t06_loops__DOT__t = ((0xffe3ffffU & t06_loops__DOT__t) 
                             | (((4U & (vlSelfRef.din 
                                        >> 9U)) | (
                                                   (2U 
                                                    & (vlSelfRef.din 
                                                       >> 0x0000000bU)) 
                                                   | (1U 
                                                      & (vlSelfRef.din 
                                                         >> 0x0dU)))) 
                                << 0x00000012U));
        t06_loops__DOT__t = ((0xff1fffffU & t06_loops__DOT__t) 
                             | (((4U & (vlSelfRef.din 
                                        >> 6U)) | (
                                                   (2U 
                                                    & (vlSelfRef.din 
                                                       >> 8U)) 
                                                   | (1U 
                                                      & (vlSelfRef.din 
                                                         >> 0x0aU)))) 
                                << 0x00000015U));
        t06_loops__DOT__t = ((0xf8ffffffU & t06_loops__DOT__t) 
                             | (((4U & (vlSelfRef.din 
                                        >> 3U)) | (
                                                   (2U 
                                                    & (vlSelfRef.din 
                                                       >> 5U)) 
                                                   | (1U 
                                                      & (vlSelfRef.din 
                                                         >> 7U)))) 
                                << 0x00000018U));
        t06_loops__DOT__t = ((0xc7ffffffU & t06_loops__DOT__t) 
                             | (((4U & vlSelfRef.din) 
                                 | ((2U & (vlSelfRef.din 
                                           >> 2U)) 
                                    | (1U & (vlSelfRef.din 
                                             >> 4U)))) 
                                << 0x0000001bU));
        t06_loops__DOT__t = ((0x3fffffffU & t06_loops__DOT__t) 
                             | (((2U & (vlSelfRef.din 
                                        << 1U)) | (1U 
                                                   & (vlSelfRef.din 
                                                      >> 1U))) 
                                << 0x0000001eU));
        vlSelfRef.rev = t06_loops__DOT__t;
        t06_loops__DOT__t = (1U & vlSelfRef.din);
        t06_loops__DOT__t = (t06_loops__DOT__t + (1U 
                                                  & (vlSelfRef.din 
                                                     >> 1U)));
        t06_loops__DOT__t = (t06_loops__DOT__t + (1U 
                                                  & (vlSelfRef.din 
                                                     >> 2U)));
        t06_loops__DOT__t = (t06_loops__DOT__t + (1U 
                                                  & (vlSelfRef.din 
                                                     >> 3U)));
        t06_loops__DOT__t = (t06_loops__DOT__t + (1U 
                                                  & (vlSelfRef.din 
                                                     >> 4U)));
        t06_loops__DOT__t = (t06_loops__DOT__t + (1U 
                                                  & (vlSelfRef.din 
                                                     >> 5U)));
        t06_loops__DOT__t = (t06_loops__DOT__t + (1U 
                                                  & (vlSelfRef.din 
                                                     >> 6U)));
        t06_loops__DOT__t = (t06_loops__DOT__t + (1U 
                                                  & (vlSelfRef.din 
                                                     >> 7U)));
        t06_loops__DOT__t = (t06_loops__DOT__t + (1U 
                                                  & (vlSelfRef.din 
                                                     >> 8U)));
        t06_loops__DOT__t = (t06_loops__DOT__t + (1U 
                                                  & (vlSelfRef.din 
                                                     >> 9U)));
        t06_loops__DOT__t = (t06_loops__DOT__t + (1U 
                                                  & (vlSelfRef.din 
                                                     >> 0x0aU)));
        t06_loops__DOT__t = (t06_loops__DOT__t + (1U 
                                                  & (vlSelfRef.din 
                                                     >> 0x0bU)));
        t06_loops__DOT__t = (t06_loops__DOT__t + (1U 
                                                  & (vlSelfRef.din 
                                                     >> 0x0cU)));
        t06_loops__DOT__t = (t06_loops__DOT__t + (1U 
                                                  & (vlSelfRef.din 
                                                     >> 0x0dU)));
        t06_loops__DOT__t = (t06_loops__DOT__t + (1U 
                                                  & (vlSelfRef.din 
                                                     >> 0x0eU)));
        t06_loops__DOT__t = (t06_loops__DOT__t + (1U 
                                                  & (vlSelfRef.din 
                                                     >> 0x0fU)));
        t06_loops__DOT__t = (t06_loops__DOT__t + (1U 
                                                  & (vlSelfRef.din 
                                                     >> 0x10U)));
        t06_loops__DOT__t = (t06_loops__DOT__t + (1U 
                                                  & (vlSelfRef.din 
                                                     >> 0x11U)));
        t06_loops__DOT__t = (t06_loops__DOT__t + (1U 
                                                  & (vlSelfRef.din 
                                                     >> 0x12U)));
        t06_loops__DOT__t = (t06_loops__DOT__t + (1U 
                                                  & (vlSelfRef.din 
                                                     >> 0x13U)));
        t06_loops__DOT__t = (t06_loops__DOT__t + (1U 
                                                  & (vlSelfRef.din 
                                                     >> 0x14U)));
        t06_loops__DOT__t = (t06_loops__DOT__t + (1U 
                                                  & (vlSelfRef.din 
                                                     >> 0x15U)));
        t06_loops__DOT__t = (t06_loops__DOT__t + (1U 
                                                  & (vlSelfRef.din 
                                                     >> 0x16U)));
        t06_loops__DOT__t = (t06_loops__DOT__t + (1U 
                                                  & (vlSelfRef.din 
                                                     >> 0x17U)));
        t06_loops__DOT__t = (t06_loops__DOT__t + (1U 
                                                  & (vlSelfRef.din 
                                                     >> 0x18U)));
        t06_loops__DOT__t = (t06_loops__DOT__t + (1U 
                                                  & (vlSelfRef.din 
                                                     >> 0x19U)));
        t06_loops__DOT__t = (t06_loops__DOT__t + (1U 
                                                  & (vlSelfRef.din 
                                                     >> 0x1aU)));
        t06_loops__DOT__t = (t06_loops__DOT__t + (1U 
                                                  & (vlSelfRef.din 
                                                     >> 0x1bU)));
        t06_loops__DOT__t = (t06_loops__DOT__t + (1U 
                                                  & (vlSelfRef.din 
                                                     >> 0x1cU)));
        t06_loops__DOT__t = (t06_loops__DOT__t + (1U 
                                                  & (vlSelfRef.din 
                                                     >> 0x1dU)));
        t06_loops__DOT__t = (t06_loops__DOT__t + (1U 
                                                  & (vlSelfRef.din 
                                                     >> 0x1eU)));
        t06_loops__DOT__t = (t06_loops__DOT__t + (vlSelfRef.din 
                                                  >> 0x1fU));
        vlSelfRef.ones = (0x0000003fU & t06_loops__DOT__t);
        t06_loops__DOT__t = vlSelfRef.crc;
        t06_loops__DOT__t = ((1U & t06_loops__DOT__t)
                              ? (0xedb88320U ^ (t06_loops__DOT__t 
                                                >> 1U))
                              : (t06_loops__DOT__t 
                                 >> 1U));
        t06_loops__DOT__t = ((1U & t06_loops__DOT__t)
                              ? (0xedb88320U ^ (t06_loops__DOT__t 
                                                >> 1U))
                              : (t06_loops__DOT__t 
                                 >> 1U));
        t06_loops__DOT__t = ((1U & t06_loops__DOT__t)
                              ? (0xedb88320U ^ (t06_loops__DOT__t 
                                                >> 1U))
                              : (t06_loops__DOT__t 
                                 >> 1U));
        t06_loops__DOT__t = ((1U & t06_loops__DOT__t)
                              ? (0xedb88320U ^ (t06_loops__DOT__t 
                                                >> 1U))
                              : (t06_loops__DOT__t 
                                 >> 1U));
        t06_loops__DOT__t = ((1U & t06_loops__DOT__t)
                              ? (0xedb88320U ^ (t06_loops__DOT__t 
                                                >> 1U))
                              : (t06_loops__DOT__t 
                                 >> 1U));
        t06_loops__DOT__t = ((1U & t06_loops__DOT__t)
                              ? (0xedb88320U ^ (t06_loops__DOT__t 
                                                >> 1U))
                              : (t06_loops__DOT__t 
                                 >> 1U));
        t06_loops__DOT__t = ((1U & t06_loops__DOT__t)
                              ? (0xedb88320U ^ (t06_loops__DOT__t 
                                                >> 1U))
                              : (t06_loops__DOT__t 
                                 >> 1U));
        t06_loops__DOT__t = ((1U & t06_loops__DOT__t)
                              ? (0xedb88320U ^ (t06_loops__DOT__t 
                                                >> 1U))
                              : (t06_loops__DOT__t 
                                 >> 1U));
        vlSelfRef.crc = (t06_loops__DOT__t ^ vlSelfRef.din);
        if (__VdlySet__t06_loops__DOT__pipe__v0) {
            vlSelfRef.t06_loops__DOT__pipe[0U] = __VdlyVal__t06_loops__DOT__pipe__v0;
            vlSelfRef.t06_loops__DOT__pipe[1U] = __VdlyVal__t06_loops__DOT__pipe__v1;
            vlSelfRef.t06_loops__DOT__pipe[2U] = __VdlyVal__t06_loops__DOT__pipe__v2;
            vlSelfRef.t06_loops__DOT__pipe[3U] = __VdlyVal__t06_loops__DOT__pipe__v3;
            vlSelfRef.t06_loops__DOT__pipe[4U] = __VdlyVal__t06_loops__DOT__pipe__v4;
            vlSelfRef.t06_loops__DOT__pipe[5U] = __VdlyVal__t06_loops__DOT__pipe__v5;
            vlSelfRef.t06_loops__DOT__pipe[6U] = __VdlyVal__t06_loops__DOT__pipe__v6;
            vlSelfRef.t06_loops__DOT__pipe[7U] = __VdlyVal__t06_loops__DOT__pipe__v7;
        }
        vlSelfRef.tap = (vlSelfRef.t06_loops__DOT__pipe[7U] 
                         ^ vlSelfRef.t06_loops__DOT__pipe[3U]);
    }

    VLCU_D static void Vt06_loops___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt06_loops___024root___eval_phase__act(Vt06_loops___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt06_loops___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vt06_loops___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vt06_loops___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vt06_loops___024root___eval_phase__nba(Vt06_loops___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vt06_loops___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    Vt06_loops___024root___nba_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
            Vt06_loops___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vt06_loops___024root___eval(Vt06_loops___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt06_loops___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t06_loops.v", 3, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vt06_loops___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t06_loops.v", 3, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vt06_loops___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vt06_loops___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vt06_loops___024root___ctor_var_reset(Vt06_loops___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        vlSelf->en = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 7710216835639188562ull);
        vlSelf->din = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 15192908731043726583ull);
        vlSelf->rev = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 12834581346094625679ull);
        vlSelf->ones = VL_SCOPED_RAND_RESET_I(6, __VscopeHash, 13651352298117903657ull);
        vlSelf->tap = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 13933190746779120222ull);
        vlSelf->crc = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 17734517729991455266ull);
        for (int __Vi0 = 0; __Vi0 < 8; ++__Vi0) {
            vlSelf->t06_loops__DOT__pipe[__Vi0] = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 10449335950031318250ull);
        }
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
        Vt06_loops___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt06_loops___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt06_loops___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt06_loops___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vt06_loops___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt06_loops___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt06_loops__cuda::info() {
    typedef Vt06_loops__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"en", "en", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_en},
        {"din", "din", VLCU_MEMBER_IN, 4, 0, 1, 32, 31, 0, T::OFF_din},
        {"rev", "rev", VLCU_MEMBER_OUT, 4, 0, 1, 32, 31, 0, T::OFF_rev},
        {"ones", "ones", VLCU_MEMBER_OUT, 1, 0, 1, 6, 5, 0, T::OFF_ones},
        {"tap", "tap", VLCU_MEMBER_OUT, 4, 0, 1, 32, 31, 0, T::OFF_tap},
        {"crc", "crc", VLCU_MEMBER_OUT, 4, 0, 1, 32, 31, 0, T::OFF_crc},
        {"__VstlFirstIteration", "__VstlFirstIteration", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlFirstIteration},
        {"__VstlPhaseResult", "__VstlPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"t06_loops__DOT__pipe", "t06_loops.pipe", VLCU_MEMBER_STATE, 4, 0, 8, 32, 31, 0, T::OFF_t06_loops__DOT__pipe},
        {"__VstlTriggered", "__VstlTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VstlTriggered},
        {"__VactTriggered", "__VactTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VactTriggered},
        {"__VnbaTriggered", "__VnbaTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VnbaTriggered},
    };
    static const char* const messages[] = {nullptr};
    static const char* const plusargs[] = {nullptr};
    static const VlcuTraceDecl traceDecls[] = {
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0}};  // (terminator, not counted)
    static const uint32_t traceOrder[] = {0};  // codes in dump order (+ terminator)
    static const uint32_t saveOrder[] = {0, 1, 4, 7, 8, 9, 10, 11, 2, 3, 5, 6, 12, 13, 14, 15, 16, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "t06_loops", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 0, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 17, 0, 0x63e44c3fb0276468ULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt06_loops__cuda VLCU_TESTBENCHES);
