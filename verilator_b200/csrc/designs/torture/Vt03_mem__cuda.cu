// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt03_mem"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt03_mem___024root final {
    CData/*0:0*/ clk;
    CData/*0:0*/ we;
    CData/*3:0*/ be;
    CData/*5:0*/ waddr;
    CData/*5:0*/ raddr;
    IData/*31:0*/ wdata;
    IData/*31:0*/ rdata;
    CData/*7:0*/ tag_out;
    SData/*9:0*/ hits;
    CData/*0:0*/ __VstlFirstIteration;
    CData/*0:0*/ __VstlPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*0:0*/ __Vtrigprevexpr___TOP__we__0;
    CData/*3:0*/ __Vtrigprevexpr___TOP__be__0;
    CData/*5:0*/ __Vtrigprevexpr___TOP__waddr__0;
    CData/*5:0*/ __Vtrigprevexpr___TOP__raddr__0;
    CData/*0:0*/ __VicoDidInit;
    CData/*0:0*/ __VicoPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__1;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    IData/*31:0*/ __Vtrigprevexpr___TOP__wdata__0;
    IData/*31:0*/ __VactIterCount;
    VlcuSharedArray<IData/*31:0*/, 64> t03_mem__DOT__ram;  // run-time indexed: shared memory [index][lane]
    VlcuLocalArray<CData/*7:0*/, 16> t03_mem__DOT__tag;  // thread-local memory
    VlUnpacked<QData/*63:0*/, 1> __VstlTriggered;
    VlUnpacked<QData/*63:0*/, 2> __VicoTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vt03_mem__cuda final {
    typedef Vt03_mem___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 314,
        OFF_we = 315,
        OFF_be = 316,
        OFF_waddr = 317,
        OFF_raddr = 318,
        OFF_wdata = 40,
        OFF_rdata = 44,
        OFF_tag_out = 319,
        OFF_hits = 312,
        OFF___VstlFirstIteration = 320,
        OFF___VstlPhaseResult = 321,
        OFF___Vtrigprevexpr___TOP__clk__0 = 322,
        OFF___Vtrigprevexpr___TOP__we__0 = 323,
        OFF___Vtrigprevexpr___TOP__be__0 = 324,
        OFF___Vtrigprevexpr___TOP__waddr__0 = 325,
        OFF___Vtrigprevexpr___TOP__raddr__0 = 326,
        OFF___VicoDidInit = 327,
        OFF___VicoPhaseResult = 328,
        OFF___Vtrigprevexpr___TOP__clk__1 = 329,
        OFF___VactPhaseResult = 330,
        OFF___VnbaPhaseResult = 331,
        OFF___Vtrigprevexpr___TOP__wdata__0 = 48,
        OFF___VactIterCount = 52,
        OFF_t03_mem__DOT__ram = 56,
        OFF_t03_mem__DOT__tag = 332,
        OFF___VstlTriggered = 0,
        OFF___VicoTriggered = 8,
        OFF___VactTriggered = 24,
        OFF___VnbaTriggered = 32,
        kStateBytes = 352,  // S_persist
        kTouchedReadBytes = 95,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 78,  // S_state_wr
    };
    struct Locals final {
        CData/*7:0*/ t03_mem__DOT__tag[16];
    };
    struct Shared final {
        IData/*31:0*/ t03_mem__DOT__ram[64][VLCU_BLOCK];
    };
    VLCU_D static void bind(Root& s, Locals& l, Shared& sh, const VlcuState& st, uint64_t i) {
        (void)l; (void)sh; (void)st; (void)i;
        s.vlNamep = "TOP";
        s.t03_mem__DOT__ram.bind(sh.t03_mem__DOT__ram, threadIdx.x);
        s.t03_mem__DOT__tag.bind(l.t03_mem__DOT__tag);
    }
    VLCU_D static void load(Root& s, const VlcuState& st, uint64_t i) {
        s.clk = st.ld<VlcuMem8>(OFF_clk, i);
        s.we = st.ld<VlcuMem8>(OFF_we, i);
        s.be = st.ld<VlcuMem8>(OFF_be, i);
        s.waddr = st.ld<VlcuMem8>(OFF_waddr, i);
        s.raddr = st.ld<VlcuMem8>(OFF_raddr, i);
        s.wdata = st.ld<IData>(OFF_wdata, i);
        s.rdata = st.ld<IData>(OFF_rdata, i);
        s.tag_out = st.ld<VlcuMem8>(OFF_tag_out, i);
        s.hits = st.ld<VlcuMem16>(OFF_hits, i);
        s.__VstlFirstIteration = st.ld<VlcuMem8>(OFF___VstlFirstIteration, i);
        s.__VstlPhaseResult = st.ld<VlcuMem8>(OFF___VstlPhaseResult, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__Vtrigprevexpr___TOP__we__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__we__0, i);
        s.__Vtrigprevexpr___TOP__be__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__be__0, i);
        s.__Vtrigprevexpr___TOP__waddr__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__waddr__0, i);
        s.__Vtrigprevexpr___TOP__raddr__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__raddr__0, i);
        s.__VicoDidInit = st.ld<VlcuMem8>(OFF___VicoDidInit, i);
        s.__VicoPhaseResult = st.ld<VlcuMem8>(OFF___VicoPhaseResult, i);
        s.__Vtrigprevexpr___TOP__clk__1 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__1, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        s.__Vtrigprevexpr___TOP__wdata__0 = st.ld<IData>(OFF___Vtrigprevexpr___TOP__wdata__0, i);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        for (uint32_t e = 0; e < 64; ++e) s.t03_mem__DOT__ram[e] = st.ldE<IData>(OFF_t03_mem__DOT__ram, e, i);
        for (uint32_t e = 0; e < 16; ++e) s.t03_mem__DOT__tag[e] = st.ldE<VlcuMem8>(OFF_t03_mem__DOT__tag, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VstlTriggered[e] = st.ldE<QData>(OFF___VstlTriggered, e, i);
        for (uint32_t e = 0; e < 2; ++e) s.__VicoTriggered[e] = st.ldE<QData>(OFF___VicoTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.st<VlcuMem8>(OFF_we, i, s.we);
        st.st<VlcuMem8>(OFF_be, i, s.be);
        st.st<VlcuMem8>(OFF_waddr, i, s.waddr);
        st.st<VlcuMem8>(OFF_raddr, i, s.raddr);
        st.st<IData>(OFF_wdata, i, s.wdata);
        st.st<IData>(OFF_rdata, i, s.rdata);
        st.st<VlcuMem8>(OFF_tag_out, i, s.tag_out);
        st.st<VlcuMem16>(OFF_hits, i, s.hits);
        st.st<VlcuMem8>(OFF___VstlFirstIteration, i, s.__VstlFirstIteration);
        st.st<VlcuMem8>(OFF___VstlPhaseResult, i, s.__VstlPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__we__0, i, s.__Vtrigprevexpr___TOP__we__0);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__be__0, i, s.__Vtrigprevexpr___TOP__be__0);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__waddr__0, i, s.__Vtrigprevexpr___TOP__waddr__0);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__raddr__0, i, s.__Vtrigprevexpr___TOP__raddr__0);
        st.st<VlcuMem8>(OFF___VicoDidInit, i, s.__VicoDidInit);
        st.st<VlcuMem8>(OFF___VicoPhaseResult, i, s.__VicoPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__1, i, s.__Vtrigprevexpr___TOP__clk__1);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.st<IData>(OFF___Vtrigprevexpr___TOP__wdata__0, i, s.__Vtrigprevexpr___TOP__wdata__0);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
        for (uint32_t e = 0; e < 64; ++e) st.stE<IData>(OFF_t03_mem__DOT__ram, e, i, s.t03_mem__DOT__ram[e]);
        for (uint32_t e = 0; e < 16; ++e) st.stE<VlcuMem8>(OFF_t03_mem__DOT__tag, e, i, s.t03_mem__DOT__tag[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VstlTriggered, e, i, s.__VstlTriggered[e]);
        for (uint32_t e = 0; e < 2; ++e) st.stE<QData>(OFF___VicoTriggered, e, i, s.__VicoTriggered[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VactTriggered, e, i, s.__VactTriggered[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VnbaTriggered, e, i, s.__VnbaTriggered[e]);
    }
    VLCU_D static void load_inputs_row(Root& s, const uint32_t* rowp, uint64_t stride, uint64_t i) {
        s.clk = static_cast<VlcuMem8>(rowp[0 * stride + i] & 0x1U);
        s.we = static_cast<VlcuMem8>(rowp[1 * stride + i] & 0x1U);
        s.be = static_cast<VlcuMem8>(rowp[2 * stride + i] & 0xfU);
        s.waddr = static_cast<VlcuMem8>(rowp[3 * stride + i] & 0x3fU);
        s.raddr = static_cast<VlcuMem8>(rowp[4 * stride + i] & 0x3fU);
        s.wdata = static_cast<IData>(rowp[5 * stride + i] & 0xffffffffU);
    }
    VLCU_D static void store_outputs_row(const Root& s, uint32_t* rowp, uint64_t stride, uint64_t i) {
        rowp[0 * stride + i] = s.rdata;
        rowp[1 * stride + i] = s.tag_out;
        rowp[2 * stride + i] = s.hits;
    }
    VLCU_D static uint64_t outputs_sig(const Root& s, uint64_t sig) {
        uint32_t acc = VLCU_SIG_ACC_INIT;
        acc = vlcu_sig_word(acc, s.rdata);
        acc = vlcu_sig_word(acc, s.tag_out);
        acc = vlcu_sig_word(acc, s.hits);
        return vlcu_sig_mix(sig, acc);
    }
    enum : uint32_t { kInWords = 6, kOutWords = 3, kInBytes = 9, kOutBytes = 7, kTraceWords = 0 };

    VLCU_D static void Vt03_mem___024root___eval_static(Vt03_mem___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
        vlSelfRef.__Vtrigprevexpr___TOP__we__0 = vlSelfRef.we;
        vlSelfRef.__Vtrigprevexpr___TOP__be__0 = vlSelfRef.be;
        vlSelfRef.__Vtrigprevexpr___TOP__waddr__0 = vlSelfRef.waddr;
        vlSelfRef.__Vtrigprevexpr___TOP__raddr__0 = vlSelfRef.raddr;
        vlSelfRef.__Vtrigprevexpr___TOP__wdata__0 = vlSelfRef.wdata;
        vlSelfRef.__Vtrigprevexpr___TOP__clk__1 = vlSelfRef.clk;
    }

    VLCU_D static void Vt03_mem___024root___eval_initial(Vt03_mem___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt03_mem___024root___eval_final(Vt03_mem___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt03_mem___024root___eval_settle(Vt03_mem___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt03_mem___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t03_mem.v", 3, "", "DIDNOTCONVERGE: Settle region did not converge after '--converge-limit' of 10000 tries");
            }
            __VstlIterCount = ((IData)(1U) + __VstlIterCount);
            vlSelfRef.__VstlPhaseResult = Vt03_mem___024root___eval_phase__stl(vlSelf, vlCtx);
            vlSelfRef.__VstlFirstIteration = 0U;
        } while (vlSelfRef.__VstlPhaseResult);
    }

    VLCU_D static bool Vt03_mem___024root___trigger_anySet__stl(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt03_mem___024root___eval_phase__stl(Vt03_mem___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt03_mem___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
        }
#endif
        __VstlExecute = Vt03_mem___024root___trigger_anySet__stl(vlSelfRef.__VstlTriggered);
        if (__VstlExecute) {
            {
                // Inlined CFunc: _eval_stl
                if ((1ULL & vlSelfRef.__VstlTriggered[0U])) {
                    {
                        // Inlined CFunc: _ico_comb__TOP__0
                        vlSelfRef.tag_out = (0x000000ffU 
                                             & ((IData)(vlSelfRef.t03_mem__DOT__tag
                                                        [
                                                        (0x0000000fU 
                                                         & (IData)(vlSelfRef.raddr))]) 
                                                + (IData)(vlSelfRef.t03_mem__DOT__tag
                                                          [
                                                          (0x0000000fU 
                                                           & (IData)(vlSelfRef.waddr))])));
                    }
                }
            }
        }
        return (__VstlExecute);
    }

    VLCU_D static bool Vt03_mem___024root___trigger_anySet__ico(const VlUnpacked<QData/*63:0*/, 2> &in) {
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

    VLCU_D static bool Vt03_mem___024root___eval_phase__ico(Vt03_mem___024root* vlSelf, VlcuContext& vlCtx) {
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
                                                            ((((vlSelfRef.wdata 
                                                                != vlSelfRef.__Vtrigprevexpr___TOP__wdata__0) 
                                                               << 5U) 
                                                              | (((IData)(vlSelfRef.raddr) 
                                                                  != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__raddr__0)) 
                                                                 << 4U)) 
                                                             | (((((IData)(vlSelfRef.waddr) 
                                                                   != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__waddr__0)) 
                                                                  << 3U) 
                                                                 | (((IData)(vlSelfRef.be) 
                                                                     != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__be__0)) 
                                                                    << 2U)) 
                                                                | ((((IData)(vlSelfRef.we) 
                                                                     != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__we__0)) 
                                                                    << 1U) 
                                                                   | ((IData)(vlSelfRef.clk) 
                                                                      != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__clk__0)))))));
            vlSelfRef.__Vtrigprevexpr___TOP__clk__0 
                = vlSelfRef.clk;
            vlSelfRef.__Vtrigprevexpr___TOP__we__0 
                = vlSelfRef.we;
            vlSelfRef.__Vtrigprevexpr___TOP__be__0 
                = vlSelfRef.be;
            vlSelfRef.__Vtrigprevexpr___TOP__waddr__0 
                = vlSelfRef.waddr;
            vlSelfRef.__Vtrigprevexpr___TOP__raddr__0 
                = vlSelfRef.raddr;
            vlSelfRef.__Vtrigprevexpr___TOP__wdata__0 
                = vlSelfRef.wdata;
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
            }
        }
#ifdef VL_DEBUG
        if (VL_UNLIKELY(vlSymsp->_vm_contextp__->debug())) {
            Vt03_mem___024root___dump_triggers__ico(vlSelfRef.__VicoTriggered, "ico"s);
        }
#endif
        __VicoExecute = Vt03_mem___024root___trigger_anySet__ico(vlSelfRef.__VicoTriggered);
        if (__VicoExecute) {
            {
                // Inlined CFunc: _eval_ico
                if ((0x0000000000000018ULL & vlSelfRef.__VicoTriggered[0U])) {
                    {
                        // Inlined CFunc: _ico_comb__TOP__0
                        vlSelfRef.tag_out = (0x000000ffU 
                                             & ((IData)(vlSelfRef.t03_mem__DOT__tag
                                                        [
                                                        (0x0000000fU 
                                                         & (IData)(vlSelfRef.raddr))]) 
                                                + (IData)(vlSelfRef.t03_mem__DOT__tag
                                                          [
                                                          (0x0000000fU 
                                                           & (IData)(vlSelfRef.waddr))])));
                    }
                }
            }
        }
        return (__VicoExecute);
    }

    VLCU_D static bool Vt03_mem___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt03_mem___024root___nba_sequent__TOP__0(Vt03_mem___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*7:0*/ __VdlyVal__t03_mem__DOT__tag__v0;
        __VdlyVal__t03_mem__DOT__tag__v0 = 0;
        CData/*3:0*/ __VdlyDim0__t03_mem__DOT__tag__v0;
        __VdlyDim0__t03_mem__DOT__tag__v0 = 0;
        CData/*0:0*/ __VdlySet__t03_mem__DOT__tag__v0;
        __VdlySet__t03_mem__DOT__tag__v0 = 0;
        CData/*7:0*/ __VdlyVal__t03_mem__DOT__ram__v0;
        __VdlyVal__t03_mem__DOT__ram__v0 = 0;
        CData/*5:0*/ __VdlyDim0__t03_mem__DOT__ram__v0;
        __VdlyDim0__t03_mem__DOT__ram__v0 = 0;
        CData/*0:0*/ __VdlySet__t03_mem__DOT__ram__v0;
        __VdlySet__t03_mem__DOT__ram__v0 = 0;
        CData/*7:0*/ __VdlyVal__t03_mem__DOT__ram__v1;
        __VdlyVal__t03_mem__DOT__ram__v1 = 0;
        CData/*5:0*/ __VdlyDim0__t03_mem__DOT__ram__v1;
        __VdlyDim0__t03_mem__DOT__ram__v1 = 0;
        CData/*0:0*/ __VdlySet__t03_mem__DOT__ram__v1;
        __VdlySet__t03_mem__DOT__ram__v1 = 0;
        CData/*7:0*/ __VdlyVal__t03_mem__DOT__ram__v2;
        __VdlyVal__t03_mem__DOT__ram__v2 = 0;
        CData/*5:0*/ __VdlyDim0__t03_mem__DOT__ram__v2;
        __VdlyDim0__t03_mem__DOT__ram__v2 = 0;
        CData/*0:0*/ __VdlySet__t03_mem__DOT__ram__v2;
        __VdlySet__t03_mem__DOT__ram__v2 = 0;
        CData/*7:0*/ __VdlyVal__t03_mem__DOT__ram__v3;
        __VdlyVal__t03_mem__DOT__ram__v3 = 0;
        CData/*5:0*/ __VdlyDim0__t03_mem__DOT__ram__v3;
        __VdlyDim0__t03_mem__DOT__ram__v3 = 0;
        CData/*0:0*/ __VdlySet__t03_mem__DOT__ram__v3;
        __VdlySet__t03_mem__DOT__ram__v3 = 0;
        IData/*31:0*/ __VdlyVal__t03_mem__DOT__ram__v4;
        __VdlyVal__t03_mem__DOT__ram__v4 = 0;
        CData/*5:0*/ __VdlyDim0__t03_mem__DOT__ram__v4;
        __VdlyDim0__t03_mem__DOT__ram__v4 = 0;
        CData/*0:0*/ __VdlySet__t03_mem__DOT__ram__v4;
        __VdlySet__t03_mem__DOT__ram__v4 = 0;
        // Body
        __VdlySet__t03_mem__DOT__tag__v0 = 0U;
        __VdlySet__t03_mem__DOT__ram__v0 = 0U;
        __VdlySet__t03_mem__DOT__ram__v1 = 0U;
        __VdlySet__t03_mem__DOT__ram__v2 = 0U;
        __VdlySet__t03_mem__DOT__ram__v3 = 0U;
        __VdlySet__t03_mem__DOT__ram__v4 = 0U;
        if (vlSelfRef.we) {
            __VdlyVal__t03_mem__DOT__tag__v0 = (0x000000ffU 
                                                & (vlSelfRef.t03_mem__DOT__tag
                                                   [
                                                   (0x0000000fU 
                                                    & (IData)(vlSelfRef.waddr))] 
                                                   ^ vlSelfRef.wdata));
            __VdlyDim0__t03_mem__DOT__tag__v0 = (0x0000000fU 
                                                 & (IData)(vlSelfRef.waddr));
            __VdlySet__t03_mem__DOT__tag__v0 = 1U;
            if ((1U & (IData)(vlSelfRef.be))) {
                __VdlyVal__t03_mem__DOT__ram__v0 = 
                    (0x000000ffU & vlSelfRef.wdata);
                __VdlyDim0__t03_mem__DOT__ram__v0 = vlSelfRef.waddr;
                __VdlySet__t03_mem__DOT__ram__v0 = 1U;
            }
            if ((2U & (IData)(vlSelfRef.be))) {
                __VdlyVal__t03_mem__DOT__ram__v1 = 
                    (0x000000ffU & (vlSelfRef.wdata 
                                    >> 8U));
                __VdlyDim0__t03_mem__DOT__ram__v1 = vlSelfRef.waddr;
                __VdlySet__t03_mem__DOT__ram__v1 = 1U;
            }
            if ((4U & (IData)(vlSelfRef.be))) {
                __VdlyVal__t03_mem__DOT__ram__v2 = 
                    (0x000000ffU & (vlSelfRef.wdata 
                                    >> 0x10U));
                __VdlyDim0__t03_mem__DOT__ram__v2 = vlSelfRef.waddr;
                __VdlySet__t03_mem__DOT__ram__v2 = 1U;
            }
            if ((8U & (IData)(vlSelfRef.be))) {
                __VdlyVal__t03_mem__DOT__ram__v3 = 
                    (vlSelfRef.wdata >> 0x18U);
                __VdlyDim0__t03_mem__DOT__ram__v3 = vlSelfRef.waddr;
                __VdlySet__t03_mem__DOT__ram__v3 = 1U;
            }
        } else {
            __VdlyVal__t03_mem__DOT__ram__v4 = ((IData)(1U) 
                                                + vlSelfRef.t03_mem__DOT__ram
                                                [vlSelfRef.raddr]);
            __VdlyDim0__t03_mem__DOT__ram__v4 = vlSelfRef.raddr;
            __VdlySet__t03_mem__DOT__ram__v4 = 1U;
        }
        vlSelfRef.rdata = vlSelfRef.t03_mem__DOT__ram
            [vlSelfRef.raddr];
        if (((0x000000ffU & vlSelfRef.t03_mem__DOT__ram
              [vlSelfRef.raddr]) == vlSelfRef.t03_mem__DOT__tag
             [(0x0000000fU & (IData)(vlSelfRef.raddr))])) {
            vlSelfRef.hits = (0x000003ffU & ((IData)(1U) 
                                             + (IData)(vlSelfRef.hits)));
        }
        if (__VdlySet__t03_mem__DOT__tag__v0) {
            vlSelfRef.t03_mem__DOT__tag[__VdlyDim0__t03_mem__DOT__tag__v0] 
                = __VdlyVal__t03_mem__DOT__tag__v0;
        }
        if (__VdlySet__t03_mem__DOT__ram__v0) {
            vlSelfRef.t03_mem__DOT__ram[__VdlyDim0__t03_mem__DOT__ram__v0] 
                = ((0xffffff00U & vlSelfRef.t03_mem__DOT__ram
                    [__VdlyDim0__t03_mem__DOT__ram__v0]) 
                   | (IData)(__VdlyVal__t03_mem__DOT__ram__v0));
        }
        if (__VdlySet__t03_mem__DOT__ram__v1) {
            vlSelfRef.t03_mem__DOT__ram[__VdlyDim0__t03_mem__DOT__ram__v1] 
                = ((0xffff00ffU & vlSelfRef.t03_mem__DOT__ram
                    [__VdlyDim0__t03_mem__DOT__ram__v1]) 
                   | ((IData)(__VdlyVal__t03_mem__DOT__ram__v1) 
                      << 8U));
        }
        if (__VdlySet__t03_mem__DOT__ram__v2) {
            vlSelfRef.t03_mem__DOT__ram[__VdlyDim0__t03_mem__DOT__ram__v2] 
                = ((0xff00ffffU & vlSelfRef.t03_mem__DOT__ram
                    [__VdlyDim0__t03_mem__DOT__ram__v2]) 
                   | ((IData)(__VdlyVal__t03_mem__DOT__ram__v2) 
                      << 0x00000010U));
        }
        if (__VdlySet__t03_mem__DOT__ram__v3) {
            vlSelfRef.t03_mem__DOT__ram[__VdlyDim0__t03_mem__DOT__ram__v3] 
                = ((0x00ffffffU & vlSelfRef.t03_mem__DOT__ram
                    [__VdlyDim0__t03_mem__DOT__ram__v3]) 
                   | ((IData)(__VdlyVal__t03_mem__DOT__ram__v3) 
                      << 0x00000018U));
        }
        if (__VdlySet__t03_mem__DOT__ram__v4) {
            vlSelfRef.t03_mem__DOT__ram[__VdlyDim0__t03_mem__DOT__ram__v4] 
                = __VdlyVal__t03_mem__DOT__ram__v4;
        }
        vlSelfRef.tag_out = (0x000000ffU & ((IData)(vlSelfRef.t03_mem__DOT__tag
                                                    [
                                                    (0x0000000fU 
                                                     & (IData)(vlSelfRef.raddr))]) 
                                            + (IData)(vlSelfRef.t03_mem__DOT__tag
                                                      [
                                                      (0x0000000fU 
                                                       & (IData)(vlSelfRef.waddr))])));
    }

    VLCU_D static void Vt03_mem___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt03_mem___024root___eval_phase__act(Vt03_mem___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt03_mem___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vt03_mem___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vt03_mem___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vt03_mem___024root___eval_phase__nba(Vt03_mem___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vt03_mem___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    Vt03_mem___024root___nba_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
            Vt03_mem___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vt03_mem___024root___eval(Vt03_mem___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt03_mem___024root___dump_triggers__ico(vlSelfRef.__VicoTriggered, "ico"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t03_mem.v", 3, "", "DIDNOTCONVERGE: Input combinational region did not converge after '--converge-limit' of 10000 tries");
            }
            __VicoIterCount = ((IData)(1U) + __VicoIterCount);
            vlSelfRef.__VicoPhaseResult = Vt03_mem___024root___eval_phase__ico(vlSelf, vlCtx);
        } while (vlSelfRef.__VicoPhaseResult);
        __VnbaIterCount = 0U;
        do {
            if (VL_UNLIKELY(((0x00002710U < __VnbaIterCount)))) {
#ifdef VL_DEBUG
                Vt03_mem___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t03_mem.v", 3, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vt03_mem___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t03_mem.v", 3, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vt03_mem___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vt03_mem___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vt03_mem___024root___ctor_var_reset(Vt03_mem___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        vlSelf->we = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 10105644630884274164ull);
        vlSelf->be = VL_SCOPED_RAND_RESET_I(4, __VscopeHash, 5293259605002647980ull);
        vlSelf->waddr = VL_SCOPED_RAND_RESET_I(6, __VscopeHash, 15530376126285439671ull);
        vlSelf->raddr = VL_SCOPED_RAND_RESET_I(6, __VscopeHash, 12822808926477430924ull);
        vlSelf->wdata = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 12890271867161903902ull);
        vlSelf->rdata = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 10065165116613087284ull);
        vlSelf->tag_out = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 4682924195059438493ull);
        vlSelf->hits = VL_SCOPED_RAND_RESET_I(10, __VscopeHash, 3088725349321236595ull);
        for (int __Vi0 = 0; __Vi0 < 64; ++__Vi0) {
            vlSelf->t03_mem__DOT__ram[__Vi0] = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 4577305574709645116ull);
        }
        for (int __Vi0 = 0; __Vi0 < 16; ++__Vi0) {
            vlSelf->t03_mem__DOT__tag[__Vi0] = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 13252865202937094894ull);
        }
        for (int __Vi0 = 0; __Vi0 < 1; ++__Vi0) {
            vlSelf->__VstlTriggered[__Vi0] = 0;
        }
        for (int __Vi0 = 0; __Vi0 < 2; ++__Vi0) {
            vlSelf->__VicoTriggered[__Vi0] = 0;
        }
        vlSelf->__Vtrigprevexpr___TOP__clk__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__we__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__be__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__waddr__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__raddr__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__wdata__0 = 0;
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
        Vt03_mem___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt03_mem___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt03_mem___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt03_mem___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vt03_mem___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt03_mem___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt03_mem__cuda::info() {
    typedef Vt03_mem__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"we", "we", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_we},
        {"be", "be", VLCU_MEMBER_IN, 1, 0, 1, 4, 3, 0, T::OFF_be},
        {"waddr", "waddr", VLCU_MEMBER_IN, 1, 0, 1, 6, 5, 0, T::OFF_waddr},
        {"raddr", "raddr", VLCU_MEMBER_IN, 1, 0, 1, 6, 5, 0, T::OFF_raddr},
        {"wdata", "wdata", VLCU_MEMBER_IN, 4, 0, 1, 32, 31, 0, T::OFF_wdata},
        {"rdata", "rdata", VLCU_MEMBER_OUT, 4, 0, 1, 32, 31, 0, T::OFF_rdata},
        {"tag_out", "tag_out", VLCU_MEMBER_OUT, 1, 0, 1, 8, 7, 0, T::OFF_tag_out},
        {"hits", "hits", VLCU_MEMBER_OUT, 2, 0, 1, 10, 9, 0, T::OFF_hits},
        {"__VstlFirstIteration", "__VstlFirstIteration", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlFirstIteration},
        {"__VstlPhaseResult", "__VstlPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__Vtrigprevexpr___TOP__we__0", "__Vtrigprevexpr___TOP__we__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__we__0},
        {"__Vtrigprevexpr___TOP__be__0", "__Vtrigprevexpr___TOP__be__0", VLCU_MEMBER_STATE, 1, 0, 1, 4, 3, 0, T::OFF___Vtrigprevexpr___TOP__be__0},
        {"__Vtrigprevexpr___TOP__waddr__0", "__Vtrigprevexpr___TOP__waddr__0", VLCU_MEMBER_STATE, 1, 0, 1, 6, 5, 0, T::OFF___Vtrigprevexpr___TOP__waddr__0},
        {"__Vtrigprevexpr___TOP__raddr__0", "__Vtrigprevexpr___TOP__raddr__0", VLCU_MEMBER_STATE, 1, 0, 1, 6, 5, 0, T::OFF___Vtrigprevexpr___TOP__raddr__0},
        {"__VicoDidInit", "__VicoDidInit", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VicoDidInit},
        {"__VicoPhaseResult", "__VicoPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VicoPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__1", "__Vtrigprevexpr___TOP__clk__1", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__1},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"__Vtrigprevexpr___TOP__wdata__0", "__Vtrigprevexpr___TOP__wdata__0", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___Vtrigprevexpr___TOP__wdata__0},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"t03_mem__DOT__ram", "t03_mem.ram", VLCU_MEMBER_STATE, 4, 0, 64, 32, 31, 0, T::OFF_t03_mem__DOT__ram},
        {"t03_mem__DOT__tag", "t03_mem.tag", VLCU_MEMBER_STATE, 1, 0, 16, 8, 7, 0, T::OFF_t03_mem__DOT__tag},
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
    static const uint32_t saveOrder[] = {0, 1, 2, 3, 4, 7, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 8, 5, 6, 21, 22, 23, 24, 25, 26, 27, 28, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "t03_mem", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 0, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 29, 0, 0xdd8f4b7deb221c6fULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt03_mem__cuda VLCU_TESTBENCHES);
