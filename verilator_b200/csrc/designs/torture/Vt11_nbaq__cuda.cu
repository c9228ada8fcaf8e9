// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

// 64 bytes of register-resident state: latency-bound, compile for full occupancy
#ifndef VLCU_MIN_BLOCKS
#define VLCU_MIN_BLOCKS 32
#endif

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt11_nbaq"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt11_nbaq___024root final {
    CData/*0:0*/ clk;
    CData/*3:0*/ n;
    CData/*3:0*/ base;
    SData/*15:0*/ d;
    SData/*15:0*/ q;
    VlWide<3>/*71:0*/ wq;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    IData/*31:0*/ __VactIterCount;
    VlcuSharedArray<SData/*15:0*/, 16> t11_nbaq__DOT__mem;  // run-time indexed: shared memory [index][lane]
    VlcuLocalArray<VlWide<3>/*71:0*/, 8> t11_nbaq__DOT__wmem;  // thread-local memory
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    VlNBACommitQueue<VlUnpacked<SData/*15:0*/, 16>, false, SData/*15:0*/, 1, 32> __VdlyCommitQueuet11_nbaq__DOT__mem;  // entries in thread-local memory (Locals)
    VlNBACommitQueue<VlUnpacked<VlWide<3>/*71:0*/, 8>, true, VlWide<3>/*71:0*/, 1, 16> __VdlyCommitQueuet11_nbaq__DOT__wmem;  // entries in thread-local memory (Locals)
    const char* vlNamep;
};

struct Vt11_nbaq__cuda final {
    typedef Vt11_nbaq___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 164,
        OFF_n = 165,
        OFF_base = 166,
        OFF_d = 128,
        OFF_q = 130,
        OFF_wq = 16,
        OFF___Vtrigprevexpr___TOP__clk__0 = 167,
        OFF___VactPhaseResult = 168,
        OFF___VnbaPhaseResult = 169,
        OFF___VactIterCount = 28,
        OFF_t11_nbaq__DOT__mem = 132,
        OFF_t11_nbaq__DOT__wmem = 32,
        OFF___VactTriggered = 0,
        OFF___VnbaTriggered = 8,
        kStateBytes = 176,  // S_persist
        kTouchedReadBytes = 70,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 51,  // S_state_wr
    };
    struct Locals final {
        VlWide<3>/*71:0*/ t11_nbaq__DOT__wmem[8];
        VlNBACommitQueue<VlUnpacked<SData/*15:0*/, 16>, false, SData/*15:0*/, 1, 32>::Store __VdlyCommitQueuet11_nbaq__DOT__mem;
        VlNBACommitQueue<VlUnpacked<VlWide<3>/*71:0*/, 8>, true, VlWide<3>/*71:0*/, 1, 16>::Store __VdlyCommitQueuet11_nbaq__DOT__wmem;
    };
    struct Shared final {
        SData/*15:0*/ t11_nbaq__DOT__mem[16][VLCU_BLOCK];
    };
    VLCU_D static void bind(Root& s, Locals& l, Shared& sh, const VlcuState& st, uint64_t i) {
        (void)l; (void)sh; (void)st; (void)i;
        s.vlNamep = "TOP";
        s.__VdlyCommitQueuet11_nbaq__DOT__mem.bind(&l.__VdlyCommitQueuet11_nbaq__DOT__mem);
        s.__VdlyCommitQueuet11_nbaq__DOT__wmem.bind(&l.__VdlyCommitQueuet11_nbaq__DOT__wmem);
        s.t11_nbaq__DOT__mem.bind(sh.t11_nbaq__DOT__mem, threadIdx.x);
        s.t11_nbaq__DOT__wmem.bind(l.t11_nbaq__DOT__wmem);
    }
    VLCU_D static void load(Root& s, const VlcuState& st, uint64_t i) {
        s.clk = st.ld<VlcuMem8>(OFF_clk, i);
        s.n = st.ld<VlcuMem8>(OFF_n, i);
        s.base = st.ld<VlcuMem8>(OFF_base, i);
        s.d = st.ld<VlcuMem16>(OFF_d, i);
        s.q = st.ld<VlcuMem16>(OFF_q, i);
        st.ldW(OFF_wq, i, s.wq);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        for (uint32_t e = 0; e < 16; ++e) s.t11_nbaq__DOT__mem[e] = st.ldE<VlcuMem16>(OFF_t11_nbaq__DOT__mem, e, i);
        for (uint32_t e = 0; e < 8; ++e) {
            for (uint32_t w = 0; w < 3; ++w) s.t11_nbaq__DOT__wmem[e][w] = st.ldE<EData>(OFF_t11_nbaq__DOT__wmem, e * 3 + w, i);
        }
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.st<VlcuMem8>(OFF_n, i, s.n);
        st.st<VlcuMem8>(OFF_base, i, s.base);
        st.st<VlcuMem16>(OFF_d, i, s.d);
        st.st<VlcuMem16>(OFF_q, i, s.q);
        st.stW(OFF_wq, i, s.wq);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
        for (uint32_t e = 0; e < 16; ++e) st.stE<VlcuMem16>(OFF_t11_nbaq__DOT__mem, e, i, s.t11_nbaq__DOT__mem[e]);
        for (uint32_t e = 0; e < 8; ++e) {
            for (uint32_t w = 0; w < 3; ++w) st.stE<EData>(OFF_t11_nbaq__DOT__wmem, e * 3 + w, i, s.t11_nbaq__DOT__wmem[e][w]);
        }
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VactTriggered, e, i, s.__VactTriggered[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VnbaTriggered, e, i, s.__VnbaTriggered[e]);
    }
    VLCU_D static void load_inputs_row(Root& s, const uint32_t* rowp, uint64_t stride, uint64_t i) {
        s.clk = static_cast<VlcuMem8>(rowp[0 * stride + i] & 0x1U);
        s.n = static_cast<VlcuMem8>(rowp[1 * stride + i] & 0xfU);
        s.base = static_cast<VlcuMem8>(rowp[2 * stride + i] & 0xfU);
        s.d = static_cast<VlcuMem16>(rowp[3 * stride + i] & 0xffffU);
    }
    VLCU_D static void store_outputs_row(const Root& s, uint32_t* rowp, uint64_t stride, uint64_t i) {
        rowp[0 * stride + i] = s.q;
        rowp[1 * stride + i] = s.wq[0];
        rowp[2 * stride + i] = s.wq[1];
        rowp[3 * stride + i] = s.wq[2];
    }
    VLCU_D static uint64_t outputs_sig(const Root& s, uint64_t sig) {
        uint32_t acc = VLCU_SIG_ACC_INIT;
        acc = vlcu_sig_word(acc, s.q);
        acc = vlcu_sig_word(acc, s.wq[0]);
        acc = vlcu_sig_word(acc, s.wq[1]);
        acc = vlcu_sig_word(acc, s.wq[2]);
        return vlcu_sig_mix(sig, acc);
    }
    enum : uint32_t { kInWords = 4, kOutWords = 4, kInBytes = 5, kOutBytes = 14, kTraceWords = 0 };

    VLCU_D static void Vt11_nbaq___024root___eval_static(Vt11_nbaq___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
    }

    VLCU_D static void Vt11_nbaq___024root___eval_initial(Vt11_nbaq___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt11_nbaq___024root___eval_final(Vt11_nbaq___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt11_nbaq___024root___eval_settle(Vt11_nbaq___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static bool Vt11_nbaq___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt11_nbaq___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt11_nbaq___024root___eval_phase__act(Vt11_nbaq___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt11_nbaq___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vt11_nbaq___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vt11_nbaq___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vt11_nbaq___024root___eval_phase__nba(Vt11_nbaq___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vt11_nbaq___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    {
                        // Inlined CFunc: _nba_sequent__TOP__0
                        IData/*31:0*/ __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0_t11_nbaq__DOT__unnamedblk1__DOT__i;
                        __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0_t11_nbaq__DOT__unnamedblk1__DOT__i = 0;
                        SData/*15:0*/ __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyVal__t11_nbaq__DOT__mem__v0;
                        __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyVal__t11_nbaq__DOT__mem__v0 = 0;
                        CData/*3:0*/ __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyDim0__t11_nbaq__DOT__mem__v0;
                        __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyDim0__t11_nbaq__DOT__mem__v0 = 0;
                        CData/*7:0*/ __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyVal__t11_nbaq__DOT__wmem__v0;
                        __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyVal__t11_nbaq__DOT__wmem__v0 = 0;
                        CData/*6:0*/ __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyLsb__t11_nbaq__DOT__wmem__v0;
                        __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyLsb__t11_nbaq__DOT__wmem__v0 = 0;
                        CData/*2:0*/ __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyDim0__t11_nbaq__DOT__wmem__v0;
                        __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyDim0__t11_nbaq__DOT__wmem__v0 = 0;
                        VlWide<3>/*71:0*/ __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyMask__t11_nbaq__DOT__wmem__v0;
                        VL_ZERO_W(72, __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyMask__t11_nbaq__DOT__wmem__v0);
                        VlWide<3>/*71:0*/ __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyElem__t11_nbaq__DOT__wmem__v0;
                        VL_ZERO_W(72, __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyElem__t11_nbaq__DOT__wmem__v0);
                        vlSelfRef.wq[0U] = vlSelfRef.t11_nbaq__DOT__wmem
                            [(7U & (IData)(vlSelfRef.base))][0U];
                        vlSelfRef.wq[1U] = vlSelfRef.t11_nbaq__DOT__wmem
                            [(7U & (IData)(vlSelfRef.base))][1U];
                        vlSelfRef.wq[2U] = vlSelfRef.t11_nbaq__DOT__wmem
                            [(7U & (IData)(vlSelfRef.base))][2U];
                        __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0_t11_nbaq__DOT__unnamedblk1__DOT__i = 0U;
                        while ((__Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0_t11_nbaq__DOT__unnamedblk1__DOT__i 
                                < (IData)(vlSelfRef.n))) {
                            __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyVal__t11_nbaq__DOT__mem__v0 
                                = (0x0000ffffU & (((IData)(vlSelfRef.t11_nbaq__DOT__mem
                                                           [
                                                           (0x0000000fU 
                                                            & ((IData)(vlSelfRef.base) 
                                                               + __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0_t11_nbaq__DOT__unnamedblk1__DOT__i))]) 
                                                   + (IData)(vlSelfRef.d)) 
                                                  + __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0_t11_nbaq__DOT__unnamedblk1__DOT__i));
                            __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyDim0__t11_nbaq__DOT__mem__v0 
                                = (0x0000000fU & ((IData)(vlSelfRef.base) 
                                                  + __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0_t11_nbaq__DOT__unnamedblk1__DOT__i));
                            vlSelfRef.__VdlyCommitQueuet11_nbaq__DOT__mem.enqueue(__Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyVal__t11_nbaq__DOT__mem__v0, __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyDim0__t11_nbaq__DOT__mem__v0);
                            __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyVal__t11_nbaq__DOT__wmem__v0 
                                = (0x000000ffU & ((IData)(vlSelfRef.d) 
                                                  ^ __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0_t11_nbaq__DOT__unnamedblk1__DOT__i));
                            __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyLsb__t11_nbaq__DOT__wmem__v0 
                                = (0x0000003fU & (IData)(vlSelfRef.d));
                            __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyDim0__t11_nbaq__DOT__wmem__v0 
                                = (7U & __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0_t11_nbaq__DOT__unnamedblk1__DOT__i);
                            __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyMask__t11_nbaq__DOT__wmem__v0[0U] = 0U;
                            __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyMask__t11_nbaq__DOT__wmem__v0[1U] = 0U;
                            __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyMask__t11_nbaq__DOT__wmem__v0[2U] = 0U;
                            VL_ASSIGNSEL_WI(72, 8, __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyLsb__t11_nbaq__DOT__wmem__v0, __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyMask__t11_nbaq__DOT__wmem__v0, 0xffU);
                            __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyElem__t11_nbaq__DOT__wmem__v0[0U] = 0U;
                            __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyElem__t11_nbaq__DOT__wmem__v0[1U] = 0U;
                            __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyElem__t11_nbaq__DOT__wmem__v0[2U] = 0U;
                            VL_ASSIGNSEL_WI(72, 8, __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyLsb__t11_nbaq__DOT__wmem__v0, __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyElem__t11_nbaq__DOT__wmem__v0, __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyVal__t11_nbaq__DOT__wmem__v0);
                            vlSelfRef.__VdlyCommitQueuet11_nbaq__DOT__wmem.enqueue(__Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyElem__t11_nbaq__DOT__wmem__v0, __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyMask__t11_nbaq__DOT__wmem__v0, __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0___VdlyDim0__t11_nbaq__DOT__wmem__v0);
                            __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0_t11_nbaq__DOT__unnamedblk1__DOT__i 
                                = ((IData)(1U) + __Vinline_0__eval_nba___Vinline_0__nba_sequent__TOP__0_t11_nbaq__DOT__unnamedblk1__DOT__i);
                        }
                        vlSelfRef.q = (vlSelfRef.t11_nbaq__DOT__mem
                                       [vlSelfRef.base] 
                                       ^ vlSelfRef.t11_nbaq__DOT__mem
                                       [vlSelfRef.n]);
                        vlSelfRef.__VdlyCommitQueuet11_nbaq__DOT__mem.commit(vlSelfRef.t11_nbaq__DOT__mem);
                        vlSelfRef.__VdlyCommitQueuet11_nbaq__DOT__wmem.commit(vlSelfRef.t11_nbaq__DOT__wmem);
                    }
                }
            }
            Vt11_nbaq___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vt11_nbaq___024root___eval(Vt11_nbaq___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt11_nbaq___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t11_nbaq.v", 4, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vt11_nbaq___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t11_nbaq.v", 4, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vt11_nbaq___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vt11_nbaq___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vt11_nbaq___024root___ctor_var_reset(Vt11_nbaq___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        vlSelf->n = VL_SCOPED_RAND_RESET_I(4, __VscopeHash, 1489474827855109852ull);
        vlSelf->base = VL_SCOPED_RAND_RESET_I(4, __VscopeHash, 8110129090154140942ull);
        vlSelf->d = VL_SCOPED_RAND_RESET_I(16, __VscopeHash, 1720370409040345145ull);
        vlSelf->q = VL_SCOPED_RAND_RESET_I(16, __VscopeHash, 8861071527689086543ull);
        VL_SCOPED_RAND_RESET_W(72, vlSelf->wq, __VscopeHash, 781485619835072487ull);
        for (int __Vi0 = 0; __Vi0 < 16; ++__Vi0) {
            vlSelf->t11_nbaq__DOT__mem[__Vi0] = VL_SCOPED_RAND_RESET_I(16, __VscopeHash, 12303923694379649196ull);
        }
        for (int __Vi0 = 0; __Vi0 < 8; ++__Vi0) {
            VL_SCOPED_RAND_RESET_W(72, vlSelf->t11_nbaq__DOT__wmem[__Vi0], __VscopeHash, 9302407211540367037ull);
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
        Vt11_nbaq___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt11_nbaq___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt11_nbaq___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt11_nbaq___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) {
        Vt11_nbaq___024root___eval(vlSelf, vlCtx);
        if (VL_UNLIKELY(vlSelf->__VdlyCommitQueuet11_nbaq__DOT__mem.overflowed())) vlCtx.fatal(0, VLCU_FATAL_NBAQ_OVERFLOW);
        if (VL_UNLIKELY(vlSelf->__VdlyCommitQueuet11_nbaq__DOT__wmem.overflowed())) vlCtx.fatal(0, VLCU_FATAL_NBAQ_OVERFLOW);
    }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt11_nbaq___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt11_nbaq__cuda::info() {
    typedef Vt11_nbaq__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"n", "n", VLCU_MEMBER_IN, 1, 0, 1, 4, 3, 0, T::OFF_n},
        {"base", "base", VLCU_MEMBER_IN, 1, 0, 1, 4, 3, 0, T::OFF_base},
        {"d", "d", VLCU_MEMBER_IN, 2, 0, 1, 16, 15, 0, T::OFF_d},
        {"q", "q", VLCU_MEMBER_OUT, 2, 0, 1, 16, 15, 0, T::OFF_q},
        {"wq", "wq", VLCU_MEMBER_OUT, 4, 0, 3, 72, 71, 0, T::OFF_wq},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"t11_nbaq__DOT__mem", "t11_nbaq.mem", VLCU_MEMBER_STATE, 2, 0, 16, 16, 15, 0, T::OFF_t11_nbaq__DOT__mem},
        {"t11_nbaq__DOT__wmem", "t11_nbaq.wmem", VLCU_MEMBER_STATE, 4, 0, 24, 72, 71, 0, T::OFF_t11_nbaq__DOT__wmem},
        {"__VactTriggered", "__VactTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VactTriggered},
        {"__VnbaTriggered", "__VnbaTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VnbaTriggered},
    };
    static const char* const messages[] = {nullptr};
    static const char* const plusargs[] = {nullptr};
    static const VlcuTraceDecl traceDecls[] = {
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0}};  // (terminator, not counted)
    static const uint32_t traceOrder[] = {0};  // codes in dump order (+ terminator)
    static const uint32_t saveOrder[] = {0, 1, 2, 6, 7, 8, 3, 4, 5, 9, 10, 11, 12, 13, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "t11_nbaq", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 0, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 14, 0, 0x32e76cb27619a710ULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt11_nbaq__cuda VLCU_TESTBENCHES);
