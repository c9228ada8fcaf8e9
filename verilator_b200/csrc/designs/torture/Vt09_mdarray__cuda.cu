// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt09_mdarray"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt09_mdarray___024root final {
    CData/*0:0*/ clk;
    CData/*0:0*/ we;
    CData/*1:0*/ i;
    CData/*2:0*/ j;
    CData/*3:0*/ k;
    CData/*7:0*/ d;
    VlWide<3>/*71:0*/ wd;
    CData/*7:0*/ q;
    SData/*15:0*/ s;
    VlWide<3>/*71:0*/ wq;
    IData/*31:0*/ big_q;
    SData/*11:0*/ cube_q;
    CData/*1:0*/ t09_mdarray__DOT__jj;
    CData/*0:0*/ __VstlFirstIteration;
    CData/*0:0*/ __VstlPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*0:0*/ __Vtrigprevexpr___TOP__we__0;
    CData/*1:0*/ __Vtrigprevexpr___TOP__i__0;
    CData/*2:0*/ __Vtrigprevexpr___TOP__j__0;
    CData/*3:0*/ __Vtrigprevexpr___TOP__k__0;
    CData/*7:0*/ __Vtrigprevexpr___TOP__d__0;
    CData/*0:0*/ __VicoDidInit;
    CData/*0:0*/ __VicoPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__1;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    VlWide<3>/*71:0*/ __Vtrigprevexpr___TOP__wd__0;
    IData/*31:0*/ __VactIterCount;
    VlcuMdArray<VlcuSharedArray<CData/*7:0*/, 32>, 4, 8> t09_mdarray__DOT__m;  // run-time indexed: shared memory [index][lane]
    VlcuMdArray<VlcuSharedArray<SData/*15:0*/, 6>, 3, 2> t09_mdarray__DOT__w;  // run-time indexed: shared memory [index][lane]
    VlcuMdArray<VlcuLocalArray<VlWide<3>/*71:0*/, 12>, 4, 3> t09_mdarray__DOT__wm;  // thread-local memory
    VlcuMdArray<VlcuGlobalArray<IData/*31:0*/, 1024>, 16, 64> t09_mdarray__DOT__big;  // stays in HBM, instance-major
    VlcuMdArray<VlcuSharedArray<SData/*11:0*/, 24>, 2, 3, 4> t09_mdarray__DOT__cube;  // run-time indexed: shared memory [index][lane]
    VlUnpacked<QData/*63:0*/, 1> __VstlTriggered;
    VlUnpacked<QData/*63:0*/, 2> __VicoTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vt09_mdarray__cuda final {
    typedef Vt09_mdarray___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 4388,
        OFF_we = 4389,
        OFF_i = 4390,
        OFF_j = 4391,
        OFF_k = 4392,
        OFF_d = 4393,
        OFF_wd = 40,
        OFF_q = 4394,
        OFF_s = 4324,
        OFF_wq = 52,
        OFF_big_q = 64,
        OFF_cube_q = 4326,
        OFF_t09_mdarray__DOT__jj = 4395,
        OFF___VstlFirstIteration = 4396,
        OFF___VstlPhaseResult = 4397,
        OFF___Vtrigprevexpr___TOP__clk__0 = 4398,
        OFF___Vtrigprevexpr___TOP__we__0 = 4399,
        OFF___Vtrigprevexpr___TOP__i__0 = 4400,
        OFF___Vtrigprevexpr___TOP__j__0 = 4401,
        OFF___Vtrigprevexpr___TOP__k__0 = 4402,
        OFF___Vtrigprevexpr___TOP__d__0 = 4403,
        OFF___VicoDidInit = 4404,
        OFF___VicoPhaseResult = 4405,
        OFF___Vtrigprevexpr___TOP__clk__1 = 4406,
        OFF___VactPhaseResult = 4407,
        OFF___VnbaPhaseResult = 4408,
        OFF___Vtrigprevexpr___TOP__wd__0 = 68,
        OFF___VactIterCount = 80,
        OFF_t09_mdarray__DOT__m = 4409,
        OFF_t09_mdarray__DOT__w = 4328,
        OFF_t09_mdarray__DOT__wm = 84,
        OFF_t09_mdarray__DOT__big = 228,
        OFF_t09_mdarray__DOT__cube = 4340,
        OFF___VstlTriggered = 0,
        OFF___VicoTriggered = 8,
        OFF___VactTriggered = 24,
        OFF___VnbaTriggered = 32,
        kStateBytes = 4448,  // S_persist
        kTouchedReadBytes = 151,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 126,  // S_state_wr
    };
    struct Locals final {
        VlWide<3>/*71:0*/ t09_mdarray__DOT__wm[12];
    };
    struct Shared final {
        CData/*7:0*/ t09_mdarray__DOT__m[32][VLCU_BLOCK];
        SData/*15:0*/ t09_mdarray__DOT__w[6][VLCU_BLOCK];
        SData/*11:0*/ t09_mdarray__DOT__cube[24][VLCU_BLOCK];
    };
    VLCU_D static void bind(Root& s, Locals& l, Shared& sh, const VlcuState& st, uint64_t i) {
        (void)l; (void)sh; (void)st; (void)i;
        s.vlNamep = "TOP";
        s.t09_mdarray__DOT__m.bind(sh.t09_mdarray__DOT__m, threadIdx.x);
        s.t09_mdarray__DOT__w.bind(sh.t09_mdarray__DOT__w, threadIdx.x);
        s.t09_mdarray__DOT__wm.bind(l.t09_mdarray__DOT__wm);
        s.t09_mdarray__DOT__big.bind(st, OFF_t09_mdarray__DOT__big, i);
        s.t09_mdarray__DOT__cube.bind(sh.t09_mdarray__DOT__cube, threadIdx.x);
    }
    VLCU_D static void load(Root& s, const VlcuState& st, uint64_t i) {
        s.clk = st.ld<VlcuMem8>(OFF_clk, i);
        s.we = st.ld<VlcuMem8>(OFF_we, i);
        s.i = st.ld<VlcuMem8>(OFF_i, i);
        s.j = st.ld<VlcuMem8>(OFF_j, i);
        s.k = st.ld<VlcuMem8>(OFF_k, i);
        s.d = st.ld<VlcuMem8>(OFF_d, i);
        st.ldW(OFF_wd, i, s.wd);
        s.q = st.ld<VlcuMem8>(OFF_q, i);
        s.s = st.ld<VlcuMem16>(OFF_s, i);
        st.ldW(OFF_wq, i, s.wq);
        s.big_q = st.ld<IData>(OFF_big_q, i);
        s.cube_q = st.ld<VlcuMem16>(OFF_cube_q, i);
        s.t09_mdarray__DOT__jj = st.ld<VlcuMem8>(OFF_t09_mdarray__DOT__jj, i);
        s.__VstlFirstIteration = st.ld<VlcuMem8>(OFF___VstlFirstIteration, i);
        s.__VstlPhaseResult = st.ld<VlcuMem8>(OFF___VstlPhaseResult, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__Vtrigprevexpr___TOP__we__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__we__0, i);
        s.__Vtrigprevexpr___TOP__i__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__i__0, i);
        s.__Vtrigprevexpr___TOP__j__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__j__0, i);
        s.__Vtrigprevexpr___TOP__k__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__k__0, i);
        s.__Vtrigprevexpr___TOP__d__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__d__0, i);
        s.__VicoDidInit = st.ld<VlcuMem8>(OFF___VicoDidInit, i);
        s.__VicoPhaseResult = st.ld<VlcuMem8>(OFF___VicoPhaseResult, i);
        s.__Vtrigprevexpr___TOP__clk__1 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__1, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        st.ldW(OFF___Vtrigprevexpr___TOP__wd__0, i, s.__Vtrigprevexpr___TOP__wd__0);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        for (uint32_t e = 0; e < 32; ++e) s.t09_mdarray__DOT__m.flat()[e] = st.ldE<VlcuMem8>(OFF_t09_mdarray__DOT__m, e, i);
        for (uint32_t e = 0; e < 6; ++e) s.t09_mdarray__DOT__w.flat()[e] = st.ldE<VlcuMem16>(OFF_t09_mdarray__DOT__w, e, i);
        for (uint32_t e = 0; e < 12; ++e) {
            for (uint32_t w = 0; w < 3; ++w) s.t09_mdarray__DOT__wm.flat()[e][w] = st.ldE<EData>(OFF_t09_mdarray__DOT__wm, e * 3 + w, i);
        }
        for (uint32_t e = 0; e < 24; ++e) s.t09_mdarray__DOT__cube.flat()[e] = st.ldE<VlcuMem16>(OFF_t09_mdarray__DOT__cube, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VstlTriggered[e] = st.ldE<QData>(OFF___VstlTriggered, e, i);
        for (uint32_t e = 0; e < 2; ++e) s.__VicoTriggered[e] = st.ldE<QData>(OFF___VicoTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.st<VlcuMem8>(OFF_we, i, s.we);
        st.st<VlcuMem8>(OFF_i, i, s.i);
        st.st<VlcuMem8>(OFF_j, i, s.j);
        st.st<VlcuMem8>(OFF_k, i, s.k);
        st.st<VlcuMem8>(OFF_d, i, s.d);
        st.stW(OFF_wd, i, s.wd);
        st.st<VlcuMem8>(OFF_q, i, s.q);
        st.st<VlcuMem16>(OFF_s, i, s.s);
        st.stW(OFF_wq, i, s.wq);
        st.st<IData>(OFF_big_q, i, s.big_q);
        st.st<VlcuMem16>(OFF_cube_q, i, s.cube_q);
        st.st<VlcuMem8>(OFF_t09_mdarray__DOT__jj, i, s.t09_mdarray__DOT__jj);
        st.st<VlcuMem8>(OFF___VstlFirstIteration, i, s.__VstlFirstIteration);
        st.st<VlcuMem8>(OFF___VstlPhaseResult, i, s.__VstlPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__we__0, i, s.__Vtrigprevexpr___TOP__we__0);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__i__0, i, s.__Vtrigprevexpr___TOP__i__0);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__j__0, i, s.__Vtrigprevexpr___TOP__j__0);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__k__0, i, s.__Vtrigprevexpr___TOP__k__0);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__d__0, i, s.__Vtrigprevexpr___TOP__d__0);
        st.st<VlcuMem8>(OFF___VicoDidInit, i, s.__VicoDidInit);
        st.st<VlcuMem8>(OFF___VicoPhaseResult, i, s.__VicoPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__1, i, s.__Vtrigprevexpr___TOP__clk__1);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.stW(OFF___Vtrigprevexpr___TOP__wd__0, i, s.__Vtrigprevexpr___TOP__wd__0);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
        for (uint32_t e = 0; e < 32; ++e) st.stE<VlcuMem8>(OFF_t09_mdarray__DOT__m, e, i, s.t09_mdarray__DOT__m.flat()[e]);
        for (uint32_t e = 0; e < 6; ++e) st.stE<VlcuMem16>(OFF_t09_mdarray__DOT__w, e, i, s.t09_mdarray__DOT__w.flat()[e]);
        for (uint32_t e = 0; e < 12; ++e) {
            for (uint32_t w = 0; w < 3; ++w) st.stE<EData>(OFF_t09_mdarray__DOT__wm, e * 3 + w, i, s.t09_mdarray__DOT__wm.flat()[e][w]);
        }
        for (uint32_t e = 0; e < 24; ++e) st.stE<VlcuMem16>(OFF_t09_mdarray__DOT__cube, e, i, s.t09_mdarray__DOT__cube.flat()[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VstlTriggered, e, i, s.__VstlTriggered[e]);
        for (uint32_t e = 0; e < 2; ++e) st.stE<QData>(OFF___VicoTriggered, e, i, s.__VicoTriggered[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VactTriggered, e, i, s.__VactTriggered[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VnbaTriggered, e, i, s.__VnbaTriggered[e]);
    }
    VLCU_D static void load_inputs_row(Root& s, const uint32_t* rowp, uint64_t stride, uint64_t i) {
        s.clk = static_cast<VlcuMem8>(rowp[0 * stride + i] & 0x1U);
        s.we = static_cast<VlcuMem8>(rowp[1 * stride + i] & 0x1U);
        s.i = static_cast<VlcuMem8>(rowp[2 * stride + i] & 0x3U);
        s.j = static_cast<VlcuMem8>(rowp[3 * stride + i] & 0x7U);
        s.k = static_cast<VlcuMem8>(rowp[4 * stride + i] & 0xfU);
        s.d = static_cast<VlcuMem8>(rowp[5 * stride + i] & 0xffU);
        s.wd[0] = rowp[6 * stride + i] & 0xffffffffU;
        s.wd[1] = rowp[7 * stride + i] & 0xffffffffU;
        s.wd[2] = rowp[8 * stride + i] & 0xffU;
    }
    VLCU_D static void store_outputs_row(const Root& s, uint32_t* rowp, uint64_t stride, uint64_t i) {
        rowp[0 * stride + i] = s.q;
        rowp[1 * stride + i] = s.s;
        rowp[2 * stride + i] = s.wq[0];
        rowp[3 * stride + i] = s.wq[1];
        rowp[4 * stride + i] = s.wq[2];
        rowp[5 * stride + i] = s.big_q;
        rowp[6 * stride + i] = s.cube_q;
    }
    VLCU_D static uint64_t outputs_sig(const Root& s, uint64_t sig) {
        uint32_t acc = VLCU_SIG_ACC_INIT;
        acc = vlcu_sig_word(acc, s.q);
        acc = vlcu_sig_word(acc, s.s);
        acc = vlcu_sig_word(acc, s.wq[0]);
        acc = vlcu_sig_word(acc, s.wq[1]);
        acc = vlcu_sig_word(acc, s.wq[2]);
        acc = vlcu_sig_word(acc, s.big_q);
        acc = vlcu_sig_word(acc, s.cube_q);
        return vlcu_sig_mix(sig, acc);
    }
    enum : uint32_t { kInWords = 9, kOutWords = 7, kInBytes = 18, kOutBytes = 21, kTraceWords = 0 };

    VLCU_D static void Vt09_mdarray___024root___eval_static(Vt09_mdarray___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
        vlSelfRef.__Vtrigprevexpr___TOP__we__0 = vlSelfRef.we;
        vlSelfRef.__Vtrigprevexpr___TOP__i__0 = vlSelfRef.i;
        vlSelfRef.__Vtrigprevexpr___TOP__j__0 = vlSelfRef.j;
        vlSelfRef.__Vtrigprevexpr___TOP__k__0 = vlSelfRef.k;
        vlSelfRef.__Vtrigprevexpr___TOP__d__0 = vlSelfRef.d;
        vlSelfRef.__Vtrigprevexpr___TOP__wd__0[0U] 
            = vlSelfRef.wd[0U];
        vlSelfRef.__Vtrigprevexpr___TOP__wd__0[1U] 
            = vlSelfRef.wd[1U];
        vlSelfRef.__Vtrigprevexpr___TOP__wd__0[2U] 
            = vlSelfRef.wd[2U];
        vlSelfRef.__Vtrigprevexpr___TOP__clk__1 = vlSelfRef.clk;
    }

    VLCU_D static void Vt09_mdarray___024root___eval_initial(Vt09_mdarray___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt09_mdarray___024root___eval_final(Vt09_mdarray___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt09_mdarray___024root___eval_settle(Vt09_mdarray___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt09_mdarray___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t09_mdarray.v", 5, "", "DIDNOTCONVERGE: Settle region did not converge after '--converge-limit' of 10000 tries");
            }
            __VstlIterCount = ((IData)(1U) + __VstlIterCount);
            vlSelfRef.__VstlPhaseResult = Vt09_mdarray___024root___eval_phase__stl(vlSelf, vlCtx);
            vlSelfRef.__VstlFirstIteration = 0U;
        } while (vlSelfRef.__VstlPhaseResult);
    }

    VLCU_D static bool Vt09_mdarray___024root___trigger_anySet__stl(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt09_mdarray___024root___eval_phase__stl(Vt09_mdarray___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt09_mdarray___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
        }
#endif
        __VstlExecute = Vt09_mdarray___024root___trigger_anySet__stl(vlSelfRef.__VstlTriggered);
        if (__VstlExecute) {
            {
                // Inlined CFunc: _eval_stl
                if ((1ULL & vlSelfRef.__VstlTriggered[0U])) {
                    {
                        // Inlined CFunc: _ico_sequent__TOP__0
                        vlSelfRef.t09_mdarray__DOT__jj 
                            = ((3U == (3U & (IData)(vlSelfRef.j)))
                                ? 2U : (3U & (IData)(vlSelfRef.j)));
                    }
                }
            }
        }
        return (__VstlExecute);
    }

    VLCU_D static bool Vt09_mdarray___024root___trigger_anySet__ico(const VlUnpacked<QData/*63:0*/, 2> &in) {
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

    VLCU_D static bool Vt09_mdarray___024root___eval_phase__ico(Vt09_mdarray___024root* vlSelf, VlcuContext& vlCtx) {
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
                                                            ((((0U 
                                                                != 
                                                                (((vlSelfRef.wd[0U] 
                                                                   ^ vlSelfRef.__Vtrigprevexpr___TOP__wd__0[0U]) 
                                                                  | (vlSelfRef.wd[1U] 
                                                                     ^ vlSelfRef.__Vtrigprevexpr___TOP__wd__0[1U])) 
                                                                 | (vlSelfRef.wd[2U] 
                                                                    ^ vlSelfRef.__Vtrigprevexpr___TOP__wd__0[2U]))) 
                                                               << 6U) 
                                                              | ((((IData)(vlSelfRef.d) 
                                                                   != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__d__0)) 
                                                                  << 5U) 
                                                                 | (((IData)(vlSelfRef.k) 
                                                                     != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__k__0)) 
                                                                    << 4U))) 
                                                             | (((((IData)(vlSelfRef.j) 
                                                                   != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__j__0)) 
                                                                  << 3U) 
                                                                 | (((IData)(vlSelfRef.i) 
                                                                     != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__i__0)) 
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
            vlSelfRef.__Vtrigprevexpr___TOP__i__0 = vlSelfRef.i;
            vlSelfRef.__Vtrigprevexpr___TOP__j__0 = vlSelfRef.j;
            vlSelfRef.__Vtrigprevexpr___TOP__k__0 = vlSelfRef.k;
            vlSelfRef.__Vtrigprevexpr___TOP__d__0 = vlSelfRef.d;
            vlSelfRef.__Vtrigprevexpr___TOP__wd__0[0U] 
                = vlSelfRef.wd[0U];
            vlSelfRef.__Vtrigprevexpr___TOP__wd__0[1U] 
                = vlSelfRef.wd[1U];
            vlSelfRef.__Vtrigprevexpr___TOP__wd__0[2U] 
                = vlSelfRef.wd[2U];
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
            }
        }
#ifdef VL_DEBUG
        if (VL_UNLIKELY(vlSymsp->_vm_contextp__->debug())) {
            Vt09_mdarray___024root___dump_triggers__ico(vlSelfRef.__VicoTriggered, "ico"s);
        }
#endif
        __VicoExecute = Vt09_mdarray___024root___trigger_anySet__ico(vlSelfRef.__VicoTriggered);
        if (__VicoExecute) {
            {
                // Inlined CFunc: _eval_ico
                if ((8ULL & vlSelfRef.__VicoTriggered[0U])) {
                    {
                        // Inlined CFunc: _ico_sequent__TOP__0
                        vlSelfRef.t09_mdarray__DOT__jj 
                            = ((3U == (3U & (IData)(vlSelfRef.j)))
                                ? 2U : (3U & (IData)(vlSelfRef.j)));
                    }
                }
            }
        }
        return (__VicoExecute);
    }

    VLCU_D static bool Vt09_mdarray___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt09_mdarray___024root___nba_sequent__TOP__0(Vt09_mdarray___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        VlWide<3>/*71:0*/ __VdlyVal__t09_mdarray__DOT__wm__v0;
        VL_ZERO_W(72, __VdlyVal__t09_mdarray__DOT__wm__v0);
        CData/*1:0*/ __VdlyDim0__t09_mdarray__DOT__wm__v0;
        __VdlyDim0__t09_mdarray__DOT__wm__v0 = 0;
        CData/*1:0*/ __VdlyDim1__t09_mdarray__DOT__wm__v0;
        __VdlyDim1__t09_mdarray__DOT__wm__v0 = 0;
        CData/*0:0*/ __VdlySet__t09_mdarray__DOT__wm__v0;
        __VdlySet__t09_mdarray__DOT__wm__v0 = 0;
        SData/*15:0*/ __VdlyVal__t09_mdarray__DOT__w__v0;
        __VdlyVal__t09_mdarray__DOT__w__v0 = 0;
        CData/*0:0*/ __VdlyDim0__t09_mdarray__DOT__w__v0;
        __VdlyDim0__t09_mdarray__DOT__w__v0 = 0;
        CData/*1:0*/ __VdlyDim1__t09_mdarray__DOT__w__v0;
        __VdlyDim1__t09_mdarray__DOT__w__v0 = 0;
        CData/*0:0*/ __VdlySet__t09_mdarray__DOT__w__v0;
        __VdlySet__t09_mdarray__DOT__w__v0 = 0;
        IData/*31:0*/ __VdlyVal__t09_mdarray__DOT__big__v0;
        __VdlyVal__t09_mdarray__DOT__big__v0 = 0;
        CData/*5:0*/ __VdlyDim0__t09_mdarray__DOT__big__v0;
        __VdlyDim0__t09_mdarray__DOT__big__v0 = 0;
        CData/*3:0*/ __VdlyDim1__t09_mdarray__DOT__big__v0;
        __VdlyDim1__t09_mdarray__DOT__big__v0 = 0;
        CData/*7:0*/ __VdlyVal__t09_mdarray__DOT__m__v0;
        __VdlyVal__t09_mdarray__DOT__m__v0 = 0;
        CData/*2:0*/ __VdlyDim0__t09_mdarray__DOT__m__v0;
        __VdlyDim0__t09_mdarray__DOT__m__v0 = 0;
        CData/*1:0*/ __VdlyDim1__t09_mdarray__DOT__m__v0;
        __VdlyDim1__t09_mdarray__DOT__m__v0 = 0;
        CData/*0:0*/ __VdlySet__t09_mdarray__DOT__m__v0;
        __VdlySet__t09_mdarray__DOT__m__v0 = 0;
        SData/*11:0*/ __VdlyVal__t09_mdarray__DOT__cube__v0;
        __VdlyVal__t09_mdarray__DOT__cube__v0 = 0;
        CData/*1:0*/ __VdlyDim0__t09_mdarray__DOT__cube__v0;
        __VdlyDim0__t09_mdarray__DOT__cube__v0 = 0;
        CData/*1:0*/ __VdlyDim1__t09_mdarray__DOT__cube__v0;
        __VdlyDim1__t09_mdarray__DOT__cube__v0 = 0;
        CData/*0:0*/ __VdlyDim2__t09_mdarray__DOT__cube__v0;
        __VdlyDim2__t09_mdarray__DOT__cube__v0 = 0;
        CData/*0:0*/ __VdlySet__t09_mdarray__DOT__cube__v0;
        __VdlySet__t09_mdarray__DOT__cube__v0 = 0;
        VlWide<3>/*95:0*/ __Vtemp_3;
        VlWide<3>/*95:0*/ __Vtemp_4;
        VlWide<3>/*95:0*/ __Vtemp_5;
        // Body
        __VdlySet__t09_mdarray__DOT__m__v0 = 0U;
        __VdlySet__t09_mdarray__DOT__cube__v0 = 0U;
        __VdlySet__t09_mdarray__DOT__wm__v0 = 0U;
        __VdlySet__t09_mdarray__DOT__w__v0 = 0U;
        if (vlSelfRef.we) {
            __VdlyVal__t09_mdarray__DOT__m__v0 = vlSelfRef.d;
            __VdlyDim0__t09_mdarray__DOT__m__v0 = vlSelfRef.j;
            __VdlyDim1__t09_mdarray__DOT__m__v0 = vlSelfRef.i;
            __VdlySet__t09_mdarray__DOT__m__v0 = 1U;
            if ((2U >= (IData)(vlSelfRef.t09_mdarray__DOT__jj))) {
                __VdlyVal__t09_mdarray__DOT__wm__v0[0U] 
                    = (vlSelfRef.wd[0U] ^ (IData)((0x0000000fffffffffULL 
                                                   & (((QData)((IData)(vlSelfRef.wq[2U])) 
                                                       << 0x0000001cU) 
                                                      | ((QData)((IData)(vlSelfRef.wq[1U])) 
                                                         >> 4U)))));
                __VdlyVal__t09_mdarray__DOT__wm__v0[1U] 
                    = (vlSelfRef.wd[1U] ^ (((IData)(
                                                    (0x0000000fffffffffULL 
                                                     & (((QData)((IData)(vlSelfRef.wq[1U])) 
                                                         << 0x00000020U) 
                                                        | (QData)((IData)(vlSelfRef.wq[0U]))))) 
                                            << 4U) 
                                           | (IData)(
                                                     ((0x0000000fffffffffULL 
                                                       & (((QData)((IData)(vlSelfRef.wq[2U])) 
                                                           << 0x0000001cU) 
                                                          | ((QData)((IData)(vlSelfRef.wq[1U])) 
                                                             >> 4U))) 
                                                      >> 0x00000020U))));
                __VdlyVal__t09_mdarray__DOT__wm__v0[2U] 
                    = (vlSelfRef.wd[2U] ^ (((IData)(
                                                    (0x0000000fffffffffULL 
                                                     & (((QData)((IData)(vlSelfRef.wq[1U])) 
                                                         << 0x00000020U) 
                                                        | (QData)((IData)(vlSelfRef.wq[0U]))))) 
                                            >> 0x0000001cU) 
                                           | ((IData)(
                                                      ((0x0000000fffffffffULL 
                                                        & (((QData)((IData)(vlSelfRef.wq[1U])) 
                                                            << 0x00000020U) 
                                                           | (QData)((IData)(vlSelfRef.wq[0U])))) 
                                                       >> 0x00000020U)) 
                                              << 4U)));
                __VdlyDim0__t09_mdarray__DOT__wm__v0 
                    = vlSelfRef.t09_mdarray__DOT__jj;
                __VdlyDim1__t09_mdarray__DOT__wm__v0 
                    = vlSelfRef.i;
                __VdlySet__t09_mdarray__DOT__wm__v0 = 1U;
            }
        }
        if ((2U >= (IData)(vlSelfRef.t09_mdarray__DOT__jj))) {
            __VdlyVal__t09_mdarray__DOT__cube__v0 = 
                (((IData)(vlSelfRef.k) << 8U) | (IData)(vlSelfRef.d));
            __VdlyDim0__t09_mdarray__DOT__cube__v0 
                = (3U & (IData)(vlSelfRef.k));
            __VdlyDim1__t09_mdarray__DOT__cube__v0 
                = vlSelfRef.t09_mdarray__DOT__jj;
            __VdlyDim2__t09_mdarray__DOT__cube__v0 
                = (1U & (IData)(vlSelfRef.i));
            __VdlySet__t09_mdarray__DOT__cube__v0 = 1U;
            __Vtemp_3[0U] = vlSelfRef.t09_mdarray__DOT__wm
                [(3U & (IData)(vlSelfRef.k))][vlSelfRef.t09_mdarray__DOT__jj][0U];
            __Vtemp_3[1U] = vlSelfRef.t09_mdarray__DOT__wm
                [(3U & (IData)(vlSelfRef.k))][vlSelfRef.t09_mdarray__DOT__jj][1U];
            __Vtemp_3[2U] = vlSelfRef.t09_mdarray__DOT__wm
                [(3U & (IData)(vlSelfRef.k))][vlSelfRef.t09_mdarray__DOT__jj][2U];
        } else {
            __Vtemp_3[0U] = 0U;
            __Vtemp_3[1U] = 0U;
            __Vtemp_3[2U] = 0U;
        }
        if ((3U != (IData)(vlSelfRef.i))) {
            if ((2U >= (IData)(vlSelfRef.i))) {
                __VdlyVal__t09_mdarray__DOT__w__v0 
                    = (((IData)(vlSelfRef.d) << 8U) 
                       | (IData)(vlSelfRef.q));
                __VdlyDim0__t09_mdarray__DOT__w__v0 
                    = (1U & (IData)(vlSelfRef.j));
                __VdlyDim1__t09_mdarray__DOT__w__v0 
                    = vlSelfRef.i;
                __VdlySet__t09_mdarray__DOT__w__v0 = 1U;
            }
        }
        vlSelfRef.s = ((vlSelfRef.t09_mdarray__DOT__w[0U][1U] 
                        ^ vlSelfRef.t09_mdarray__DOT__w[2U][0U]) 
                       ^ vlSelfRef.t09_mdarray__DOT__w
                       [((2U >= ((3U == (IData)(vlSelfRef.i))
                                  ? 0U : (IData)(vlSelfRef.i)))
                          ? ((3U == (IData)(vlSelfRef.i))
                              ? 0U : (IData)(vlSelfRef.i))
                          : 0U)][(1U & (IData)(vlSelfRef.j))]);
        vlSelfRef.cube_q = (0x00000fffU & ((IData)(vlSelfRef.t09_mdarray__DOT__cube
                                                   [
                                                   (1U 
                                                    & (IData)(vlSelfRef.k))]
                                                   [
                                                   ((2U 
                                                     >= (IData)(vlSelfRef.t09_mdarray__DOT__jj))
                                                     ? (IData)(vlSelfRef.t09_mdarray__DOT__jj)
                                                     : 0U)]
                                                   [vlSelfRef.i]) 
                                           + (IData)(vlSelfRef.t09_mdarray__DOT__cube[1U][2U][3U])));
        __VdlyVal__t09_mdarray__DOT__big__v0 = (vlSelfRef.t09_mdarray__DOT__big
                                                [vlSelfRef.k]
                                                [(0x0000003fU 
                                                  & (((IData)(vlSelfRef.j) 
                                                      << 3U) 
                                                     | (IData)(vlSelfRef.j)))] 
                                                + (IData)(vlSelfRef.d));
        __VdlyDim0__t09_mdarray__DOT__big__v0 = (((IData)(vlSelfRef.i) 
                                                  << 4U) 
                                                 | (IData)(vlSelfRef.k));
        __VdlyDim1__t09_mdarray__DOT__big__v0 = vlSelfRef.k;
        vlSelfRef.big_q = (vlSelfRef.t09_mdarray__DOT__big
                           [(0x0000000fU & (((IData)(vlSelfRef.i) 
                                             << 2U) 
                                            | (IData)(vlSelfRef.i)))]
                           [(((IData)(vlSelfRef.k) 
                              << 2U) | (IData)(vlSelfRef.i))] 
                           ^ vlSelfRef.t09_mdarray__DOT__big
                           [vlSelfRef.k][(((IData)(vlSelfRef.i) 
                                           << 4U) | (IData)(vlSelfRef.k))]);
        if (__VdlySet__t09_mdarray__DOT__w__v0) {
            vlSelfRef.t09_mdarray__DOT__w[__VdlyDim1__t09_mdarray__DOT__w__v0][__VdlyDim0__t09_mdarray__DOT__w__v0] 
                = __VdlyVal__t09_mdarray__DOT__w__v0;
        }
        if (__VdlySet__t09_mdarray__DOT__cube__v0) {
            vlSelfRef.t09_mdarray__DOT__cube[__VdlyDim2__t09_mdarray__DOT__cube__v0][__VdlyDim1__t09_mdarray__DOT__cube__v0][__VdlyDim0__t09_mdarray__DOT__cube__v0] 
                = __VdlyVal__t09_mdarray__DOT__cube__v0;
        }
        vlSelfRef.t09_mdarray__DOT__big[__VdlyDim1__t09_mdarray__DOT__big__v0][__VdlyDim0__t09_mdarray__DOT__big__v0] 
            = __VdlyVal__t09_mdarray__DOT__big__v0;
        __Vtemp_4[0U] = vlSelfRef.d;
        __Vtemp_4[1U] = 0U;
        __Vtemp_4[2U] = 0U;
        VL_ADD_W(3, __Vtemp_5, __Vtemp_3, __Vtemp_4);
        vlSelfRef.wq[0U] = __Vtemp_5[0U];
        vlSelfRef.wq[1U] = __Vtemp_5[1U];
        vlSelfRef.wq[2U] = (0x000000ffU & __Vtemp_5[2U]);
        vlSelfRef.q = (0x000000ffU & ((IData)(vlSelfRef.t09_mdarray__DOT__m
                                              [(3U 
                                                & (IData)(vlSelfRef.j))]
                                              [vlSelfRef.i]) 
                                      + (IData)(vlSelfRef.t09_mdarray__DOT__m
                                                [vlSelfRef.i]
                                                [vlSelfRef.j])));
        if (__VdlySet__t09_mdarray__DOT__wm__v0) {
            vlSelfRef.t09_mdarray__DOT__wm[__VdlyDim1__t09_mdarray__DOT__wm__v0][__VdlyDim0__t09_mdarray__DOT__wm__v0][0U] 
                = __VdlyVal__t09_mdarray__DOT__wm__v0[0U];
            vlSelfRef.t09_mdarray__DOT__wm[__VdlyDim1__t09_mdarray__DOT__wm__v0][__VdlyDim0__t09_mdarray__DOT__wm__v0][1U] 
                = __VdlyVal__t09_mdarray__DOT__wm__v0[1U];
            vlSelfRef.t09_mdarray__DOT__wm[__VdlyDim1__t09_mdarray__DOT__wm__v0][__VdlyDim0__t09_mdarray__DOT__wm__v0][2U] 
                = __VdlyVal__t09_mdarray__DOT__wm__v0[2U];
        }
        if (__VdlySet__t09_mdarray__DOT__m__v0) {
            vlSelfRef.t09_mdarray__DOT__m[__VdlyDim1__t09_mdarray__DOT__m__v0][__VdlyDim0__t09_mdarray__DOT__m__v0] 
                = __VdlyVal__t09_mdarray__DOT__m__v0;
        }
    }

    VLCU_D static void Vt09_mdarray___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt09_mdarray___024root___eval_phase__act(Vt09_mdarray___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt09_mdarray___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vt09_mdarray___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vt09_mdarray___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vt09_mdarray___024root___eval_phase__nba(Vt09_mdarray___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vt09_mdarray___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    Vt09_mdarray___024root___nba_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
            Vt09_mdarray___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vt09_mdarray___024root___eval(Vt09_mdarray___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt09_mdarray___024root___dump_triggers__ico(vlSelfRef.__VicoTriggered, "ico"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t09_mdarray.v", 5, "", "DIDNOTCONVERGE: Input combinational region did not converge after '--converge-limit' of 10000 tries");
            }
            __VicoIterCount = ((IData)(1U) + __VicoIterCount);
            vlSelfRef.__VicoPhaseResult = Vt09_mdarray___024root___eval_phase__ico(vlSelf, vlCtx);
        } while (vlSelfRef.__VicoPhaseResult);
        __VnbaIterCount = 0U;
        do {
            if (VL_UNLIKELY(((0x00002710U < __VnbaIterCount)))) {
#ifdef VL_DEBUG
                Vt09_mdarray___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t09_mdarray.v", 5, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vt09_mdarray___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t09_mdarray.v", 5, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vt09_mdarray___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vt09_mdarray___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vt09_mdarray___024root___ctor_var_reset(Vt09_mdarray___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        vlSelf->we = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 10105644630884274164ull);
        vlSelf->i = VL_SCOPED_RAND_RESET_I(2, __VscopeHash, 15817570140490810055ull);
        vlSelf->j = VL_SCOPED_RAND_RESET_I(3, __VscopeHash, 15917291556903334389ull);
        vlSelf->k = VL_SCOPED_RAND_RESET_I(4, __VscopeHash, 8565857220533584986ull);
        vlSelf->d = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 1720370409040345145ull);
        VL_SCOPED_RAND_RESET_W(72, vlSelf->wd, __VscopeHash, 18210509495843279068ull);
        vlSelf->q = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 8861071527689086543ull);
        vlSelf->s = VL_SCOPED_RAND_RESET_I(16, __VscopeHash, 16562859848569467201ull);
        VL_SCOPED_RAND_RESET_W(72, vlSelf->wq, __VscopeHash, 781485619835072487ull);
        vlSelf->big_q = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 14407086883661648589ull);
        vlSelf->cube_q = VL_SCOPED_RAND_RESET_I(12, __VscopeHash, 4543577453763574990ull);
        for (int __Vi0 = 0; __Vi0 < 4; ++__Vi0) {
            for (int __Vi1 = 0; __Vi1 < 8; ++__Vi1) {
                vlSelf->t09_mdarray__DOT__m[__Vi0][__Vi1] = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 2905778969064215382ull);
            }
        }
        for (int __Vi0 = 0; __Vi0 < 3; ++__Vi0) {
            for (int __Vi1 = 0; __Vi1 < 2; ++__Vi1) {
                vlSelf->t09_mdarray__DOT__w[__Vi0][__Vi1] = VL_SCOPED_RAND_RESET_I(16, __VscopeHash, 17271229708630961747ull);
            }
        }
        for (int __Vi0 = 0; __Vi0 < 4; ++__Vi0) {
            for (int __Vi1 = 0; __Vi1 < 3; ++__Vi1) {
                VL_SCOPED_RAND_RESET_W(72, vlSelf->t09_mdarray__DOT__wm[__Vi0][__Vi1], __VscopeHash, 9129087561986828150ull);
            }
        }
        for (int __Vi0 = 0; __Vi0 < 16; ++__Vi0) {
            for (int __Vi1 = 0; __Vi1 < 64; ++__Vi1) {
                vlSelf->t09_mdarray__DOT__big[__Vi0][__Vi1] = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 4123647679829478792ull);
            }
        }
        for (int __Vi0 = 0; __Vi0 < 2; ++__Vi0) {
            for (int __Vi1 = 0; __Vi1 < 3; ++__Vi1) {
                for (int __Vi2 = 0; __Vi2 < 4; ++__Vi2) {
                    vlSelf->t09_mdarray__DOT__cube[__Vi0][__Vi1][__Vi2] = VL_SCOPED_RAND_RESET_I(12, __VscopeHash, 6161902636421314286ull);
                }
            }
        }
        vlSelf->t09_mdarray__DOT__jj = VL_SCOPED_RAND_RESET_I(2, __VscopeHash, 594793112020309423ull);
        for (int __Vi0 = 0; __Vi0 < 1; ++__Vi0) {
            vlSelf->__VstlTriggered[__Vi0] = 0;
        }
        for (int __Vi0 = 0; __Vi0 < 2; ++__Vi0) {
            vlSelf->__VicoTriggered[__Vi0] = 0;
        }
        vlSelf->__Vtrigprevexpr___TOP__clk__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__we__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__i__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__j__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__k__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__d__0 = 0;
        VL_ZERO_RESET_W(72, vlSelf->__Vtrigprevexpr___TOP__wd__0);
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
        Vt09_mdarray___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt09_mdarray___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt09_mdarray___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt09_mdarray___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vt09_mdarray___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt09_mdarray___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt09_mdarray__cuda::info() {
    typedef Vt09_mdarray__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"we", "we", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_we},
        {"i", "i", VLCU_MEMBER_IN, 1, 0, 1, 2, 1, 0, T::OFF_i},
        {"j", "j", VLCU_MEMBER_IN, 1, 0, 1, 3, 2, 0, T::OFF_j},
        {"k", "k", VLCU_MEMBER_IN, 1, 0, 1, 4, 3, 0, T::OFF_k},
        {"d", "d", VLCU_MEMBER_IN, 1, 0, 1, 8, 7, 0, T::OFF_d},
        {"wd", "wd", VLCU_MEMBER_IN, 4, 0, 3, 72, 71, 0, T::OFF_wd},
        {"q", "q", VLCU_MEMBER_OUT, 1, 0, 1, 8, 7, 0, T::OFF_q},
        {"s", "s", VLCU_MEMBER_OUT, 2, 0, 1, 16, 15, 0, T::OFF_s},
        {"wq", "wq", VLCU_MEMBER_OUT, 4, 0, 3, 72, 71, 0, T::OFF_wq},
        {"big_q", "big_q", VLCU_MEMBER_OUT, 4, 0, 1, 32, 31, 0, T::OFF_big_q},
        {"cube_q", "cube_q", VLCU_MEMBER_OUT, 2, 0, 1, 12, 11, 0, T::OFF_cube_q},
        {"t09_mdarray__DOT__jj", "t09_mdarray.jj", VLCU_MEMBER_STATE, 1, 0, 1, 2, 1, 0, T::OFF_t09_mdarray__DOT__jj},
        {"__VstlFirstIteration", "__VstlFirstIteration", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlFirstIteration},
        {"__VstlPhaseResult", "__VstlPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__Vtrigprevexpr___TOP__we__0", "__Vtrigprevexpr___TOP__we__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__we__0},
        {"__Vtrigprevexpr___TOP__i__0", "__Vtrigprevexpr___TOP__i__0", VLCU_MEMBER_STATE, 1, 0, 1, 2, 1, 0, T::OFF___Vtrigprevexpr___TOP__i__0},
        {"__Vtrigprevexpr___TOP__j__0", "__Vtrigprevexpr___TOP__j__0", VLCU_MEMBER_STATE, 1, 0, 1, 3, 2, 0, T::OFF___Vtrigprevexpr___TOP__j__0},
        {"__Vtrigprevexpr___TOP__k__0", "__Vtrigprevexpr___TOP__k__0", VLCU_MEMBER_STATE, 1, 0, 1, 4, 3, 0, T::OFF___Vtrigprevexpr___TOP__k__0},
        {"__Vtrigprevexpr___TOP__d__0", "__Vtrigprevexpr___TOP__d__0", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF___Vtrigprevexpr___TOP__d__0},
        {"__VicoDidInit", "__VicoDidInit", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VicoDidInit},
        {"__VicoPhaseResult", "__VicoPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VicoPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__1", "__Vtrigprevexpr___TOP__clk__1", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__1},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"__Vtrigprevexpr___TOP__wd__0", "__Vtrigprevexpr___TOP__wd__0", VLCU_MEMBER_STATE, 4, 0, 3, 72, 71, 0, T::OFF___Vtrigprevexpr___TOP__wd__0},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"t09_mdarray__DOT__m", "t09_mdarray.m", VLCU_MEMBER_STATE, 1, 0, 32, 8, 7, 0, T::OFF_t09_mdarray__DOT__m},
        {"t09_mdarray__DOT__w", "t09_mdarray.w", VLCU_MEMBER_STATE, 2, 0, 6, 16, 15, 0, T::OFF_t09_mdarray__DOT__w},
        {"t09_mdarray__DOT__wm", "t09_mdarray.wm", VLCU_MEMBER_STATE, 4, 0, 36, 72, 71, 0, T::OFF_t09_mdarray__DOT__wm},
        {"t09_mdarray__DOT__big", "t09_mdarray.big", VLCU_MEMBER_MEMORY, 4, 0, 1024, 32, 31, 0, T::OFF_t09_mdarray__DOT__big},
        {"t09_mdarray__DOT__cube", "t09_mdarray.cube", VLCU_MEMBER_STATE, 2, 0, 24, 12, 11, 0, T::OFF_t09_mdarray__DOT__cube},
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
    static const uint32_t saveOrder[] = {0, 1, 2, 3, 4, 5, 7, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 8, 11, 6, 9, 10, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 36, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "t09_mdarray", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 0, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 37, 0, 0xeae58fda201bc3b0ULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt09_mdarray__cuda VLCU_TESTBENCHES);
