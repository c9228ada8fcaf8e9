// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt08_widemem"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt08_widemem___024root final {
    CData/*0:0*/ clk;
    CData/*0:0*/ we0;
    CData/*0:0*/ we1;
    CData/*3:0*/ wa0;
    CData/*3:0*/ wa1;
    CData/*3:0*/ ra;
    VlWide<4>/*99:0*/ wd;
    QData/*63:0*/ q;
    VlWide<4>/*99:0*/ rd;
    QData/*63:0*/ qsum;
    SData/*15:0*/ part;
    CData/*0:0*/ same;
    QData/*63:0*/ sdiv;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    IData/*31:0*/ __VactIterCount;
    VlcuLocalArray<VlWide<4>/*99:0*/, 16> t08_widemem__DOT__wmem;  // thread-local memory
    VlcuSharedArray<QData/*63:0*/, 8> t08_widemem__DOT__qmem;  // run-time indexed: shared memory [index][lane]
    VlcuLocalArray<VlWide<4>/*99:0*/, 4> t08_widemem__DOT__shadow;  // thread-local memory
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vt08_widemem__cuda final {
    typedef Vt08_widemem___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 462,
        OFF_we0 = 463,
        OFF_we1 = 464,
        OFF_wa0 = 465,
        OFF_wa1 = 466,
        OFF_ra = 467,
        OFF_wd = 104,
        OFF_q = 0,
        OFF_rd = 120,
        OFF_qsum = 8,
        OFF_part = 460,
        OFF_same = 468,
        OFF_sdiv = 16,
        OFF___Vtrigprevexpr___TOP__clk__0 = 469,
        OFF___VactPhaseResult = 470,
        OFF___VnbaPhaseResult = 471,
        OFF___VactIterCount = 136,
        OFF_t08_widemem__DOT__wmem = 140,
        OFF_t08_widemem__DOT__qmem = 24,
        OFF_t08_widemem__DOT__shadow = 396,
        OFF___VactTriggered = 88,
        OFF___VnbaTriggered = 96,
        kStateBytes = 472,  // S_persist
        kTouchedReadBytes = 421,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 258,  // S_state_wr
    };
    struct Locals final {
        VlWide<4>/*99:0*/ t08_widemem__DOT__wmem[16];
        VlWide<4>/*99:0*/ t08_widemem__DOT__shadow[4];
    };
    struct Shared final {
        QData/*63:0*/ t08_widemem__DOT__qmem[8][VLCU_BLOCK];
    };
    VLCU_D static void bind(Root& s, Locals& l, Shared& sh, const VlcuState& st, uint64_t i) {
        (void)l; (void)sh; (void)st; (void)i;
        s.vlNamep = "TOP";
        s.t08_widemem__DOT__wmem.bind(l.t08_widemem__DOT__wmem);
        s.t08_widemem__DOT__qmem.bind(sh.t08_widemem__DOT__qmem, threadIdx.x);
        s.t08_widemem__DOT__shadow.bind(l.t08_widemem__DOT__shadow);
    }
    VLCU_D static void load(Root& s, const VlcuState& st, uint64_t i) {
        s.clk = st.ld<VlcuMem8>(OFF_clk, i);
        s.we0 = st.ld<VlcuMem8>(OFF_we0, i);
        s.we1 = st.ld<VlcuMem8>(OFF_we1, i);
        s.wa0 = st.ld<VlcuMem8>(OFF_wa0, i);
        s.wa1 = st.ld<VlcuMem8>(OFF_wa1, i);
        s.ra = st.ld<VlcuMem8>(OFF_ra, i);
        st.ldW(OFF_wd, i, s.wd);
        s.q = st.ld<QData>(OFF_q, i);
        st.ldW(OFF_rd, i, s.rd);
        s.qsum = st.ld<QData>(OFF_qsum, i);
        s.part = st.ld<VlcuMem16>(OFF_part, i);
        s.same = st.ld<VlcuMem8>(OFF_same, i);
        s.sdiv = st.ld<QData>(OFF_sdiv, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        for (uint32_t e = 0; e < 16; ++e) {
            for (uint32_t w = 0; w < 4; ++w) s.t08_widemem__DOT__wmem[e][w] = st.ldE<EData>(OFF_t08_widemem__DOT__wmem, e * 4 + w, i);
        }
        for (uint32_t e = 0; e < 8; ++e) s.t08_widemem__DOT__qmem[e] = st.ldE<QData>(OFF_t08_widemem__DOT__qmem, e, i);
        for (uint32_t e = 0; e < 4; ++e) {
            for (uint32_t w = 0; w < 4; ++w) s.t08_widemem__DOT__shadow[e][w] = st.ldE<EData>(OFF_t08_widemem__DOT__shadow, e * 4 + w, i);
        }
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.st<VlcuMem8>(OFF_we0, i, s.we0);
        st.st<VlcuMem8>(OFF_we1, i, s.we1);
        st.st<VlcuMem8>(OFF_wa0, i, s.wa0);
        st.st<VlcuMem8>(OFF_wa1, i, s.wa1);
        st.st<VlcuMem8>(OFF_ra, i, s.ra);
        st.stW(OFF_wd, i, s.wd);
        st.st<QData>(OFF_q, i, s.q);
        st.stW(OFF_rd, i, s.rd);
        st.st<QData>(OFF_qsum, i, s.qsum);
        st.st<VlcuMem16>(OFF_part, i, s.part);
        st.st<VlcuMem8>(OFF_same, i, s.same);
        st.st<QData>(OFF_sdiv, i, s.sdiv);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
        for (uint32_t e = 0; e < 16; ++e) {
            for (uint32_t w = 0; w < 4; ++w) st.stE<EData>(OFF_t08_widemem__DOT__wmem, e * 4 + w, i, s.t08_widemem__DOT__wmem[e][w]);
        }
        for (uint32_t e = 0; e < 8; ++e) st.stE<QData>(OFF_t08_widemem__DOT__qmem, e, i, s.t08_widemem__DOT__qmem[e]);
        for (uint32_t e = 0; e < 4; ++e) {
            for (uint32_t w = 0; w < 4; ++w) st.stE<EData>(OFF_t08_widemem__DOT__shadow, e * 4 + w, i, s.t08_widemem__DOT__shadow[e][w]);
        }
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VactTriggered, e, i, s.__VactTriggered[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VnbaTriggered, e, i, s.__VnbaTriggered[e]);
    }
    VLCU_D static void load_inputs_row(Root& s, const uint32_t* rowp, uint64_t stride, uint64_t i) {
        s.clk = static_cast<VlcuMem8>(rowp[0 * stride + i] & 0x1U);
        s.we0 = static_cast<VlcuMem8>(rowp[1 * stride + i] & 0x1U);
        s.we1 = static_cast<VlcuMem8>(rowp[2 * stride + i] & 0x1U);
        s.wa0 = static_cast<VlcuMem8>(rowp[3 * stride + i] & 0xfU);
        s.wa1 = static_cast<VlcuMem8>(rowp[4 * stride + i] & 0xfU);
        s.ra = static_cast<VlcuMem8>(rowp[5 * stride + i] & 0xfU);
        s.wd[0] = rowp[6 * stride + i] & 0xffffffffU;
        s.wd[1] = rowp[7 * stride + i] & 0xffffffffU;
        s.wd[2] = rowp[8 * stride + i] & 0xffffffffU;
        s.wd[3] = rowp[9 * stride + i] & 0xfU;
        s.q = (static_cast<QData>(rowp[11 * stride + i] & 0xffffffffU) << 32) | rowp[10 * stride + i];
    }
    VLCU_D static void store_outputs_row(const Root& s, uint32_t* rowp, uint64_t stride, uint64_t i) {
        rowp[0 * stride + i] = s.rd[0];
        rowp[1 * stride + i] = s.rd[1];
        rowp[2 * stride + i] = s.rd[2];
        rowp[3 * stride + i] = s.rd[3];
        rowp[4 * stride + i] = static_cast<uint32_t>(s.qsum);
        rowp[5 * stride + i] = static_cast<uint32_t>(s.qsum >> 32);
        rowp[6 * stride + i] = s.part;
        rowp[7 * stride + i] = s.same;
        rowp[8 * stride + i] = static_cast<uint32_t>(s.sdiv);
        rowp[9 * stride + i] = static_cast<uint32_t>(s.sdiv >> 32);
    }
    VLCU_D static uint64_t outputs_sig(const Root& s, uint64_t sig) {
        uint32_t acc = VLCU_SIG_ACC_INIT;
        acc = vlcu_sig_word(acc, s.rd[0]);
        acc = vlcu_sig_word(acc, s.rd[1]);
        acc = vlcu_sig_word(acc, s.rd[2]);
        acc = vlcu_sig_word(acc, s.rd[3]);
        acc = vlcu_sig_word(acc, static_cast<uint32_t>(s.qsum));
        acc = vlcu_sig_word(acc, static_cast<uint32_t>(s.qsum >> 32));
        acc = vlcu_sig_word(acc, s.part);
        acc = vlcu_sig_word(acc, s.same);
        acc = vlcu_sig_word(acc, static_cast<uint32_t>(s.sdiv));
        acc = vlcu_sig_word(acc, static_cast<uint32_t>(s.sdiv >> 32));
        return vlcu_sig_mix(sig, acc);
    }
    enum : uint32_t { kInWords = 12, kOutWords = 10, kInBytes = 30, kOutBytes = 35, kTraceWords = 0 };

    VLCU_D static void Vt08_widemem___024root___eval_static(Vt08_widemem___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
    }

    VLCU_D static void Vt08_widemem___024root___eval_initial(Vt08_widemem___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt08_widemem___024root___eval_final(Vt08_widemem___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt08_widemem___024root___eval_settle(Vt08_widemem___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static bool Vt08_widemem___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt08_widemem___024root___nba_sequent__TOP__0(Vt08_widemem___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        VlWide<4>/*99:0*/ __VdlyVal__t08_widemem__DOT__wmem__v0;
        VL_ZERO_W(100, __VdlyVal__t08_widemem__DOT__wmem__v0);
        CData/*3:0*/ __VdlyDim0__t08_widemem__DOT__wmem__v0;
        __VdlyDim0__t08_widemem__DOT__wmem__v0 = 0;
        CData/*0:0*/ __VdlySet__t08_widemem__DOT__wmem__v0;
        __VdlySet__t08_widemem__DOT__wmem__v0 = 0;
        VlWide<4>/*99:0*/ __VdlyVal__t08_widemem__DOT__wmem__v1;
        VL_ZERO_W(100, __VdlyVal__t08_widemem__DOT__wmem__v1);
        CData/*3:0*/ __VdlyDim0__t08_widemem__DOT__wmem__v1;
        __VdlyDim0__t08_widemem__DOT__wmem__v1 = 0;
        CData/*0:0*/ __VdlySet__t08_widemem__DOT__wmem__v1;
        __VdlySet__t08_widemem__DOT__wmem__v1 = 0;
        QData/*63:0*/ __VdlyVal__t08_widemem__DOT__qmem__v0;
        __VdlyVal__t08_widemem__DOT__qmem__v0 = 0;
        CData/*2:0*/ __VdlyDim0__t08_widemem__DOT__qmem__v0;
        __VdlyDim0__t08_widemem__DOT__qmem__v0 = 0;
        VlWide<4>/*99:0*/ __VdlyVal__t08_widemem__DOT__shadow__v0;
        VL_ZERO_W(100, __VdlyVal__t08_widemem__DOT__shadow__v0);
        CData/*1:0*/ __VdlyDim0__t08_widemem__DOT__shadow__v0;
        __VdlyDim0__t08_widemem__DOT__shadow__v0 = 0;
        VlWide<4>/*127:0*/ __Vtemp_2;
        VlWide<4>/*127:0*/ __Vtemp_3;
        VlWide<4>/*127:0*/ __Vtemp_4;
        // Body
        __VdlySet__t08_widemem__DOT__wmem__v0 = 0U;
        __VdlySet__t08_widemem__DOT__wmem__v1 = 0U;
        if (vlSelfRef.we0) {
            __VdlyVal__t08_widemem__DOT__wmem__v0[0U] 
                = vlSelfRef.wd[0U];
            __VdlyVal__t08_widemem__DOT__wmem__v0[1U] 
                = vlSelfRef.wd[1U];
            __VdlyVal__t08_widemem__DOT__wmem__v0[2U] 
                = vlSelfRef.wd[2U];
            __VdlyVal__t08_widemem__DOT__wmem__v0[3U] 
                = vlSelfRef.wd[3U];
            __VdlyDim0__t08_widemem__DOT__wmem__v0 
                = vlSelfRef.wa0;
            __VdlySet__t08_widemem__DOT__wmem__v0 = 1U;
        }
        if (vlSelfRef.we1) {
            __Vtemp_2[0U] = (~ vlSelfRef.wd[0U]);
            __Vtemp_2[1U] = (~ vlSelfRef.wd[1U]);
            __Vtemp_2[2U] = (~ vlSelfRef.wd[2U]);
            __Vtemp_2[3U] = (~ vlSelfRef.wd[3U]);
            __Vtemp_3[0U] = (IData)(vlSelfRef.q);
            __Vtemp_3[1U] = (IData)((vlSelfRef.q >> 0x00000020U));
            __Vtemp_3[2U] = 0U;
            __Vtemp_3[3U] = 0U;
            VL_ADD_W(4, __Vtemp_4, __Vtemp_2, __Vtemp_3);
            __VdlyVal__t08_widemem__DOT__wmem__v1[0U] 
                = __Vtemp_4[0U];
            __VdlyVal__t08_widemem__DOT__wmem__v1[1U] 
                = __Vtemp_4[1U];
            __VdlyVal__t08_widemem__DOT__wmem__v1[2U] 
                = __Vtemp_4[2U];
            __VdlyVal__t08_widemem__DOT__wmem__v1[3U] 
                = (0x0000000fU & __Vtemp_4[3U]);
            __VdlyDim0__t08_widemem__DOT__wmem__v1 
                = vlSelfRef.wa1;
            __VdlySet__t08_widemem__DOT__wmem__v1 = 1U;
        }
        __VdlyVal__t08_widemem__DOT__qmem__v0 = (vlSelfRef.t08_widemem__DOT__qmem
                                                 [(7U 
                                                   & (IData)(vlSelfRef.wa1))] 
                                                 + vlSelfRef.q);
        __VdlyDim0__t08_widemem__DOT__qmem__v0 = (7U 
                                                  & (IData)(vlSelfRef.wa0));
        vlSelfRef.qsum = (vlSelfRef.t08_widemem__DOT__qmem
                          [(7U & (IData)(vlSelfRef.ra))] 
                          - vlSelfRef.t08_widemem__DOT__qmem
                          [(7U & (IData)(vlSelfRef.wa0))]);
        vlSelfRef.sdiv = (VL_SHIFTRS_QQI(64,64,4, vlSelfRef.q, (IData)(vlSelfRef.wa0)) 
                          + VL_DIVS_QQQ(64, vlSelfRef.t08_widemem__DOT__qmem
                                        [(7U & (IData)(vlSelfRef.ra))], 
                                        (1ULL | (QData)((IData)(vlSelfRef.wa1)))));
        vlSelfRef.rd[0U] = (vlSelfRef.t08_widemem__DOT__wmem
                            [vlSelfRef.ra][0U] ^ (IData)(
                                                         (0x0003ffffffffffffULL 
                                                          & (((QData)((IData)(vlSelfRef.t08_widemem__DOT__wmem
                                                                              [vlSelfRef.wa1][3U])) 
                                                              << 0x0000002eU) 
                                                             | (((QData)((IData)(vlSelfRef.t08_widemem__DOT__wmem
                                                                                [vlSelfRef.wa1][2U])) 
                                                                 << 0x0000000eU) 
                                                                | ((QData)((IData)(vlSelfRef.t08_widemem__DOT__wmem
                                                                                [vlSelfRef.wa1][1U])) 
                                                                   >> 0x00000012U))))));
        vlSelfRef.rd[1U] = (vlSelfRef.t08_widemem__DOT__wmem
                            [vlSelfRef.ra][1U] ^ (((IData)(
                                                           (0x0003ffffffffffffULL 
                                                            & (((QData)((IData)(vlSelfRef.t08_widemem__DOT__wmem
                                                                                [vlSelfRef.wa0][1U])) 
                                                                << 0x00000020U) 
                                                               | (QData)((IData)(vlSelfRef.t08_widemem__DOT__wmem
                                                                                [vlSelfRef.wa0][0U]))))) 
                                                   << 0x00000012U) 
                                                  | (IData)(
                                                            ((0x0003ffffffffffffULL 
                                                              & (((QData)((IData)(vlSelfRef.t08_widemem__DOT__wmem
                                                                                [vlSelfRef.wa1][3U])) 
                                                                  << 0x0000002eU) 
                                                                 | (((QData)((IData)(vlSelfRef.t08_widemem__DOT__wmem
                                                                                [vlSelfRef.wa1][2U])) 
                                                                     << 0x0000000eU) 
                                                                    | ((QData)((IData)(vlSelfRef.t08_widemem__DOT__wmem
                                                                                [vlSelfRef.wa1][1U])) 
                                                                       >> 0x00000012U)))) 
                                                             >> 0x00000020U))));
        vlSelfRef.rd[2U] = (vlSelfRef.t08_widemem__DOT__wmem
                            [vlSelfRef.ra][2U] ^ (((IData)(
                                                           (0x0003ffffffffffffULL 
                                                            & (((QData)((IData)(vlSelfRef.t08_widemem__DOT__wmem
                                                                                [vlSelfRef.wa0][1U])) 
                                                                << 0x00000020U) 
                                                               | (QData)((IData)(vlSelfRef.t08_widemem__DOT__wmem
                                                                                [vlSelfRef.wa0][0U]))))) 
                                                   >> 0x0000000eU) 
                                                  | ((IData)(
                                                             ((0x0003ffffffffffffULL 
                                                               & (((QData)((IData)(vlSelfRef.t08_widemem__DOT__wmem
                                                                                [vlSelfRef.wa0][1U])) 
                                                                   << 0x00000020U) 
                                                                  | (QData)((IData)(vlSelfRef.t08_widemem__DOT__wmem
                                                                                [vlSelfRef.wa0][0U])))) 
                                                              >> 0x00000020U)) 
                                                     << 0x00000012U)));
        vlSelfRef.rd[3U] = (vlSelfRef.t08_widemem__DOT__wmem
                            [vlSelfRef.ra][3U] ^ ((IData)(
                                                          ((0x0003ffffffffffffULL 
                                                            & (((QData)((IData)(vlSelfRef.t08_widemem__DOT__wmem
                                                                                [vlSelfRef.wa0][1U])) 
                                                                << 0x00000020U) 
                                                               | (QData)((IData)(vlSelfRef.t08_widemem__DOT__wmem
                                                                                [vlSelfRef.wa0][0U])))) 
                                                           >> 0x00000020U)) 
                                                  >> 0x0000000eU));
        vlSelfRef.part = (0x0000ffffU & (((0U == (0x0000001fU 
                                                  & (IData)(vlSelfRef.q)))
                                           ? 0U : (vlSelfRef.t08_widemem__DOT__wmem
                                                   [vlSelfRef.ra]
                                                   [
                                                   (((IData)(0x0000000fU) 
                                                     + 
                                                     (0x0000003fU 
                                                      & (IData)(vlSelfRef.q))) 
                                                    >> 5U)] 
                                                   << 
                                                   ((IData)(0x00000020U) 
                                                    - 
                                                    (0x0000001fU 
                                                     & (IData)(vlSelfRef.q))))) 
                                         | (vlSelfRef.t08_widemem__DOT__wmem
                                            [vlSelfRef.ra]
                                            [(1U & 
                                              ((IData)(vlSelfRef.q) 
                                               >> 5U))] 
                                            >> (0x0000001fU 
                                                & (IData)(vlSelfRef.q)))));
        vlSelfRef.same = ((0U == ((((vlSelfRef.t08_widemem__DOT__wmem
                                     [vlSelfRef.ra][0U] 
                                     ^ vlSelfRef.t08_widemem__DOT__wmem
                                     [vlSelfRef.wa0][0U]) 
                                    | (vlSelfRef.t08_widemem__DOT__wmem
                                       [vlSelfRef.ra][1U] 
                                       ^ vlSelfRef.t08_widemem__DOT__wmem
                                       [vlSelfRef.wa0][1U])) 
                                   | (vlSelfRef.t08_widemem__DOT__wmem
                                      [vlSelfRef.ra][2U] 
                                      ^ vlSelfRef.t08_widemem__DOT__wmem
                                      [vlSelfRef.wa0][2U])) 
                                  | (vlSelfRef.t08_widemem__DOT__wmem
                                     [vlSelfRef.ra][3U] 
                                     ^ vlSelfRef.t08_widemem__DOT__wmem
                                     [vlSelfRef.wa0][3U]))) 
                          | VL_LT_W(4, vlSelfRef.t08_widemem__DOT__wmem
                                    [vlSelfRef.wa1], vlSelfRef.wd));
        if ((0U != ((((vlSelfRef.t08_widemem__DOT__shadow
                       [(3U & (IData)(vlSelfRef.wa0))][0U] 
                       ^ vlSelfRef.t08_widemem__DOT__shadow
                       [(3U & (IData)(vlSelfRef.wa1))][0U]) 
                      | (vlSelfRef.t08_widemem__DOT__shadow
                         [(3U & (IData)(vlSelfRef.wa0))][1U] 
                         ^ vlSelfRef.t08_widemem__DOT__shadow
                         [(3U & (IData)(vlSelfRef.wa1))][1U])) 
                     | (vlSelfRef.t08_widemem__DOT__shadow
                        [(3U & (IData)(vlSelfRef.wa0))][2U] 
                        ^ vlSelfRef.t08_widemem__DOT__shadow
                        [(3U & (IData)(vlSelfRef.wa1))][2U])) 
                    | (vlSelfRef.t08_widemem__DOT__shadow
                       [(3U & (IData)(vlSelfRef.wa0))][3U] 
                       ^ vlSelfRef.t08_widemem__DOT__shadow
                       [(3U & (IData)(vlSelfRef.wa1))][3U])))) {
            vlSelfRef.rd[0U] = ((0xfffffffeU & vlSelfRef.rd[0U]) 
                                | (1U & (~ vlSelfRef.wd[0U])));
        }
        __VdlyVal__t08_widemem__DOT__shadow__v0[0U] 
            = vlSelfRef.t08_widemem__DOT__wmem[vlSelfRef.ra][0U];
        __VdlyVal__t08_widemem__DOT__shadow__v0[1U] 
            = vlSelfRef.t08_widemem__DOT__wmem[vlSelfRef.ra][1U];
        __VdlyVal__t08_widemem__DOT__shadow__v0[2U] 
            = vlSelfRef.t08_widemem__DOT__wmem[vlSelfRef.ra][2U];
        __VdlyVal__t08_widemem__DOT__shadow__v0[3U] 
            = vlSelfRef.t08_widemem__DOT__wmem[vlSelfRef.ra][3U];
        __VdlyDim0__t08_widemem__DOT__shadow__v0 = 
            (3U & (IData)(vlSelfRef.ra));
        vlSelfRef.t08_widemem__DOT__qmem[__VdlyDim0__t08_widemem__DOT__qmem__v0] 
            = __VdlyVal__t08_widemem__DOT__qmem__v0;
        vlSelfRef.t08_widemem__DOT__shadow[__VdlyDim0__t08_widemem__DOT__shadow__v0][0U] 
            = __VdlyVal__t08_widemem__DOT__shadow__v0[0U];
        vlSelfRef.t08_widemem__DOT__shadow[__VdlyDim0__t08_widemem__DOT__shadow__v0][1U] 
            = __VdlyVal__t08_widemem__DOT__shadow__v0[1U];
        vlSelfRef.t08_widemem__DOT__shadow[__VdlyDim0__t08_widemem__DOT__shadow__v0][2U] 
            = __VdlyVal__t08_widemem__DOT__shadow__v0[2U];
        vlSelfRef.t08_widemem__DOT__shadow[__VdlyDim0__t08_widemem__DOT__shadow__v0][3U] 
            = __VdlyVal__t08_widemem__DOT__shadow__v0[3U];
        if (__VdlySet__t08_widemem__DOT__wmem__v0) {
            vlSelfRef.t08_widemem__DOT__wmem[__VdlyDim0__t08_widemem__DOT__wmem__v0][0U] 
                = __VdlyVal__t08_widemem__DOT__wmem__v0[0U];
            vlSelfRef.t08_widemem__DOT__wmem[__VdlyDim0__t08_widemem__DOT__wmem__v0][1U] 
                = __VdlyVal__t08_widemem__DOT__wmem__v0[1U];
            vlSelfRef.t08_widemem__DOT__wmem[__VdlyDim0__t08_widemem__DOT__wmem__v0][2U] 
                = __VdlyVal__t08_widemem__DOT__wmem__v0[2U];
            vlSelfRef.t08_widemem__DOT__wmem[__VdlyDim0__t08_widemem__DOT__wmem__v0][3U] 
                = __VdlyVal__t08_widemem__DOT__wmem__v0[3U];
        }
        if (__VdlySet__t08_widemem__DOT__wmem__v1) {
            vlSelfRef.t08_widemem__DOT__wmem[__VdlyDim0__t08_widemem__DOT__wmem__v1][0U] 
                = __VdlyVal__t08_widemem__DOT__wmem__v1[0U];
            vlSelfRef.t08_widemem__DOT__wmem[__VdlyDim0__t08_widemem__DOT__wmem__v1][1U] 
                = __VdlyVal__t08_widemem__DOT__wmem__v1[1U];
            vlSelfRef.t08_widemem__DOT__wmem[__VdlyDim0__t08_widemem__DOT__wmem__v1][2U] 
                = __VdlyVal__t08_widemem__DOT__wmem__v1[2U];
            vlSelfRef.t08_widemem__DOT__wmem[__VdlyDim0__t08_widemem__DOT__wmem__v1][3U] 
                = __VdlyVal__t08_widemem__DOT__wmem__v1[3U];
        }
    }

    VLCU_D static void Vt08_widemem___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt08_widemem___024root___eval_phase__act(Vt08_widemem___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt08_widemem___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vt08_widemem___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vt08_widemem___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vt08_widemem___024root___eval_phase__nba(Vt08_widemem___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vt08_widemem___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    Vt08_widemem___024root___nba_sequent__TOP__0(vlSelf, vlCtx);
                }
            }
            Vt08_widemem___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vt08_widemem___024root___eval(Vt08_widemem___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt08_widemem___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t08_widemem.v", 4, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vt08_widemem___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t08_widemem.v", 4, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vt08_widemem___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vt08_widemem___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vt08_widemem___024root___ctor_var_reset(Vt08_widemem___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        vlSelf->we0 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 15191994881821842139ull);
        vlSelf->we1 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 8911534591852561079ull);
        vlSelf->wa0 = VL_SCOPED_RAND_RESET_I(4, __VscopeHash, 13614112491244487440ull);
        vlSelf->wa1 = VL_SCOPED_RAND_RESET_I(4, __VscopeHash, 12979143277029874325ull);
        vlSelf->ra = VL_SCOPED_RAND_RESET_I(4, __VscopeHash, 397185324164439704ull);
        VL_SCOPED_RAND_RESET_W(100, vlSelf->wd, __VscopeHash, 18210509495843279068ull);
        vlSelf->q = VL_SCOPED_RAND_RESET_Q(64, __VscopeHash, 8861071527689086543ull);
        VL_SCOPED_RAND_RESET_W(100, vlSelf->rd, __VscopeHash, 10303652599525973633ull);
        vlSelf->qsum = VL_SCOPED_RAND_RESET_Q(64, __VscopeHash, 16335810068866503488ull);
        vlSelf->part = VL_SCOPED_RAND_RESET_I(16, __VscopeHash, 14213356343002679141ull);
        vlSelf->same = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 5563288323055129069ull);
        vlSelf->sdiv = VL_SCOPED_RAND_RESET_Q(64, __VscopeHash, 13644446615494747826ull);
        for (int __Vi0 = 0; __Vi0 < 16; ++__Vi0) {
            VL_SCOPED_RAND_RESET_W(100, vlSelf->t08_widemem__DOT__wmem[__Vi0], __VscopeHash, 12427546446607791109ull);
        }
        for (int __Vi0 = 0; __Vi0 < 8; ++__Vi0) {
            vlSelf->t08_widemem__DOT__qmem[__Vi0] = VL_SCOPED_RAND_RESET_Q(64, __VscopeHash, 8083714073266708241ull);
        }
        for (int __Vi0 = 0; __Vi0 < 4; ++__Vi0) {
            VL_SCOPED_RAND_RESET_W(100, vlSelf->t08_widemem__DOT__shadow[__Vi0], __VscopeHash, 11783151868218910088ull);
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
        Vt08_widemem___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt08_widemem___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt08_widemem___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt08_widemem___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vt08_widemem___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt08_widemem___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt08_widemem__cuda::info() {
    typedef Vt08_widemem__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"we0", "we0", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_we0},
        {"we1", "we1", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_we1},
        {"wa0", "wa0", VLCU_MEMBER_IN, 1, 0, 1, 4, 3, 0, T::OFF_wa0},
        {"wa1", "wa1", VLCU_MEMBER_IN, 1, 0, 1, 4, 3, 0, T::OFF_wa1},
        {"ra", "ra", VLCU_MEMBER_IN, 1, 0, 1, 4, 3, 0, T::OFF_ra},
        {"wd", "wd", VLCU_MEMBER_IN, 4, 0, 4, 100, 99, 0, T::OFF_wd},
        {"q", "q", VLCU_MEMBER_IN, 8, 0, 1, 64, 63, 0, T::OFF_q},
        {"rd", "rd", VLCU_MEMBER_OUT, 4, 0, 4, 100, 99, 0, T::OFF_rd},
        {"qsum", "qsum", VLCU_MEMBER_OUT, 8, 0, 1, 64, 63, 0, T::OFF_qsum},
        {"part", "part", VLCU_MEMBER_OUT, 2, 0, 1, 16, 15, 0, T::OFF_part},
        {"same", "same", VLCU_MEMBER_OUT, 1, 0, 1, 1, 0, 0, T::OFF_same},
        {"sdiv", "sdiv", VLCU_MEMBER_OUT, 8, 0, 1, 64, 63, 0, T::OFF_sdiv},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"t08_widemem__DOT__wmem", "t08_widemem.wmem", VLCU_MEMBER_STATE, 4, 0, 64, 100, 99, 0, T::OFF_t08_widemem__DOT__wmem},
        {"t08_widemem__DOT__qmem", "t08_widemem.qmem", VLCU_MEMBER_STATE, 8, 0, 8, 64, 63, 0, T::OFF_t08_widemem__DOT__qmem},
        {"t08_widemem__DOT__shadow", "t08_widemem.shadow", VLCU_MEMBER_STATE, 4, 0, 16, 100, 99, 0, T::OFF_t08_widemem__DOT__shadow},
        {"__VactTriggered", "__VactTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VactTriggered},
        {"__VnbaTriggered", "__VnbaTriggered", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___VnbaTriggered},
    };
    static const char* const messages[] = {nullptr};
    static const char* const plusargs[] = {nullptr};
    static const VlcuTraceDecl traceDecls[] = {
        {VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0}};  // (terminator, not counted)
    static const uint32_t traceOrder[] = {0};  // codes in dump order (+ terminator)
    static const uint32_t saveOrder[] = {0, 1, 2, 3, 4, 5, 11, 13, 14, 15, 10, 6, 8, 16, 7, 9, 12, 17, 18, 19, 20, 21, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "t08_widemem", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 0, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 22, 0, 0xc8b7f9412b1760ddULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt08_widemem__cuda VLCU_TESTBENCHES);
