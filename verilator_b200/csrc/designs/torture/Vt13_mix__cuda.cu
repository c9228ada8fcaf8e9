// Verilated -*- C++ -*-
// DESCRIPTION: Verilator output: Batched CUDA model (one instance per thread) for include/verilated_cuda.h
// Design '$root' printed by V3EmitCuda; compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include

#include "verilated_cuda_launch.h"

#include <vector>

// Name the design registers under with vlcu_create()
#ifndef VLCU_DESIGN_NAME
#define VLCU_DESIGN_NAME "Vt13_mix"
#endif

// The model image of one instance (Vtop___024root of the C++ model), held by one thread
struct Vt13_mix___024root final {
    CData/*0:0*/ clk;
    CData/*0:0*/ arst;
    CData/*0:0*/ en;
    CData/*6:0*/ sel;
    VlWide<4>/*99:0*/ wa;
    VlWide<4>/*99:0*/ wb;
    QData/*63:0*/ qa;
    SData/*15:0*/ d;
    VlWide<4>/*127:0*/ wr;
    QData/*63:0*/ qr;
    VlWide<4>/*99:0*/ sprod;
    VlWide<4>/*99:0*/ sdiv;
    CData/*7:0*/ cmp;
    SData/*15:0*/ lat;
    CData/*7:0*/ cnt_div;
    CData/*7:0*/ cnt_gate;
    VlWide<4>/*99:0*/ pw;
    CData/*6:0*/ lg;
    CData/*0:0*/ t13_mix__DOT__div2;
    CData/*7:0*/ t13_mix__DOT__hi8;
    CData/*0:0*/ __VstlFirstIteration;
    CData/*0:0*/ __VstlPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData/*0:0*/ __Vtrigprevexpr___TOP__arst__0;
    CData/*0:0*/ __Vtrigprevexpr___TOP__en__0;
    CData/*6:0*/ __Vtrigprevexpr___TOP__sel__0;
    CData/*0:0*/ __VicoDidInit;
    CData/*0:0*/ __VicoPhaseResult;
    CData/*0:0*/ __Vtrigprevexpr___TOP__arst__1;
    CData/*0:0*/ __Vtrigprevexpr___TOP__clk__1;
    CData/*0:0*/ __Vtrigprevexpr___TOP__t13_mix__DOT__div2__0;
    CData/*0:0*/ __Vtrigprevexpr_h364b5a28__1;
    CData/*0:0*/ __VactPhaseResult;
    CData/*0:0*/ __VnbaPhaseResult;
    SData/*15:0*/ __Vtrigprevexpr___TOP__d__0;
    VlWide<4>/*99:0*/ __Vtrigprevexpr___TOP__wa__0;
    VlWide<4>/*99:0*/ __Vtrigprevexpr___TOP__wb__0;
    IData/*31:0*/ __VactIterCount;
    QData/*63:0*/ __Vtrigprevexpr___TOP__qa__0;
    VlUnpacked<QData/*63:0*/, 1> __VstlTriggered;
    VlUnpacked<QData/*63:0*/, 2> __VicoTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData/*63:0*/, 1> __VnbaTriggered;
    const char* vlNamep;
};

struct Vt13_mix__cuda final {
    typedef Vt13_mix___024root Root;
    static constexpr bool kTiled = false;
    // Structure-of-arrays slot table: byte offset per instance of every persistent member
    enum : uint32_t {
        OFF_clk = 202,
        OFF_arst = 203,
        OFF_en = 204,
        OFF_sel = 205,
        OFF_wa = 64,
        OFF_wb = 80,
        OFF_qa = 0,
        OFF_d = 196,
        OFF_wr = 96,
        OFF_qr = 8,
        OFF_sprod = 112,
        OFF_sdiv = 128,
        OFF_cmp = 206,
        OFF_lat = 198,
        OFF_cnt_div = 207,
        OFF_cnt_gate = 208,
        OFF_pw = 144,
        OFF_lg = 209,
        OFF_t13_mix__DOT__div2 = 210,
        OFF_t13_mix__DOT__hi8 = 211,
        OFF___VstlFirstIteration = 212,
        OFF___VstlPhaseResult = 213,
        OFF___Vtrigprevexpr___TOP__clk__0 = 214,
        OFF___Vtrigprevexpr___TOP__arst__0 = 215,
        OFF___Vtrigprevexpr___TOP__en__0 = 216,
        OFF___Vtrigprevexpr___TOP__sel__0 = 217,
        OFF___VicoDidInit = 218,
        OFF___VicoPhaseResult = 219,
        OFF___Vtrigprevexpr___TOP__arst__1 = 220,
        OFF___Vtrigprevexpr___TOP__clk__1 = 221,
        OFF___Vtrigprevexpr___TOP__t13_mix__DOT__div2__0 = 222,
        OFF___Vtrigprevexpr_h364b5a28__1 = 223,
        OFF___VactPhaseResult = 224,
        OFF___VnbaPhaseResult = 225,
        OFF___Vtrigprevexpr___TOP__d__0 = 200,
        OFF___Vtrigprevexpr___TOP__wa__0 = 160,
        OFF___Vtrigprevexpr___TOP__wb__0 = 176,
        OFF___VactIterCount = 192,
        OFF___Vtrigprevexpr___TOP__qa__0 = 16,
        OFF___VstlTriggered = 24,
        OFF___VicoTriggered = 32,
        OFF___VactTriggered = 48,
        OFF___VnbaTriggered = 56,
        kStateBytes = 232,  // S_persist
        kTouchedReadBytes = 165,  // S_state_rd: persistent bytes one eval can read
        kTouchedWriteBytes = 170,  // S_state_wr
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
        s.arst = st.ld<VlcuMem8>(OFF_arst, i);
        s.en = st.ld<VlcuMem8>(OFF_en, i);
        s.sel = st.ld<VlcuMem8>(OFF_sel, i);
        st.ldW(OFF_wa, i, s.wa);
        st.ldW(OFF_wb, i, s.wb);
        s.qa = st.ld<QData>(OFF_qa, i);
        s.d = st.ld<VlcuMem16>(OFF_d, i);
        st.ldW(OFF_wr, i, s.wr);
        s.qr = st.ld<QData>(OFF_qr, i);
        st.ldW(OFF_sprod, i, s.sprod);
        st.ldW(OFF_sdiv, i, s.sdiv);
        s.cmp = st.ld<VlcuMem8>(OFF_cmp, i);
        s.lat = st.ld<VlcuMem16>(OFF_lat, i);
        s.cnt_div = st.ld<VlcuMem8>(OFF_cnt_div, i);
        s.cnt_gate = st.ld<VlcuMem8>(OFF_cnt_gate, i);
        st.ldW(OFF_pw, i, s.pw);
        s.lg = st.ld<VlcuMem8>(OFF_lg, i);
        s.t13_mix__DOT__div2 = st.ld<VlcuMem8>(OFF_t13_mix__DOT__div2, i);
        s.t13_mix__DOT__hi8 = st.ld<VlcuMem8>(OFF_t13_mix__DOT__hi8, i);
        s.__VstlFirstIteration = st.ld<VlcuMem8>(OFF___VstlFirstIteration, i);
        s.__VstlPhaseResult = st.ld<VlcuMem8>(OFF___VstlPhaseResult, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__Vtrigprevexpr___TOP__arst__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__arst__0, i);
        s.__Vtrigprevexpr___TOP__en__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__en__0, i);
        s.__Vtrigprevexpr___TOP__sel__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__sel__0, i);
        s.__VicoDidInit = st.ld<VlcuMem8>(OFF___VicoDidInit, i);
        s.__VicoPhaseResult = st.ld<VlcuMem8>(OFF___VicoPhaseResult, i);
        s.__Vtrigprevexpr___TOP__arst__1 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__arst__1, i);
        s.__Vtrigprevexpr___TOP__clk__1 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__1, i);
        s.__Vtrigprevexpr___TOP__t13_mix__DOT__div2__0 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr___TOP__t13_mix__DOT__div2__0, i);
        s.__Vtrigprevexpr_h364b5a28__1 = st.ld<VlcuMem8>(OFF___Vtrigprevexpr_h364b5a28__1, i);
        s.__VactPhaseResult = st.ld<VlcuMem8>(OFF___VactPhaseResult, i);
        s.__VnbaPhaseResult = st.ld<VlcuMem8>(OFF___VnbaPhaseResult, i);
        s.__Vtrigprevexpr___TOP__d__0 = st.ld<VlcuMem16>(OFF___Vtrigprevexpr___TOP__d__0, i);
        st.ldW(OFF___Vtrigprevexpr___TOP__wa__0, i, s.__Vtrigprevexpr___TOP__wa__0);
        st.ldW(OFF___Vtrigprevexpr___TOP__wb__0, i, s.__Vtrigprevexpr___TOP__wb__0);
        s.__VactIterCount = st.ld<IData>(OFF___VactIterCount, i);
        s.__Vtrigprevexpr___TOP__qa__0 = st.ld<QData>(OFF___Vtrigprevexpr___TOP__qa__0, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VstlTriggered[e] = st.ldE<QData>(OFF___VstlTriggered, e, i);
        for (uint32_t e = 0; e < 2; ++e) s.__VicoTriggered[e] = st.ldE<QData>(OFF___VicoTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VactTriggered[e] = st.ldE<QData>(OFF___VactTriggered, e, i);
        for (uint32_t e = 0; e < 1; ++e) s.__VnbaTriggered[e] = st.ldE<QData>(OFF___VnbaTriggered, e, i);
    }
    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<VlcuMem8>(OFF_clk, i, s.clk);
        st.st<VlcuMem8>(OFF_arst, i, s.arst);
        st.st<VlcuMem8>(OFF_en, i, s.en);
        st.st<VlcuMem8>(OFF_sel, i, s.sel);
        st.stW(OFF_wa, i, s.wa);
        st.stW(OFF_wb, i, s.wb);
        st.st<QData>(OFF_qa, i, s.qa);
        st.st<VlcuMem16>(OFF_d, i, s.d);
        st.stW(OFF_wr, i, s.wr);
        st.st<QData>(OFF_qr, i, s.qr);
        st.stW(OFF_sprod, i, s.sprod);
        st.stW(OFF_sdiv, i, s.sdiv);
        st.st<VlcuMem8>(OFF_cmp, i, s.cmp);
        st.st<VlcuMem16>(OFF_lat, i, s.lat);
        st.st<VlcuMem8>(OFF_cnt_div, i, s.cnt_div);
        st.st<VlcuMem8>(OFF_cnt_gate, i, s.cnt_gate);
        st.stW(OFF_pw, i, s.pw);
        st.st<VlcuMem8>(OFF_lg, i, s.lg);
        st.st<VlcuMem8>(OFF_t13_mix__DOT__div2, i, s.t13_mix__DOT__div2);
        st.st<VlcuMem8>(OFF_t13_mix__DOT__hi8, i, s.t13_mix__DOT__hi8);
        st.st<VlcuMem8>(OFF___VstlFirstIteration, i, s.__VstlFirstIteration);
        st.st<VlcuMem8>(OFF___VstlPhaseResult, i, s.__VstlPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__arst__0, i, s.__Vtrigprevexpr___TOP__arst__0);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__en__0, i, s.__Vtrigprevexpr___TOP__en__0);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__sel__0, i, s.__Vtrigprevexpr___TOP__sel__0);
        st.st<VlcuMem8>(OFF___VicoDidInit, i, s.__VicoDidInit);
        st.st<VlcuMem8>(OFF___VicoPhaseResult, i, s.__VicoPhaseResult);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__arst__1, i, s.__Vtrigprevexpr___TOP__arst__1);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__clk__1, i, s.__Vtrigprevexpr___TOP__clk__1);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr___TOP__t13_mix__DOT__div2__0, i, s.__Vtrigprevexpr___TOP__t13_mix__DOT__div2__0);
        st.st<VlcuMem8>(OFF___Vtrigprevexpr_h364b5a28__1, i, s.__Vtrigprevexpr_h364b5a28__1);
        st.st<VlcuMem8>(OFF___VactPhaseResult, i, s.__VactPhaseResult);
        st.st<VlcuMem8>(OFF___VnbaPhaseResult, i, s.__VnbaPhaseResult);
        st.st<VlcuMem16>(OFF___Vtrigprevexpr___TOP__d__0, i, s.__Vtrigprevexpr___TOP__d__0);
        st.stW(OFF___Vtrigprevexpr___TOP__wa__0, i, s.__Vtrigprevexpr___TOP__wa__0);
        st.stW(OFF___Vtrigprevexpr___TOP__wb__0, i, s.__Vtrigprevexpr___TOP__wb__0);
        st.st<IData>(OFF___VactIterCount, i, s.__VactIterCount);
        st.st<QData>(OFF___Vtrigprevexpr___TOP__qa__0, i, s.__Vtrigprevexpr___TOP__qa__0);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VstlTriggered, e, i, s.__VstlTriggered[e]);
        for (uint32_t e = 0; e < 2; ++e) st.stE<QData>(OFF___VicoTriggered, e, i, s.__VicoTriggered[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VactTriggered, e, i, s.__VactTriggered[e]);
        for (uint32_t e = 0; e < 1; ++e) st.stE<QData>(OFF___VnbaTriggered, e, i, s.__VnbaTriggered[e]);
    }
    VLCU_D static void load_inputs_row(Root& s, const uint32_t* rowp, uint64_t stride, uint64_t i) {
        s.clk = static_cast<VlcuMem8>(rowp[0 * stride + i] & 0x1U);
        s.arst = static_cast<VlcuMem8>(rowp[1 * stride + i] & 0x1U);
        s.en = static_cast<VlcuMem8>(rowp[2 * stride + i] & 0x1U);
        s.sel = static_cast<VlcuMem8>(rowp[3 * stride + i] & 0x7fU);
        s.wa[0] = rowp[4 * stride + i] & 0xffffffffU;
        s.wa[1] = rowp[5 * stride + i] & 0xffffffffU;
        s.wa[2] = rowp[6 * stride + i] & 0xffffffffU;
        s.wa[3] = rowp[7 * stride + i] & 0xfU;
        s.wb[0] = rowp[8 * stride + i] & 0xffffffffU;
        s.wb[1] = rowp[9 * stride + i] & 0xffffffffU;
        s.wb[2] = rowp[10 * stride + i] & 0xffffffffU;
        s.wb[3] = rowp[11 * stride + i] & 0xfU;
        s.qa = (static_cast<QData>(rowp[13 * stride + i] & 0xffffffffU) << 32) | rowp[12 * stride + i];
        s.d = static_cast<VlcuMem16>(rowp[14 * stride + i] & 0xffffU);
    }
    VLCU_D static void store_outputs_row(const Root& s, uint32_t* rowp, uint64_t stride, uint64_t i) {
        rowp[0 * stride + i] = s.wr[0];
        rowp[1 * stride + i] = s.wr[1];
        rowp[2 * stride + i] = s.wr[2];
        rowp[3 * stride + i] = s.wr[3];
        rowp[4 * stride + i] = static_cast<uint32_t>(s.qr);
        rowp[5 * stride + i] = static_cast<uint32_t>(s.qr >> 32);
        rowp[6 * stride + i] = s.sprod[0];
        rowp[7 * stride + i] = s.sprod[1];
        rowp[8 * stride + i] = s.sprod[2];
        rowp[9 * stride + i] = s.sprod[3];
        rowp[10 * stride + i] = s.sdiv[0];
        rowp[11 * stride + i] = s.sdiv[1];
        rowp[12 * stride + i] = s.sdiv[2];
        rowp[13 * stride + i] = s.sdiv[3];
        rowp[14 * stride + i] = s.cmp;
        rowp[15 * stride + i] = s.lat;
        rowp[16 * stride + i] = s.cnt_div;
        rowp[17 * stride + i] = s.cnt_gate;
        rowp[18 * stride + i] = s.pw[0];
        rowp[19 * stride + i] = s.pw[1];
        rowp[20 * stride + i] = s.pw[2];
        rowp[21 * stride + i] = s.pw[3];
        rowp[22 * stride + i] = s.lg;
    }
    VLCU_D static uint64_t outputs_sig(const Root& s, uint64_t sig) {
        uint32_t acc = VLCU_SIG_ACC_INIT;
        acc = vlcu_sig_word(acc, s.wr[0]);
        acc = vlcu_sig_word(acc, s.wr[1]);
        acc = vlcu_sig_word(acc, s.wr[2]);
        acc = vlcu_sig_word(acc, s.wr[3]);
        acc = vlcu_sig_word(acc, static_cast<uint32_t>(s.qr));
        acc = vlcu_sig_word(acc, static_cast<uint32_t>(s.qr >> 32));
        acc = vlcu_sig_word(acc, s.sprod[0]);
        acc = vlcu_sig_word(acc, s.sprod[1]);
        acc = vlcu_sig_word(acc, s.sprod[2]);
        acc = vlcu_sig_word(acc, s.sprod[3]);
        acc = vlcu_sig_word(acc, s.sdiv[0]);
        acc = vlcu_sig_word(acc, s.sdiv[1]);
        acc = vlcu_sig_word(acc, s.sdiv[2]);
        acc = vlcu_sig_word(acc, s.sdiv[3]);
        acc = vlcu_sig_word(acc, s.cmp);
        acc = vlcu_sig_word(acc, s.lat);
        acc = vlcu_sig_word(acc, s.cnt_div);
        acc = vlcu_sig_word(acc, s.cnt_gate);
        acc = vlcu_sig_word(acc, s.pw[0]);
        acc = vlcu_sig_word(acc, s.pw[1]);
        acc = vlcu_sig_word(acc, s.pw[2]);
        acc = vlcu_sig_word(acc, s.pw[3]);
        acc = vlcu_sig_word(acc, s.lg);
        return vlcu_sig_mix(sig, acc);
    }
    enum : uint32_t { kInWords = 15, kOutWords = 23, kInBytes = 46, kOutBytes = 78, kTraceWords = 0 };

    VLCU_D static void Vt13_mix___024root___eval_static(Vt13_mix___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
        vlSelfRef.__Vtrigprevexpr___TOP__arst__0 = vlSelfRef.arst;
        vlSelfRef.__Vtrigprevexpr___TOP__en__0 = vlSelfRef.en;
        vlSelfRef.__Vtrigprevexpr___TOP__sel__0 = vlSelfRef.sel;
        vlSelfRef.__Vtrigprevexpr___TOP__wa__0[0U] 
            = vlSelfRef.wa[0U];
        vlSelfRef.__Vtrigprevexpr___TOP__wa__0[1U] 
            = vlSelfRef.wa[1U];
        vlSelfRef.__Vtrigprevexpr___TOP__wa__0[2U] 
            = vlSelfRef.wa[2U];
        vlSelfRef.__Vtrigprevexpr___TOP__wa__0[3U] 
            = vlSelfRef.wa[3U];
        vlSelfRef.__Vtrigprevexpr___TOP__wb__0[0U] 
            = vlSelfRef.wb[0U];
        vlSelfRef.__Vtrigprevexpr___TOP__wb__0[1U] 
            = vlSelfRef.wb[1U];
        vlSelfRef.__Vtrigprevexpr___TOP__wb__0[2U] 
            = vlSelfRef.wb[2U];
        vlSelfRef.__Vtrigprevexpr___TOP__wb__0[3U] 
            = vlSelfRef.wb[3U];
        vlSelfRef.__Vtrigprevexpr___TOP__qa__0 = vlSelfRef.qa;
        vlSelfRef.__Vtrigprevexpr___TOP__d__0 = vlSelfRef.d;
        vlSelfRef.__Vtrigprevexpr___TOP__arst__1 = vlSelfRef.arst;
        vlSelfRef.__Vtrigprevexpr___TOP__clk__1 = vlSelfRef.clk;
        vlSelfRef.__Vtrigprevexpr___TOP__t13_mix__DOT__div2__0 
            = vlSelfRef.t13_mix__DOT__div2;
        vlSelfRef.__Vtrigprevexpr_h364b5a28__1 = ((IData)(vlSelfRef.clk) 
                                                  & (IData)(vlSelfRef.en));
    }

    VLCU_D static void Vt13_mix___024root___eval_initial(Vt13_mix___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt13_mix___024root___eval_final(Vt13_mix___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
    }

    VLCU_D static void Vt13_mix___024root___eval_settle(Vt13_mix___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt13_mix___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t13_mix.v", 5, "", "DIDNOTCONVERGE: Settle region did not converge after '--converge-limit' of 10000 tries");
            }
            __VstlIterCount = ((IData)(1U) + __VstlIterCount);
            vlSelfRef.__VstlPhaseResult = Vt13_mix___024root___eval_phase__stl(vlSelf, vlCtx);
            vlSelfRef.__VstlFirstIteration = 0U;
        } while (vlSelfRef.__VstlPhaseResult);
    }

    VLCU_D static bool Vt13_mix___024root___trigger_anySet__stl(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt13_mix___024root___eval_phase__stl(Vt13_mix___024root* vlSelf, VlcuContext& vlCtx) {
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
            Vt13_mix___024root___dump_triggers__stl(vlSelfRef.__VstlTriggered, "stl"s);
        }
#endif
        __VstlExecute = Vt13_mix___024root___trigger_anySet__stl(vlSelfRef.__VstlTriggered);
        if (__VstlExecute) {
            {
                // Inlined CFunc: _eval_stl
                if ((1ULL & vlSelfRef.__VstlTriggered[0U])) {
                    {
                        // Inlined CFunc: _ico_comb__TOP__0
                        if (vlSelfRef.en) {
                            vlSelfRef.lat = ((IData)(vlSelfRef.d) 
                                             ^ (((IData)(vlSelfRef.cnt_div) 
                                                 << 8U) 
                                                | (IData)(vlSelfRef.cnt_gate)));
                        }
                    }
                }
            }
        }
        return (__VstlExecute);
    }

    VLCU_D static void Vt13_mix___024root___eval_triggers_vec__ico(Vt13_mix___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        vlSelfRef.__VicoTriggered[0U] = (QData)((IData)(
                                                        (((((((IData)(vlSelfRef.d) 
                                                              != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__d__0)) 
                                                             << 3U) 
                                                            | ((vlSelfRef.qa 
                                                                != vlSelfRef.__Vtrigprevexpr___TOP__qa__0) 
                                                               << 2U)) 
                                                           | (((0U 
                                                                != 
                                                                ((((vlSelfRef.wb[0U] 
                                                                    ^ vlSelfRef.__Vtrigprevexpr___TOP__wb__0[0U]) 
                                                                   | (vlSelfRef.wb[1U] 
                                                                      ^ vlSelfRef.__Vtrigprevexpr___TOP__wb__0[1U])) 
                                                                  | (vlSelfRef.wb[2U] 
                                                                     ^ vlSelfRef.__Vtrigprevexpr___TOP__wb__0[2U])) 
                                                                 | (vlSelfRef.wb[3U] 
                                                                    ^ vlSelfRef.__Vtrigprevexpr___TOP__wb__0[3U]))) 
                                                               << 1U) 
                                                              | (0U 
                                                                 != 
                                                                 ((((vlSelfRef.wa[0U] 
                                                                     ^ vlSelfRef.__Vtrigprevexpr___TOP__wa__0[0U]) 
                                                                    | (vlSelfRef.wa[1U] 
                                                                       ^ vlSelfRef.__Vtrigprevexpr___TOP__wa__0[1U])) 
                                                                   | (vlSelfRef.wa[2U] 
                                                                      ^ vlSelfRef.__Vtrigprevexpr___TOP__wa__0[2U])) 
                                                                  | (vlSelfRef.wa[3U] 
                                                                     ^ vlSelfRef.__Vtrigprevexpr___TOP__wa__0[3U]))))) 
                                                          << 4U) 
                                                         | (((((IData)(vlSelfRef.sel) 
                                                               != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__sel__0)) 
                                                              << 3U) 
                                                             | (((IData)(vlSelfRef.en) 
                                                                 != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__en__0)) 
                                                                << 2U)) 
                                                            | ((((IData)(vlSelfRef.arst) 
                                                                 != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__arst__0)) 
                                                                << 1U) 
                                                               | ((IData)(vlSelfRef.clk) 
                                                                  != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__clk__0)))))));
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
        vlSelfRef.__Vtrigprevexpr___TOP__arst__0 = vlSelfRef.arst;
        vlSelfRef.__Vtrigprevexpr___TOP__en__0 = vlSelfRef.en;
        vlSelfRef.__Vtrigprevexpr___TOP__sel__0 = vlSelfRef.sel;
        vlSelfRef.__Vtrigprevexpr___TOP__wa__0[0U] 
            = vlSelfRef.wa[0U];
        vlSelfRef.__Vtrigprevexpr___TOP__wa__0[1U] 
            = vlSelfRef.wa[1U];
        vlSelfRef.__Vtrigprevexpr___TOP__wa__0[2U] 
            = vlSelfRef.wa[2U];
        vlSelfRef.__Vtrigprevexpr___TOP__wa__0[3U] 
            = vlSelfRef.wa[3U];
        vlSelfRef.__Vtrigprevexpr___TOP__wb__0[0U] 
            = vlSelfRef.wb[0U];
        vlSelfRef.__Vtrigprevexpr___TOP__wb__0[1U] 
            = vlSelfRef.wb[1U];
        vlSelfRef.__Vtrigprevexpr___TOP__wb__0[2U] 
            = vlSelfRef.wb[2U];
        vlSelfRef.__Vtrigprevexpr___TOP__wb__0[3U] 
            = vlSelfRef.wb[3U];
        vlSelfRef.__Vtrigprevexpr___TOP__qa__0 = vlSelfRef.qa;
        vlSelfRef.__Vtrigprevexpr___TOP__d__0 = vlSelfRef.d;
        if (VL_UNLIKELY(((1U & (~ (IData)(vlSelfRef.__VicoDidInit)))))) {
            vlSelfRef.__VicoDidInit = 1U;
            vlSelfRef.__VicoTriggered[0U] = (1ULL | vlSelfRef.__VicoTriggered[0U]);
            vlSelfRef.__VicoTriggered[0U] = (2ULL | vlSelfRef.__VicoTriggered[0U]);
            vlSelfRef.__VicoTriggered[0U] = (4ULL | vlSelfRef.__VicoTriggered[0U]);
            vlSelfRef.__VicoTriggered[0U] = (8ULL | vlSelfRef.__VicoTriggered[0U]);
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

    VLCU_D static bool Vt13_mix___024root___trigger_anySet__ico(const VlUnpacked<QData/*63:0*/, 2> &in) {
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

    VLCU_D static bool Vt13_mix___024root___eval_phase__ico(Vt13_mix___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VicoExecute;
        // Body
        Vt13_mix___024root___eval_triggers_vec__ico(vlSelf, vlCtx);
#ifdef VL_DEBUG
        if (VL_UNLIKELY(vlSymsp->_vm_contextp__->debug())) {
            Vt13_mix___024root___dump_triggers__ico(vlSelfRef.__VicoTriggered, "ico"s);
        }
#endif
        __VicoExecute = Vt13_mix___024root___trigger_anySet__ico(vlSelfRef.__VicoTriggered);
        if (__VicoExecute) {
            {
                // Inlined CFunc: _eval_ico
                if ((0x0000000000000084ULL & vlSelfRef.__VicoTriggered[0U])) {
                    {
                        // Inlined CFunc: _ico_comb__TOP__0
                        if (vlSelfRef.en) {
                            vlSelfRef.lat = ((IData)(vlSelfRef.d) 
                                             ^ (((IData)(vlSelfRef.cnt_div) 
                                                 << 8U) 
                                                | (IData)(vlSelfRef.cnt_gate)));
                        }
                    }
                }
            }
        }
        return (__VicoExecute);
    }

    VLCU_D static bool Vt13_mix___024root___trigger_anySet__act(const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static void Vt13_mix___024root___nba_sequent__TOP__3(Vt13_mix___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*7:0*/ __Vdly__t13_mix__DOT__hi8;
        __Vdly__t13_mix__DOT__hi8 = 0;
        VlWide<4>/*127:0*/ __Vtemp_2;
        VlWide<4>/*127:0*/ __Vtemp_4;
        VlWide<4>/*127:0*/ __Vtemp_5;
        VlWide<4>/*127:0*/ __Vtemp_8;
        VlWide<4>/*127:0*/ __Vtemp_9;
        VlWide<4>/*127:0*/ __Vtemp_11;
        VlWide<4>/*127:0*/ __Vtemp_12;
        VlWide<4>/*127:0*/ __Vtemp_14;
        // Body
        __Vdly__t13_mix__DOT__hi8 = vlSelfRef.t13_mix__DOT__hi8;
        vlSelfRef.lg = (0x0000007fU & VL_CLOG2_Q(vlSelfRef.qa));
        VL_MULS_WWW(100, __Vtemp_2, vlSelfRef.wa, vlSelfRef.wb);
        vlSelfRef.sprod[0U] = __Vtemp_2[0U];
        vlSelfRef.sprod[1U] = __Vtemp_2[1U];
        vlSelfRef.sprod[2U] = __Vtemp_2[2U];
        vlSelfRef.sprod[3U] = (0x0000000fU & __Vtemp_2[3U]);
        vlSelfRef.wr[((IData)(vlSelfRef.sel) >> 5U)] 
            = (((~ ((IData)(1U) << (0x0000001fU & (IData)(vlSelfRef.sel)))) 
                & vlSelfRef.wr[((IData)(vlSelfRef.sel) 
                                >> 5U)]) | ((1U & (IData)(vlSelfRef.d)) 
                                            << (0x0000001fU 
                                                & (IData)(vlSelfRef.sel))));
        VL_ASSIGNSEL_WI(128, 8, (0x00000078U & ((IData)(vlSelfRef.sel) 
                                                << 3U)), vlSelfRef.wr, 
                        (0x000000ffU & ((IData)(vlSelfRef.d) 
                                        >> 8U)));
        __Vtemp_4[0U] = (0x000000ffU & vlSelfRef.wa[0U]);
        __Vtemp_4[1U] = 0U;
        __Vtemp_4[2U] = 0U;
        __Vtemp_4[3U] = 0U;
        VL_POW_WWI(100,100,4, __Vtemp_5, __Vtemp_4, 
                   (0x0000000fU & (IData)(vlSelfRef.sel)));
        vlSelfRef.pw[0U] = __Vtemp_5[0U];
        vlSelfRef.pw[1U] = __Vtemp_5[1U];
        vlSelfRef.pw[2U] = __Vtemp_5[2U];
        vlSelfRef.pw[3U] = (0x0000000fU & __Vtemp_5[3U]);
        vlSelfRef.qr = (((~ (1ULL << (0x0000003fU & (IData)(vlSelfRef.sel)))) 
                         & vlSelfRef.qr) | ((QData)((IData)(
                                                            (1U 
                                                             & ((IData)(vlSelfRef.d) 
                                                                >> 1U)))) 
                                            << (0x0000003fU 
                                                & (IData)(vlSelfRef.sel))));
        vlSelfRef.qr = (((~ (0x00000000000000ffULL 
                             << (0x00000038U & ((IData)(vlSelfRef.sel) 
                                                << 3U)))) 
                         & vlSelfRef.qr) | ((QData)((IData)(
                                                            (0x000000ffU 
                                                             & ((IData)(vlSelfRef.d) 
                                                                ^ (IData)(vlSelfRef.qa))))) 
                                            << (0x00000038U 
                                                & ((IData)(vlSelfRef.sel) 
                                                   << 3U))));
        __Vtemp_8[0U] = (1U | vlSelfRef.wb[0U]);
        __Vtemp_8[1U] = vlSelfRef.wb[1U];
        __Vtemp_8[2U] = vlSelfRef.wb[2U];
        __Vtemp_8[3U] = vlSelfRef.wb[3U];
        VL_DIVS_WWW(100, __Vtemp_9, vlSelfRef.wa, __Vtemp_8);
        __Vtemp_11[0U] = (3U | vlSelfRef.wb[0U]);
        __Vtemp_11[1U] = vlSelfRef.wb[1U];
        __Vtemp_11[2U] = vlSelfRef.wb[2U];
        __Vtemp_11[3U] = vlSelfRef.wb[3U];
        VL_MODDIVS_WWW(100, __Vtemp_12, vlSelfRef.wa, __Vtemp_11);
        VL_SHIFTRS_WWI(100,100,6, __Vtemp_14, vlSelfRef.wa, 
                       (0x0000003fU & (IData)(vlSelfRef.sel)));
        vlSelfRef.sdiv[0U] = ((__Vtemp_9[0U] ^ __Vtemp_12[0U]) 
                              ^ __Vtemp_14[0U]);
        vlSelfRef.sdiv[1U] = ((__Vtemp_9[1U] ^ __Vtemp_12[1U]) 
                              ^ __Vtemp_14[1U]);
        vlSelfRef.sdiv[2U] = ((__Vtemp_9[2U] ^ __Vtemp_12[2U]) 
                              ^ __Vtemp_14[2U]);
        vlSelfRef.sdiv[3U] = (0x0000000fU & ((__Vtemp_9[3U] 
                                              ^ __Vtemp_12[3U]) 
                                             ^ __Vtemp_14[3U]));
        __Vdly__t13_mix__DOT__hi8 = (0x000000ffU & 
                                     (((IData)(vlSelfRef.d) 
                                       + (IData)(vlSelfRef.t13_mix__DOT__hi8)) 
                                      >> 4U));
        vlSelfRef.cmp = ((0xf0U & (IData)(vlSelfRef.cmp)) 
                         | (0x0000000fU & ((IData)(vlSelfRef.d) 
                                           + (IData)(vlSelfRef.t13_mix__DOT__hi8))));
        vlSelfRef.cmp = ((0x0fU & (IData)(vlSelfRef.cmp)) 
                         | ((((VL_LTS_IWW(100, vlSelfRef.wa, vlSelfRef.wb) 
                               << 3U) | (VL_GTES_IWW(100, vlSelfRef.wa, vlSelfRef.wb) 
                                         << 2U)) | 
                             ((VL_GTS_IQQ(64, vlSelfRef.qa, 
                                          (((QData)((IData)(vlSelfRef.wa[1U])) 
                                            << 0x00000020U) 
                                           | (QData)((IData)(vlSelfRef.wa[0U])))) 
                               << 1U) | (0U == ((((vlSelfRef.wa[0U] 
                                                   ^ vlSelfRef.wb[0U]) 
                                                  | (vlSelfRef.wa[1U] 
                                                     ^ vlSelfRef.wb[1U])) 
                                                 | (vlSelfRef.wa[2U] 
                                                    ^ vlSelfRef.wb[2U])) 
                                                | (vlSelfRef.wa[3U] 
                                                   ^ vlSelfRef.wb[3U]))))) 
                            << 4U));
        vlSelfRef.t13_mix__DOT__hi8 = __Vdly__t13_mix__DOT__hi8;
    }

    VLCU_D static void Vt13_mix___024root___trigger_orInto__act_vec_vec(VlUnpacked<QData/*63:0*/, 1> &out, const VlUnpacked<QData/*63:0*/, 1> &in) {
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

    VLCU_D static bool Vt13_mix___024root___eval_phase__act(Vt13_mix___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        {
            // Inlined CFunc: _eval_triggers_vec__act
            CData/*0:0*/ __Vinline_0__eval_triggers_vec__act___Vtrigprevexpr_h364b5a28__0;
            __Vinline_0__eval_triggers_vec__act___Vtrigprevexpr_h364b5a28__0 = 0;
            __Vinline_0__eval_triggers_vec__act___Vtrigprevexpr_h364b5a28__0 
                = ((IData)(vlSelfRef.clk) & (IData)(vlSelfRef.en));
            vlSelfRef.__VactTriggered[0U] = (QData)((IData)(
                                                            ((((__Vinline_0__eval_triggers_vec__act___Vtrigprevexpr_h364b5a28__0 
                                                                & (~ (IData)(vlSelfRef.__Vtrigprevexpr_h364b5a28__1))) 
                                                               << 3U) 
                                                              | (((IData)(vlSelfRef.t13_mix__DOT__div2) 
                                                                  & (~ (IData)(vlSelfRef.__Vtrigprevexpr___TOP__t13_mix__DOT__div2__0))) 
                                                                 << 2U)) 
                                                             | ((((IData)(vlSelfRef.clk) 
                                                                  & (~ (IData)(vlSelfRef.__Vtrigprevexpr___TOP__clk__1))) 
                                                                 << 1U) 
                                                                | ((IData)(vlSelfRef.arst) 
                                                                   & (~ (IData)(vlSelfRef.__Vtrigprevexpr___TOP__arst__1)))))));
            vlSelfRef.__Vtrigprevexpr___TOP__arst__1 
                = vlSelfRef.arst;
            vlSelfRef.__Vtrigprevexpr___TOP__clk__1 
                = vlSelfRef.clk;
            vlSelfRef.__Vtrigprevexpr___TOP__t13_mix__DOT__div2__0 
                = vlSelfRef.t13_mix__DOT__div2;
            vlSelfRef.__Vtrigprevexpr_h364b5a28__1 
                = __Vinline_0__eval_triggers_vec__act___Vtrigprevexpr_h364b5a28__0;
        }
#ifdef VL_DEBUG
        if (VL_UNLIKELY(vlSymsp->_vm_contextp__->debug())) {
            Vt13_mix___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
        }
#endif
        Vt13_mix___024root___trigger_orInto__act_vec_vec(vlSelfRef.__VnbaTriggered, vlSelfRef.__VactTriggered);
        return (0U);
    }

    VLCU_D static void Vt13_mix___024root___trigger_clear__act(VlUnpacked<QData/*63:0*/, 1> &out) {
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

    VLCU_D static bool Vt13_mix___024root___eval_phase__nba(Vt13_mix___024root* vlSelf, VlcuContext& vlCtx) {
        typedef bool VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Locals
        CData/*0:0*/ __VnbaExecute;
        // Body
        __VnbaExecute = Vt13_mix___024root___trigger_anySet__act(vlSelfRef.__VnbaTriggered);
        if (__VnbaExecute) {
            {
                // Inlined CFunc: _eval_nba
                if ((3ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    {
                        // Inlined CFunc: _nba_sequent__TOP__0
                        vlSelfRef.t13_mix__DOT__div2 
                            = (1U & ((~ (IData)(vlSelfRef.arst)) 
                                     & (~ (IData)(vlSelfRef.t13_mix__DOT__div2))));
                    }
                }
                if ((5ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    {
                        // Inlined CFunc: _nba_sequent__TOP__1
                        vlSelfRef.cnt_div = ((IData)(vlSelfRef.arst)
                                              ? 0U : 
                                             (0x000000ffU 
                                              & ((IData)(1U) 
                                                 + (IData)(vlSelfRef.cnt_div))));
                    }
                }
                if ((8ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    {
                        // Inlined CFunc: _nba_sequent__TOP__2
                        vlSelfRef.cnt_gate = (0x000000ffU 
                                              & ((IData)(vlSelfRef.cnt_gate) 
                                                 + (IData)(vlSelfRef.d)));
                    }
                }
                if ((2ULL & vlSelfRef.__VnbaTriggered[0U])) {
                    Vt13_mix___024root___nba_sequent__TOP__3(vlSelf, vlCtx);
                }
                if ((0x000000000000000dULL & vlSelfRef.__VnbaTriggered[0U])) {
                    {
                        // Inlined CFunc: _ico_comb__TOP__0
                        if (vlSelfRef.en) {
                            vlSelfRef.lat = ((IData)(vlSelfRef.d) 
                                             ^ (((IData)(vlSelfRef.cnt_div) 
                                                 << 8U) 
                                                | (IData)(vlSelfRef.cnt_gate)));
                        }
                    }
                }
            }
            Vt13_mix___024root___trigger_clear__act(vlSelfRef.__VnbaTriggered);
        }
        return (__VnbaExecute);
    }

    VLCU_D static void Vt13_mix___024root___eval(Vt13_mix___024root* vlSelf, VlcuContext& vlCtx) {
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
                Vt13_mix___024root___dump_triggers__ico(vlSelfRef.__VicoTriggered, "ico"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t13_mix.v", 5, "", "DIDNOTCONVERGE: Input combinational region did not converge after '--converge-limit' of 10000 tries");
            }
            __VicoIterCount = ((IData)(1U) + __VicoIterCount);
            vlSelfRef.__VicoPhaseResult = Vt13_mix___024root___eval_phase__ico(vlSelf, vlCtx);
        } while (vlSelfRef.__VicoPhaseResult);
        __VnbaIterCount = 0U;
        do {
            if (VL_UNLIKELY(((0x00002710U < __VnbaIterCount)))) {
#ifdef VL_DEBUG
                Vt13_mix___024root___dump_triggers__act(vlSelfRef.__VnbaTriggered, "nba"s);
#endif
                VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t13_mix.v", 5, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            }
            __VnbaIterCount = ((IData)(1U) + __VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
#ifdef VL_DEBUG
                    Vt13_mix___024root___dump_triggers__act(vlSelfRef.__VactTriggered, "act"s);
#endif
                    VL_FATAL_MT("../verilator_b200/csrc/designs/torture/t13_mix.v", 5, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) 
                                             + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = Vt13_mix___024root___eval_phase__act(vlSelf, vlCtx);
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = Vt13_mix___024root___eval_phase__nba(vlSelf, vlCtx);
        } while (vlSelfRef.__VnbaPhaseResult);
    }

    VLCU_D static void Vt13_mix___024root___ctor_var_reset(Vt13_mix___024root* vlSelf, VlcuContext& vlCtx) {
        typedef void VlcuFuncRet;
        auto& vlSelfRef = *vlSelf;
        const VlcuSymsShim vlSymsShim{&vlCtx};
        const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;
        // Body
        const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
        vlSelf->clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
        vlSelf->arst = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 14097772074827770664ull);
        vlSelf->en = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 7710216835639188562ull);
        vlSelf->sel = VL_SCOPED_RAND_RESET_I(7, __VscopeHash, 14660815779230110548ull);
        VL_SCOPED_RAND_RESET_W(100, vlSelf->wa, __VscopeHash, 15817334002654144152ull);
        VL_SCOPED_RAND_RESET_W(100, vlSelf->wb, __VscopeHash, 17970158761716929289ull);
        vlSelf->qa = VL_SCOPED_RAND_RESET_Q(64, __VscopeHash, 13914198460801392675ull);
        vlSelf->d = VL_SCOPED_RAND_RESET_I(16, __VscopeHash, 1720370409040345145ull);
        VL_SCOPED_RAND_RESET_W(128, vlSelf->wr, __VscopeHash, 10619988472972731121ull);
        vlSelf->qr = VL_SCOPED_RAND_RESET_Q(64, __VscopeHash, 9270906056362523178ull);
        VL_SCOPED_RAND_RESET_W(100, vlSelf->sprod, __VscopeHash, 8658572552303323694ull);
        VL_SCOPED_RAND_RESET_W(100, vlSelf->sdiv, __VscopeHash, 13644446615494747826ull);
        vlSelf->cmp = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 10061869555615269896ull);
        vlSelf->lat = VL_SCOPED_RAND_RESET_I(16, __VscopeHash, 11408774834842292007ull);
        vlSelf->cnt_div = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 4232671579786978952ull);
        vlSelf->cnt_gate = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 9840192267010412661ull);
        VL_SCOPED_RAND_RESET_W(100, vlSelf->pw, __VscopeHash, 5549622911285396126ull);
        vlSelf->lg = VL_SCOPED_RAND_RESET_I(7, __VscopeHash, 722646293792979593ull);
        vlSelf->t13_mix__DOT__div2 = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 3290552772298103798ull);
        vlSelf->t13_mix__DOT__hi8 = VL_SCOPED_RAND_RESET_I(8, __VscopeHash, 16372052012946030958ull);
        for (int __Vi0 = 0; __Vi0 < 1; ++__Vi0) {
            vlSelf->__VstlTriggered[__Vi0] = 0;
        }
        for (int __Vi0 = 0; __Vi0 < 2; ++__Vi0) {
            vlSelf->__VicoTriggered[__Vi0] = 0;
        }
        vlSelf->__Vtrigprevexpr___TOP__clk__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__arst__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__en__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__sel__0 = 0;
        VL_ZERO_RESET_W(100, vlSelf->__Vtrigprevexpr___TOP__wa__0);
        VL_ZERO_RESET_W(100, vlSelf->__Vtrigprevexpr___TOP__wb__0);
        vlSelf->__Vtrigprevexpr___TOP__qa__0 = 0;
        vlSelf->__Vtrigprevexpr___TOP__d__0 = 0;
        vlSelf->__VicoDidInit = 0;
        for (int __Vi0 = 0; __Vi0 < 1; ++__Vi0) {
            vlSelf->__VactTriggered[__Vi0] = 0;
        }
        vlSelf->__Vtrigprevexpr___TOP__arst__1 = 0;
        vlSelf->__Vtrigprevexpr___TOP__clk__1 = 0;
        vlSelf->__Vtrigprevexpr___TOP__t13_mix__DOT__div2__0 = 0;
        vlSelf->__Vtrigprevexpr_h364b5a28__1 = 0;
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
        Vt13_mix___024root___ctor_var_reset(vlSelf, vlCtx);
    }
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext& vlCtx) { Vt13_mix___024root___eval_static(vlSelf, vlCtx); }
    VLCU_D static void eval_initial(Root* vlSelf, VlcuContext& vlCtx) { Vt13_mix___024root___eval_initial(vlSelf, vlCtx); }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) { Vt13_mix___024root___eval_settle(vlSelf, vlCtx); }
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) { Vt13_mix___024root___eval(vlSelf, vlCtx); }
    VLCU_D static void eval_final(Root* vlSelf, VlcuContext& vlCtx) { Vt13_mix___024root___eval_final(vlSelf, vlCtx); }

    static const VlcuDesignInfo& info();
};

//======================================================================
// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime

const VlcuDesignInfo& Vt13_mix__cuda::info() {
    typedef Vt13_mix__cuda T;
    // name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset
    static const VlcuMemberDesc members[] = {
        {"clk", "clk", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_clk},
        {"arst", "arst", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_arst},
        {"en", "en", VLCU_MEMBER_IN, 1, 0, 1, 1, 0, 0, T::OFF_en},
        {"sel", "sel", VLCU_MEMBER_IN, 1, 0, 1, 7, 6, 0, T::OFF_sel},
        {"wa", "wa", VLCU_MEMBER_IN, 4, 0, 4, 100, 99, 0, T::OFF_wa},
        {"wb", "wb", VLCU_MEMBER_IN, 4, 0, 4, 100, 99, 0, T::OFF_wb},
        {"qa", "qa", VLCU_MEMBER_IN, 8, 0, 1, 64, 63, 0, T::OFF_qa},
        {"d", "d", VLCU_MEMBER_IN, 2, 0, 1, 16, 15, 0, T::OFF_d},
        {"wr", "wr", VLCU_MEMBER_OUT, 4, 0, 4, 128, 127, 0, T::OFF_wr},
        {"qr", "qr", VLCU_MEMBER_OUT, 8, 0, 1, 64, 63, 0, T::OFF_qr},
        {"sprod", "sprod", VLCU_MEMBER_OUT, 4, 0, 4, 100, 99, 0, T::OFF_sprod},
        {"sdiv", "sdiv", VLCU_MEMBER_OUT, 4, 0, 4, 100, 99, 0, T::OFF_sdiv},
        {"cmp", "cmp", VLCU_MEMBER_OUT, 1, 0, 1, 8, 7, 0, T::OFF_cmp},
        {"lat", "lat", VLCU_MEMBER_OUT, 2, 0, 1, 16, 15, 0, T::OFF_lat},
        {"cnt_div", "cnt_div", VLCU_MEMBER_OUT, 1, 0, 1, 8, 7, 0, T::OFF_cnt_div},
        {"cnt_gate", "cnt_gate", VLCU_MEMBER_OUT, 1, 0, 1, 8, 7, 0, T::OFF_cnt_gate},
        {"pw", "pw", VLCU_MEMBER_OUT, 4, 0, 4, 100, 99, 0, T::OFF_pw},
        {"lg", "lg", VLCU_MEMBER_OUT, 1, 0, 1, 7, 6, 0, T::OFF_lg},
        {"t13_mix__DOT__div2", "t13_mix.div2", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF_t13_mix__DOT__div2},
        {"t13_mix__DOT__hi8", "t13_mix.hi8", VLCU_MEMBER_STATE, 1, 0, 1, 8, 7, 0, T::OFF_t13_mix__DOT__hi8},
        {"__VstlFirstIteration", "__VstlFirstIteration", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlFirstIteration},
        {"__VstlPhaseResult", "__VstlPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VstlPhaseResult},
        {"__Vtrigprevexpr___TOP__clk__0", "__Vtrigprevexpr___TOP__clk__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__0},
        {"__Vtrigprevexpr___TOP__arst__0", "__Vtrigprevexpr___TOP__arst__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__arst__0},
        {"__Vtrigprevexpr___TOP__en__0", "__Vtrigprevexpr___TOP__en__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__en__0},
        {"__Vtrigprevexpr___TOP__sel__0", "__Vtrigprevexpr___TOP__sel__0", VLCU_MEMBER_STATE, 1, 0, 1, 7, 6, 0, T::OFF___Vtrigprevexpr___TOP__sel__0},
        {"__VicoDidInit", "__VicoDidInit", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VicoDidInit},
        {"__VicoPhaseResult", "__VicoPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VicoPhaseResult},
        {"__Vtrigprevexpr___TOP__arst__1", "__Vtrigprevexpr___TOP__arst__1", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__arst__1},
        {"__Vtrigprevexpr___TOP__clk__1", "__Vtrigprevexpr___TOP__clk__1", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__clk__1},
        {"__Vtrigprevexpr___TOP__t13_mix__DOT__div2__0", "__Vtrigprevexpr___TOP__t13_mix.div2__0", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr___TOP__t13_mix__DOT__div2__0},
        {"__Vtrigprevexpr_h364b5a28__1", "__Vtrigprevexpr_h364b5a28__1", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___Vtrigprevexpr_h364b5a28__1},
        {"__VactPhaseResult", "__VactPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VactPhaseResult},
        {"__VnbaPhaseResult", "__VnbaPhaseResult", VLCU_MEMBER_STATE, 1, 0, 1, 1, 0, 0, T::OFF___VnbaPhaseResult},
        {"__Vtrigprevexpr___TOP__d__0", "__Vtrigprevexpr___TOP__d__0", VLCU_MEMBER_STATE, 2, 0, 1, 16, 15, 0, T::OFF___Vtrigprevexpr___TOP__d__0},
        {"__Vtrigprevexpr___TOP__wa__0", "__Vtrigprevexpr___TOP__wa__0", VLCU_MEMBER_STATE, 4, 0, 4, 100, 99, 0, T::OFF___Vtrigprevexpr___TOP__wa__0},
        {"__Vtrigprevexpr___TOP__wb__0", "__Vtrigprevexpr___TOP__wb__0", VLCU_MEMBER_STATE, 4, 0, 4, 100, 99, 0, T::OFF___Vtrigprevexpr___TOP__wb__0},
        {"__VactIterCount", "__VactIterCount", VLCU_MEMBER_STATE, 4, 0, 1, 32, 31, 0, T::OFF___VactIterCount},
        {"__Vtrigprevexpr___TOP__qa__0", "__Vtrigprevexpr___TOP__qa__0", VLCU_MEMBER_STATE, 8, 0, 1, 64, 63, 0, T::OFF___Vtrigprevexpr___TOP__qa__0},
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
    static const uint32_t saveOrder[] = {0, 1, 2, 3, 12, 14, 15, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 7, 13, 34, 4, 5, 8, 10, 11, 16, 35, 36, 37, 6, 9, 38, 39, 40, 41, 42, 0};  // (+ terminator)
    static const VlcuDesignInfo s = {
        VLCU_DESIGN_NAME, "t13_mix", members, sizeof(members) / sizeof(members[0]),
        T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, T::kTouchedWriteBytes,
        nullptr, 0, 0, 0,
        messages, 0, plusargs, 0,
        nullptr, 0, 0, nullptr, 0,
        saveOrder, 43, 0, 0xd18a6f46a954cc30ULL, nullptr, 12, 12, 0, nullptr, 0};
    return s;
}

// The device testbenches of this design (the user's C++ main loop, on the GPU)
#ifdef VLCU_TESTBENCH_INCLUDE
#include VLCU_TESTBENCH_INCLUDE
#endif
#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB
#define VLCU_TESTBENCHES
#endif
VLCU_REGISTER_DESIGN_X(Vt13_mix__cuda VLCU_TESTBENCHES);
