// -*- mode: C++; c-file-style: "cc-mode" -*-
// DESCRIPTION: batched CUDA model of examples/make_tracing_c (top.v + sub.v),
// written BY HAND in the form V3EmitCuda prints (see src/V3EmitCuda.cpp), because
// verilator_bin cannot be built in this image (needs bison/flex).  It follows
// the generator rules of the reference for
//   -cc -x-assign fast --assert [--coverage] examples/make_tracing_c/top.v :
//   state members            src/V3EmitCHeaders.cpp:73-143, src/V3AstNodes.cpp:1175-1300
//   eval loop nest           src/V3Sched.cpp:124-252 (loops), :498-654 (ico), :671-884 (_eval)
//   trigger computation      src/V3SchedTrigger.cpp:560-583,696-804; src/V3SenExprBuilder.h:234-314
//   trigger-gated leaves     src/V3OrderCFuncEmitter.h:86-184, src/V3SchedUtil.cpp:246-263
//   NBA shadow variable      src/V3Delayed.cpp:48-57
//   toggle/line coverage     src/V3Coverage.cpp:395-460, src/V3Clock.cpp:69-98,
//                            src/V3EmitCFunc.h:928-967
//   reset values             src/V3EmitCFunc.cpp:600-677
// Included twice by Vtop__cuda.cu: with and without --coverage (VTOP_COVERAGE).
//
// Expected macros: VTOP (struct name), VTOP_ROOT (root struct name), VTOP_COVERAGE (0/1),
// VTOP_DESIGN_NAME (string)

struct VTOP_ROOT final {
    // DESIGN SPECIFIC STATE (ordered by storage size: QData, IData/EData, CData)
    QData /*39:0*/ in_quad;
    QData /*39:0*/ out_quad;
    QData /*39:0*/ __Vtrigprevexpr___TOP__in_quad__0;
    VlWide<3> /*69:0*/ in_wide;
    VlWide<3> /*69:0*/ out_wide;
    VlWide<3> /*69:0*/ __Vtrigprevexpr___TOP__in_wide__0;
    IData /*31:0*/ top__DOT__sub__DOT__count_c;
    CData /*0:0*/ clk;
    CData /*0:0*/ reset_l;
    CData /*1:0*/ in_small;
    CData /*1:0*/ out_small;
    CData /*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData /*0:0*/ __Vtrigprevexpr___TOP__reset_l__0;
    CData /*1:0*/ __Vtrigprevexpr___TOP__in_small__0;
    CData /*0:0*/ __Vtrigprevexpr___TOP__clk__1;
    CData /*0:0*/ __VicoDidInit;
#if VTOP_COVERAGE
    QData /*39:0*/ top__DOT____Vtogcov__out_quad;
    QData /*39:0*/ top__DOT____Vtogcov__in_quad;
    VlWide<3> /*69:0*/ top__DOT____Vtogcov__out_wide;
    VlWide<3> /*69:0*/ top__DOT____Vtogcov__in_wide;
    IData /*31:0*/ top__DOT__sub__DOT____Vtogcov__count_c;
    CData /*0:0*/ top__DOT____Vtogcov__clk;
    CData /*0:0*/ top__DOT____Vtogcov__reset_l;
    CData /*1:0*/ top__DOT____Vtogcov__out_small;
    CData /*1:0*/ top__DOT____Vtogcov__in_small;
    CData /*0:0*/ top__DOT__sub__DOT____Vtogcov__clk;
    CData /*0:0*/ top__DOT__sub__DOT____Vtogcov__reset_l;
    VlcuLocalArray<uint32_t, 527> __Vcoverage;  // run-time indexed: storage in Locals
#endif
    // Scheduling temporaries (noReset; never stored to HBM)
    VlUnpacked<QData /*63:0*/, 1> __VstlTriggered;
    VlUnpacked<QData /*63:0*/, 2> __VicoTriggered;
    VlUnpacked<QData /*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData /*63:0*/, 1> __VnbaTriggered;
    IData /*31:0*/ __VstlIterCount;
    IData /*31:0*/ __VicoIterCount;
    IData /*31:0*/ __VactIterCount;
    IData /*31:0*/ __VnbaIterCount;
    CData /*0:0*/ __VstlFirstIteration;
    CData /*0:0*/ __VstlPhaseResult;
    CData /*0:0*/ __VstlExecute;
    CData /*0:0*/ __VicoFirstIteration;
    CData /*0:0*/ __VicoPhaseResult;
    CData /*0:0*/ __VicoExecute;
    CData /*0:0*/ __VactFirstIteration;
    CData /*0:0*/ __VactPhaseResult;
    CData /*0:0*/ __VactExecute;
    CData /*0:0*/ __VnbaFirstIteration;
    CData /*0:0*/ __VnbaPhaseResult;
    CData /*0:0*/ __VnbaExecute;
};

struct VTOP final {
    typedef VTOP_ROOT Root;

    //=====================================================================
    // SoA layout: byte offset of each persistent member per instance
    enum : uint32_t {
        OFF_in_quad = 0,
        OFF_out_quad = 8,
        OFF___Vtrigprevexpr___TOP__in_quad__0 = 16,
        OFF_in_wide = 24,
        OFF_out_wide = 36,
        OFF___Vtrigprevexpr___TOP__in_wide__0 = 48,
        OFF_top__DOT__sub__DOT__count_c = 60,
#if VTOP_COVERAGE
        OFF_top__DOT____Vtogcov__out_quad = 64,
        OFF_top__DOT____Vtogcov__in_quad = 72,
        OFF_top__DOT____Vtogcov__out_wide = 80,
        OFF_top__DOT____Vtogcov__in_wide = 92,
        OFF_top__DOT__sub__DOT____Vtogcov__count_c = 104,
        OFF___Vcoverage = 108,
        OFF_clk = 108 + 527 * 4,
#else
        OFF_clk = 64,
#endif
        OFF_reset_l = OFF_clk + 1,
        OFF_in_small = OFF_clk + 2,
        OFF_out_small = OFF_clk + 3,
        OFF___Vtrigprevexpr___TOP__clk__0 = OFF_clk + 4,
        OFF___Vtrigprevexpr___TOP__reset_l__0 = OFF_clk + 5,
        OFF___Vtrigprevexpr___TOP__in_small__0 = OFF_clk + 6,
        OFF___Vtrigprevexpr___TOP__clk__1 = OFF_clk + 7,
        OFF___VicoDidInit = OFF_clk + 8,
#if VTOP_COVERAGE
        OFF_top__DOT____Vtogcov__clk = OFF_clk + 9,
        OFF_top__DOT____Vtogcov__reset_l = OFF_clk + 10,
        OFF_top__DOT____Vtogcov__out_small = OFF_clk + 11,
        OFF_top__DOT____Vtogcov__in_small = OFF_clk + 12,
        OFF_top__DOT__sub__DOT____Vtogcov__clk = OFF_clk + 13,
        OFF_top__DOT__sub__DOT____Vtogcov__reset_l = OFF_clk + 14,
        kStateBytesRaw = OFF_clk + 15,
#else
        kStateBytesRaw = OFF_clk + 9,
#endif
        kStateBytes = (kStateBytesRaw + 7u) & ~7u,  // S_persist (8-byte multiple: records stay aligned)
        kInWords = 8,  // clk, reset_l, in_small, in_quad(2), in_wide(3)
        kOutWords = 6,  // out_small, out_quad(2), out_wide(3)
        kInBytes = 23,  // S_in  (C storage)
        kOutBytes = 21,  // S_out
        kCoverBins = VTOP_COVERAGE ? 527 : 0,
    };

    struct Locals final {
#if VTOP_COVERAGE
        uint32_t __Vcoverage[527];
#endif
    };
    struct Shared final {};
    VLCU_D static void bind(Root& s, Locals& l, Shared&, const VlcuState&, uint64_t) {
#if VTOP_COVERAGE
        s.__Vcoverage.bind(l.__Vcoverage);
#else
        (void)s, (void)l;
#endif
    }

    VLCU_D static void load(Root& s, const VlcuState& st, uint64_t i) {
        s.in_quad = st.ld<QData>(OFF_in_quad, i);
        s.out_quad = st.ld<QData>(OFF_out_quad, i);
        s.__Vtrigprevexpr___TOP__in_quad__0 = st.ld<QData>(OFF___Vtrigprevexpr___TOP__in_quad__0, i);
        st.ldW(OFF_in_wide, i, s.in_wide);
        st.ldW(OFF_out_wide, i, s.out_wide);
        st.ldW(OFF___Vtrigprevexpr___TOP__in_wide__0, i, s.__Vtrigprevexpr___TOP__in_wide__0);
        s.top__DOT__sub__DOT__count_c = st.ld<IData>(OFF_top__DOT__sub__DOT__count_c, i);
        s.clk = st.ld<CData>(OFF_clk, i);
        s.reset_l = st.ld<CData>(OFF_reset_l, i);
        s.in_small = st.ld<CData>(OFF_in_small, i);
        s.out_small = st.ld<CData>(OFF_out_small, i);
        s.__Vtrigprevexpr___TOP__clk__0 = st.ld<CData>(OFF___Vtrigprevexpr___TOP__clk__0, i);
        s.__Vtrigprevexpr___TOP__reset_l__0 = st.ld<CData>(OFF___Vtrigprevexpr___TOP__reset_l__0, i);
        s.__Vtrigprevexpr___TOP__in_small__0 = st.ld<CData>(OFF___Vtrigprevexpr___TOP__in_small__0, i);
        s.__Vtrigprevexpr___TOP__clk__1 = st.ld<CData>(OFF___Vtrigprevexpr___TOP__clk__1, i);
        s.__VicoDidInit = st.ld<CData>(OFF___VicoDidInit, i);
#if VTOP_COVERAGE
        s.top__DOT____Vtogcov__out_quad = st.ld<QData>(OFF_top__DOT____Vtogcov__out_quad, i);
        s.top__DOT____Vtogcov__in_quad = st.ld<QData>(OFF_top__DOT____Vtogcov__in_quad, i);
        st.ldW(OFF_top__DOT____Vtogcov__out_wide, i, s.top__DOT____Vtogcov__out_wide);
        st.ldW(OFF_top__DOT____Vtogcov__in_wide, i, s.top__DOT____Vtogcov__in_wide);
        s.top__DOT__sub__DOT____Vtogcov__count_c
            = st.ld<IData>(OFF_top__DOT__sub__DOT____Vtogcov__count_c, i);
        s.top__DOT____Vtogcov__clk = st.ld<CData>(OFF_top__DOT____Vtogcov__clk, i);
        s.top__DOT____Vtogcov__reset_l = st.ld<CData>(OFF_top__DOT____Vtogcov__reset_l, i);
        s.top__DOT____Vtogcov__out_small = st.ld<CData>(OFF_top__DOT____Vtogcov__out_small, i);
        s.top__DOT____Vtogcov__in_small = st.ld<CData>(OFF_top__DOT____Vtogcov__in_small, i);
        s.top__DOT__sub__DOT____Vtogcov__clk = st.ld<CData>(OFF_top__DOT__sub__DOT____Vtogcov__clk, i);
        s.top__DOT__sub__DOT____Vtogcov__reset_l
            = st.ld<CData>(OFF_top__DOT__sub__DOT____Vtogcov__reset_l, i);
        for (uint32_t b = 0; b < 527; ++b) s.__Vcoverage[b] = st.ldE<uint32_t>(OFF___Vcoverage, b, i);
#endif
    }

    VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {
        st.st<QData>(OFF_in_quad, i, s.in_quad);
        st.st<QData>(OFF_out_quad, i, s.out_quad);
        st.st<QData>(OFF___Vtrigprevexpr___TOP__in_quad__0, i, s.__Vtrigprevexpr___TOP__in_quad__0);
        st.stW(OFF_in_wide, i, s.in_wide);
        st.stW(OFF_out_wide, i, s.out_wide);
        st.stW(OFF___Vtrigprevexpr___TOP__in_wide__0, i, s.__Vtrigprevexpr___TOP__in_wide__0);
        st.st<IData>(OFF_top__DOT__sub__DOT__count_c, i, s.top__DOT__sub__DOT__count_c);
        st.st<CData>(OFF_clk, i, s.clk);
        st.st<CData>(OFF_reset_l, i, s.reset_l);
        st.st<CData>(OFF_in_small, i, s.in_small);
        st.st<CData>(OFF_out_small, i, s.out_small);
        st.st<CData>(OFF___Vtrigprevexpr___TOP__clk__0, i, s.__Vtrigprevexpr___TOP__clk__0);
        st.st<CData>(OFF___Vtrigprevexpr___TOP__reset_l__0, i, s.__Vtrigprevexpr___TOP__reset_l__0);
        st.st<CData>(OFF___Vtrigprevexpr___TOP__in_small__0, i, s.__Vtrigprevexpr___TOP__in_small__0);
        st.st<CData>(OFF___Vtrigprevexpr___TOP__clk__1, i, s.__Vtrigprevexpr___TOP__clk__1);
        st.st<CData>(OFF___VicoDidInit, i, s.__VicoDidInit);
#if VTOP_COVERAGE
        st.st<QData>(OFF_top__DOT____Vtogcov__out_quad, i, s.top__DOT____Vtogcov__out_quad);
        st.st<QData>(OFF_top__DOT____Vtogcov__in_quad, i, s.top__DOT____Vtogcov__in_quad);
        st.stW(OFF_top__DOT____Vtogcov__out_wide, i, s.top__DOT____Vtogcov__out_wide);
        st.stW(OFF_top__DOT____Vtogcov__in_wide, i, s.top__DOT____Vtogcov__in_wide);
        st.st<IData>(OFF_top__DOT__sub__DOT____Vtogcov__count_c, i,
                     s.top__DOT__sub__DOT____Vtogcov__count_c);
        st.st<CData>(OFF_top__DOT____Vtogcov__clk, i, s.top__DOT____Vtogcov__clk);
        st.st<CData>(OFF_top__DOT____Vtogcov__reset_l, i, s.top__DOT____Vtogcov__reset_l);
        st.st<CData>(OFF_top__DOT____Vtogcov__out_small, i, s.top__DOT____Vtogcov__out_small);
        st.st<CData>(OFF_top__DOT____Vtogcov__in_small, i, s.top__DOT____Vtogcov__in_small);
        st.st<CData>(OFF_top__DOT__sub__DOT____Vtogcov__clk, i, s.top__DOT__sub__DOT____Vtogcov__clk);
        st.st<CData>(OFF_top__DOT__sub__DOT____Vtogcov__reset_l, i,
                     s.top__DOT__sub__DOT____Vtogcov__reset_l);
        for (uint32_t b = 0; b < 527; ++b) st.stE<uint32_t>(OFF___Vcoverage, b, i, s.__Vcoverage[b]);
#endif
    }

    // Stimulus row: one 32-bit word per <=32-bit input, two per QData, N per VlWide,
    // in port declaration order; values are masked to the port width on ingest
    VLCU_D static void load_inputs_row(Root& s, const uint32_t* rowp, uint64_t stride, uint64_t i) {
        s.clk = static_cast<CData>(rowp[0 * stride + i] & 1U);
        s.reset_l = static_cast<CData>(rowp[1 * stride + i] & 1U);
        s.in_small = static_cast<CData>(rowp[2 * stride + i] & 3U);
        s.in_quad = (static_cast<QData>(rowp[4 * stride + i] & 0xffU) << 32) | rowp[3 * stride + i];
        s.in_wide[0] = rowp[5 * stride + i];
        s.in_wide[1] = rowp[6 * stride + i];
        s.in_wide[2] = rowp[7 * stride + i] & 0x3fU;
    }
    VLCU_D static void store_outputs_row(const Root& s, uint32_t* rowp, uint64_t stride, uint64_t i) {
        rowp[0 * stride + i] = s.out_small;
        rowp[1 * stride + i] = static_cast<uint32_t>(s.out_quad);
        rowp[2 * stride + i] = static_cast<uint32_t>(s.out_quad >> 32);
        rowp[3 * stride + i] = s.out_wide[0];
        rowp[4 * stride + i] = s.out_wide[1];
        rowp[5 * stride + i] = s.out_wide[2];
    }
    VLCU_D static uint64_t outputs_sig(const Root& s, uint64_t sig) {
        sig = vlcu_sig_mix(sig, s.out_small);
        sig = vlcu_sig_mix(sig, static_cast<uint32_t>(s.out_quad));
        sig = vlcu_sig_mix(sig, static_cast<uint32_t>(s.out_quad >> 32));
        sig = vlcu_sig_mix(sig, s.out_wide[0]);
        sig = vlcu_sig_mix(sig, s.out_wide[1]);
        sig = vlcu_sig_mix(sig, s.out_wide[2]);
        return sig;
    }

    //=====================================================================
    // Vtop___024root___ctor_var_reset: --x-initial unique, so every variable takes
    // VL_SCOPED_RAND_RESET_*(width, hash(instance scope name), hash(var prettyName))
    // (zero unless the context's randReset is set).  Scope name of the root is "TOP".
    VLCU_D static void ctor_var_reset(Root* vlSelf, VlcuContext& vlCtx) {
        auto& vlSelfRef = *vlSelf;
        const uint64_t __VscopeHash = 1586713752769493771ULL;  // VL_MURMUR64_HASH(vlSelf->vlNamep), vlNamep == "TOP"
        vlSelfRef.clk = VL_SCOPED_RAND_RESET_I(vlCtx, 1, __VscopeHash, 16707436170211756652ull);
        vlSelfRef.reset_l = VL_SCOPED_RAND_RESET_I(vlCtx, 1, __VscopeHash, 12926602469485068081ull);
        vlSelfRef.out_small = VL_SCOPED_RAND_RESET_I(vlCtx, 2, __VscopeHash, 14267498806251343750ull);
        vlSelfRef.out_quad = VL_SCOPED_RAND_RESET_Q(vlCtx, 40, __VscopeHash, 3771664055402206494ull);
        VL_SCOPED_RAND_RESET_W(vlCtx, 70, vlSelfRef.out_wide, __VscopeHash, 6124341913962757201ull);
        vlSelfRef.in_small = VL_SCOPED_RAND_RESET_I(vlCtx, 2, __VscopeHash, 4925076677430572012ull);
        vlSelfRef.in_quad = VL_SCOPED_RAND_RESET_Q(vlCtx, 40, __VscopeHash, 7333266129396445036ull);
        VL_SCOPED_RAND_RESET_W(vlCtx, 70, vlSelfRef.in_wide, __VscopeHash, 18101672252301227225ull);
        vlSelfRef.top__DOT__sub__DOT__count_c
            = VL_SCOPED_RAND_RESET_I(vlCtx, 32, __VscopeHash, 2802051788925154131ull);
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = 0;
        vlSelfRef.__Vtrigprevexpr___TOP__reset_l__0 = 0;
        vlSelfRef.__Vtrigprevexpr___TOP__in_small__0 = 0;
        vlSelfRef.__Vtrigprevexpr___TOP__in_quad__0 = 0;
        VL_ZERO_W(70, vlSelfRef.__Vtrigprevexpr___TOP__in_wide__0);
        vlSelfRef.__Vtrigprevexpr___TOP__clk__1 = 0;
        vlSelfRef.__VicoDidInit = 0;
#if VTOP_COVERAGE
        vlSelfRef.top__DOT____Vtogcov__clk = 0;
        vlSelfRef.top__DOT____Vtogcov__reset_l = 0;
        vlSelfRef.top__DOT____Vtogcov__out_small = 0;
        vlSelfRef.top__DOT____Vtogcov__out_quad = 0;
        VL_ZERO_W(70, vlSelfRef.top__DOT____Vtogcov__out_wide);
        vlSelfRef.top__DOT____Vtogcov__in_small = 0;
        vlSelfRef.top__DOT____Vtogcov__in_quad = 0;
        VL_ZERO_W(70, vlSelfRef.top__DOT____Vtogcov__in_wide);
        vlSelfRef.top__DOT__sub__DOT____Vtogcov__clk = 0;
        vlSelfRef.top__DOT__sub__DOT____Vtogcov__reset_l = 0;
        vlSelfRef.top__DOT__sub__DOT____Vtogcov__count_c = 0;
        for (int __Vi0 = 0; __Vi0 < 527; ++__Vi0) vlSelfRef.__Vcoverage[__Vi0] = 0;
#endif
    }

    //=====================================================================
    // Combinational logic of top.v:25-27 (shared by the 'stl' and 'ico' copies)
    VLCU_D static void comb_out_small(Root& vlSelfRef) {
        vlSelfRef.out_small = ((1U & (~(IData)(vlSelfRef.reset_l)))
                                   ? 0U
                                   : (3U & ((IData)(1U) + (IData)(vlSelfRef.in_small))));
    }
    VLCU_D static void comb_out_quad(Root& vlSelfRef) {
        vlSelfRef.out_quad = ((1U & (~(IData)(vlSelfRef.reset_l)))
                                  ? 0ULL
                                  : (0x000000ffffffffffULL & (1ULL + vlSelfRef.in_quad)));
    }
    VLCU_D static void comb_out_wide(Root& vlSelfRef) {
        VlWide<3> /*95:0*/ __Vtemp_1;
        VlWide<3> /*95:0*/ __Vtemp_2;
        if ((1U & (~(IData)(vlSelfRef.reset_l)))) {
            vlSelfRef.out_wide[0U] = 0U;
            vlSelfRef.out_wide[1U] = 0U;
            vlSelfRef.out_wide[2U] = 0U;
        } else {
            VL_CONST_W_1X(70, __Vtemp_1, 0x00000001);
            VL_ADD_W(3, __Vtemp_2, vlSelfRef.in_wide, __Vtemp_1);
            vlSelfRef.out_wide[0U] = __Vtemp_2[0U];
            vlSelfRef.out_wide[1U] = __Vtemp_2[1U];
            vlSelfRef.out_wide[2U] = (0x0000003fU & __Vtemp_2[2U]);
        }
    }
#if VTOP_COVERAGE
    // Toggle coverage: IF(ORIG ^ CHANGE) { INC; CHANGE = ORIG; }   (V3Clock.cpp:69-98)
    VLCU_D static void togcov_clk(Root& vlSelfRef) {
        if ((IData)(vlSelfRef.clk) ^ (IData)(vlSelfRef.top__DOT____Vtogcov__clk)) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelfRef.__Vcoverage + 0, vlSelfRef.clk,
                                   vlSelfRef.top__DOT____Vtogcov__clk);
            vlSelfRef.top__DOT____Vtogcov__clk = vlSelfRef.clk;
        }
        if ((IData)(vlSelfRef.clk) ^ (IData)(vlSelfRef.top__DOT__sub__DOT____Vtogcov__clk)) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelfRef.__Vcoverage + 452, vlSelfRef.clk,
                                   vlSelfRef.top__DOT__sub__DOT____Vtogcov__clk);
            vlSelfRef.top__DOT__sub__DOT____Vtogcov__clk = vlSelfRef.clk;
        }
    }
    VLCU_D static void togcov_reset_l(Root& vlSelfRef) {
        if ((IData)(vlSelfRef.reset_l) ^ (IData)(vlSelfRef.top__DOT____Vtogcov__reset_l)) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelfRef.__Vcoverage + 2, vlSelfRef.reset_l,
                                   vlSelfRef.top__DOT____Vtogcov__reset_l);
            vlSelfRef.top__DOT____Vtogcov__reset_l = vlSelfRef.reset_l;
        }
        if ((IData)(vlSelfRef.reset_l) ^ (IData)(vlSelfRef.top__DOT__sub__DOT____Vtogcov__reset_l)) {
            VL_COV_TOGGLE_CHG_ST_I(1, vlSelfRef.__Vcoverage + 454, vlSelfRef.reset_l,
                                   vlSelfRef.top__DOT__sub__DOT____Vtogcov__reset_l);
            vlSelfRef.top__DOT__sub__DOT____Vtogcov__reset_l = vlSelfRef.reset_l;
        }
    }
    VLCU_D static void togcov_in_small(Root& vlSelfRef) {
        if ((IData)(vlSelfRef.in_small) ^ (IData)(vlSelfRef.top__DOT____Vtogcov__in_small)) {
            VL_COV_TOGGLE_CHG_ST_I(2, vlSelfRef.__Vcoverage + 228, vlSelfRef.in_small,
                                   vlSelfRef.top__DOT____Vtogcov__in_small);
            vlSelfRef.top__DOT____Vtogcov__in_small = vlSelfRef.in_small;
        }
    }
    VLCU_D static void togcov_in_quad(Root& vlSelfRef) {
        if (0ULL != (vlSelfRef.in_quad ^ vlSelfRef.top__DOT____Vtogcov__in_quad)) {
            VL_COV_TOGGLE_CHG_ST_Q(40, vlSelfRef.__Vcoverage + 232, vlSelfRef.in_quad,
                                   vlSelfRef.top__DOT____Vtogcov__in_quad);
            vlSelfRef.top__DOT____Vtogcov__in_quad = vlSelfRef.in_quad;
        }
    }
    VLCU_D static void togcov_in_wide(Root& vlSelfRef) {
        if (VL_CHANGEXOR_W(3, vlSelfRef.in_wide, vlSelfRef.top__DOT____Vtogcov__in_wide)) {
            VL_COV_TOGGLE_CHG_ST_W(70, vlSelfRef.__Vcoverage + 312, vlSelfRef.in_wide,
                                   vlSelfRef.top__DOT____Vtogcov__in_wide);
            VL_ASSIGN_W(70, vlSelfRef.top__DOT____Vtogcov__in_wide, vlSelfRef.in_wide);
        }
    }
    VLCU_D static void togcov_out_small(Root& vlSelfRef) {
        if ((IData)(vlSelfRef.out_small) ^ (IData)(vlSelfRef.top__DOT____Vtogcov__out_small)) {
            VL_COV_TOGGLE_CHG_ST_I(2, vlSelfRef.__Vcoverage + 4, vlSelfRef.out_small,
                                   vlSelfRef.top__DOT____Vtogcov__out_small);
            vlSelfRef.top__DOT____Vtogcov__out_small = vlSelfRef.out_small;
        }
    }
    VLCU_D static void togcov_out_quad(Root& vlSelfRef) {
        if (0ULL != (vlSelfRef.out_quad ^ vlSelfRef.top__DOT____Vtogcov__out_quad)) {
            VL_COV_TOGGLE_CHG_ST_Q(40, vlSelfRef.__Vcoverage + 8, vlSelfRef.out_quad,
                                   vlSelfRef.top__DOT____Vtogcov__out_quad);
            vlSelfRef.top__DOT____Vtogcov__out_quad = vlSelfRef.out_quad;
        }
    }
    VLCU_D static void togcov_out_wide(Root& vlSelfRef) {
        if (VL_CHANGEXOR_W(3, vlSelfRef.out_wide, vlSelfRef.top__DOT____Vtogcov__out_wide)) {
            VL_COV_TOGGLE_CHG_ST_W(70, vlSelfRef.__Vcoverage + 88, vlSelfRef.out_wide,
                                   vlSelfRef.top__DOT____Vtogcov__out_wide);
            VL_ASSIGN_W(70, vlSelfRef.top__DOT____Vtogcov__out_wide, vlSelfRef.out_wide);
        }
    }
    VLCU_D static void togcov_count_c(Root& vlSelfRef) {
        if (vlSelfRef.top__DOT__sub__DOT__count_c ^ vlSelfRef.top__DOT__sub__DOT____Vtogcov__count_c) {
            VL_COV_TOGGLE_CHG_ST_I(32, vlSelfRef.__Vcoverage + 456,
                                   vlSelfRef.top__DOT__sub__DOT__count_c,
                                   vlSelfRef.top__DOT__sub__DOT____Vtogcov__count_c);
            vlSelfRef.top__DOT__sub__DOT____Vtogcov__count_c = vlSelfRef.top__DOT__sub__DOT__count_c;
        }
    }
#endif

    //=====================================================================
    // _eval_static: trigger 'prev' initialisers (V3SenExprBuilder.h:193-201 -> V3Sched.cpp:958)
    VLCU_D static void eval_static(Root* vlSelf, VlcuContext&) {
        auto& vlSelfRef = *vlSelf;
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
        vlSelfRef.__Vtrigprevexpr___TOP__reset_l__0 = vlSelfRef.reset_l;
        vlSelfRef.__Vtrigprevexpr___TOP__in_small__0 = vlSelfRef.in_small;
        vlSelfRef.__Vtrigprevexpr___TOP__in_quad__0 = vlSelfRef.in_quad;
        VL_ASSIGN_W(70, vlSelfRef.__Vtrigprevexpr___TOP__in_wide__0, vlSelfRef.in_wide);
        vlSelfRef.__Vtrigprevexpr___TOP__clk__1 = vlSelfRef.clk;
    }
    // _eval_initial: top.v:36-43
    VLCU_D static void eval_initial(Root*, VlcuContext& vlCtx) {
        if (vlCtx.testPlusArg(0)) {  // $test$plusargs("trace")
            vlCtx.writef(0, 1, vlCtx.time());  // "[%0t] Tracing to logs/vlt_dump.vcd...\n"
            // $dumpfile / $dumpvars: waveform output is host side (vlcu_run trace samples)
        }
        vlCtx.writef(1, 1, vlCtx.time());  // "[%0t] Model running...\n"
    }
    // _eval_settle: 'stl' loop with the single "first iteration" trigger (V3Sched.cpp:433-493)
    VLCU_D static void stl_sequent__TOP__0(Root& vlSelfRef) {
        comb_out_small(vlSelfRef);
        comb_out_quad(vlSelfRef);
        comb_out_wide(vlSelfRef);
#if VTOP_COVERAGE
        togcov_clk(vlSelfRef);
        togcov_reset_l(vlSelfRef);
        togcov_in_small(vlSelfRef);
        togcov_in_quad(vlSelfRef);
        togcov_in_wide(vlSelfRef);
        togcov_out_small(vlSelfRef);
        togcov_out_quad(vlSelfRef);
        togcov_out_wide(vlSelfRef);
        togcov_count_c(vlSelfRef);
#endif
    }
    VLCU_D static bool eval_phase__stl(Root* vlSelf, VlcuContext&) {
        auto& vlSelfRef = *vlSelf;
        // _eval_triggers_vec__stl: bit 0 = first iteration
        vlSelfRef.__VstlTriggered[0U] = (QData)((IData)(vlSelfRef.__VstlFirstIteration));
        vlSelfRef.__VstlExecute = (0ULL != vlSelfRef.__VstlTriggered[0U]);
        if (vlSelfRef.__VstlExecute) {
            if ((1ULL & vlSelfRef.__VstlTriggered[0U])) stl_sequent__TOP__0(vlSelfRef);
        }
        return vlSelfRef.__VstlExecute;
    }
    VLCU_D static void eval_settle(Root* vlSelf, VlcuContext& vlCtx) {
        auto& vlSelfRef = *vlSelf;
        vlSelfRef.__VstlIterCount = 0U;
        vlSelfRef.__VstlFirstIteration = 1U;
        do {
            if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VstlIterCount)))) {
                VLCU_FATAL(vlCtx, 11, VLCU_FATAL_DIDNOTCONVERGE);
                break;  // the host runtime never returns from VL_FATAL_MT
            }
            vlSelfRef.__VstlIterCount = ((IData)(1U) + vlSelfRef.__VstlIterCount);
            vlSelfRef.__VstlPhaseResult = eval_phase__stl(vlSelf, vlCtx);
            vlSelfRef.__VstlFirstIteration = 0U;
        } while (vlSelfRef.__VstlPhaseResult);
    }

    //=====================================================================
    // 'ico' region
    VLCU_D static void eval_triggers_vec__ico(Root* vlSelf) {
        auto& vlSelfRef = *vlSelf;
        vlSelfRef.__VicoTriggered[0U] = (QData)((IData)(
            ((((IData)(VL_NEQ_W(3, vlSelfRef.in_wide, vlSelfRef.__Vtrigprevexpr___TOP__in_wide__0)) << 4U)
              | ((IData)((vlSelfRef.in_quad != vlSelfRef.__Vtrigprevexpr___TOP__in_quad__0)) << 3U))
             | ((((IData)(((IData)(vlSelfRef.in_small) != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__in_small__0))) << 2U)
                 | ((IData)(((IData)(vlSelfRef.reset_l) != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__reset_l__0))) << 1U))
                | (IData)(((IData)(vlSelfRef.clk) != (IData)(vlSelfRef.__Vtrigprevexpr___TOP__clk__0)))))));
        vlSelfRef.__VicoTriggered[1U] = 0ULL;
        vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
        vlSelfRef.__Vtrigprevexpr___TOP__reset_l__0 = vlSelfRef.reset_l;
        vlSelfRef.__Vtrigprevexpr___TOP__in_small__0 = vlSelfRef.in_small;
        vlSelfRef.__Vtrigprevexpr___TOP__in_quad__0 = vlSelfRef.in_quad;
        VL_ASSIGN_W(70, vlSelfRef.__Vtrigprevexpr___TOP__in_wide__0, vlSelfRef.in_wide);
        if (VL_UNLIKELY(((1U & (~(IData)(vlSelfRef.__VicoDidInit)))))) {
            vlSelfRef.__VicoDidInit = 1U;
            vlSelfRef.__VicoTriggered[0U] = (0x000000000000001fULL | vlSelfRef.__VicoTriggered[0U]);
        }
    }
    VLCU_D static void eval_ico(Root* vlSelf) {
        auto& vlSelfRef = *vlSelf;
#if VTOP_COVERAGE
        if ((1ULL & vlSelfRef.__VicoTriggered[0U])) togcov_clk(vlSelfRef);
        if ((2ULL & vlSelfRef.__VicoTriggered[0U])) togcov_reset_l(vlSelfRef);
        if ((4ULL & vlSelfRef.__VicoTriggered[0U])) togcov_in_small(vlSelfRef);
        if ((8ULL & vlSelfRef.__VicoTriggered[0U])) togcov_in_quad(vlSelfRef);
        if ((0x10ULL & vlSelfRef.__VicoTriggered[0U])) togcov_in_wide(vlSelfRef);
#endif
        if ((6ULL & vlSelfRef.__VicoTriggered[0U])) {
            comb_out_small(vlSelfRef);
#if VTOP_COVERAGE
            togcov_out_small(vlSelfRef);
#endif
        }
        if ((0x000000000000000aULL & vlSelfRef.__VicoTriggered[0U])) {
            comb_out_quad(vlSelfRef);
#if VTOP_COVERAGE
            togcov_out_quad(vlSelfRef);
#endif
        }
        if ((0x0000000000000012ULL & vlSelfRef.__VicoTriggered[0U])) {
            comb_out_wide(vlSelfRef);
#if VTOP_COVERAGE
            togcov_out_wide(vlSelfRef);
#endif
        }
    }
    VLCU_D static bool eval_phase__ico(Root* vlSelf) {
        auto& vlSelfRef = *vlSelf;
        eval_triggers_vec__ico(vlSelf);
        vlSelfRef.__VicoExecute
            = (0ULL != (vlSelfRef.__VicoTriggered[0U] | vlSelfRef.__VicoTriggered[1U]));
        if (vlSelfRef.__VicoExecute) eval_ico(vlSelf);
        return vlSelfRef.__VicoExecute;
    }

    //=====================================================================
    // 'act' / 'nba' regions
    VLCU_D static void eval_triggers_vec__act(Root* vlSelf) {
        auto& vlSelfRef = *vlSelf;
        // bit 0: @(posedge clk)
        vlSelfRef.__VactTriggered[0U]
            = (QData)((IData)(((IData)(vlSelfRef.clk)
                               & (~(IData)(vlSelfRef.__Vtrigprevexpr___TOP__clk__1)))));
        vlSelfRef.__Vtrigprevexpr___TOP__clk__1 = vlSelfRef.clk;
    }
    // sub.v:15-39 under posedge clk
    VLCU_D static void nba_sequent__TOP__0(Root* vlSelf, VlcuContext& vlCtx) {
        auto& vlSelfRef = *vlSelf;
        IData /*31:0*/ __Vdly__top__DOT__sub__DOT__count_c;
        __Vdly__top__DOT__sub__DOT__count_c = vlSelfRef.top__DOT__sub__DOT__count_c;
#if VTOP_COVERAGE
        if ((3U == vlSelfRef.top__DOT__sub__DOT__count_c)) ++(vlSelfRef.__Vcoverage[526]);  // cover property
        ++(vlSelfRef.__Vcoverage[525]);  // sub.v:34 block
#endif
        if (VL_UNLIKELY(((1U & (~((~(IData)(vlSelfRef.reset_l))
                                  | (0x00000064U > vlSelfRef.top__DOT__sub__DOT__count_c))))))) {
            // assert (!reset_l || count_c < 100)   sub.v:35, --assert
            vlCtx.writef(3, 1, vlCtx.time());
            VLCU_STOP(vlCtx, 35, false);
        }
#if VTOP_COVERAGE
        ++(vlSelfRef.__Vcoverage[520]);  // sub.v:15 block
#endif
        if (vlSelfRef.reset_l) {
#if VTOP_COVERAGE
            ++(vlSelfRef.__Vcoverage[522]);  // sub.v:22 else
#endif
            __Vdly__top__DOT__sub__DOT__count_c
                = ((IData)(1U) + vlSelfRef.top__DOT__sub__DOT__count_c);
            if ((3U <= vlSelfRef.top__DOT__sub__DOT__count_c)) {
#if VTOP_COVERAGE
                ++(vlSelfRef.__Vcoverage[523]);  // sub.v:24 if
#endif
                vlCtx.writef(2);  // "*-* All Finished *-*\n"
                VLCU_FINISH(vlCtx, 28);
            }
#if VTOP_COVERAGE
            else {
                ++(vlSelfRef.__Vcoverage[524]);  // sub.v:24 else
            }
#endif
        } else {
#if VTOP_COVERAGE
            ++(vlSelfRef.__Vcoverage[521]);  // sub.v:16 if
#endif
            __Vdly__top__DOT__sub__DOT__count_c = 0U;
        }
        vlSelfRef.top__DOT__sub__DOT__count_c = __Vdly__top__DOT__sub__DOT__count_c;
    }
    VLCU_D static void nba_comb__TOP__0(Root* vlSelf) {
#if VTOP_COVERAGE
        togcov_count_c(*vlSelf);
#else
        (void)vlSelf;
#endif
    }
    VLCU_D static void eval_nba(Root* vlSelf, VlcuContext& vlCtx) {
        auto& vlSelfRef = *vlSelf;
        if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
            nba_sequent__TOP__0(vlSelf, vlCtx);
            nba_comb__TOP__0(vlSelf);
        }
    }
    VLCU_D static bool eval_phase__act(Root* vlSelf) {
        auto& vlSelfRef = *vlSelf;
        eval_triggers_vec__act(vlSelf);
        // _trigger_orInto__act_vec_vec: latch 'act' triggers under 'nba'
        vlSelfRef.__VnbaTriggered[0U] = (vlSelfRef.__VnbaTriggered[0U] | vlSelfRef.__VactTriggered[0U]);
        vlSelfRef.__VactExecute = (0ULL != vlSelfRef.__VactTriggered[0U]);
        if (vlSelfRef.__VactExecute) {
            // _eval_act: no logic in the 'act' region (the only clock is a top-level input)
        }
        return vlSelfRef.__VactExecute;
    }
    VLCU_D static bool eval_phase__nba(Root* vlSelf, VlcuContext& vlCtx) {
        auto& vlSelfRef = *vlSelf;
        vlSelfRef.__VnbaExecute = (0ULL != vlSelfRef.__VnbaTriggered[0U]);
        if (vlSelfRef.__VnbaExecute) {
            eval_nba(vlSelf, vlCtx);
            vlSelfRef.__VnbaTriggered[0U] = 0ULL;  // _trigger_clear__act
        }
        return vlSelfRef.__VnbaExecute;
    }

    // Vtop___024root___eval  (V3Sched.cpp:869-883)
    VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) {
        auto& vlSelfRef = *vlSelf;
        vlSelfRef.__VicoIterCount = 0U;
        vlSelfRef.__VicoFirstIteration = 1U;
        do {
            if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VicoIterCount)))) {
                VLCU_FATAL(vlCtx, 11, VLCU_FATAL_DIDNOTCONVERGE);
                break;
            }
            vlSelfRef.__VicoIterCount = ((IData)(1U) + vlSelfRef.__VicoIterCount);
            vlSelfRef.__VicoPhaseResult = eval_phase__ico(vlSelf);
            vlSelfRef.__VicoFirstIteration = 0U;
        } while (vlSelfRef.__VicoPhaseResult);
        vlSelfRef.__VnbaIterCount = 0U;
        vlSelfRef.__VnbaFirstIteration = 1U;
        // __VnbaTriggered is always clear when _eval returns (_trigger_clear__act below), so it
        // is demoted from a persistent member to a temporary that starts each eval at zero
        vlSelfRef.__VnbaTriggered[0U] = 0ULL;
        do {
            if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VnbaIterCount)))) {
                VLCU_FATAL(vlCtx, 11, VLCU_FATAL_DIDNOTCONVERGE);
                break;
            }
            vlSelfRef.__VnbaIterCount = ((IData)(1U) + vlSelfRef.__VnbaIterCount);
            vlSelfRef.__VactIterCount = 0U;
            vlSelfRef.__VactFirstIteration = 1U;
            do {
                if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
                    VLCU_FATAL(vlCtx, 11, VLCU_FATAL_DIDNOTCONVERGE);
                    break;
                }
                vlSelfRef.__VactIterCount = ((IData)(1U) + vlSelfRef.__VactIterCount);
                vlSelfRef.__VactPhaseResult = eval_phase__act(vlSelf);
                vlSelfRef.__VactFirstIteration = 0U;
            } while (vlSelfRef.__VactPhaseResult);
            vlSelfRef.__VnbaPhaseResult = eval_phase__nba(vlSelf, vlCtx);
            vlSelfRef.__VnbaFirstIteration = 0U;
        } while (vlSelfRef.__VnbaPhaseResult);
    }
    VLCU_D static void eval_final(Root*, VlcuContext&) {}

    static const VlcuDesignInfo& info();  // host tables, Vtop__cuda.cu
};
