// TEST INFRASTRUCTURE (not product code) -- design-level ORACLE, *restated*.
//
// Hand restatement of the C++ Verilator emits for examples/make_tracing_c
// (see Vtop.h).  Every function cites the generator rule it follows.  The
// VL_* operators, VerilatedContext, VL_SCOPED_RAND_RESET_*, VL_COVER_INSERT,
// VerilatedVcd and VL_FINISH_MT/VL_STOP_MT/VL_WRITEF_NX used below are the
// reference's own (compiled from /root/reference/include).
#include "Vtop.h"

#include "verilated_vcd_c.h"

#include "Vtop___024root.h"

//============================================================
// Symbol table (src/V3EmitCSyms.cpp:923-982, ctor :1503-)

Vtop__Syms::Vtop__Syms(VerilatedContext* contextp, const char* namep, Vtop* modelp)
    : VerilatedSyms{contextp}
    , __Vm_modelp{modelp}
    , TOP{this, namep} {
    // Configure time unit / time precision
    _vm_contextp__->timeunit(-12);
    _vm_contextp__->timeprecision(-12);
    // Setup each module's pointer back to symbol table (for public functions)
    TOP.__Vconfigure(true);
}
Vtop__Syms::~Vtop__Syms() {}

//============================================================
// Root module constructor (src/V3EmitCImp.cpp:91-172)

Vtop___024root::Vtop___024root(Vtop__Syms* symsp, const char* namep) {
    vlSymsp = symsp;
    vlNamep = strdup(namep);  // modp->isTop()
    // Reset structure values
    Vtop___024root___ctor_var_reset(this);
}
Vtop___024root::~Vtop___024root() { VL_DO_DANGLING(free(const_cast<char*>(vlNamep)), vlNamep); }

// V3EmitCImp.cpp:179-215
void Vtop___024root::__vlCoverInsert(uint32_t* countp, bool enable, bool localCounter,
                                     const char* filenamep, int lineno, int column,
                                     const char* hierp, const char* pagep, const char* commentp,
                                     const char* linescovp) {
    uint32_t* count32p = countp;
    static uint32_t fake_zero_count = 0;
    std::string fullhier = std::string{vlNamep} + hierp;
    if (!fullhier.empty() && fullhier[0] == '.') fullhier = fullhier.substr(1);
    if (!enable && !localCounter) count32p = &fake_zero_count;
    *count32p = 0;
    VL_COVER_INSERT(vlSymsp->_vm_contextp__->coveragep(), vlNamep, count32p, "filename", filenamep,
                    "lineno", lineno, "column", column, "hier", fullhier.c_str(), "page", pagep,
                    "comment", commentp, (linescovp[0] ? "linescov" : ""), linescovp);
}
// V3EmitCImp.cpp:216-262
void Vtop___024root::__vlCoverToggleInsert(int begin, int end, bool ranged, uint32_t* countp,
                                           bool enable, bool localCounter, const char* filenamep,
                                           int lineno, int column, const char* hierp,
                                           const char* pagep, const char* commentp) {
    int step = (end >= begin) ? 1 : -1;
    for (int i = begin; i != end + step; i += step) {
        for (int j = 0; j < 2; j++) {
            uint32_t* count32p = countp;
            static uint32_t fake_zero_count = 0;
            std::string fullhier = std::string{vlNamep} + hierp;
            if (!fullhier.empty() && fullhier[0] == '.') fullhier = fullhier.substr(1);
            std::string commentWithIndex = commentp;
            if (ranged) commentWithIndex += '[' + std::to_string(i) + ']';
            commentWithIndex += j ? ":0->1" : ":1->0";
            if (!enable && !localCounter) count32p = &fake_zero_count;
            *count32p = 0;
            VL_COVER_INSERT(vlSymsp->_vm_contextp__->coveragep(), vlNamep, count32p, "filename",
                            filenamep, "lineno", lineno, "column", column, "hier",
                            fullhier.c_str(), "page", pagep, "comment", commentWithIndex.c_str(),
                            "", "");
            ++countp;
        }
    }
}

// __Vconfigure: coverage registrations (V3EmitCImp.cpp:157-172; V3EmitCFunc.h:179-215,895-927)
void Vtop___024root::__Vconfigure(bool first) {
    (void)first;
    Vtop___024root* const vlSelf = this;
    vlSelf->__vlCoverToggleInsert(0, 0, 0, vlSelf->__Vcoverage + 0, first, true, "top.v", 13, 11, ".top", "v_toggle/top", "clk");
    vlSelf->__vlCoverToggleInsert(0, 0, 0, vlSelf->__Vcoverage + 2, first, true, "top.v", 14, 11, ".top", "v_toggle/top", "reset_l");
    vlSelf->__vlCoverToggleInsert(0, 1, 1, vlSelf->__Vcoverage + 4, first, true, "top.v", 16, 23, ".top", "v_toggle/top", "out_small");
    vlSelf->__vlCoverToggleInsert(0, 39, 1, vlSelf->__Vcoverage + 8, first, true, "top.v", 17, 24, ".top", "v_toggle/top", "out_quad");
    vlSelf->__vlCoverToggleInsert(0, 69, 1, vlSelf->__Vcoverage + 88, first, true, "top.v", 18, 24, ".top", "v_toggle/top", "out_wide");
    vlSelf->__vlCoverToggleInsert(0, 1, 1, vlSelf->__Vcoverage + 228, first, true, "top.v", 19, 17, ".top", "v_toggle/top", "in_small");
    vlSelf->__vlCoverToggleInsert(0, 39, 1, vlSelf->__Vcoverage + 232, first, true, "top.v", 20, 18, ".top", "v_toggle/top", "in_quad");
    vlSelf->__vlCoverToggleInsert(0, 69, 1, vlSelf->__Vcoverage + 312, first, true, "top.v", 21, 18, ".top", "v_toggle/top", "in_wide");
    vlSelf->__vlCoverToggleInsert(0, 0, 0, vlSelf->__Vcoverage + 452, first, true, "sub.v", 9, 11, ".top.sub", "v_toggle/sub", "clk");
    vlSelf->__vlCoverToggleInsert(0, 0, 0, vlSelf->__Vcoverage + 454, first, true, "sub.v", 10, 11, ".top.sub", "v_toggle/sub", "reset_l");
    vlSelf->__vlCoverToggleInsert(0, 31, 1, vlSelf->__Vcoverage + 456, first, true, "sub.v", 14, 14, ".top.sub", "v_toggle/sub", "count_c");
    vlSelf->__vlCoverInsert(vlSelf->__Vcoverage + 520, first, true, "sub.v", 15, 3, ".top.sub", "v_line/sub", "block", "15");
    vlSelf->__vlCoverInsert(vlSelf->__Vcoverage + 521, first, true, "sub.v", 16, 5, ".top.sub", "v_line/sub", "if", "16,19");
    vlSelf->__vlCoverInsert(vlSelf->__Vcoverage + 522, first, true, "sub.v", 16, 6, ".top.sub", "v_line/sub", "else", "23");
    vlSelf->__vlCoverInsert(vlSelf->__Vcoverage + 523, first, true, "sub.v", 24, 7, ".top.sub", "v_line/sub", "if", "24,27-28");
    vlSelf->__vlCoverInsert(vlSelf->__Vcoverage + 524, first, true, "sub.v", 24, 8, ".top.sub", "v_line/sub", "else", "");
    vlSelf->__vlCoverInsert(vlSelf->__Vcoverage + 525, first, true, "sub.v", 34, 3, ".top.sub", "v_line/sub", "block", "34-35");
    vlSelf->__vlCoverInsert(vlSelf->__Vcoverage + 526, first, true, "sub.v", 39, 3, ".top.sub", "v_user/sub", "cover", "");
}

//============================================================
// _ctor_var_reset (src/V3CCtors.cpp:52-236, src/V3EmitCFunc.cpp:524-677):
// --x-initial unique -> VL_SCOPED_RAND_RESET_* for user variables, zero for temporaries

void Vtop___024root___ctor_var_reset(Vtop___024root* vlSelf) {
    auto& vlSelfRef = std::ref(*vlSelf).get();
    const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
    vlSelfRef.clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
    vlSelfRef.reset_l = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 12926602469485068081ull);
    vlSelfRef.out_small = VL_SCOPED_RAND_RESET_I(2, __VscopeHash, 14267498806251343750ull);
    vlSelfRef.out_quad = VL_SCOPED_RAND_RESET_Q(40, __VscopeHash, 3771664055402206494ull);
    VL_SCOPED_RAND_RESET_W(70, vlSelfRef.out_wide, __VscopeHash, 6124341913962757201ull);
    vlSelfRef.in_small = VL_SCOPED_RAND_RESET_I(2, __VscopeHash, 4925076677430572012ull);
    vlSelfRef.in_quad = VL_SCOPED_RAND_RESET_Q(40, __VscopeHash, 7333266129396445036ull);
    VL_SCOPED_RAND_RESET_W(70, vlSelfRef.in_wide, __VscopeHash, 18101672252301227225ull);
    vlSelfRef.top__DOT__sub__DOT__count_c
        = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 2802051788925154131ull);
    vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = 0;
    vlSelfRef.__Vtrigprevexpr___TOP__reset_l__0 = 0;
    vlSelfRef.__Vtrigprevexpr___TOP__in_small__0 = 0;
    vlSelfRef.__Vtrigprevexpr___TOP__in_quad__0 = 0;
    VL_ZERO_RESET_W(70, vlSelfRef.__Vtrigprevexpr___TOP__in_wide__0);
    vlSelfRef.__Vtrigprevexpr___TOP__clk__1 = 0;
    vlSelfRef.__VicoDidInit = 0;
    vlSelfRef.top__DOT____Vtogcov__clk = 0;
    vlSelfRef.top__DOT____Vtogcov__reset_l = 0;
    vlSelfRef.top__DOT____Vtogcov__out_small = 0;
    vlSelfRef.top__DOT____Vtogcov__out_quad = 0;
    VL_ZERO_RESET_W(70, vlSelfRef.top__DOT____Vtogcov__out_wide);
    vlSelfRef.top__DOT____Vtogcov__in_small = 0;
    vlSelfRef.top__DOT____Vtogcov__in_quad = 0;
    VL_ZERO_RESET_W(70, vlSelfRef.top__DOT____Vtogcov__in_wide);
    vlSelfRef.top__DOT__sub__DOT____Vtogcov__clk = 0;
    vlSelfRef.top__DOT__sub__DOT____Vtogcov__reset_l = 0;
    vlSelfRef.top__DOT__sub__DOT____Vtogcov__count_c = 0;
    for (int __Vi0 = 0; __Vi0 < 1; ++__Vi0) vlSelfRef.__VstlTriggered[__Vi0] = 0;
    for (int __Vi0 = 0; __Vi0 < 2; ++__Vi0) vlSelfRef.__VicoTriggered[__Vi0] = 0;
    for (int __Vi0 = 0; __Vi0 < 1; ++__Vi0) vlSelfRef.__VactTriggered[__Vi0] = 0;
    for (int __Vi0 = 0; __Vi0 < 1; ++__Vi0) vlSelfRef.__VnbaTriggered[__Vi0] = 0;
}

//============================================================
// Combinational logic top.v:25-27 and toggle coverage (V3Clock.cpp:69-98)

static void comb_out_small(Vtop___024root& vlSelfRef) {
    vlSelfRef.out_small = ((1U & (~(IData)(vlSelfRef.reset_l)))
                               ? 0U
                               : (3U & ((IData)(1U) + (IData)(vlSelfRef.in_small))));
}
static void comb_out_quad(Vtop___024root& vlSelfRef) {
    vlSelfRef.out_quad = ((1U & (~(IData)(vlSelfRef.reset_l)))
                              ? 0ULL
                              : (0x000000ffffffffffULL & (1ULL + vlSelfRef.in_quad)));
}
static void comb_out_wide(Vtop___024root& vlSelfRef) {
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
static void togcov_clk(Vtop___024root& vlSelfRef) {
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
static void togcov_reset_l(Vtop___024root& vlSelfRef) {
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
static void togcov_in_small(Vtop___024root& vlSelfRef) {
    if ((IData)(vlSelfRef.in_small) ^ (IData)(vlSelfRef.top__DOT____Vtogcov__in_small)) {
        VL_COV_TOGGLE_CHG_ST_I(2, vlSelfRef.__Vcoverage + 228, vlSelfRef.in_small,
                               vlSelfRef.top__DOT____Vtogcov__in_small);
        vlSelfRef.top__DOT____Vtogcov__in_small = vlSelfRef.in_small;
    }
}
static void togcov_in_quad(Vtop___024root& vlSelfRef) {
    if (0ULL != (vlSelfRef.in_quad ^ vlSelfRef.top__DOT____Vtogcov__in_quad)) {
        VL_COV_TOGGLE_CHG_ST_Q(40, vlSelfRef.__Vcoverage + 232, vlSelfRef.in_quad,
                               vlSelfRef.top__DOT____Vtogcov__in_quad);
        vlSelfRef.top__DOT____Vtogcov__in_quad = vlSelfRef.in_quad;
    }
}
static void togcov_in_wide(Vtop___024root& vlSelfRef) {
    if (VL_CHANGEXOR_W(3, vlSelfRef.in_wide, vlSelfRef.top__DOT____Vtogcov__in_wide)) {
        VL_COV_TOGGLE_CHG_ST_W(70, vlSelfRef.__Vcoverage + 312, vlSelfRef.in_wide,
                               vlSelfRef.top__DOT____Vtogcov__in_wide);
        VL_ASSIGN_W(70, vlSelfRef.top__DOT____Vtogcov__in_wide, vlSelfRef.in_wide);
    }
}
static void togcov_out_small(Vtop___024root& vlSelfRef) {
    if ((IData)(vlSelfRef.out_small) ^ (IData)(vlSelfRef.top__DOT____Vtogcov__out_small)) {
        VL_COV_TOGGLE_CHG_ST_I(2, vlSelfRef.__Vcoverage + 4, vlSelfRef.out_small,
                               vlSelfRef.top__DOT____Vtogcov__out_small);
        vlSelfRef.top__DOT____Vtogcov__out_small = vlSelfRef.out_small;
    }
}
static void togcov_out_quad(Vtop___024root& vlSelfRef) {
    if (0ULL != (vlSelfRef.out_quad ^ vlSelfRef.top__DOT____Vtogcov__out_quad)) {
        VL_COV_TOGGLE_CHG_ST_Q(40, vlSelfRef.__Vcoverage + 8, vlSelfRef.out_quad,
                               vlSelfRef.top__DOT____Vtogcov__out_quad);
        vlSelfRef.top__DOT____Vtogcov__out_quad = vlSelfRef.out_quad;
    }
}
static void togcov_out_wide(Vtop___024root& vlSelfRef) {
    if (VL_CHANGEXOR_W(3, vlSelfRef.out_wide, vlSelfRef.top__DOT____Vtogcov__out_wide)) {
        VL_COV_TOGGLE_CHG_ST_W(70, vlSelfRef.__Vcoverage + 88, vlSelfRef.out_wide,
                               vlSelfRef.top__DOT____Vtogcov__out_wide);
        VL_ASSIGN_W(70, vlSelfRef.top__DOT____Vtogcov__out_wide, vlSelfRef.out_wide);
    }
}
static void togcov_count_c(Vtop___024root& vlSelfRef) {
    if (vlSelfRef.top__DOT__sub__DOT__count_c ^ vlSelfRef.top__DOT__sub__DOT____Vtogcov__count_c) {
        VL_COV_TOGGLE_CHG_ST_I(32, vlSelfRef.__Vcoverage + 456,
                               vlSelfRef.top__DOT__sub__DOT__count_c,
                               vlSelfRef.top__DOT__sub__DOT____Vtogcov__count_c);
        vlSelfRef.top__DOT__sub__DOT____Vtogcov__count_c = vlSelfRef.top__DOT__sub__DOT__count_c;
    }
}

//============================================================
// _eval_static / _eval_initial / _eval_settle (V3Sched.cpp:958-984, 433-493)

void Vtop___024root___eval_static(Vtop___024root* vlSelf) {
    auto& vlSelfRef = std::ref(*vlSelf).get();
    vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = vlSelfRef.clk;
    vlSelfRef.__Vtrigprevexpr___TOP__reset_l__0 = vlSelfRef.reset_l;
    vlSelfRef.__Vtrigprevexpr___TOP__in_small__0 = vlSelfRef.in_small;
    vlSelfRef.__Vtrigprevexpr___TOP__in_quad__0 = vlSelfRef.in_quad;
    VL_ASSIGN_W(70, vlSelfRef.__Vtrigprevexpr___TOP__in_wide__0, vlSelfRef.in_wide);
    vlSelfRef.__Vtrigprevexpr___TOP__clk__1 = vlSelfRef.clk;
}

void Vtop___024root___eval_initial(Vtop___024root* vlSelf) {
    Vtop__Syms* const __restrict vlSymsp VL_ATTR_UNUSED = vlSelf->vlSymsp;
    // top.v:36-43
    if ((0U != VL_TESTPLUSARGS_I(std::string{"trace"}))) {
        VL_WRITEF_NX("[%0t] Tracing to logs/vlt_dump.vcd...\n\n", 2, 'T', -12, '#', 64, VL_TIME_UNITED_Q(1));
        // $dumpfile/$dumpvars: the oracle harness attaches the VCD writer itself
    }
    VL_WRITEF_NX("[%0t] Model running...\n\n", 2, 'T', -12, '#', 64, VL_TIME_UNITED_Q(1));
}

static void stl_sequent__TOP__0(Vtop___024root& vlSelfRef) {
    comb_out_small(vlSelfRef);
    comb_out_quad(vlSelfRef);
    comb_out_wide(vlSelfRef);
    togcov_clk(vlSelfRef);
    togcov_reset_l(vlSelfRef);
    togcov_in_small(vlSelfRef);
    togcov_in_quad(vlSelfRef);
    togcov_in_wide(vlSelfRef);
    togcov_out_small(vlSelfRef);
    togcov_out_quad(vlSelfRef);
    togcov_out_wide(vlSelfRef);
    togcov_count_c(vlSelfRef);
}
static bool eval_phase__stl(Vtop___024root* vlSelf) {
    auto& vlSelfRef = std::ref(*vlSelf).get();
    vlSelfRef.__VstlTriggered[0U] = (QData)((IData)(vlSelfRef.__VstlFirstIteration));
    vlSelfRef.__VstlExecute = (0ULL != vlSelfRef.__VstlTriggered[0U]);
    if (vlSelfRef.__VstlExecute) {
        if ((1ULL & vlSelfRef.__VstlTriggered[0U])) stl_sequent__TOP__0(vlSelfRef);
    }
    return vlSelfRef.__VstlExecute;
}
void Vtop___024root___eval_settle(Vtop___024root* vlSelf) {
    auto& vlSelfRef = std::ref(*vlSelf).get();
    vlSelfRef.__VstlIterCount = 0U;
    vlSelfRef.__VstlFirstIteration = 1U;
    do {
        if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VstlIterCount)))) {
            VL_FATAL_MT("top.v", 11, "", "DIDNOTCONVERGE: Settle region did not converge after '--converge-limit' of 10000 tries");
            break;
        }
        vlSelfRef.__VstlIterCount = ((IData)(1U) + vlSelfRef.__VstlIterCount);
        vlSelfRef.__VstlPhaseResult = eval_phase__stl(vlSelf);
        vlSelfRef.__VstlFirstIteration = 0U;
    } while (vlSelfRef.__VstlPhaseResult);
}

//============================================================
// 'ico' region (V3Sched.cpp:498-654; V3SchedTrigger.cpp:560-583,696-804)

static void eval_triggers_vec__ico(Vtop___024root* vlSelf) {
    auto& vlSelfRef = std::ref(*vlSelf).get();
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
static void eval_ico(Vtop___024root* vlSelf) {
    auto& vlSelfRef = std::ref(*vlSelf).get();
    if ((1ULL & vlSelfRef.__VicoTriggered[0U])) togcov_clk(vlSelfRef);
    if ((2ULL & vlSelfRef.__VicoTriggered[0U])) togcov_reset_l(vlSelfRef);
    if ((4ULL & vlSelfRef.__VicoTriggered[0U])) togcov_in_small(vlSelfRef);
    if ((8ULL & vlSelfRef.__VicoTriggered[0U])) togcov_in_quad(vlSelfRef);
    if ((0x10ULL & vlSelfRef.__VicoTriggered[0U])) togcov_in_wide(vlSelfRef);
    if ((6ULL & vlSelfRef.__VicoTriggered[0U])) {
        comb_out_small(vlSelfRef);
        togcov_out_small(vlSelfRef);
    }
    if ((0x000000000000000aULL & vlSelfRef.__VicoTriggered[0U])) {
        comb_out_quad(vlSelfRef);
        togcov_out_quad(vlSelfRef);
    }
    if ((0x0000000000000012ULL & vlSelfRef.__VicoTriggered[0U])) {
        comb_out_wide(vlSelfRef);
        togcov_out_wide(vlSelfRef);
    }
}
static bool eval_phase__ico(Vtop___024root* vlSelf) {
    auto& vlSelfRef = std::ref(*vlSelf).get();
    eval_triggers_vec__ico(vlSelf);
    vlSelfRef.__VicoExecute
        = (0ULL != (vlSelfRef.__VicoTriggered[0U] | vlSelfRef.__VicoTriggered[1U]));
    if (vlSelfRef.__VicoExecute) eval_ico(vlSelf);
    return vlSelfRef.__VicoExecute;
}

//============================================================
// 'act' / 'nba' regions (V3Sched.cpp:671-884)

static void eval_triggers_vec__act(Vtop___024root* vlSelf) {
    auto& vlSelfRef = std::ref(*vlSelf).get();
    vlSelfRef.__VactTriggered[0U]
        = (QData)((IData)(((IData)(vlSelfRef.clk)
                           & (~(IData)(vlSelfRef.__Vtrigprevexpr___TOP__clk__1)))));
    vlSelfRef.__Vtrigprevexpr___TOP__clk__1 = vlSelfRef.clk;
}
// sub.v:15-39 under posedge clk; NBA through a shadow variable (V3Delayed.cpp:48-57)
static void nba_sequent__TOP__0(Vtop___024root* vlSelf) {
    Vtop__Syms* const __restrict vlSymsp VL_ATTR_UNUSED = vlSelf->vlSymsp;
    auto& vlSelfRef = std::ref(*vlSelf).get();
    IData /*31:0*/ __Vdly__top__DOT__sub__DOT__count_c;
    __Vdly__top__DOT__sub__DOT__count_c = vlSelfRef.top__DOT__sub__DOT__count_c;
    if ((3U == vlSelfRef.top__DOT__sub__DOT__count_c)) ++(vlSelf->__Vcoverage[526]);
    ++(vlSelf->__Vcoverage[525]);
    if (VL_UNLIKELY(((1U & (~((~(IData)(vlSelfRef.reset_l))
                              | (0x00000064U > vlSelfRef.top__DOT__sub__DOT__count_c))))))) {
        VL_WRITEF_NX("[%0t] %%Error: sub.v:35: Assertion failed in TOP.top.sub.AssertionExample\n",
                     2, 'T', -12, '#', 64, VL_TIME_UNITED_Q(1));
        VL_STOP_MT("sub.v", 35, "");
    }
    ++(vlSelf->__Vcoverage[520]);
    if (vlSelfRef.reset_l) {
        ++(vlSelf->__Vcoverage[522]);
        __Vdly__top__DOT__sub__DOT__count_c = ((IData)(1U) + vlSelfRef.top__DOT__sub__DOT__count_c);
        if ((3U <= vlSelfRef.top__DOT__sub__DOT__count_c)) {
            ++(vlSelf->__Vcoverage[523]);
            VL_WRITEF_NX("*-* All Finished *-*\n", 0);
            VL_FINISH_MT("sub.v", 28, "");
        } else {
            ++(vlSelf->__Vcoverage[524]);
        }
    } else {
        ++(vlSelf->__Vcoverage[521]);
        __Vdly__top__DOT__sub__DOT__count_c = 0U;
    }
    vlSelfRef.top__DOT__sub__DOT__count_c = __Vdly__top__DOT__sub__DOT__count_c;
}
static void nba_comb__TOP__0(Vtop___024root* vlSelf) { togcov_count_c(*vlSelf); }
static void eval_nba(Vtop___024root* vlSelf) {
    auto& vlSelfRef = std::ref(*vlSelf).get();
    if ((1ULL & vlSelfRef.__VnbaTriggered[0U])) {
        nba_sequent__TOP__0(vlSelf);
        nba_comb__TOP__0(vlSelf);
        vlSelf->vlSymsp->__Vm_traceActivity = true;
    }
}
static bool eval_phase__act(Vtop___024root* vlSelf) {
    auto& vlSelfRef = std::ref(*vlSelf).get();
    eval_triggers_vec__act(vlSelf);
    vlSelfRef.__VnbaTriggered[0U] = (vlSelfRef.__VnbaTriggered[0U] | vlSelfRef.__VactTriggered[0U]);
    vlSelfRef.__VactExecute = (0ULL != vlSelfRef.__VactTriggered[0U]);
    if (vlSelfRef.__VactExecute) {
        // _eval_act: empty
    }
    return vlSelfRef.__VactExecute;
}
static bool eval_phase__nba(Vtop___024root* vlSelf) {
    auto& vlSelfRef = std::ref(*vlSelf).get();
    vlSelfRef.__VnbaExecute = (0ULL != vlSelfRef.__VnbaTriggered[0U]);
    if (vlSelfRef.__VnbaExecute) {
        eval_nba(vlSelf);
        vlSelfRef.__VnbaTriggered[0U] = 0ULL;
    }
    return vlSelfRef.__VnbaExecute;
}

void Vtop___024root___eval(Vtop___024root* vlSelf) {
    auto& vlSelfRef = std::ref(*vlSelf).get();
    vlSelfRef.__VicoIterCount = 0U;
    vlSelfRef.__VicoFirstIteration = 1U;
    do {
        if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VicoIterCount)))) {
            VL_FATAL_MT("top.v", 11, "", "DIDNOTCONVERGE: Input combinational region did not converge after '--converge-limit' of 10000 tries");
            break;
        }
        vlSelfRef.__VicoIterCount = ((IData)(1U) + vlSelfRef.__VicoIterCount);
        vlSelfRef.__VicoPhaseResult = eval_phase__ico(vlSelf);
        vlSelfRef.__VicoFirstIteration = 0U;
    } while (vlSelfRef.__VicoPhaseResult);
    vlSelfRef.__VnbaIterCount = 0U;
    vlSelfRef.__VnbaFirstIteration = 1U;
    do {
        if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VnbaIterCount)))) {
            VL_FATAL_MT("top.v", 11, "", "DIDNOTCONVERGE: NBA region did not converge after '--converge-limit' of 10000 tries");
            break;
        }
        vlSelfRef.__VnbaIterCount = ((IData)(1U) + vlSelfRef.__VnbaIterCount);
        vlSelfRef.__VactIterCount = 0U;
        vlSelfRef.__VactFirstIteration = 1U;
        do {
            if (VL_UNLIKELY(((0x00002710U < vlSelfRef.__VactIterCount)))) {
                VL_FATAL_MT("top.v", 11, "", "DIDNOTCONVERGE: Active region did not converge after '--converge-limit' of 10000 tries");
                break;
            }
            vlSelfRef.__VactIterCount = ((IData)(1U) + vlSelfRef.__VactIterCount);
            vlSelfRef.__VactPhaseResult = eval_phase__act(vlSelf);
            vlSelfRef.__VactFirstIteration = 0U;
        } while (vlSelfRef.__VactPhaseResult);
        vlSelfRef.__VnbaPhaseResult = eval_phase__nba(vlSelf);
        vlSelfRef.__VnbaFirstIteration = 0U;
    } while (vlSelfRef.__VnbaPhaseResult);
}
void Vtop___024root___eval_final(Vtop___024root*) {}

//============================================================
// Tracing (src/V3TraceDecl.cpp, src/V3Trace.cpp:17-35, src/V3EmitCImp.cpp:643-880)

static void trace_init_top(Vtop___024root* vlSelf, VerilatedVcd* tracep) {
    (void)vlSelf;
    const int c = vlSelf->vlSymsp->__Vm_baseCode;
    VL_TRACE_DECL_BIT(tracep, c + 1, 0, "clk", -1, VerilatedTraceSigDirection::INPUT, VerilatedTraceSigKind::WIRE, VerilatedTraceSigType::LOGIC);
    VL_TRACE_DECL_BIT(tracep, c + 2, 0, "reset_l", -1, VerilatedTraceSigDirection::INPUT, VerilatedTraceSigKind::WIRE, VerilatedTraceSigType::LOGIC);
    VL_TRACE_DECL_BUS(tracep, c + 3, 0, "out_small", -1, VerilatedTraceSigDirection::OUTPUT, VerilatedTraceSigKind::WIRE, VerilatedTraceSigType::LOGIC, 1, 0);
    VL_TRACE_DECL_QUAD(tracep, c + 4, 0, "out_quad", -1, VerilatedTraceSigDirection::OUTPUT, VerilatedTraceSigKind::WIRE, VerilatedTraceSigType::LOGIC, 39, 0);
    VL_TRACE_DECL_WIDE(tracep, c + 6, 0, "out_wide", -1, VerilatedTraceSigDirection::OUTPUT, VerilatedTraceSigKind::WIRE, VerilatedTraceSigType::LOGIC, 69, 0);
    VL_TRACE_DECL_BUS(tracep, c + 9, 0, "in_small", -1, VerilatedTraceSigDirection::INPUT, VerilatedTraceSigKind::WIRE, VerilatedTraceSigType::LOGIC, 1, 0);
    VL_TRACE_DECL_QUAD(tracep, c + 10, 0, "in_quad", -1, VerilatedTraceSigDirection::INPUT, VerilatedTraceSigKind::WIRE, VerilatedTraceSigType::LOGIC, 39, 0);
    VL_TRACE_DECL_WIDE(tracep, c + 12, 0, "in_wide", -1, VerilatedTraceSigDirection::INPUT, VerilatedTraceSigKind::WIRE, VerilatedTraceSigType::LOGIC, 69, 0);
    VL_TRACE_PUSH_PREFIX(tracep, "top", VerilatedTracePrefixType::SCOPE_MODULE, 0, 0);
    VL_TRACE_DECL_BIT(tracep, c + 1, 0, "clk", -1, VerilatedTraceSigDirection::INPUT, VerilatedTraceSigKind::WIRE, VerilatedTraceSigType::LOGIC);
    VL_TRACE_DECL_BIT(tracep, c + 2, 0, "reset_l", -1, VerilatedTraceSigDirection::INPUT, VerilatedTraceSigKind::WIRE, VerilatedTraceSigType::LOGIC);
    VL_TRACE_DECL_BUS(tracep, c + 3, 0, "out_small", -1, VerilatedTraceSigDirection::OUTPUT, VerilatedTraceSigKind::WIRE, VerilatedTraceSigType::LOGIC, 1, 0);
    VL_TRACE_DECL_QUAD(tracep, c + 4, 0, "out_quad", -1, VerilatedTraceSigDirection::OUTPUT, VerilatedTraceSigKind::WIRE, VerilatedTraceSigType::LOGIC, 39, 0);
    VL_TRACE_DECL_WIDE(tracep, c + 6, 0, "out_wide", -1, VerilatedTraceSigDirection::OUTPUT, VerilatedTraceSigKind::WIRE, VerilatedTraceSigType::LOGIC, 69, 0);
    VL_TRACE_DECL_BUS(tracep, c + 9, 0, "in_small", -1, VerilatedTraceSigDirection::INPUT, VerilatedTraceSigKind::WIRE, VerilatedTraceSigType::LOGIC, 1, 0);
    VL_TRACE_DECL_QUAD(tracep, c + 10, 0, "in_quad", -1, VerilatedTraceSigDirection::INPUT, VerilatedTraceSigKind::WIRE, VerilatedTraceSigType::LOGIC, 39, 0);
    VL_TRACE_DECL_WIDE(tracep, c + 12, 0, "in_wide", -1, VerilatedTraceSigDirection::INPUT, VerilatedTraceSigKind::WIRE, VerilatedTraceSigType::LOGIC, 69, 0);
    VL_TRACE_PUSH_PREFIX(tracep, "sub", VerilatedTracePrefixType::SCOPE_MODULE, 0, 0);
    VL_TRACE_DECL_BIT(tracep, c + 1, 0, "clk", -1, VerilatedTraceSigDirection::INPUT, VerilatedTraceSigKind::WIRE, VerilatedTraceSigType::LOGIC);
    VL_TRACE_DECL_BIT(tracep, c + 2, 0, "reset_l", -1, VerilatedTraceSigDirection::INPUT, VerilatedTraceSigKind::WIRE, VerilatedTraceSigType::LOGIC);
    VL_TRACE_DECL_BUS(tracep, c + 15, 0, "count_c", -1, VerilatedTraceSigDirection::NONE, VerilatedTraceSigKind::VAR, VerilatedTraceSigType::LOGIC, 31, 0);
    VL_TRACE_POP_PREFIX(tracep);
    VL_TRACE_POP_PREFIX(tracep);
}

static void trace_init(void* voidSelf, VerilatedVcd* tracep, uint32_t code) {
    Vtop___024root* const __restrict vlSelf VL_ATTR_UNUSED = static_cast<Vtop___024root*>(voidSelf);
    Vtop__Syms* const __restrict vlSymsp VL_ATTR_UNUSED = vlSelf->vlSymsp;
    if (!vlSymsp->_vm_contextp__->calcUnusedSigs()) {
        VL_FATAL_MT(__FILE__, __LINE__, __FILE__,
                    "Turning on wave traces requires Verilated::traceEverOn(true) call before time 0.");
    }
    vlSymsp->__Vm_baseCode = code;
    tracep->pushPrefix(vlSymsp->name(), VerilatedTracePrefixType::SCOPE_MODULE);
    trace_init_top(vlSelf, tracep);
    tracep->popPrefix();
}

static void trace_full_0(void* voidSelf, VerilatedVcd::Buffer* bufp) {
    Vtop___024root* const __restrict vlSelf VL_ATTR_UNUSED = static_cast<Vtop___024root*>(voidSelf);
    Vtop__Syms* const __restrict vlSymsp VL_ATTR_UNUSED = vlSelf->vlSymsp;
    auto& vlSelfRef = std::ref(*vlSelf).get();
    uint32_t* const oldp VL_ATTR_UNUSED = bufp->oldp(vlSymsp->__Vm_baseCode);
    bufp->fullBit(oldp + 1, (vlSelfRef.clk));
    bufp->fullBit(oldp + 2, (vlSelfRef.reset_l));
    bufp->fullCData(oldp + 3, (vlSelfRef.out_small), 2);
    bufp->fullQData(oldp + 4, (vlSelfRef.out_quad), 40);
    bufp->fullWData(oldp + 6, (vlSelfRef.out_wide), 70);
    bufp->fullCData(oldp + 9, (vlSelfRef.in_small), 2);
    bufp->fullQData(oldp + 10, (vlSelfRef.in_quad), 40);
    bufp->fullWData(oldp + 12, (vlSelfRef.in_wide), 70);
    bufp->fullIData(oldp + 15, (vlSelfRef.top__DOT__sub__DOT__count_c), 32);
}
static void trace_chg_0(void* voidSelf, VerilatedVcd::Buffer* bufp) {
    Vtop___024root* const __restrict vlSelf VL_ATTR_UNUSED = static_cast<Vtop___024root*>(voidSelf);
    Vtop__Syms* const __restrict vlSymsp VL_ATTR_UNUSED = vlSelf->vlSymsp;
    auto& vlSelfRef = std::ref(*vlSelf).get();
    if (VL_UNLIKELY(!vlSymsp->__Vm_activity)) return;
    uint32_t* const oldp VL_ATTR_UNUSED = bufp->oldp(vlSymsp->__Vm_baseCode + 1);
    bufp->chgBit(oldp + 0, (vlSelfRef.clk));
    bufp->chgBit(oldp + 1, (vlSelfRef.reset_l));
    bufp->chgCData(oldp + 2, (vlSelfRef.out_small), 2);
    bufp->chgQData(oldp + 3, (vlSelfRef.out_quad), 40);
    bufp->chgWData(oldp + 5, (vlSelfRef.out_wide), 70);
    bufp->chgCData(oldp + 8, (vlSelfRef.in_small), 2);
    bufp->chgQData(oldp + 9, (vlSelfRef.in_quad), 40);
    bufp->chgWData(oldp + 11, (vlSelfRef.in_wide), 70);
    // count_c only changes in the 'nba' region: gated by its activity flag (V3Trace.cpp:17-35)
    if (VL_UNLIKELY(vlSymsp->__Vm_traceActivity)) {
        bufp->chgIData(oldp + 14, (vlSelfRef.top__DOT__sub__DOT__count_c), 32);
    }
}
static void trace_cleanup(void* voidSelf, VerilatedVcd* /*unused*/) {
    Vtop___024root* const __restrict vlSelf VL_ATTR_UNUSED = static_cast<Vtop___024root*>(voidSelf);
    Vtop__Syms* const __restrict vlSymsp VL_ATTR_UNUSED = vlSelf->vlSymsp;
    vlSymsp->__Vm_activity = false;
    vlSymsp->__Vm_traceActivity = false;
}
static void trace_register(Vtop___024root* vlSelf, VerilatedVcd* tracep) {
    tracep->addFullCb(&trace_full_0, 0, vlSelf);
    tracep->addChgCb(&trace_chg_0, 0, vlSelf);
    tracep->addCleanupCb(&trace_cleanup, vlSelf);
}

//============================================================
// Model class (src/V3EmitCModel.cpp:275-359, 403-456)

Vtop::Vtop(VerilatedContext* _vcontextp__, const char* _vcname__)
    : VerilatedModel{*_vcontextp__}
    , vlSymsp{new Vtop__Syms(contextp(), _vcname__, this)}
    , clk{vlSymsp->TOP.clk}
    , reset_l{vlSymsp->TOP.reset_l}
    , out_small{vlSymsp->TOP.out_small}
    , in_small{vlSymsp->TOP.in_small}
    , out_quad{vlSymsp->TOP.out_quad}
    , in_quad{vlSymsp->TOP.in_quad}
    , out_wide{vlSymsp->TOP.out_wide}
    , in_wide{vlSymsp->TOP.in_wide}
    , rootp{&(vlSymsp->TOP)} {
    // Register model with the context
    contextp()->addModel(this);
    contextp()->traceBaseModelCbAdd(
        [this](VerilatedTraceBaseC* tfp, int levels, int options) { traceBaseModel(tfp, levels, options); });
}
Vtop::Vtop(const char* _vcname__)
    : Vtop(Verilated::threadContextp(), _vcname__) {}
Vtop::~Vtop() { delete vlSymsp; }

void Vtop::eval_step() {
    vlSymsp->__Vm_activity = true;
    if (VL_UNLIKELY(!vlSymsp->__Vm_didInit)) {
        Vtop___024root___eval_static(&(vlSymsp->TOP));
        Vtop___024root___eval_initial(&(vlSymsp->TOP));
        Vtop___024root___eval_settle(&(vlSymsp->TOP));
        vlSymsp->__Vm_didInit = true;
    }
    Vtop___024root___eval(&(vlSymsp->TOP));
    // Evaluate cleanup
    Verilated::endOfEval(vlSymsp->__Vm_evalMsgQp);
}
uint64_t Vtop::nextTimeSlot() {
    VL_FATAL_MT(__FILE__, __LINE__, "", "No delays in the design");
    return 0;
}
const char* Vtop::name() const { return vlSymsp->name(); }
void Vtop::final() {
    contextp()->executingFinal(true);
    Vtop___024root___eval_final(&(vlSymsp->TOP));
    contextp()->executingFinal(false);
}
const char* Vtop::hierName() const { return vlSymsp->name(); }
const char* Vtop::modelName() const { return "Vtop"; }
unsigned Vtop::threads() const { return 1; }
std::unique_ptr<VerilatedTraceConfig> Vtop::traceConfig() const {
    return std::unique_ptr<VerilatedTraceConfig>{new VerilatedTraceConfig{false}};
}
void Vtop::traceBaseModel(VerilatedTraceBaseC* tfp, int levels, int options) {
    (void)levels;
    (void)options;
    VerilatedVcdC* const stfp = dynamic_cast<VerilatedVcdC*>(tfp);
    if (VL_UNLIKELY(!stfp)) {
        vl_fatal(__FILE__, __LINE__, __FILE__, "'Vtop::trace()' called on non-VerilatedVcdC object;");
    }
    stfp->spTrace()->addModel(this);
    stfp->spTrace()->addInitCb(&trace_init, &(vlSymsp->TOP), name(), false, 16);
    trace_register(&(vlSymsp->TOP), stfp->spTrace());
}
