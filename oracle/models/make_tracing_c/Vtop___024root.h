// TEST INFRASTRUCTURE (not product code) -- design-level ORACLE, *restated*.
// Root module class and symbol table as Verilator declares them
// (src/V3EmitCHeaders.cpp:73-143 members; src/V3EmitCSyms.cpp:923-982 syms).
#ifndef VERILATED_VTOP___024ROOT_H_
#define VERILATED_VTOP___024ROOT_H_

#include "verilated.h"
#include "verilated_cov.h"

class Vtop;
class Vtop__Syms;

class alignas(VL_CACHE_LINE_BYTES) Vtop___024root final {
public:
    // DESIGN SPECIFIC STATE
    VL_IN8(clk, 0, 0);
    VL_IN8(reset_l, 0, 0);
    VL_OUT8(out_small, 1, 0);
    VL_IN8(in_small, 1, 0);
    CData /*0:0*/ top__DOT____Vtogcov__clk;
    CData /*0:0*/ top__DOT____Vtogcov__reset_l;
    CData /*1:0*/ top__DOT____Vtogcov__out_small;
    CData /*1:0*/ top__DOT____Vtogcov__in_small;
    CData /*0:0*/ top__DOT__sub__DOT____Vtogcov__clk;
    CData /*0:0*/ top__DOT__sub__DOT____Vtogcov__reset_l;
    CData /*0:0*/ __VstlFirstIteration;
    CData /*0:0*/ __VstlPhaseResult;
    CData /*0:0*/ __VstlExecute;
    CData /*0:0*/ __VicoFirstIteration;
    CData /*0:0*/ __VicoPhaseResult;
    CData /*0:0*/ __VicoExecute;
    CData /*0:0*/ __VicoDidInit;
    CData /*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData /*0:0*/ __Vtrigprevexpr___TOP__reset_l__0;
    CData /*1:0*/ __Vtrigprevexpr___TOP__in_small__0;
    CData /*0:0*/ __Vtrigprevexpr___TOP__clk__1;
    CData /*0:0*/ __VactFirstIteration;
    CData /*0:0*/ __VactPhaseResult;
    CData /*0:0*/ __VactExecute;
    CData /*0:0*/ __VnbaFirstIteration;
    CData /*0:0*/ __VnbaPhaseResult;
    CData /*0:0*/ __VnbaExecute;
    IData /*31:0*/ top__DOT__sub__DOT__count_c;
    IData /*31:0*/ top__DOT__sub__DOT____Vtogcov__count_c;
    IData /*31:0*/ __VstlIterCount;
    IData /*31:0*/ __VicoIterCount;
    IData /*31:0*/ __VactIterCount;
    IData /*31:0*/ __VnbaIterCount;
    VL_OUT64(out_quad, 39, 0);
    VL_IN64(in_quad, 39, 0);
    QData /*39:0*/ top__DOT____Vtogcov__out_quad;
    QData /*39:0*/ top__DOT____Vtogcov__in_quad;
    QData /*39:0*/ __Vtrigprevexpr___TOP__in_quad__0;
    VL_OUTW(out_wide, 69, 0, 3);
    VL_INW(in_wide, 69, 0, 3);
    VlWide<3> /*69:0*/ top__DOT____Vtogcov__out_wide;
    VlWide<3> /*69:0*/ top__DOT____Vtogcov__in_wide;
    VlWide<3> /*69:0*/ __Vtrigprevexpr___TOP__in_wide__0;
    VlUnpacked<QData /*63:0*/, 1> __VstlTriggered;
    VlUnpacked<QData /*63:0*/, 2> __VicoTriggered;
    VlUnpacked<QData /*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData /*63:0*/, 1> __VnbaTriggered;

    // INTERNAL VARIABLES
    Vtop__Syms* vlSymsp;
    const char* vlNamep;
    uint32_t __Vcoverage[527];

    // CONSTRUCTORS
    Vtop___024root(Vtop__Syms* symsp, const char* namep);
    ~Vtop___024root();
    VL_UNCOPYABLE(Vtop___024root);

    // INTERNAL METHODS
    void __Vconfigure(bool first);
    void __vlCoverInsert(uint32_t* countp, bool enable, bool localCounter, const char* filenamep,
                         int lineno, int column, const char* hierp, const char* pagep,
                         const char* commentp, const char* linescovp);
    void __vlCoverToggleInsert(int begin, int end, bool ranged, uint32_t* countp, bool enable,
                               bool localCounter, const char* filenamep, int lineno, int column,
                               const char* hierp, const char* pagep, const char* commentp);
};

void Vtop___024root___ctor_var_reset(Vtop___024root* vlSelf);
void Vtop___024root___eval_static(Vtop___024root* vlSelf);
void Vtop___024root___eval_initial(Vtop___024root* vlSelf);
void Vtop___024root___eval_settle(Vtop___024root* vlSelf);
void Vtop___024root___eval(Vtop___024root* vlSelf);
void Vtop___024root___eval_final(Vtop___024root* vlSelf);

class alignas(VL_CACHE_LINE_BYTES) Vtop__Syms final : public VerilatedSyms {
public:
    // INTERNAL STATE
    Vtop* const __Vm_modelp;
    bool __Vm_activity = false;  ///< Used by trace routines to determine change occurred
    bool __Vm_traceActivity = false;  ///< 'nba' region activity flag (V3Trace.cpp:17-35)
    uint32_t __Vm_baseCode = 0;  ///< Used by trace routines when tracing multiple models
    bool __Vm_didInit = false;

    // MODULE INSTANCE STATE
    Vtop___024root TOP;

    // CONSTRUCTORS
    Vtop__Syms(VerilatedContext* contextp, const char* namep, Vtop* modelp);
    ~Vtop__Syms();

    // METHODS
    const char* name() const override { return TOP.vlNamep; }
};

#endif  // guard
