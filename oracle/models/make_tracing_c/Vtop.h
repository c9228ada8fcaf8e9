// TEST INFRASTRUCTURE (not product code) -- design-level ORACLE, *restated*.
//
// The reference's generator (verilator_bin) cannot be built in this image
// (bison/flex absent), so this is a hand restatement of the C++ that
//   verilator -cc -x-assign fast --trace-vcd --assert --coverage top.v
// emits for examples/make_tracing_c, following the generator sources:
//   model class / eval_step   src/V3EmitCModel.cpp:55-269,275-359,403-437
//   root class members        src/V3EmitCHeaders.cpp:73-143
//   symbol table              src/V3EmitCSyms.cpp:923-982
// It is compiled against the GENUINE runtime (include/verilated*.cpp under
// /root/reference), so VL_* operators, VerilatedContext, the coverage writer
// and the VCD writer are the reference's own.  Label: "restated reference".
#ifndef VERILATED_VTOP_H_
#define VERILATED_VTOP_H_

#include "verilated.h"
#include "verilated_cov.h"

class Vtop__Syms;
class Vtop___024root;
class VerilatedVcdC;

class alignas(VL_CACHE_LINE_BYTES) Vtop VL_NOT_FINAL : public VerilatedModel {
private:
    Vtop__Syms* const vlSymsp;

public:
    // PORTS (V3EmitCModel.cpp:108-117)
    VL_IN8(&clk, 0, 0);
    VL_IN8(&reset_l, 0, 0);
    VL_OUT8(&out_small, 1, 0);
    VL_IN8(&in_small, 1, 0);
    VL_OUT64(&out_quad, 39, 0);
    VL_IN64(&in_quad, 39, 0);
    VL_OUTW(&out_wide, 69, 0, 3);
    VL_INW(&in_wide, 69, 0, 3);
    Vtop___024root* const rootp;

    explicit Vtop(VerilatedContext* contextp, const char* name = "TOP");
    explicit Vtop(const char* name = "TOP");
    virtual ~Vtop();

private:
    VL_UNCOPYABLE(Vtop);

public:
    void eval() { eval_step(); }
    void eval_step();
    void eval_end_step() {}
    void final();
    bool eventsPending() { return false; }
    uint64_t nextTimeSlot();
    void trace(VerilatedTraceBaseC* tfp, int levels, int options = 0) {
        contextp()->trace(tfp, levels, options);
    }
    const char* name() const;
    const char* hierName() const override final;
    const char* modelName() const override final;
    unsigned threads() const override final;
    std::unique_ptr<VerilatedTraceConfig> traceConfig() const override final;

private:
    void traceBaseModel(VerilatedTraceBaseC* tfp, int levels, int options);
};

#endif  // guard
