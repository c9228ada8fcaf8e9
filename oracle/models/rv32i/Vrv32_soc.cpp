// TEST INFRASTRUCTURE (not product code) -- design-level ORACLE for rv32_soc.v, *restated*.
//
// Host model of verilator_b200/csrc/designs/rv32i/rv32_soc.v in the form V3EmitCFunc /
// V3EmitCModel / V3EmitCSyms print, compiled against the GENUINE runtime (VL_* operators,
// VerilatedContext, VL_SCOPED_RAND_RESET_*, VlRNG from /root/reference/include).  The scheduled
// function bodies come from the shared hand-scheduled "AST print" Vrv32_soc__body.inc (see the
// header of that file for why one body serves both printers); everything around it -- class
// layout, construction/reset, eval_step, the testbench loop -- is restated here from:
//   src/V3EmitCHeaders.cpp:73-143, src/V3EmitCImp.cpp:91-152, src/V3EmitCModel.cpp:275-437,
//   src/V3EmitCFunc.cpp:524-677.
// The design logic is pinned independently by tests/rv32_iss.py.
#include "verilated.h"
#include "ref_timing.h"

#include <atomic>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "ref_events.h"

class Vrv32_soc__Syms;

class alignas(VL_CACHE_LINE_BYTES) Vrv32_soc___024root final {
public:
    VL_IN8(clk, 0, 0);
    VL_IN8(resetn, 0, 0);
    VL_OUT8(trap, 0, 0);
    VL_OUT8(mem_valid, 0, 0);
    VL_OUT8(mem_instr, 0, 0);
    VL_OUT8(mem_wstrb, 3, 0);
    CData /*1:0*/ rv32_soc__DOT__state;
    CData /*0:0*/ __VstlFirstIteration, __VstlPhaseResult, __VstlExecute;
    CData /*0:0*/ __Vtrigprevexpr___TOP__clk__0;
    CData /*0:0*/ __VactFirstIteration, __VactPhaseResult, __VactExecute;
    CData /*0:0*/ __VnbaFirstIteration, __VnbaPhaseResult, __VnbaExecute;
    VL_OUT(mem_addr, 31, 0);
    VL_OUT(mem_wdata, 31, 0);
    VL_OUT(mem_rdata, 31, 0);
    VL_OUT(retired, 31, 0);
    IData /*31:0*/ rv32_soc__DOT__pc;
    IData /*31:0*/ rv32_soc__DOT__insn;
    IData /*31:0*/ rv32_soc__DOT__ld_addr;
    IData /*31:0*/ rv32_soc__DOT__instret;
    IData /*31:0*/ __VstlIterCount, __VactIterCount, __VnbaIterCount;
    VlUnpacked<IData /*31:0*/, 1024> rv32_soc__DOT__mem;
    VlUnpacked<IData /*31:0*/, 32> rv32_soc__DOT__regs;
    VlUnpacked<QData /*63:0*/, 1> __VstlTriggered;
    VlUnpacked<QData /*63:0*/, 1> __VactTriggered;
    VlUnpacked<QData /*63:0*/, 1> __VnbaTriggered;

    Vrv32_soc__Syms* vlSymsp;
    const char* vlNamep;

    Vrv32_soc___024root(Vrv32_soc__Syms* symsp, const char* namep);
    ~Vrv32_soc___024root() { VL_DO_DANGLING(free(const_cast<char*>(vlNamep)), vlNamep); }
    VL_UNCOPYABLE(Vrv32_soc___024root);
};

class alignas(VL_CACHE_LINE_BYTES) Vrv32_soc__Syms final : public VerilatedSyms {
public:
    bool __Vm_didInit = false;
    Vrv32_soc___024root TOP;
    Vrv32_soc__Syms(VerilatedContext* contextp, const char* namep)
        : VerilatedSyms{contextp}
        , TOP{this, namep} {
        _vm_contextp__->timeunit(-12);
        _vm_contextp__->timeprecision(-12);
    }
    const char* name() const override { return TOP.vlNamep; }
};

// _ctor_var_reset: --x-initial unique; one (scope, salt) pair per variable, array elements
// all draw the same value (V3EmitCFunc.cpp:560-575 loops over elements with one salt)
static void Vrv32_soc___024root___ctor_var_reset(Vrv32_soc___024root* vlSelf) {
    auto& vlSelfRef = std::ref(*vlSelf).get();
    const uint64_t __VscopeHash = VL_MURMUR64_HASH(vlSelf->vlNamep);
    vlSelfRef.clk = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 16707436170211756652ull);
    vlSelfRef.resetn = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 8624841754543469506ull);
    vlSelfRef.trap = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 18214934560881419504ull);
    vlSelfRef.mem_valid = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 1551974250180885553ull);
    vlSelfRef.mem_instr = VL_SCOPED_RAND_RESET_I(1, __VscopeHash, 1406578376079151150ull);
    vlSelfRef.mem_addr = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 326597072690670135ull);
    vlSelfRef.mem_wdata = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 5431754401481461448ull);
    vlSelfRef.mem_wstrb = VL_SCOPED_RAND_RESET_I(4, __VscopeHash, 8859681292774497410ull);
    vlSelfRef.mem_rdata = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 9659133473039683418ull);
    vlSelfRef.retired = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 9241213606999339685ull);
    for (int __Vi0 = 0; __Vi0 < 1024; ++__Vi0) {
        vlSelfRef.rv32_soc__DOT__mem[__Vi0] = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 157536494676281617ull);
    }
    for (int __Vi0 = 0; __Vi0 < 32; ++__Vi0) {
        vlSelfRef.rv32_soc__DOT__regs[__Vi0] = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 11523106369459488163ull);
    }
    vlSelfRef.rv32_soc__DOT__pc = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 11841601083557326370ull);
    vlSelfRef.rv32_soc__DOT__insn = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 6790827016312138654ull);
    vlSelfRef.rv32_soc__DOT__ld_addr = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 18411059245981549376ull);
    vlSelfRef.rv32_soc__DOT__instret = VL_SCOPED_RAND_RESET_I(32, __VscopeHash, 1322355925642281443ull);
    vlSelfRef.rv32_soc__DOT__state = VL_SCOPED_RAND_RESET_I(2, __VscopeHash, 769792777424239494ull);
    vlSelfRef.__Vtrigprevexpr___TOP__clk__0 = 0;
    vlSelfRef.__VstlTriggered[0] = 0;
    vlSelfRef.__VactTriggered[0] = 0;
    vlSelfRef.__VnbaTriggered[0] = 0;
}
Vrv32_soc___024root::Vrv32_soc___024root(Vrv32_soc__Syms* symsp, const char* namep) {
    vlSymsp = symsp;
    vlNamep = strdup(namep);
    Vrv32_soc___024root___ctor_var_reset(this);
}

namespace {
#define VLB_FN static inline
#define VLB_ROOT Vrv32_soc___024root
#define VLB_CTXPARAM
#define VLB_FATAL_DIDNOTCONVERGE(region) \
    VL_FATAL_MT("rv32_soc.v", 13, "", "DIDNOTCONVERGE: " region " region did not converge after '--converge-limit' of 10000 tries")
#include "Vrv32_soc__body.inc"
#undef VLB_FN
#undef VLB_ROOT
#undef VLB_CTXPARAM
#undef VLB_FATAL_DIDNOTCONVERGE
}  // namespace

// The model class (src/V3EmitCModel.cpp:55-269, 275-437), reduced to what a testbench uses
class Vrv32_soc final : public VerilatedModel {
    Vrv32_soc__Syms* const vlSymsp;

public:
    CData& clk;
    CData& resetn;
    CData& trap;
    CData& mem_valid;
    CData& mem_instr;
    IData& mem_addr;
    IData& mem_wdata;
    CData& mem_wstrb;
    IData& mem_rdata;
    IData& retired;
    Vrv32_soc___024root* const rootp;
    explicit Vrv32_soc(VerilatedContext* contextp, const char* name = "TOP")
        : VerilatedModel{*contextp}
        , vlSymsp{new Vrv32_soc__Syms(contextp, name)}
        , clk{vlSymsp->TOP.clk}
        , resetn{vlSymsp->TOP.resetn}
        , trap{vlSymsp->TOP.trap}
        , mem_valid{vlSymsp->TOP.mem_valid}
        , mem_instr{vlSymsp->TOP.mem_instr}
        , mem_addr{vlSymsp->TOP.mem_addr}
        , mem_wdata{vlSymsp->TOP.mem_wdata}
        , mem_wstrb{vlSymsp->TOP.mem_wstrb}
        , mem_rdata{vlSymsp->TOP.mem_rdata}
        , retired{vlSymsp->TOP.retired}
        , rootp{&(vlSymsp->TOP)} {
        contextp->addModel(this);
    }
    ~Vrv32_soc() override { delete vlSymsp; }
    void eval() { eval_step(); }
    void eval_step() {
        if (VL_UNLIKELY(!vlSymsp->__Vm_didInit)) {
            rv32_eval_static(vlSymsp->TOP);
            // _eval_initial: no initial blocks
            rv32_eval_settle(vlSymsp->TOP);
            vlSymsp->__Vm_didInit = true;
        }
        rv32_eval(vlSymsp->TOP);
        Verilated::endOfEval(vlSymsp->__Vm_evalMsgQp);
    }
    void final() {}
    const char* hierName() const override { return vlSymsp->name(); }
    const char* modelName() const override { return "Vrv32_soc"; }
    unsigned threads() const override { return 1; }
};

//=============================================================================
// Harness: N instances, the stimulus of the device testbenches TbRv32Rand / TbRv32Host

namespace {
inline uint64_t sigMix(uint64_t sig, uint32_t word) { return (sig ^ word) * 0x100000001b3ULL; }

struct Rv32Params {
    uint64_t n, instBase, seedBase;
    int randReset;
    uint64_t nEvals, tbSeed;
    uint32_t resetTime;
    const uint32_t* memImages;  // [n][1024] or null -> generated from tbSeed ^ global index
    uint32_t* obs;  // [eval][8][n]
    uint64_t obsEvals;
    uint64_t* sigs;
    uint64_t* times;
    uint32_t* finalRegs;  // [n][32]
    uint32_t* finalMem;  // [n][1024]
    uint32_t* finalScalars;  // [n][8]: pc, insn, ld_addr, instret, state, trap, retired, clk
    // Campaign mode (tb flag RV32_TBF_RESET_PER_RUN of the device testbench): `rounds` runs of
    // nEvals evals on the same model; each run begins with clk = resetn = 0, a fresh program
    // from (tbSeed + round) ^ index, and resetn low for the first resetTime evals OF THAT RUN
    uint32_t rounds;
};

void runOne(const Rv32Params& p, uint64_t i) {
    const std::unique_ptr<VerilatedContext> contextp{new VerilatedContext};
    Verilated::threadContextp(contextp.get());
    contextp->randReset(p.randReset);
    contextp->randSeed(static_cast<int>(p.seedBase ^ (p.instBase + i)));
    const char* argv[] = {"oracle"};
    contextp->commandArgs(1, argv);
    const std::unique_ptr<Vrv32_soc> top{new Vrv32_soc{contextp.get(), "TOP"}};
    const uint64_t resetTime = p.resetTime ? p.resetTime : 16;
    uint64_t sig = 0xcbf29ce484222325ULL;
    const uint32_t rounds = p.rounds ? p.rounds : 1;
    for (uint32_t round = 0; round < rounds; ++round) {
        top->clk = 0;
        top->resetn = 0;
        if (p.memImages) {
            for (int w = 0; w < 1024; ++w) top->rootp->rv32_soc__DOT__mem[w] = p.memImages[i * 1024 + w];
        } else {
            rv32_gen_program(top->rootp->rv32_soc__DOT__mem, (p.tbSeed + round) ^ (p.instBase + i));
        }
        const VlrefLoopTimer timer;
        for (uint64_t e = 0; e < p.nEvals; ++e) {
            contextp->timeInc(1);
            top->clk = !top->clk;
            top->resetn = p.rounds ? (e >= resetTime) : (contextp->time() > resetTime);
            top->eval();
            sig = sigMix(sig, (static_cast<uint32_t>(top->trap) << 6) | (static_cast<uint32_t>(top->mem_valid) << 5)
                                  | (static_cast<uint32_t>(top->mem_instr) << 4) | top->mem_wstrb);
            sig = sigMix(sig, top->mem_addr);
            sig = sigMix(sig, top->mem_wdata);
            sig = sigMix(sig, top->mem_rdata);
            sig = sigMix(sig, top->retired);
            const uint64_t ge = round * p.nEvals + e;
            if (p.obs && ge < p.obsEvals) {
                uint32_t* const rowp = p.obs + ge * 8 * p.n;
                rowp[0 * p.n + i] = top->trap;
                rowp[1 * p.n + i] = top->mem_valid;
                rowp[2 * p.n + i] = top->mem_instr;
                rowp[3 * p.n + i] = top->mem_addr;
                rowp[4 * p.n + i] = top->mem_wdata;
                rowp[5 * p.n + i] = top->mem_wstrb;
                rowp[6 * p.n + i] = top->mem_rdata;
                rowp[7 * p.n + i] = top->retired;
            }
        }
    }
    top->final();
    if (p.sigs) p.sigs[i] = sig;
    if (p.times) p.times[i] = contextp->time();
    const Vrv32_soc___024root& r = *top->rootp;
    if (p.finalRegs) {
        for (int k = 0; k < 32; ++k) p.finalRegs[i * 32 + k] = r.rv32_soc__DOT__regs[k];
    }
    if (p.finalMem) {
        for (int k = 0; k < 1024; ++k) p.finalMem[i * 1024 + k] = r.rv32_soc__DOT__mem[k];
    }
    if (p.finalScalars) {
        uint32_t* const s = p.finalScalars + i * 8;
        s[0] = r.rv32_soc__DOT__pc, s[1] = r.rv32_soc__DOT__insn, s[2] = r.rv32_soc__DOT__ld_addr;
        s[3] = r.rv32_soc__DOT__instret, s[4] = r.rv32_soc__DOT__state, s[5] = r.trap;
        s[6] = r.retired, s[7] = r.clk;
    }
}
}  // namespace

extern "C" int vlref_rv32_run(uint64_t n, uint64_t instBase, uint64_t seedBase, int randReset,
                              uint64_t nEvals, uint64_t tbSeed, uint32_t resetTime,
                              const uint32_t* memImages, uint32_t* obs, uint64_t obsEvals,
                              uint64_t* sigs, uint64_t* times, uint32_t* finalRegs,
                              uint32_t* finalMem, uint32_t* finalScalars, int threads,
                              uint32_t rounds) {
    const Rv32Params p{n, instBase, seedBase, randReset, nEvals, tbSeed, resetTime, memImages,
                       obs, obsEvals, sigs, times, finalRegs, finalMem, finalScalars, rounds};
    vlref_loop_seconds_begin();
    if (threads <= 1) {
        for (uint64_t i = 0; i < n; ++i) runOne(p, i);
        vlref_loop_seconds_thread_done();
    } else {
        std::atomic<uint64_t> next{0};
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t) {
            pool.emplace_back([&]() {
                for (;;) {
                    const uint64_t i = next.fetch_add(1);
                    if (i >= n) break;
                    runOne(p, i);
                }
                vlref_loop_seconds_thread_done();
            });
        }
        for (auto& th : pool) th.join();
    }
    return 0;
}

// The program generator alone (for the instruction-set-simulator cross check)
extern "C" void vlref_rv32_gen_program(uint64_t seed, uint32_t* out1024) {
    VlUnpacked<IData, 1024> mem;
    rv32_gen_program(mem, seed);
    for (int k = 0; k < 1024; ++k) out1024[k] = mem[k];
}
