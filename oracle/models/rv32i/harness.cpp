// TEST INFRASTRUCTURE (not product code) -- oracle-side testbench harness for rv32_soc.v.
//
// Runs N independent instances of the GENUINE Verilated model (oracle/_ref/gen/rv32i/
// Vrv32_soc*.cpp, written by the reference's own scheduler and emitter through
// oracle/_ref/verilator_lite, linked with the reference's runtime), one VerilatedContext +
// Vrv32_soc per instance (docs/guide/verilating.rst:232-236), with the stimulus of the device
// testbenches TbRv32Rand / TbRv32Host (verilator_b200/csrc/designs/rv32i/rv32_tb.h).
// Compiled twice: as is for the plain model (vlref_rv32_run), and with
//   -DRV32_COV=1 for the --coverage build Vrv32_cov (vlref_rv32cov_run), which also returns the
//   per-bin sums of the instances' coverage counters and can write one instance's coverage.dat
//   through the reference's own writer.
#ifdef RV32_COV
#include "Vrv32_cov.h"
#include "Vrv32_cov___024root.h"
#include "verilated_cov.h"
#define Vrv32_soc Vrv32_cov
#define Vrv32_soc___024root Vrv32_cov___024root
#define vlref_rv32_run vlref_rv32cov_run
#elif defined(RV32_THR)
// the --threads 4 build: ONE instance spread over 4 host threads (SURVEY 8(d) CPU baseline ii)
#include "Vrv32_thr.h"
#include "Vrv32_thr___024root.h"
#define Vrv32_soc Vrv32_thr
#define Vrv32_soc___024root Vrv32_thr___024root
#define vlref_rv32_run vlref_rv32thr_run
#else
#include "Vrv32_soc.h"
#include "Vrv32_soc___024root.h"
#include "verilated_save.h"
#endif
#include "verilated_vcd_c.h"
#include "ref_timing.h"

#include <atomic>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "rv32_program.h"

//=============================================================================
// Harness: N instances, the stimulus of the device testbenches TbRv32Rand / TbRv32Host

namespace {
// the signature of include/verilated_cuda.h (vlcu_sig_word / vlcu_sig_mix)
inline uint32_t sigWord(uint32_t acc, uint32_t word) { return (acc ^ word) * 0x01000193U; }
inline uint64_t sigMix(uint64_t sig, uint32_t acc) { return (sig ^ acc) * 0x100000001b3ULL; }

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
    const char* vcdPath;  // waveform of instance vcdInstance through the reference's VerilatedVcdC
    int64_t vcdInstance;
    uint64_t* covSums;  // (RV32_COV) [bins] summed over the instances, or null
    const char* covPath;  // (RV32_COV) coverage.dat of instance covInstance
    int64_t covInstance;
    uint32_t sameProgram;  // tb flag RV32_TBF_SAME_PROGRAM: the program seed ignores the instance index
};
#ifdef RV32_COV
constexpr int kCovBins = sizeof(Vrv32_soc___024root::__Vcoverage) / sizeof(uint32_t);
std::atomic<uint64_t> g_cov[kCovBins];
#endif

void runOne(const Rv32Params& p, uint64_t i) {
    const std::unique_ptr<VerilatedContext> contextp{new VerilatedContext};
    Verilated::threadContextp(contextp.get());
    contextp->randReset(p.randReset);
    contextp->randSeed(static_cast<int>(p.seedBase ^ (p.instBase + i)));
    const char* argv[] = {"oracle"};
    contextp->commandArgs(1, argv);
    contextp->traceEverOn(true);
#ifdef RV32_THR
    contextp->threads(4);
#endif
    const std::unique_ptr<Vrv32_soc> top{new Vrv32_soc{contextp.get(), "TOP"}};
    std::unique_ptr<VerilatedVcdC> tfp;
#ifndef RV32_THR
    if (p.vcdPath && p.vcdInstance == static_cast<int64_t>(i)) {
        tfp.reset(new VerilatedVcdC);
        top->trace(tfp.get(), 99);
        tfp->open(p.vcdPath);
    }
#endif
    const uint64_t resetTime = p.resetTime ? p.resetTime : 16;
    uint64_t sig = 0xcbf29ce484222325ULL;
    const uint32_t rounds = p.rounds ? p.rounds : 1;
    for (uint32_t round = 0; round < rounds; ++round) {
        top->clk = 0;
        top->resetn = 0;
        if (p.memImages) {
            for (int w = 0; w < 1024; ++w) top->rootp->rv32_soc__DOT__mem[w] = p.memImages[i * 1024 + w];
        } else {
            rv32_gen_program(top->rootp->rv32_soc__DOT__mem,
                             p.sameProgram ? (p.tbSeed + round) : ((p.tbSeed + round) ^ (p.instBase + i)));
        }
        const VlrefLoopTimer timer;
        for (uint64_t e = 0; e < p.nEvals; ++e) {
            contextp->timeInc(1);
            top->clk = !top->clk;
            top->resetn = p.rounds ? (e >= resetTime) : (contextp->time() > resetTime);
            top->eval();
            if (tfp) tfp->dump(contextp->time());
            // every output port in source order, as the emitted outputs_sig() does
            uint32_t acc = 0x811c9dc5U;
            acc = sigWord(acc, top->trap);
            acc = sigWord(acc, top->mem_valid);
            acc = sigWord(acc, top->mem_instr);
            acc = sigWord(acc, top->mem_addr);
            acc = sigWord(acc, top->mem_wdata);
            acc = sigWord(acc, top->mem_wstrb);
            acc = sigWord(acc, top->mem_rdata);
            acc = sigWord(acc, top->retired);
            sig = sigMix(sig, acc);
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
    if (tfp) tfp->close();
#ifdef RV32_COV
    if (p.covSums) {
        for (int b = 0; b < kCovBins; ++b) g_cov[b].fetch_add(top->rootp->__Vcoverage[b], std::memory_order_relaxed);
    }
    if (p.covPath && p.covInstance == static_cast<int64_t>(i)) contextp->coveragep()->write(p.covPath);
#endif
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
                              uint32_t rounds, const char* vcdPath, int64_t vcdInstance,
                              uint64_t* covSums, const char* covPath, int64_t covInstance,
                              uint32_t sameProgram) {
#ifdef RV32_COV
    for (int b = 0; b < kCovBins; ++b) g_cov[b] = 0;
#endif
    const Rv32Params p{n, instBase, seedBase, randReset, nEvals, tbSeed, resetTime, memImages,
                       obs, obsEvals, sigs, times, finalRegs, finalMem, finalScalars, rounds, vcdPath, vcdInstance,
                       covSums, covPath, covInstance, sameProgram};
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
#ifdef RV32_COV
    if (covSums) {
        for (int b = 0; b < kCovBins; ++b) covSums[b] = g_cov[b];
    }
#endif
    return 0;
}

#ifdef RV32_COV
extern "C" int vlref_rv32cov_bins() { return kCovBins; }
#elif defined(RV32_THR)
#else
// Save / restore through the reference's own verilated_save (the model is built --savable):
// one instance, either constructed like instance `globalIndex` of vlref_rv32_run or restored
// from `restorePath`; runs nEvals more evals of the TbRv32Rand stimulus (resetn by time) and
// optionally saves the model after the last one.
extern "C" int vlref_rv32_saverun(uint64_t globalIndex, uint64_t seedBase, int randReset,
                                  uint64_t tbSeed, uint32_t resetTime, const char* restorePath,
                                  uint64_t nEvals, const char* savePath, uint32_t* finalRegs,
                                  uint32_t* finalScalars, uint32_t* finalMem, uint64_t* timeOut) {
    const std::unique_ptr<VerilatedContext> contextp{new VerilatedContext};
    Verilated::threadContextp(contextp.get());
    contextp->randReset(randReset);
    contextp->randSeed(static_cast<int>(seedBase ^ globalIndex));
    const char* argv[] = {"oracle"};
    contextp->commandArgs(1, argv);
    const std::unique_ptr<Vrv32_soc> top{new Vrv32_soc{contextp.get(), "TOP"}};
    if (restorePath) {
        VerilatedRestore os;
        os.open(restorePath);
        os >> *top;
        os.close();
    } else {
        top->clk = 0;
        top->resetn = 0;
        rv32_gen_program(top->rootp->rv32_soc__DOT__mem, tbSeed ^ globalIndex);
    }
    const uint64_t resetAt = resetTime ? resetTime : 16;
    for (uint64_t e = 0; e < nEvals; ++e) {
        contextp->timeInc(1);
        top->clk = !top->clk;
        top->resetn = contextp->time() > resetAt;
        top->eval();
    }
    if (savePath) {
        VerilatedSave os;
        os.open(savePath);
        os << *top;
        os.close();
    }
    const Vrv32_soc___024root& r = *top->rootp;
    if (finalRegs) {
        for (int k = 0; k < 32; ++k) finalRegs[k] = r.rv32_soc__DOT__regs[k];
    }
    if (finalMem) {
        for (int k = 0; k < 1024; ++k) finalMem[k] = r.rv32_soc__DOT__mem[k];
    }
    if (finalScalars) {
        uint32_t* const s = finalScalars;
        s[0] = r.rv32_soc__DOT__pc, s[1] = r.rv32_soc__DOT__insn, s[2] = r.rv32_soc__DOT__ld_addr;
        s[3] = r.rv32_soc__DOT__instret, s[4] = r.rv32_soc__DOT__state, s[5] = r.trap;
        s[6] = r.retired, s[7] = r.clk;
    }
    if (timeOut) *timeOut = contextp->time();
    return 0;
}
// The serialized part of a freshly constructed VerilatedContext (verilated.h:402-441,777-778)
extern "C" int vlref_context_image(void* out, size_t outBytes) {
    VerilatedContext context;
    const size_t n = VerilatedContext::serialized1Size();
    if (outBytes < n) return -1;
    std::memcpy(out, context.serialized1Ptr(), n);
    return static_cast<int>(n);
}
// The program generator alone (for the instruction-set-simulator cross check)
extern "C" void vlref_rv32_gen_program(uint64_t seed, uint32_t* out1024) {
    VlUnpacked<IData, 1024> mem;
    rv32_gen_program(mem, seed);
    for (int k = 0; k < 1024; ++k) out1024[k] = mem[k];
}
#endif
