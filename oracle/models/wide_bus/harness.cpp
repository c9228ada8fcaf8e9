// TEST INFRASTRUCTURE (not product code) -- oracle-side testbench harness for wide_bus.v.
//
// Runs N independent instances of the GENUINE Verilated model (oracle/_ref/gen/wide_bus/
// Vwide_bus*.cpp, written by the reference's own scheduler and emitter through
// oracle/_ref/verilator_lite, linked with the reference's runtime), one VerilatedContext +
// Vwide_bus per instance, with the stimulus of the device testbench TbWideRand
// (verilator_b200/csrc/designs/wide_bus/wide_tb.h).
#include "Vwide_bus.h"
#include "Vwide_bus___024root.h"
#include "verilated_vcd_c.h"
#include "ref_timing.h"

#include <atomic>
#include <memory>
#include <thread>
#include <vector>

#include "wide_stim.h"

namespace {
// the signature of include/verilated_cuda.h (vlcu_sig_word / vlcu_sig_mix)
inline uint32_t sigWord(uint32_t acc, uint32_t word) { return (acc ^ word) * 0x01000193U; }
inline uint64_t sigMix(uint64_t sig, uint32_t acc) { return (sig ^ acc) * 0x100000001b3ULL; }

constexpr int kOutWords = 19;

struct WideParams {
    uint64_t n, instBase, seedBase;
    int randReset;
    uint64_t nEvals, tbSeed;
    uint32_t resetTime;
    uint32_t* obs;  // [eval][19][n]: out_valid out_lt out_lts out_ones slice[3] hash[4] data[8]
    uint64_t obsEvals;
    uint64_t* sigs;
    uint32_t* finalAcc;  // [n][8]
    const char* vcdPath;
    int64_t vcdInstance;
};

void outputs(const Vwide_bus& top, uint32_t* w) {
    w[0] = top.out_valid, w[1] = top.out_lt, w[2] = top.out_lts, w[3] = top.out_ones;
    for (int k = 0; k < 3; ++k) w[4 + k] = top.out_slice[k];
    for (int k = 0; k < 4; ++k) w[7 + k] = top.out_hash[k];
    for (int k = 0; k < 8; ++k) w[11 + k] = top.out_data[k];
}

void runOne(const WideParams& p, uint64_t i) {
    const std::unique_ptr<VerilatedContext> contextp{new VerilatedContext};
    Verilated::threadContextp(contextp.get());
    contextp->randReset(p.randReset);
    const uint64_t instSeed = p.seedBase ^ (p.instBase + i);
    contextp->randSeed(static_cast<int>(instSeed));
    contextp->traceEverOn(true);
    const char* argv[] = {"oracle"};
    contextp->commandArgs(1, argv);
    const std::unique_ptr<Vwide_bus> top{new Vwide_bus{contextp.get(), "TOP"}};
    std::unique_ptr<VerilatedVcdC> tfp;
    if (p.vcdPath && p.vcdInstance == static_cast<int64_t>(i)) {
        tfp.reset(new VerilatedVcdC);
        top->trace(tfp.get(), 99);
        tfp->open(p.vcdPath);
    }
    top->clk = 0;
    top->rst_n = 0;
    top->in_valid = 0;
    top->in_op = 0;
    top->in_shift = 0;
    for (int w = 0; w < 8; ++w) top->in_data[w] = 0;
    const uint64_t resetTime = p.resetTime ? p.resetTime : 10;
    uint64_t sig = 0xcbf29ce484222325ULL;
    {
        const VlrefLoopTimer timer;
        for (uint64_t e = 0; e < p.nEvals; ++e) {
            contextp->timeInc(1);
            top->clk = !top->clk;
            if (!top->clk) {
                top->rst_n = contextp->time() > resetTime;
                // the device context's randSeed() is the 64-bit seed_base ^ global index
                wide_drive_inputs(*top, p.tbSeed ^ instSeed, e);
            }
            top->eval();
            if (tfp) tfp->dump(contextp->time());
            uint32_t w[kOutWords];
            outputs(*top, w);
            uint32_t acc = 0x811c9dc5U;
            for (int k = 0; k < kOutWords; ++k) acc = sigWord(acc, w[k]);
            sig = sigMix(sig, acc);
            if (p.obs && e < p.obsEvals) {
                uint32_t* const rowp = p.obs + e * kOutWords * p.n;
                for (int k = 0; k < kOutWords; ++k) rowp[k * p.n + i] = w[k];
            }
        }
    }
    top->final();
    if (tfp) tfp->close();
    if (p.sigs) p.sigs[i] = sig;
    if (p.finalAcc) {
        for (int k = 0; k < 8; ++k) p.finalAcc[i * 8 + k] = top->rootp->wide_bus__DOT__acc[k];
    }
}
}  // namespace

extern "C" int vlref_wide_run(uint64_t n, uint64_t instBase, uint64_t seedBase, int randReset,
                              uint64_t nEvals, uint64_t tbSeed, uint32_t resetTime, uint32_t* obs,
                              uint64_t obsEvals, uint64_t* sigs, uint32_t* finalAcc,
                              const char* vcdPath, int64_t vcdInstance, int threads) {
    const WideParams p{n, instBase, seedBase, randReset, nEvals, tbSeed, resetTime, obs,
                       obsEvals, sigs, finalAcc, vcdPath, vcdInstance};
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
