// TEST INFRASTRUCTURE (not product code) -- oracle-side testbench harness.
//
// Runs N independent instances of the GENUINE Verilated model of examples/make_tracing_c
// (oracle/_ref/gen/make_tracing_c/Vtop*.cpp, written by the reference's own V3Sched/V3Order/
// V3EmitC through oracle/_ref/verilator_lite, linked with the reference's runtime) through
// the stimulus of examples/make_tracing_c/sim_main.cpp:52-106, one
// VerilatedContext + Vtop per instance as docs/guide/verilating.rst:232-236
// allows, and records what the GPU path is compared against: every output
// port after every eval, the running output signature, the context flags,
// the coverage counters, and (optionally) the VCD and coverage.dat written
// by the reference's own writers.
#include "Vtop.h"
#include "Vtop___024root.h"
#include "verilated_vcd_c.h"
#include "ref_timing.h"

#include <atomic>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "ref_events.h"

namespace {

struct StatusRec {  // == vlcu_status_t (include/verilated_cuda_abi.h)
    uint64_t time;
    uint64_t first_finish_time, first_stop_time, first_fatal_time;
    uint32_t status;
    uint32_t error_count;
    uint32_t finish_count, stop_count, fatal_count;
    uint32_t first_stop_line, first_fatal_line, fatal_code;
    uint32_t msg_count;
    uint32_t pad_;
};

constexpr int kCovBins = sizeof(Vtop___024root::__Vcoverage) / sizeof(uint32_t);

// the signature of include/verilated_cuda.h (vlcu_sig_word / vlcu_sig_mix)
inline uint32_t sigWord(uint32_t acc, uint32_t word) { return (acc ^ word) * 0x01000193U; }
inline uint64_t sigMix(uint64_t sig, uint32_t acc) { return (sig ^ acc) * 0x100000001b3ULL; }

struct Params {
    uint64_t n, instBase, seedBase;
    int randReset;
    uint64_t nEvals;
    uint32_t runFlags;  // bit1: stop on finish
    uint64_t tbSeed;
    uint32_t tbFlags;  // bit0: per-instance offsets
    uint32_t resetPeriod;
    int tracePlusarg;
    uint32_t* obs;  // [eval][6][n]
    uint64_t obsEvals;
    uint64_t* sigs;
    StatusRec* status;
    uint32_t* finalState;  // [n][4]: count_c, clk, reset_l, in_quad lo (spot checks)
    uint64_t* covSums;  // [vlref_mtc_cov_bins()]
    const char* vcdPath;
    int64_t vcdInstance;
    const char* covPath;  // coverage.dat of instance covInstance via the genuine writer
    int64_t covInstance;
    std::string* console;  // $display text of instance 0
};

void runOne(const Params& p, uint64_t i, std::vector<uint64_t>& covLocal) {
    const std::unique_ptr<VerilatedContext> contextp{new VerilatedContext};
    Verilated::threadContextp(contextp.get());
    VlrefEvents events;
    vlref_set_events(&events);
    std::string console;
    vlref_set_capture(&console);
    contextp->debug(0);
    contextp->threads(1);
    contextp->randReset(p.randReset);
    contextp->randSeed(static_cast<int>(p.seedBase ^ (p.instBase + i)));
    contextp->traceEverOn(true);
    contextp->fatalOnError(false);
    {  // sim_main.cpp:49  contextp->commandArgs(argc, argv)
        const char* argv[] = {"oracle", "+trace"};
        contextp->commandArgs(p.tracePlusarg ? 2 : 1, argv);
    }
    const std::unique_ptr<Vtop> top{new Vtop{contextp.get(), "TOP"}};
    std::unique_ptr<VerilatedVcdC> tfp;
    if (p.vcdPath && p.vcdInstance == static_cast<int64_t>(i)) {
        tfp.reset(new VerilatedVcdC);
        top->trace(tfp.get(), 99);
        tfp->open(p.vcdPath);
    }
    // sim_main.cpp:53-61
    top->reset_l = !0;
    top->clk = 0;
    top->in_small = 1;
    top->in_quad = 0x1234;
    top->in_wide[0] = 0x11111111;
    top->in_wide[1] = 0x22222222;
    top->in_wide[2] = 0x3;
    if (p.tbFlags & 1U) {
        VlRNG rng{p.tbSeed ^ (p.instBase + i)};
        const uint64_t r0 = rng.rand64();
        const uint64_t r1 = rng.rand64();
        const uint64_t r2 = rng.rand64();
        const uint64_t r3 = rng.rand64();
        top->in_small = static_cast<CData>((top->in_small + r0) & 3U);
        top->in_quad = (top->in_quad + r1) & 0x000000ffffffffffULL;
        VlWide<3> off;
        off[0] = static_cast<EData>(r2);
        off[1] = static_cast<EData>(r2 >> 32);
        off[2] = static_cast<EData>(r3) & 0x3fU;
        VlWide<3> sum;
        VL_ADD_W(3, sum, top->in_wide, off);
        top->in_wide[0] = sum[0];
        top->in_wide[1] = sum[1];
        top->in_wide[2] = sum[2] & 0x3fU;
    }
    uint64_t sig = 0xcbf29ce484222325ULL;
    uint32_t msgCount = 0;
    const VlrefLoopTimer* const timerp = new VlrefLoopTimer;
    for (uint64_t e = 0; e < p.nEvals; ++e) {
        if ((p.runFlags & 2U) && contextp->gotFinish()) break;
        contextp->timeInc(1);
        top->clk = !top->clk;
        if (!top->clk) {
            const uint64_t t = p.resetPeriod ? (contextp->time() % p.resetPeriod) : contextp->time();
            if (t > 1 && t < 10) {
                top->reset_l = !1;
            } else {
                top->reset_l = !0;
            }
            top->in_quad = (top->in_quad + 0x12) & 0x000000ffffffffffULL;
        }
        const size_t before = console.size();
        top->eval();
        // one VL_WRITEF_NX -> one VL_PRINTF: count newline-terminated messages by call
        if (console.size() != before) {
            // "*-* All Finished *-*\n" and the [%0t] lines each arrive in one call
            const char* s = console.c_str() + before;
            // messages in this design end with '\n' and contain no other '\n' except the
            // trailing blank line of the two initial-block $displays
            while (*s) {
                const char* nl = strchr(s, '\n');
                if (!nl) break;
                ++msgCount;
                s = nl + 1;
                if (*s == '\n') ++s;  // "...\n\n"
            }
        }
        if (tfp) tfp->dump(contextp->time());
        uint32_t acc = 0x811c9dc5U;
        acc = sigWord(acc, top->out_small);
        acc = sigWord(acc, static_cast<uint32_t>(top->out_quad));
        acc = sigWord(acc, static_cast<uint32_t>(top->out_quad >> 32));
        acc = sigWord(acc, top->out_wide[0]);
        acc = sigWord(acc, top->out_wide[1]);
        acc = sigWord(acc, top->out_wide[2]);
        sig = sigMix(sig, acc);
        if (p.obs && e < p.obsEvals) {
            uint32_t* const rowp = p.obs + e * 6 * p.n;
            rowp[0 * p.n + i] = top->out_small;
            rowp[1 * p.n + i] = static_cast<uint32_t>(top->out_quad);
            rowp[2 * p.n + i] = static_cast<uint32_t>(top->out_quad >> 32);
            rowp[3 * p.n + i] = top->out_wide[0];
            rowp[4 * p.n + i] = top->out_wide[1];
            rowp[5 * p.n + i] = top->out_wide[2];
        }
    }
    delete timerp;
    top->final();
    if (tfp) tfp->close();
    if (p.sigs) p.sigs[i] = sig;
    if (p.status) {
        StatusRec& s = p.status[i];
        memset(&s, 0, sizeof(s));
        s.time = contextp->time();
        s.status = 1U /*DIDINIT*/ | (contextp->gotFinish() ? 2U : 0U) | (contextp->gotError() ? 4U : 0U)
                   | (events.stop_count ? 8U : 0U) | (events.fatal_count ? 16U : 0U);
        if (p.nEvals == 0) s.status &= ~1U;
        s.error_count = contextp->errorCount();
        s.finish_count = static_cast<uint32_t>(events.finish_count);
        s.stop_count = static_cast<uint32_t>(events.stop_count);
        s.fatal_count = static_cast<uint32_t>(events.fatal_count);
        s.first_finish_time = events.first_finish_time;
        s.first_stop_time = events.first_stop_time;
        s.first_fatal_time = events.first_fatal_time;
        s.first_stop_line = events.first_stop_line;
        s.first_fatal_line = events.first_fatal_line;
        s.msg_count = msgCount;
    }
    if (p.finalState) {
        p.finalState[i * 4 + 0] = top->rootp->top__DOT__sub__DOT__count_c;
        p.finalState[i * 4 + 1] = top->clk;
        p.finalState[i * 4 + 2] = top->reset_l;
        p.finalState[i * 4 + 3] = static_cast<uint32_t>(top->in_quad);
    }
    for (int b = 0; b < kCovBins; ++b) covLocal[b] += top->rootp->__Vcoverage[b];
    if (p.covPath && p.covInstance == static_cast<int64_t>(i)) {
        contextp->coveragep()->write(p.covPath);
    }
    if (p.console && i == 0) *p.console = console;
    vlref_set_events(nullptr);
    vlref_set_capture(nullptr);
}

}  // namespace

extern "C" int vlref_mtc_cov_bins() { return kCovBins; }

extern "C" int vlref_mtc_run(uint64_t n, uint64_t instBase, uint64_t seedBase, int randReset,
                             uint64_t nEvals, uint32_t runFlags, uint64_t tbSeed, uint32_t tbFlags,
                             uint32_t resetPeriod, int tracePlusarg, uint32_t* obs,
                             uint64_t obsEvals, uint64_t* sigs, void* status, uint32_t* finalState,
                             uint64_t* covSums, const char* vcdPath, int64_t vcdInstance,
                             const char* covPath, int64_t covInstance, char* consoleOut,
                             uint64_t consoleCap, int threads) {
    std::string console;
    Params p{n, instBase, seedBase, randReset, nEvals, runFlags, tbSeed, tbFlags, resetPeriod,
             tracePlusarg, obs, obsEvals, sigs, static_cast<StatusRec*>(status), finalState,
             covSums, vcdPath, vcdInstance, covPath, covInstance, &console};
    if (threads < 1) threads = 1;
    std::vector<std::vector<uint64_t>> cov(threads, std::vector<uint64_t>(kCovBins, 0));
    vlref_loop_seconds_begin();
    if (threads == 1) {
        for (uint64_t i = 0; i < n; ++i) runOne(p, i, cov[0]);
        vlref_loop_seconds_thread_done();
    } else {
        std::atomic<uint64_t> next{0};
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t) {
            pool.emplace_back([&, t]() {
                for (;;) {
                    const uint64_t i = next.fetch_add(1);
                    if (i >= n) break;
                    runOne(p, i, cov[t]);
                }
                vlref_loop_seconds_thread_done();
            });
        }
        for (auto& th : pool) th.join();
    }
    if (covSums) {
        for (int b = 0; b < kCovBins; ++b) {
            covSums[b] = 0;
            for (int t = 0; t < threads; ++t) covSums[b] += cov[t][b];
        }
    }
    if (consoleOut && consoleCap) {
        strncpy(consoleOut, console.c_str(), consoleCap - 1);
        consoleOut[consoleCap - 1] = 0;
    }
    return 0;
}
