// TEST INFRASTRUCTURE (not product code) -- design-independent oracle harness.
//
// Compiled once per design with
//   -DVTOP=V<design>  -DVTOP_H='"V<design>.h"'  -DPORTS_INC='"V<design>__ports.inc"'
//   -DHARNESS_FN=vlref_<design>_run
// against the GENUINE Verilated model written by oracle/_ref/verilator_lite.  It plays the
// built-in "trace" testbench of the CUDA runtime (include/verilated_cuda_launch.h VlcuTbTrace):
// before eval e every input port is loaded from stimulus row e ([eval][input word][instance],
// ports in source order, one 32-bit word per limb, values masked to the port width), time
// advances by one unit, eval(), and every output port is stored to observation row e.
// The second entry point (-DHARNESS_FUZZ_FN=vlref_<design>_fuzz) plays the "fuzz" testbench
// (VlcuTbFuzz): the stimulus rows are made here, per instance and eval, from the same hash.
#include VTOP_H
#include "ref_timing.h"

#include <atomic>
#include <memory>
#include <thread>
#include <vector>

namespace {
inline uint32_t sigWord(uint32_t acc, uint32_t word) { return (acc ^ word) * 0x01000193U; }
inline uint64_t sigMix(uint64_t sig, uint32_t acc) { return (sig ^ acc) * 0x100000001b3ULL; }

inline uint32_t maskOf(int bits) { return bits >= 32 ? 0xffffffffU : ((1U << bits) - 1U); }

// ---- port <-> row words -------------------------------------------------------------------
template <typename T>
void setPort(T& port, int width, const uint32_t* wp, uint64_t stride) {
    port = static_cast<T>(wp[0] & maskOf(width));
}
void setPort(QData& port, int width, const uint32_t* wp, uint64_t stride) {
    port = (static_cast<QData>(wp[stride] & maskOf(width - 32)) << 32) | wp[0];
}
template <size_t N>
void setPort(VlWide<N>& port, int width, const uint32_t* wp, uint64_t stride) {
    for (size_t w = 0; w < N; ++w) {
        const int bits = (w == N - 1) ? (width - 32 * static_cast<int>(w)) : 32;
        port[w] = wp[w * stride] & maskOf(bits);
    }
}
template <typename T>
void getPort(const T& port, uint32_t* out) {
    out[0] = static_cast<uint32_t>(port);
}
void getPort(const QData& port, uint32_t* out) {
    out[0] = static_cast<uint32_t>(port);
    out[1] = static_cast<uint32_t>(port >> 32);
}
template <size_t N>
void getPort(const VlWide<N>& port, uint32_t* out) {
    for (size_t w = 0; w < N; ++w) out[w] = port[w];
}

// an inout port is driven from the stimulus row like an input and observed like an output
#define VLCU_PORT_INOUT(name, width, words) VLCU_PORT_IN(name, width, words) VLCU_PORT_OUT(name, width, words)

constexpr int kInWords = 0
#define VLCU_PORT_IN(name, width, words) +words
#define VLCU_PORT_OUT(name, width, words)
#include PORTS_INC
#undef VLCU_PORT_IN
#undef VLCU_PORT_OUT
    ;
constexpr int kOutWords = 0
#define VLCU_PORT_IN(name, width, words)
#define VLCU_PORT_OUT(name, width, words) +words
#include PORTS_INC
#undef VLCU_PORT_IN
#undef VLCU_PORT_OUT
    ;

struct Params {
    uint64_t n, instBase, seedBase;
    int randReset;
    uint64_t nEvals;
    const uint32_t* stim;  // [eval][kInWords][n], or null: fuzz
    uint32_t* obs;  // [eval][kOutWords][n]
    uint64_t* sigs;
    // fuzz testbench (VlcuTbFuzz of include/verilated_cuda_launch.h, restated)
    uint64_t tbSeed = 0;
    uint32_t tbFlags = 0, resetPeriods = 0, sparseMask = 0;
    uint64_t obsEvals = ~0ULL;
};

inline uint64_t fuzzMix(uint64_t x) {
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
    return x ^ (x >> 31);
}
void fuzzRow(const Params& p, uint64_t randSeed, uint64_t e, uint32_t* row) {
    const uint64_t period = e >> 1;
    const uint64_t key = p.tbSeed ^ ((p.tbFlags & 4U) ? 0 : fuzzMix(randSeed + 1U));
    for (int w = 0; w < kInWords; ++w) {
        const uint64_t draw
            = fuzzMix(key + period * 0x9e3779b97f4a7c15ULL + (static_cast<uint64_t>(w) + 1U) * 0xd1342543de82ef95ULL);
        const bool sparse = w < 32 && ((p.sparseMask >> w) & 1U);
        row[w] = (sparse && (draw & 0x1f) != 0) ? 0U : static_cast<uint32_t>(draw >> 32);
    }
    row[0] = static_cast<uint32_t>(e & 1U) ^ 1U;
    if (kInWords > 1 && (p.tbFlags & 1U)) row[1] = period >= p.resetPeriods;
    if (kInWords > 1 && (p.tbFlags & 2U)) row[1] = period < p.resetPeriods;
}

void runOne(const Params& p, uint64_t i) {
    const std::unique_ptr<VerilatedContext> contextp{new VerilatedContext};
    Verilated::threadContextp(contextp.get());
    contextp->randReset(p.randReset);
    contextp->randSeed(static_cast<int>(p.seedBase ^ (p.instBase + i)));
    const char* argv[] = {"oracle"};
    contextp->commandArgs(1, argv);
    const std::unique_ptr<VTOP> top{new VTOP{contextp.get(), "TOP"}};
    uint64_t sig = 0xcbf29ce484222325ULL;
    const VlrefLoopTimer timer;
    for (uint64_t e = 0; e < p.nEvals; ++e) {
        contextp->timeInc(1);
        uint32_t fuzz[kInWords > 0 ? kInWords : 1];
        const uint32_t* rowp = p.stim ? p.stim + e * kInWords * p.n + i : fuzz;
        const uint64_t stride = p.stim ? p.n : 1;
        if (!p.stim) fuzzRow(p, p.seedBase ^ (p.instBase + i), e, fuzz);
        int k = 0;
#define VLCU_PORT_IN(name, width, words) \
    setPort(top->name, width, rowp + static_cast<uint64_t>(k) * stride, stride); \
    k += words;
#define VLCU_PORT_OUT(name, width, words)
#include PORTS_INC
#undef VLCU_PORT_IN
#undef VLCU_PORT_OUT
        top->eval();
        uint32_t w[kOutWords > 0 ? kOutWords : 1];
        k = 0;
#define VLCU_PORT_IN(name, width, words)
#define VLCU_PORT_OUT(name, width, words) \
    getPort(top->name, w + k); \
    k += words;
#include PORTS_INC
#undef VLCU_PORT_IN
#undef VLCU_PORT_OUT
        uint32_t acc = 0x811c9dc5U;
        for (int j = 0; j < kOutWords; ++j) acc = sigWord(acc, w[j]);
        sig = sigMix(sig, acc);
        if (p.obs && e < p.obsEvals) {
            uint32_t* const orowp = p.obs + e * kOutWords * p.n + i;
            for (int j = 0; j < kOutWords; ++j) orowp[static_cast<uint64_t>(j) * p.n] = w[j];
        }
    }
    top->final();
    if (p.sigs) p.sigs[i] = sig;
}
}  // namespace

namespace {
int runAll(const Params& p, int threads);
}
extern "C" int HARNESS_FN(uint64_t n, uint64_t instBase, uint64_t seedBase, int randReset,
                          uint64_t nEvals, const uint32_t* stim, uint32_t* obs, uint64_t* sigs,
                          int* inWords, int* outWords, int threads) {
    if (inWords) *inWords = kInWords;
    if (outWords) *outWords = kOutWords;
    if (!stim) return 0;  // query only
    const Params p{n, instBase, seedBase, randReset, nEvals, stim, obs, sigs};
    return runAll(p, threads);
}
#ifdef HARNESS_FUZZ_FN
extern "C" int HARNESS_FUZZ_FN(uint64_t n, uint64_t instBase, uint64_t seedBase, int randReset,
                               uint64_t nEvals, uint64_t tbSeed, uint32_t tbFlags, uint32_t resetPeriods,
                               uint32_t sparseMask, uint32_t* obs, uint64_t obsEvals, uint64_t* sigs,
                               int threads) {
    Params p{n, instBase, seedBase, randReset, nEvals, nullptr, obs, sigs};
    p.tbSeed = tbSeed;
    p.tbFlags = tbFlags;
    p.resetPeriods = resetPeriods;
    p.sparseMask = sparseMask;
    p.obsEvals = obsEvals;
    return runAll(p, threads);
}
#endif
#ifdef HARNESS_COV_FN
// --coverage builds: a fuzz campaign of n instances whose coverage is MERGED by the reference's own
// writer -- all models live in ONE VerilatedContext (each constructed with its own randSeed, so the
// --x-initial values are the per-instance ones), their points register in that context's coverage
// database, and VerilatedCovContext::write sums the counters of points with the same key
// (include/verilated_cov.cpp:368-438).  The file is what BASELINE config 5 asks the GPU campaign
// to reproduce: "one merged coverage.dat equal to the sum of per-instance CPU runs".
#include "verilated_cov.h"
extern "C" int HARNESS_COV_FN(uint64_t n, uint64_t instBase, uint64_t seedBase, int randReset,
                              uint64_t nEvals, uint64_t tbSeed, uint32_t tbFlags, uint32_t resetPeriods,
                              uint32_t sparseMask, uint64_t* sigs, const char* covPath) {
    Params p{n, instBase, seedBase, randReset, nEvals, nullptr, nullptr, sigs};
    p.tbSeed = tbSeed;
    p.tbFlags = tbFlags;
    p.resetPeriods = resetPeriods;
    p.sparseMask = sparseMask;
    const std::unique_ptr<VerilatedContext> contextp{new VerilatedContext};
    Verilated::threadContextp(contextp.get());
    contextp->randReset(randReset);
    const char* argv[] = {"oracle"};
    contextp->commandArgs(1, argv);
    std::vector<std::unique_ptr<VTOP>> tops;
    for (uint64_t i = 0; i < n; ++i) {
        contextp->randSeed(static_cast<int>(seedBase ^ (instBase + i)));
        tops.emplace_back(new VTOP{contextp.get(), "TOP"});
        VTOP* const top = tops.back().get();
        uint64_t sig = 0xcbf29ce484222325ULL;
        for (uint64_t e = 0; e < nEvals; ++e) {
            contextp->timeInc(1);
            uint32_t fuzz[kInWords > 0 ? kInWords : 1];
            fuzzRow(p, seedBase ^ (instBase + i), e, fuzz);
            const uint32_t* rowp = fuzz;
            const uint64_t stride = 1;
            int k = 0;
#define VLCU_PORT_IN(name, width, words) \
    setPort(top->name, width, rowp + static_cast<uint64_t>(k) * stride, stride); \
    k += words;
#define VLCU_PORT_OUT(name, width, words)
#include PORTS_INC
#undef VLCU_PORT_IN
#undef VLCU_PORT_OUT
            top->eval();
            uint32_t w[kOutWords > 0 ? kOutWords : 1];
            k = 0;
#define VLCU_PORT_IN(name, width, words)
#define VLCU_PORT_OUT(name, width, words) \
    getPort(top->name, w + k); \
    k += words;
#include PORTS_INC
#undef VLCU_PORT_IN
#undef VLCU_PORT_OUT
            uint32_t acc = 0x811c9dc5U;
            for (int j = 0; j < kOutWords; ++j) acc = sigWord(acc, w[j]);
            sig = sigMix(sig, acc);
        }
        if (sigs) sigs[i] = sig;
    }
    if (covPath) contextp->coveragep()->write(covPath);
    for (auto& top : tops) top->final();
    return 0;
}
#endif
namespace {
int runAll(const Params& p, int threads) {
    const uint64_t n = p.n;
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
}  // namespace
