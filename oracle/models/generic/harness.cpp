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
    const uint32_t* stim;  // [eval][kInWords][n]
    uint32_t* obs;  // [eval][kOutWords][n]
    uint64_t* sigs;
};

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
        const uint32_t* rowp = p.stim + e * kInWords * p.n + i;
        int k = 0;
#define VLCU_PORT_IN(name, width, words) \
    setPort(top->name, width, rowp + static_cast<uint64_t>(k) * p.n, p.n); \
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
        if (p.obs) {
            uint32_t* const orowp = p.obs + e * kOutWords * p.n + i;
            for (int j = 0; j < kOutWords; ++j) orowp[static_cast<uint64_t>(j) * p.n] = w[j];
        }
    }
    top->final();
    if (p.sigs) p.sigs[i] = sig;
}
}  // namespace

extern "C" int HARNESS_FN(uint64_t n, uint64_t instBase, uint64_t seedBase, int randReset,
                          uint64_t nEvals, const uint32_t* stim, uint32_t* obs, uint64_t* sigs,
                          int* inWords, int* outWords, int threads) {
    if (inWords) *inWords = kInWords;
    if (outWords) *outWords = kOutWords;
    if (!stim) return 0;  // query only
    const Params p{n, instBase, seedBase, randReset, nEvals, stim, obs, sigs};
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
