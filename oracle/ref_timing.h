// TEST INFRASTRUCTURE: wall time of the cycle loops only (construction, program generation and
// final() excluded, as BASELINE.md section 3 asks), per worker thread; the harness reports the
// slowest thread's total through vlref_last_loop_seconds().
#ifndef VLREF_TIMING_H_
#define VLREF_TIMING_H_
#include <atomic>
#include <chrono>
struct VlrefLoopTimer final {
    static double& threadTotal() {
        static thread_local double t = 0.0;
        return t;
    }
    std::chrono::steady_clock::time_point m_t0 = std::chrono::steady_clock::now();
    ~VlrefLoopTimer() {
        threadTotal() += std::chrono::duration<double>(std::chrono::steady_clock::now() - m_t0).count();
    }
};
void vlref_loop_seconds_begin();  // resets the result of the current call
void vlref_loop_seconds_thread_done();  // folds this thread's total into the call's maximum
#endif
