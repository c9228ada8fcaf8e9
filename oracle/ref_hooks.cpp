// TEST INFRASTRUCTURE (not product code).
// User-definable runtime hooks of the reference (include/verilated.cpp:165-240:
// vl_finish / vl_stop / vl_fatal are overridable with VL_USER_FINISH/STOP/FATAL)
// restated so that one process can run many instances without exiting:
// the flags are set exactly as the default hooks set them, but events are
// counted per thread instead of printed, and vl_fatal does not exit().
#include "ref_hooks.h"
#include "verilated.h"

#include <cstdio>
#include <string>

#include "ref_events.h"

static thread_local VlrefEvents* t_eventsp = nullptr;
static thread_local std::string* t_capturep = nullptr;

void vlref_set_events(VlrefEvents* p) { t_eventsp = p; }
void vlref_set_capture(std::string* p) { t_capturep = p; }

extern "C" int vlref_vprintf(const char* fmt, va_list ap) {
    char buf[4096];
    const int n = vsnprintf(buf, sizeof(buf), fmt, ap);
    if (t_capturep) {
        t_capturep->append(buf);
    }  // else: discarded (quiet oracle)
    return n;
}
extern "C" int vlref_printf(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    const int n = vlref_vprintf(fmt, ap);
    va_end(ap);
    return n;
}

// Default behaviour (verilated.cpp:166-171) minus the console message
void vl_finish(const char* filename, int linenum, const char* hier) VL_MT_UNSAFE {
    (void)filename, (void)linenum, (void)hier;
    VerilatedContext* const ctxp = Verilated::threadContextp();
    if (t_eventsp) {
        if (!t_eventsp->finish_count) t_eventsp->first_finish_time = ctxp->time();
        ++t_eventsp->finish_count;
    }
    ctxp->gotFinish(true);
}
// Default behaviour (verilated.cpp:175-190) minus printing and the fatal exit
void vl_stop(const char* filename, int linenum, const char* hier) VL_MT_UNSAFE {
    (void)filename, (void)hier;
    VerilatedContext* const ctxp = Verilated::threadContextp();
    if (ctxp->gotFinish() && !ctxp->executingFinal()) return;
    ctxp->gotError(true);
    ctxp->gotFinish(true);
    if (t_eventsp) {
        if (!t_eventsp->stop_count) {
            t_eventsp->first_stop_time = ctxp->time();
            t_eventsp->first_stop_line = linenum;
        }
        ++t_eventsp->stop_count;
    }
}
void vl_fatal(const char* filename, int linenum, const char* hier, const char* msg) VL_MT_UNSAFE {
    (void)hier;
    VerilatedContext* const ctxp = Verilated::threadContextp();
    if (t_capturep) {
        t_capturep->append(std::string{"%Fatal: "} + filename + ":" + std::to_string(linenum) + ": " + msg + "\n");
    }
    ctxp->gotError(true);
    ctxp->gotFinish(true);
    if (t_eventsp) {
        if (!t_eventsp->fatal_count) {
            t_eventsp->first_fatal_time = ctxp->time();
            t_eventsp->first_fatal_line = linenum;
        }
        ++t_eventsp->fatal_count;
    }
}

// Cycle-loop timing shared by the model harnesses (ref_timing.h)
#include "ref_timing.h"
#include <mutex>
static std::mutex s_loopMutex;
static double s_loopSeconds = 0.0;
void vlref_loop_seconds_begin() {
    const std::lock_guard<std::mutex> lock{s_loopMutex};
    s_loopSeconds = 0.0;
}
void vlref_loop_seconds_thread_done() {
    const std::lock_guard<std::mutex> lock{s_loopMutex};
    if (VlrefLoopTimer::threadTotal() > s_loopSeconds) s_loopSeconds = VlrefLoopTimer::threadTotal();
    VlrefLoopTimer::threadTotal() = 0.0;
}
extern "C" double vlref_last_loop_seconds() { return s_loopSeconds; }
