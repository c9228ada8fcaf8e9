// -*- mode: C++; c-file-style: "cc-mode" -*-
//*************************************************************************
//
// verilated_cuda_launch.h -- kernels and host-side launch glue that bind one
// emitted design (a "T_Design" policy struct printed by V3EmitCuda) to the
// batch runtime (verilated_cuda.cu).  nvcc only.
//
// A T_Design provides (all static, __device__):
//     typedef ... Root;                      the per-instance model image
//                                            (Vtop___024root, V3EmitCHeaders.cpp:73-143)
//     typedef ... Locals;                    per-thread storage of run-time indexed small arrays
//     typedef ... Shared;                    per-CTA (shared memory) storage of the same kind
//     bind(Root&, Locals&, Shared&, VlcuState, inst)   attach array handles (local/shared/HBM)
//     load / store(Root&, VlcuState, inst)   SoA <-> Root for all persistent members
//     load_inputs_row / store_outputs_row    packed 32-bit port rows (stimulus / observation)
//     outputs_sig(Root, sig)                 fold the output ports into a signature
//     ctor_var_reset                         V3CCtors.cpp:52-236 / V3EmitCFunc.cpp:524-666
//     eval_static/_initial/_settle/eval/_final   V3Sched.cpp:934-1213
// and a static host-side `info()` returning the VlcuDesignInfo tables.
//
// A testbench policy T_Tb is the device-side analogue of the user's C++
// main loop (examples/make_tracing_c/sim_main.cpp:62-106): begin() sets the
// initial port values, drive() changes ports before each eval().
//
//*************************************************************************
#ifndef VERILATOR_VERILATED_CUDA_LAUNCH_H_
#define VERILATOR_VERILATED_CUDA_LAUNCH_H_

#include "verilated_cuda.h"
#include "verilated_cuda_host.h"

#include <cuda_runtime.h>
#include <string.h>

//=========================================================================
// Built-in testbenches usable with any design

// No stimulus: ports were written by vlcu_set_input(); time does not advance.
// This is what vlcu_eval() (== Vtop::eval()) runs.
struct VlcuTbNone final {
    static constexpr const char* name() { return "none"; }
    template <class T_Design>
    VLCU_D static void begin(typename T_Design::Root&, VlcuContext&, const VlcuRunDev&, uint64_t,
                             uint64_t) {}
    template <class T_Design>
    VLCU_D static void drive(typename T_Design::Root&, VlcuContext&, const VlcuRunDev&, uint64_t,
                             uint64_t) {}
};

// Stimulus rows from memory: before eval e every input port is loaded from
// row e of r.stim ([eval][input word][instance], or [eval][input word] when
// broadcast), and time advances by one unit, as a host testbench that pokes
// all ports and calls timeInc(1); eval() would do.
struct VlcuTbTrace final {
    static constexpr const char* name() { return "trace"; }
    template <class T_Design>
    VLCU_D static void begin(typename T_Design::Root&, VlcuContext&, const VlcuRunDev&, uint64_t,
                             uint64_t) {}
    template <class T_Design>
    VLCU_D static void drive(typename T_Design::Root& root, VlcuContext& ctx, const VlcuRunDev& r,
                             uint64_t inst, uint64_t e) {
        ctx.timeInc(1);
        const uint64_t stride = r.stimBroadcast ? 1 : r.npad;
        const uint32_t* rowp = r.stim + e * T_Design::kInWords * stride;
        T_Design::load_inputs_row(root, rowp, stride, r.stimBroadcast ? 0 : inst);
    }
};

//=========================================================================
// Kernels

// Resident CTAs per SM the register allocation must allow: 7 x 64 threads x 148 SMs = 66,304
// >= 65,536 instances (BASELINE.json config 2), i.e. the whole batch is one co-resident wave
// (<= 146 registers per thread).
#ifndef VLCU_MIN_BLOCKS
#define VLCU_MIN_BLOCKS 7
#endif

// Construct every instance: Vtop___024root___ctor_var_reset + zeroed context
template <class T_Design>
__global__ void __launch_bounds__(VLCU_BLOCK) vlcu_k_reset(VlcuModelDev m) {
    __shared__ typename T_Design::Shared shared;
    const uint64_t inst = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (inst >= m.n) return;
    VlcuCtxRec* const recp = &m.ctxRecs[inst];
    *recp = VlcuCtxRec{};
    VlcuContext ctx;
    vlcu_ctx_load(ctx, m, inst);
    typename T_Design::Root root;
    typename T_Design::Locals locals;
    T_Design::bind(root, locals, shared, m.state, inst);
    T_Design::ctor_var_reset(&root, ctx);
    T_Design::store(root, m.state, inst);
    vlcu_ctx_store(ctx);
}

// The cycle loop: nEvals x { testbench drive; eval_step; observe }
// kObserve = false is the throughput form: no per-eval output capture, no waveform sampling,
// no stop-on-finish test -- the loop body is drive + eval_step + signature only, with a 32-bit
// trip count; vlcu_run picks it whenever the run descriptor asks for none of those.
template <class T_Design, class T_Tb, bool kObserve>
__global__ void __launch_bounds__(VLCU_BLOCK, VLCU_MIN_BLOCKS) vlcu_k_run(VlcuModelDev m, VlcuRunDev r) {
    __shared__ typename T_Design::Shared shared;
    const uint64_t inst = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (inst >= m.n) return;
    typename T_Design::Root root;
    typename T_Design::Locals locals;
    VlcuContext ctx;
    vlcu_ctx_load(ctx, m, inst);
    T_Design::bind(root, locals, shared, m.state, inst);
    T_Design::load(root, m.state, inst);
    uint64_t sig = r.sigs[inst];
    if (r.flags & VLCU_RUN_TB_BEGIN) {
        T_Tb::template begin<T_Design>(root, ctx, r, inst, m.instBase + inst);
    }
    uint64_t evalsDone = 0;
    if (!kObserve) {
        // Runs longer than 2^32 evals are split by the host (vlcu_run)
        const uint32_t nEvals = static_cast<uint32_t>(r.nEvals);
        for (uint32_t e = 0; e < nEvals; ++e) {
            T_Tb::template drive<T_Design>(root, ctx, r, inst, r.evalBase + e);
            vlcu_eval_step<T_Design>(&root, ctx);
            sig = T_Design::outputs_sig(root, sig);
        }
        evalsDone = nEvals;
    } else {
        // Is this instance sampled for waveform tracing?
        int slot = -1;
        for (uint32_t s = 0; s < r.nSamples; ++s) {
            if (r.sampleIds[s] == inst) slot = static_cast<int>(s);
        }
        for (uint64_t e = 0; e < r.nEvals; ++e) {
            // A host testbench loops "while (!contextp->gotFinish())" (sim_main.cpp:71)
            if ((r.flags & VLCU_RUN_STOP_ON_FINISH) && ctx.gotFinish()) break;
            T_Tb::template drive<T_Design>(root, ctx, r, inst, r.evalBase + e);
            vlcu_eval_step<T_Design>(&root, ctx);
            ++evalsDone;
            sig = T_Design::outputs_sig(root, sig);
            if (r.obs && e < r.obsEvals) {
                T_Design::store_outputs_row(root, r.obs + e * T_Design::kOutWords * r.npad, r.npad,
                                            inst);
            }
            if (slot >= 0 && e < r.traceEvals) {
                constexpr uint64_t kRecBytes
                    = VLCU_TRACE_HDR + T_Design::kStateBytes + 4ULL * VLCU_TRACE_WORDS_PADDED(T_Design::kTraceWords);
                uint8_t* const recp = r.trace + (e * r.nSamples + slot) * kRecBytes;
                reinterpret_cast<uint64_t*>(recp)[0] = ctx.time();
                reinterpret_cast<uint32_t*>(recp)[2] = ctx.m_status;
                reinterpret_cast<uint32_t*>(recp)[3] = 1;  // record valid
                T_Design::store(root, VlcuState{recp + VLCU_TRACE_HDR, 1}, 0);
                // the model's trace_full functions, storing by trace code (tfp->dump(time))
                T_Design::trace_sample(&root, ctx,
                                       reinterpret_cast<uint32_t*>(recp + VLCU_TRACE_HDR + T_Design::kStateBytes));
            }
        }
    }
    T_Design::store(root, m.state, inst);
    vlcu_ctx_store(ctx);
    r.sigs[inst] = sig;
    if (r.evalsDone) r.evalsDone[inst] += evalsDone;
}

// Vtop::final(): run the design's final blocks
template <class T_Design>
__global__ void __launch_bounds__(VLCU_BLOCK) vlcu_k_final(VlcuModelDev m) {
    __shared__ typename T_Design::Shared shared;
    const uint64_t inst = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (inst >= m.n) return;
    typename T_Design::Root root;
    typename T_Design::Locals locals;
    VlcuContext ctx;
    vlcu_ctx_load(ctx, m, inst);
    T_Design::bind(root, locals, shared, m.state, inst);
    T_Design::load(root, m.state, inst);
    T_Design::eval_final(&root, ctx);
    T_Design::store(root, m.state, inst);
    vlcu_ctx_store(ctx);
}

//=========================================================================
// Host-side binding of a design + its testbenches to the runtime registry

template <class T_Design, class... T_Tbs>
class VlcuDesignImpl final : public VlcuDesignOps {
    static unsigned blocksFor(uint64_t n) {
        return static_cast<unsigned>((n + VLCU_BLOCK - 1) / VLCU_BLOCK);
    }
    template <class T_Tb>
    bool tryRun(const char* tb, const VlcuModelDev& m, const VlcuRunDev& r, cudaStream_t stream,
                cudaError_t& err) const {
        if (strcmp(tb, T_Tb::name()) != 0) return false;
        const bool observe = r.obs || r.nSamples || (r.flags & VLCU_RUN_STOP_ON_FINISH)
                             || r.nEvals > 0xffffffffULL;
        if (observe) {
            vlcu_k_run<T_Design, T_Tb, true><<<blocksFor(m.n), VLCU_BLOCK, 0, stream>>>(m, r);
        } else {
            vlcu_k_run<T_Design, T_Tb, false><<<blocksFor(m.n), VLCU_BLOCK, 0, stream>>>(m, r);
        }
        err = cudaGetLastError();
        return true;
    }

public:
    const VlcuDesignInfo& info() const override { return T_Design::info(); }
    cudaError_t launchReset(const VlcuModelDev& m, cudaStream_t stream) const override {
        vlcu_k_reset<T_Design><<<blocksFor(m.n), VLCU_BLOCK, 0, stream>>>(m);
        return cudaGetLastError();
    }
    cudaError_t launchFinal(const VlcuModelDev& m, cudaStream_t stream) const override {
        vlcu_k_final<T_Design><<<blocksFor(m.n), VLCU_BLOCK, 0, stream>>>(m);
        return cudaGetLastError();
    }
    cudaError_t launchRun(const char* tb, const VlcuModelDev& m, const VlcuRunDev& r,
                          cudaStream_t stream) const override {
        cudaError_t err = cudaErrorInvalidValue;
        const bool found = ((tryRun<VlcuTbNone>(tb, m, r, stream, err)
                             || tryRun<VlcuTbTrace>(tb, m, r, stream, err))
                            || ... || tryRun<T_Tbs>(tb, m, r, stream, err));
        return found ? err : cudaErrorInvalidValue;
    }
    int tbCount() const override { return 2 + static_cast<int>(sizeof...(T_Tbs)); }
    const char* tbName(int i) const override {
        const char* names[] = {VlcuTbNone::name(), VlcuTbTrace::name(), T_Tbs::name()...};
        return (i >= 0 && i < tbCount()) ? names[i] : nullptr;
    }
};

// One static instance per emitted design registers it at load time
#define VLCU_REGISTER_DESIGN(Design, ...) \
    static VlcuDesignImpl<Design, ##__VA_ARGS__> s_vlcuDesign_##Design; \
    static const int s_vlcuDesignReg_##Design = vlcu_register_design(&s_vlcuDesign_##Design)

// Same, with the argument list macro-expanded first (the emitted model passes the testbench
// list as a macro: VLCU_REGISTER_DESIGN_X(Vtop__cuda VLCU_TESTBENCHES))
#define VLCU_REGISTER_DESIGN_X(...) VLCU_REGISTER_DESIGN(__VA_ARGS__)

#endif  // guard
