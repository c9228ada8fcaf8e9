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
#include <stdlib.h>
#include <string.h>
#include <type_traits>

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
        // rows are relative to this run: row 0 drives the run's first eval
        const uint32_t* rowp = r.stim + (e - r.evalBase) * T_Design::kInWords * stride;
        T_Design::load_inputs_row(root, rowp, stride, r.stimBroadcast ? 0 : inst);
    }
};

// Random stimulus made in the kernel -- a fuzz campaign needs no stimulus rows at all.  Before
// eval e (clock period c = e / 2) input word w of instance i is a hash of (tbSeed, i's seed, c, w):
//   word 0    the clock: 1 on even evals, 0 on odd ones (rising edge first, sim_main.cpp:73-76)
//   word 1    with VLCU_FUZZ_RESET_LOW / _HIGH: the reset, asserted for the first tbArg[0] periods
//   tbArg[1]  bit w set: word w is "sparse", non-zero in 1 period out of 32 (interrupt pins)
//   VLCU_FUZZ_UNIFORM: every instance sees the same stimulus (they differ by --x-initial only,
//   BASELINE config 3)
// Ports are masked to their widths by load_inputs_row.  The hash is restated in
// tests/oracle_util.py (fuzz_rows) for the CPU side.
enum : uint32_t {
    VLCU_FUZZ_RESET_LOW = 1u << 0,
    VLCU_FUZZ_RESET_HIGH = 1u << 1,
    VLCU_FUZZ_UNIFORM = 1u << 2,
};
VLCU_HD uint64_t vlcu_fuzz_mix(uint64_t x) {
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
    return x ^ (x >> 31);
}
VLCU_HD uint64_t vlcu_fuzz_draw(uint64_t key, uint64_t period, uint32_t word) {
    return vlcu_fuzz_mix(key + period * 0x9e3779b97f4a7c15ULL + (static_cast<uint64_t>(word) + 1U) * 0xd1342543de82ef95ULL);
}
struct VlcuTbFuzz final {
    static constexpr const char* name() { return "fuzz"; }
    template <class T_Design>
    VLCU_D static void begin(typename T_Design::Root&, VlcuContext&, const VlcuRunDev&, uint64_t,
                             uint64_t) {}
    template <class T_Design>
    VLCU_D static void drive(typename T_Design::Root& root, VlcuContext& ctx, const VlcuRunDev& r,
                             uint64_t, uint64_t e) {
        ctx.timeInc(1);
        constexpr uint32_t kWords = T_Design::kInWords ? T_Design::kInWords : 1;
        uint32_t row[kWords];
        const uint64_t period = e >> 1;
        const uint64_t key = r.tbSeed ^ ((r.tbFlags & VLCU_FUZZ_UNIFORM) ? 0 : vlcu_fuzz_mix(ctx.randSeed() + 1U));
        VLCU_UNROLL
        for (uint32_t w = 0; w < kWords; ++w) {
            const uint64_t draw = vlcu_fuzz_draw(key, period, w);
            const bool sparse = w < 32 && ((r.tbArg[1] >> w) & 1U);
            row[w] = (sparse && (draw & 0x1f) != 0) ? 0U : static_cast<uint32_t>(draw >> 32);
        }
        row[0] = static_cast<uint32_t>(e & 1U) ^ 1U;
        if (kWords > 1 && (r.tbFlags & VLCU_FUZZ_RESET_LOW)) row[kWords > 1 ? 1 : 0] = period >= r.tbArg[0];
        if (kWords > 1 && (r.tbFlags & VLCU_FUZZ_RESET_HIGH)) row[kWords > 1 ? 1 : 0] = period < r.tbArg[0];
        T_Design::load_inputs_row(root, row, 1, 0);
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
    vlcu_rng_seed(m.rng + 2 * inst, m.seedBase ^ (m.instBase + inst));
    VlcuContext ctx;
    vlcu_ctx_load(ctx, m, inst);
    typename T_Design::Root root;
    typename T_Design::Locals locals;
    T_Design::bind(root, locals, shared, m.state, inst);
    T_Design::ctor_var_reset(&root, ctx);
    T_Design::store(root, m.state, inst);
    vlcu_ctx_store(ctx);
}

// tfp->dump(time) of one sampled instance as a change record: the model's
// trace functions store the current value of every trace code, the words that differ from the
// previous dump are appended to the instance's stream.  Nothing but the changes crosses PCIe.
template <class T_Design>
VLCU_D void vlcu_trace_changes(typename T_Design::Root* rootp, VlcuContext& ctx, const VlcuRunDev& r,
                               uint32_t slot) {
    constexpr uint32_t kWords = VLCU_TRACE_WORDS_PADDED(T_Design::kTraceWords);
    uint32_t* const streamp = reinterpret_cast<uint32_t*>(r.trace) + static_cast<uint64_t>(slot) * r.traceCapWords;
    uint32_t* const curp = r.traceScratch + static_cast<uint64_t>(slot) * 2U * kWords;
    uint32_t* const prevp = curp + kWords;
    T_Design::trace_sample(rootp, ctx, curp);
    const bool first = streamp[2] == 0;
    uint32_t used = streamp[0] ? streamp[0] : VLCU_TRACE_STREAM_HDR;  // (the buffer starts zeroed)
    uint32_t changed = 0;
    for (uint32_t w = 0; w < T_Design::kTraceWords; ++w) changed += (first || curp[w] != prevp[w]);
    if (used + 3U + 2U * changed > r.traceCapWords) {
        streamp[1] += 1U;  // no room: the dump is dropped, prev keeps the last RECORDED values
        return;
    }
    streamp[used++] = VLCU_TRACE_DUMP_MARK | changed;
    streamp[used++] = static_cast<uint32_t>(ctx.time());
    streamp[used++] = static_cast<uint32_t>(ctx.time() >> 32);
    for (uint32_t w = 0; w < T_Design::kTraceWords; ++w) {
        const uint32_t v = curp[w];
        if (first || v != prevp[w]) {
            streamp[used++] = w;
            streamp[used++] = v;
            prevp[w] = v;
        }
    }
    streamp[0] = used;
    streamp[2] += 1U;
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
        // (tryRun sends runs of 2^32 evals or more to the kObserve form, whose trip count is 64 bits)
        const uint32_t nEvals = static_cast<uint32_t>(r.nEvals);
        // A tiled design calls real functions; lanes that leave one of them apart (a cover point
        // only some of them passed) stay apart in the caller -- nothing there waits for them --
        // and would run the rest of the LAUNCH a few lanes at a time.  All lanes of the warp that
        // are there run the same number of evals: they meet again at the top of each.
        const unsigned live = T_Design::kTiled ? __activemask() : 0U;  // (nothing at all for a register-form design)
        for (uint32_t e = 0; e < nEvals; ++e) {
            if (T_Design::kTiled) __syncwarp(live);
#ifdef VLCU_EVAL_SYNC
            // (experiment) keep the CTA's warps at the same place in the code: a tiled design's
            // eval is tens of thousands of straight-line instructions, fetched once per warp
            if (T_Design::kTiled) __syncthreads();
#endif
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
            if (slot >= 0 && e < r.traceEvals && (r.flags & VLCU_RUN_TRACE_CHANGES)) {
                vlcu_trace_changes<T_Design>(&root, ctx, r, static_cast<uint32_t>(slot));
            } else if (slot >= 0 && e < r.traceEvals) {
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

//=========================================================================
// Port streaming: "trace" testbench with per-instance stimulus rows and (optionally) every
// output port captured after every eval -- the per-cycle port parity path, and the one workload
// whose bytes really stream through HBM every eval (S_in + S_out per instance-eval).
//
// Rows are staged through shared memory by the TMA unit: one elected thread per CTA issues 1-D
// bulk copies (cp.async.bulk, SASS UBLKCP) of the CTA's 64-instance segment of each input word
// kDepth evals ahead, completion is signalled on one mbarrier per stage; output words are
// written to a shared stage and leave with bulk stores (cp.async.bulk.global.shared::cta), so
// the eval loop itself never waits on a global load and issues no global store.

VLCU_D uint32_t vlcu_smem_addr(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
VLCU_D void vlcu_mbar_init(uint64_t* barp, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(vlcu_smem_addr(barp)), "r"(count)
                 : "memory");
}
VLCU_D void vlcu_mbar_expect_tx(uint64_t* barp, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(vlcu_smem_addr(barp)),
                 "r"(bytes)
                 : "memory");
}
VLCU_D void vlcu_mbar_wait(uint64_t* barp, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(vlcu_smem_addr(barp)), "r"(parity)
            : "memory");
    } while (!done);
}
// global -> shared, completion counted in bytes on the mbarrier
VLCU_D void vlcu_bulk_load(void* smemDst, const void* gsrc, uint32_t bytes,
                                         uint64_t* barp) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            vlcu_smem_addr(smemDst)),
        "l"(gsrc), "r"(bytes), "r"(vlcu_smem_addr(barp))
        : "memory");
}
// shared -> global, tracked by the thread's bulk async-group
VLCU_D void vlcu_bulk_store(void* gdst, const void* smemSrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                 "r"(vlcu_smem_addr(smemSrc)), "r"(bytes)
                 : "memory");
}
// 2-D tensor tile: one TMA instruction moves the CTA's segment of every word of one eval
VLCU_D void vlcu_tma_load_2d(void* smemDst, const void* tmapp, uint32_t x, uint32_t y, uint64_t* barp) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            vlcu_smem_addr(smemDst)),
        "l"(tmapp), "r"(x), "r"(y), "r"(vlcu_smem_addr(barp))
        : "memory");
}
VLCU_D void vlcu_tma_store_2d(const void* tmapp, uint32_t x, uint32_t y, const void* smemSrc) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tmapp),
                 "r"(x), "r"(y), "r"(vlcu_smem_addr(smemSrc))
                 : "memory");
}
VLCU_D void vlcu_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
VLCU_D void vlcu_bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
VLCU_D void vlcu_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (TMA)
VLCU_D void vlcu_fence_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// Pipeline depth: as many evals in flight as fit in 28 KB of shared memory per CTA (7 CTAs/SM)
template <class T_Design>
struct VlcuStreamCfg final {
    static constexpr uint32_t kInStage = (T_Design::kInWords ? T_Design::kInWords : 1) * VLCU_BLOCK;  // words
    static constexpr uint32_t kOutStage = (T_Design::kOutWords ? T_Design::kOutWords : 1) * VLCU_BLOCK;
    static constexpr uint32_t kStageBytes = 4 * (kInStage + kOutStage);
    // per-CTA budget: 28 KB at the default 7 CTAs per SM, less when more CTAs are to be resident
    // (what the design's own shared-memory arrays take comes off the budget; statically allocated
    // shared memory is limited to 48 KB per CTA)
    static constexpr uint32_t kShared = static_cast<uint32_t>(sizeof(typename T_Design::Shared));
    static constexpr uint32_t kGross = VLCU_MIN_BLOCKS <= 7 ? 28672U : (227U * 1024U / VLCU_MIN_BLOCKS - 1024U - 512U);
    static constexpr uint32_t kBudget = kGross > kShared ? kGross - kShared : 0U;
    static constexpr int kDepth = kStageBytes * 8 <= kBudget ? 8 : kStageBytes * 4 <= kBudget ? 4 : 2;
    // a tiled design keeps its ports in the tile image: rows are read / written in place
    static constexpr bool kFits = !T_Design::kTiled && kStageBytes * kDepth + kShared + 256U <= 48U * 1024U;
};

template <class T_Design, bool kObs, bool kTensor>
__global__ void __launch_bounds__(VLCU_BLOCK, VLCU_MIN_BLOCKS)
    vlcu_k_stream(VlcuModelDev m, VlcuRunDev r, const __grid_constant__ VlcuRowMaps maps) {
    typedef VlcuStreamCfg<T_Design> Cfg;
    constexpr int kDepth = Cfg::kDepth;
    constexpr uint32_t kSegBytes = VLCU_BLOCK * 4;  // one word of the CTA's instances
    __shared__ typename T_Design::Shared shared;
    __shared__ __align__(128) uint32_t s_in[kDepth][Cfg::kInStage];
    __shared__ __align__(128) uint32_t s_out[kObs ? kDepth : 1][Cfg::kOutStage];
    __shared__ __align__(8) uint64_t s_full[kDepth];
    const uint32_t tid = threadIdx.x;
    const uint64_t ctaBase = blockIdx.x * static_cast<uint64_t>(VLCU_BLOCK);
    const uint64_t inst = ctaBase + tid;
    const bool valid = inst < m.n;  // rows are padded to npad, so the copies stay in bounds
    const uint32_t nEvals = static_cast<uint32_t>(r.nEvals);
    const uint32_t* const stimBase = r.stim + ctaBase;
    uint32_t* const obsBase = r.obs + ctaBase;

    // producer: one stage = the CTA's segment of every input word of one eval
    const auto issueLoads = [&](uint32_t e) {
        const uint32_t s = e % kDepth;
        vlcu_mbar_expect_tx(&s_full[s], T_Design::kInWords * kSegBytes);
        if (kTensor) {
            vlcu_tma_load_2d(&s_in[s][0], &maps.in, static_cast<uint32_t>(ctaBase),
                             e * T_Design::kInWords, &s_full[s]);
            return;
        }
        const uint32_t* const rowp = stimBase + static_cast<uint64_t>(e) * T_Design::kInWords * r.npad;
#pragma unroll
        for (uint32_t w = 0; w < T_Design::kInWords; ++w) {
            vlcu_bulk_load(&s_in[s][w * VLCU_BLOCK], rowp + w * r.npad, kSegBytes, &s_full[s]);
        }
    };
    if (tid == 0) {
        for (int s = 0; s < kDepth; ++s) vlcu_mbar_init(&s_full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        vlcu_fence_async_smem();
    }
    __syncthreads();
    if (tid == 0 && T_Design::kInWords) {
        for (uint32_t e = 0; e < static_cast<uint32_t>(kDepth) && e < nEvals; ++e) issueLoads(e);
    }

    typename T_Design::Root root;
    typename T_Design::Locals locals;
    VlcuContext ctx;
    const uint64_t instc = valid ? inst : 0;  // padding lanes shadow instance 0 and store nothing
    vlcu_ctx_load(ctx, m, instc);
    T_Design::bind(root, locals, shared, m.state, instc);
    T_Design::load(root, m.state, instc);
    uint64_t sig = r.sigs[instc];
    for (uint32_t e = 0; e < nEvals; ++e) {
        const uint32_t s = e % kDepth;
        if (T_Design::kInWords) {
            vlcu_mbar_wait(&s_full[s], (e / kDepth) & 1U);
            ctx.timeInc(1);
            T_Design::load_inputs_row(root, &s_in[s][0], VLCU_BLOCK, tid);
        } else {
            ctx.timeInc(1);
        }
        if (valid) {  // padding lanes must not bump coverage counters or memories
            vlcu_eval_step<T_Design>(&root, ctx);
            sig = T_Design::outputs_sig(root, sig);
        }
        if (kObs) {
            T_Design::store_outputs_row(root, &s_out[kObs ? s : 0][0], VLCU_BLOCK, tid);
            vlcu_fence_async_smem();
            // stage s+1 is written next iteration: its bulk stores (issued kDepth-1 evals ago)
            // must have finished reading shared memory
            if (tid == 0) vlcu_bulk_wait_read<kDepth - 2>();
        }
        __syncthreads();  // inputs of stage s consumed by all, outputs of stage s complete
        if (tid == 0) {
            if (kObs && kTensor) {
                vlcu_tma_store_2d(&maps.out, static_cast<uint32_t>(ctaBase), e * T_Design::kOutWords,
                                  &s_out[kObs ? s : 0][0]);
                vlcu_bulk_commit();
            } else if (kObs) {
                uint32_t* const rowp = obsBase + static_cast<uint64_t>(e) * T_Design::kOutWords * r.npad;
#pragma unroll
                for (uint32_t w = 0; w < T_Design::kOutWords; ++w) {
                    vlcu_bulk_store(rowp + w * r.npad, &s_out[kObs ? s : 0][w * VLCU_BLOCK], kSegBytes);
                }
                vlcu_bulk_commit();
            }
            if (T_Design::kInWords && e + kDepth < nEvals) issueLoads(e + kDepth);
        }
    }
    if (kObs && tid == 0) vlcu_bulk_wait_all();
    if (valid) {
        T_Design::store(root, m.state, inst);
        vlcu_ctx_store(ctx);
        r.sigs[inst] = sig;
        if (r.evalsDone) r.evalsDone[inst] += nEvals;
    }
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
    // The TMA-staged port streaming kernel serves "trace" runs that capture either every eval's
    // outputs or none and sample no waveforms (VLCU_RUNF_NO_STREAM forces the direct-access kernel)
    static bool tryStream(const VlcuModelDev& m, const VlcuRunDev& r, cudaStream_t stream) {
        return tryStreamImpl(m, r, stream, std::integral_constant<bool, VlcuStreamCfg<T_Design>::kFits>{});
    }
    static bool tryStreamImpl(const VlcuModelDev&, const VlcuRunDev&, cudaStream_t, std::false_type) {
        return false;
    }
    static bool tryStreamImpl(const VlcuModelDev& m, const VlcuRunDev& r, cudaStream_t stream,
                              std::true_type) {
        if ((r.flags & VLCU_RUN_NO_STREAM) || r.stimBroadcast || r.nSamples || (r.flags & VLCU_RUN_STOP_ON_FINISH)
            || r.nEvals > 0xffffffffULL || r.nEvals == 0) {
            return false;
        }
        const bool obs = r.obs != nullptr;
        if (obs && r.obsEvals != r.nEvals) return false;
        const bool tensor = r.rowMaps && !(r.flags & VLCU_RUN_BULK_1D);
        const VlcuRowMaps maps = r.rowMaps ? *r.rowMaps : VlcuRowMaps{};
        const unsigned grid = blocksFor(m.n);
        if (obs && tensor) {
            vlcu_k_stream<T_Design, true, true><<<grid, VLCU_BLOCK, 0, stream>>>(m, r, maps);
        } else if (obs) {
            vlcu_k_stream<T_Design, true, false><<<grid, VLCU_BLOCK, 0, stream>>>(m, r, maps);
        } else if (tensor) {
            vlcu_k_stream<T_Design, false, true><<<grid, VLCU_BLOCK, 0, stream>>>(m, r, maps);
        } else {
            vlcu_k_stream<T_Design, false, false><<<grid, VLCU_BLOCK, 0, stream>>>(m, r, maps);
        }
        return true;
    }
    template <class T_Tb>
    bool tryRun(const char* tb, const VlcuModelDev& m, const VlcuRunDev& r, cudaStream_t stream,
                cudaError_t& err) const {
        if (strcmp(tb, T_Tb::name()) != 0) return false;
        if (std::is_same<T_Tb, VlcuTbTrace>::value && tryStream(m, r, stream)) {
            err = cudaGetLastError();
            return true;
        }
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
                             || tryRun<VlcuTbTrace>(tb, m, r, stream, err)
                             || tryRun<VlcuTbFuzz>(tb, m, r, stream, err))
                            || ... || tryRun<T_Tbs>(tb, m, r, stream, err));
        return found ? err : cudaErrorInvalidValue;
    }
    int tbCount() const override { return 3 + static_cast<int>(sizeof...(T_Tbs)); }
    const char* tbName(int i) const override {
        const char* names[] = {VlcuTbNone::name(), VlcuTbTrace::name(), VlcuTbFuzz::name(), T_Tbs::name()...};
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
