// -*- mode: C++; c-file-style: "cc-mode" -*-
//*************************************************************************
//
// verilated_cuda.h -- device-side runtime of the batched Verilated model.
//
// This is the header that code printed by V3EmitCuda includes (the sibling
// of "verilated.h" for code printed by V3EmitCFunc).  It provides, per CUDA
// thread = per model instance:
//
//   * the VL_* operator library               (verilated_cuda_funcs.h)
//   * VlWide / VlUnpacked / VlNBACommitQueue  (verilated_cuda_types.h)
//   * VlRNG, VL_SCOPED_RAND_RESET_*, VL_MURMUR64_HASH
//                 reference: include/verilated.cpp:428-487, 577-635, 2173-2210
//   * VL_COV_TOGGLE_CHG_ST_*   reference: include/verilated_cov.h:78-105
//   * VlcuContext: the per-instance slice of VerilatedContext
//                 reference: include/verilated.h:371 (time, gotFinish, gotError,
//                 errorCount/Limit, randReset/randSeed), verilated.cpp:165-232
//                 (vl_finish / vl_stop / vl_stop_maybe / vl_fatal semantics)
//   * the structure-of-arrays state accessors and the kernel templates that
//     drive an emitted design (reset / eval / K-cycle run).
//
// State layout in HBM: for a model with N instances (padded to
// a multiple of 128), persistent member m with element type T and E elements
// lives at   base + soaOffset(m) * Npad + (e * Npad + instance) * sizeof(T)
// i.e. "T sig[element][instance]", instance fastest, so a warp touches one
// 32*sizeof(T)-byte contiguous run per element.
//
//*************************************************************************
#ifndef VERILATOR_VERILATED_CUDA_H_
#define VERILATOR_VERILATED_CUDA_H_

#include "verilated_cuda_funcs.h"

//=========================================================================
// PRNG: splitmix64 seeding -> xoroshiro128** (verilated.cpp:428-487)

VLCU_HD uint64_t vlcu_rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

// One splitmix64 step: advances 'x' and returns the mixed output
VLCU_HD uint64_t vlcu_splitmix64(uint64_t& x) {
    x += 0x9e3779b97f4a7c15ULL;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

class VlRNG final {
    uint64_t m_s0, m_s1;

public:
    VLCU_HD explicit VlRNG(uint64_t seed) {
        uint64_t x = seed;
        m_s0 = vlcu_splitmix64(x);
        m_s1 = vlcu_splitmix64(x);
    }
    VLCU_HD VlRNG(uint64_t s0, uint64_t s1)
        : m_s0{s0}
        , m_s1{s1} {}
    VLCU_HD uint64_t s0() const { return m_s0; }
    VLCU_HD uint64_t s1() const { return m_s1; }
    VLCU_HD uint64_t rand64() {
        const uint64_t result = vlcu_rotl64(m_s0 * 5, 7) * 9;
        const uint64_t t = m_s1 ^ m_s0;
        m_s0 = vlcu_rotl64(m_s0, 24) ^ t ^ (t << 16);
        m_s1 = vlcu_rotl64(t, 37);
        return result;
    }
};

// MurmurHash64A with seed 0 over a NUL terminated string (verilated.cpp:2173-2210).
// Scope names are known when the model is built, so this normally runs on the
// host at create time; it is device-capable for the emitter's convenience.
VLCU_HD uint64_t VL_MURMUR64_HASH(const char* key) {
    size_t len = 0;
    while (key[len]) ++len;
    const uint64_t m = 0xc6a4a7935bd1e995ULL;
    const int r = 47;
    uint64_t h = 0 ^ (len * m);
    const size_t nblocks = len / 8;
    for (size_t b = 0; b < nblocks; ++b) {
        uint64_t k = 0;
        for (int j = 7; j >= 0; --j) k = (k << 8) | static_cast<uint8_t>(key[b * 8 + j]);
        k *= m;
        k ^= k >> r;
        k *= m;
        h ^= k;
        h *= m;
    }
    const size_t tail = len & 7;
    if (tail) {
        uint64_t k = 0;
        for (size_t j = tail; j > 0; --j) k = (k << 8) | static_cast<uint8_t>(key[nblocks * 8 + j - 1]);
        h ^= k;
        h *= m;
    }
    h ^= h >> r;
    h *= m;
    h ^= h >> r;
    return h;
}

//=========================================================================
// Coverage counters (uint32_t __Vcoverage[nBins] of the C++ model, V3EmitCHeaders.cpp:153-158).
// They stay in HBM, in the SoA block as cov[bin][instance]: a warp bumping bin n for its 32
// instances touches one 128-byte line, and the bump is a fire-and-forget reduction (RED.ADD, no
// return value), so the eval code never waits for a counter.  Keeping them thread-local instead
// makes every `++__Vcoverage[n]` a dependent load-add-store that misses L1 as soon as a design has
// more than a few hundred points (1,664 for the RV32I core = 6.6 KB per thread).
struct VlcuCovRef final {
    uint32_t* m_p;
    VLCU_HD void operator++() const {
#ifdef __CUDA_ARCH__
        atomicAdd(m_p, 1U);  // result unused: compiles to RED; the address is private to this thread
#else
        ++*m_p;
#endif
    }
    VLCU_HD void operator=(uint32_t value) const { *m_p = value; }
    VLCU_HD operator uint32_t() const { return *m_p; }
};
struct VlcuCovPtr {
    uint32_t* m_basep;  // counter 0 of this instance
    uint32_t m_stride;  // padded instance count (elements between consecutive bins)
    // The stride as seen at a point of use.  The empty volatile asm keeps the compiler from
    // treating the counter addresses as loop invariants of the cycle loop: hoisted, the thousands
    // of 64-bit addresses of a covered design would all be live across the loop and spill.
    VLCU_HD uint32_t stride() const {
        uint32_t s = m_stride;
#ifdef __CUDA_ARCH__
        asm volatile("" : "+r"(s));
#endif
        return s;
    }
    // 32 x 32 -> 64-bit products: one IMAD.WIDE per counter address
    VLCU_HD VlcuCovRef operator[](size_t bin) const {
        return VlcuCovRef{m_basep + static_cast<uint64_t>(stride()) * static_cast<uint32_t>(bin)};
    }
    VLCU_HD VlcuCovPtr operator+(size_t bin) const {
        return VlcuCovPtr{m_basep + static_cast<uint64_t>(stride()) * static_cast<uint32_t>(bin), m_stride};
    }
};
// (VlcuCovArray, the bindable member type, is defined with VlcuState below)

VLCU_HD int vlcu_ctz32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __ffs(static_cast<int>(x)) - 1;
#else
    return __builtin_ctz(x);
#endif
}
VLCU_HD int vlcu_ctz64(uint64_t x) {
#ifdef __CUDA_ARCH__
    return __ffsll(static_cast<long long>(x)) - 1;
#else
    return __builtin_ctzll(x);
#endif
}

//=========================================================================
// Toggle coverage (verilated_cov.h:78-105): two counters per bit, index
// 2*bit + new value, bumped when the bit changed.  Counters wrap (uint32).

// The host loops over all `width` bits and adds 0 or 1; here only the bits that changed are
// visited (same counters, same increments).

// Counter 2*i + (bit i of newData) for every set bit i of `chg` (covp points at the signal's
// counter 0).  One loop over the changed bits of both polarities -- the trip count of a warp is
// the largest number of changed bits in any lane, so splitting rising and falling bits into two
// loops costs more than it saves -- with a minimal body: find the highest set bit, its new value,
// one multiply-add for the address, the RED.
// (A tiled design built --coverage has tens of thousands of toggle points: the loop is a real
// function there, VLCU_COV_NOINLINE defined by the emitted model, or nvcc needs half an hour.)
#if defined(VLCU_COV_NOINLINE) && defined(__CUDACC__)
#define VLCU_COV_FN __device__ __noinline__ static
#define VLCU_COV_CALLER __device__ inline
#else
#define VLCU_COV_FN VLCU_HD
#define VLCU_COV_CALLER VLCU_HD
#endif
VLCU_COV_FN void vlcu_cov_bump_bits(VlcuCovPtr covp, uint32_t chg, uint32_t newData) {
#ifdef __CUDA_ARCH__
    // bytes between consecutive counters: fits 32 bits for any shard below 2^30 instances; the
    // product with the counter number is formed in 64 bits
    const uint32_t step = covp.stride() * 4U;
    char* const basep = reinterpret_cast<char*>(covp.m_basep);
#ifdef VLCU_COV_UNION_LOOP
    // (A/B, round 2) One pass per bit position that changed in ANY lane of the warp, the same
    // position in all lanes, so that every RED instruction is one or two coalesced 128-byte
    // requests instead of 32 scattered L2 atomics.  Measured SLOWER (rv32i_cov campaign: 0.37 vs
    // 0.69 G instance-cycles/s): the union of the lanes' changed bits is ~1.5x the largest single
    // lane's, and the kernel is bound by instruction issue, not by L2 atomic throughput.
    const unsigned lanes = __activemask();
    uint32_t any = __reduce_or_sync(lanes, chg);
    while (any) {
        uint32_t i;
        asm("bfind.u32 %0, %1;" : "=r"(i) : "r"(any));
        any ^= 1U << i;
        if ((chg >> i) & 1U) {
            const uint32_t counter = 2U * i + ((newData >> i) & 1U);
            atomicAdd(reinterpret_cast<uint32_t*>(basep + static_cast<uint64_t>(counter) * step), 1U);
        }
    }
#else
    while (chg) {
        uint32_t i;  // index of the most significant set bit (FLO)
        asm("bfind.u32 %0, %1;" : "=r"(i) : "r"(chg));
        chg ^= 1U << i;
        const uint32_t counter = 2U * i + ((newData >> i) & 1U);
        atomicAdd(reinterpret_cast<uint32_t*>(basep + static_cast<uint64_t>(counter) * step), 1U);
    }
#endif
#else
    for (; chg; chg &= chg - 1U) {
        const int i = vlcu_ctz32(chg);
        ++covp[2 * i + ((newData >> i) & 1U)];
    }
#endif
}
VLCU_HD void vlcu_cov_bump_bits(uint32_t* covp, uint32_t chg, uint32_t newData) {
    for (; chg; chg &= chg - 1U) {
        const int i = vlcu_ctz32(chg);
        ++covp[2 * i + ((newData >> i) & 1U)];
    }
}
template <class T_CovP>
VLCU_HD void VL_COV_TOGGLE_CHG_ST_I(int width, T_CovP covp, IData newData, IData oldData) {
    IData chg = newData ^ oldData;
    if (width < 32) chg &= (1U << width) - 1U;
    if (!chg) return;
    vlcu_cov_bump_bits(covp, chg, newData);
}
template <class T_CovP>
VLCU_HD void VL_COV_TOGGLE_CHG_ST_Q(int width, T_CovP covp, QData newData, QData oldData) {
    QData chg = newData ^ oldData;
    if (width < 64) chg &= (1ULL << width) - 1ULL;
    if (!chg) return;
    const uint32_t chgLo = static_cast<uint32_t>(chg), chgHi = static_cast<uint32_t>(chg >> 32);
    const uint32_t newLo = static_cast<uint32_t>(newData), newHi = static_cast<uint32_t>(newData >> 32);
    vlcu_cov_bump_bits(covp, chgLo, newLo);
    vlcu_cov_bump_bits(covp + 64, chgHi, newHi);
}
// NOTE (bug-compatible): the host runtime's inner loop runs j up to width - 32*w, i.e. past
// 31 for every word but the last, and shifts a 32-bit word by j.  That is undefined in C++;
// as built by the reference's default flags (scalar x86 `shr`, count taken mod 32) the low
// words are therefore counted again into the counters of higher bits: changed bit b of word w
// bumps the counters of bits w*32 + b, w*32 + b + 32, ... below `width`.  Reproduced here.
template <class T_CovP>
VLCU_HD void VL_COV_TOGGLE_CHG_ST_W(int width, T_CovP covp, WDataInP newData, WDataInP oldData) {
    for (int w = 0; w < VL_WORDS_I(width); ++w) {
        EData chg = newData[w] ^ oldData[w];
        const int limit = width - w * 32;  // j < limit
        while (chg) {
            const int b = vlcu_ctz32(chg);
            chg &= chg - 1U;
            const uint32_t bit = (newData[w] >> b) & 1U;
            for (int j = b; j < limit; j += 32) ++covp[(w * 32 + j) * 2 + bit];
        }
    }
}

//=========================================================================
// Per-instance simulation context (the batch dimension of VerilatedContext)

// Status flag bits, also what vlcu_status() reports to the host
enum : uint32_t {
    VLCU_ST_DIDINIT = 1u << 0,  // first eval's static/initial/settle has run (__Vm_didInit)
    VLCU_ST_FINISH = 1u << 1,  // VerilatedContext::gotFinish()
    VLCU_ST_ERROR = 1u << 2,  // VerilatedContext::gotError()
    VLCU_ST_STOPPED = 1u << 3,  // a $stop / failed assertion reached vl_stop
    VLCU_ST_FATAL = 1u << 4,  // VL_FATAL_MT (e.g. DIDNOTCONVERGE)
    VLCU_ST_MSGDROP = 1u << 5,  // more $display/$write messages than the log holds
};
// Fatal codes
enum : uint32_t {
    VLCU_FATAL_NONE = 0,
    VLCU_FATAL_DIDNOTCONVERGE = 1,  // V3SchedUtil.cpp:80-104
    VLCU_FATAL_NBAQ_OVERFLOW = 2,  // VlNBACommitQueue capacity exceeded
    VLCU_FATAL_USER = 3,  // $fatal
};

#define VLCU_MSG_ARGS 8
struct VlcuMsg {  // one logged $display/$write, formatted later on the host
    uint64_t time;
    uint32_t id;  // index into the design's message-format table
    uint32_t nargs;
    uint64_t args[VLCU_MSG_ARGS];
};

// Per-instance context record as stored in HBM (array of structs: touched only at
// kernel entry/exit and on rare events, so not worth a SoA)
struct VlcuCtxRec {
    uint64_t time;
    uint64_t firstFinishTime, firstStopTime, firstFatalTime;
    uint32_t status;
    uint32_t errorCount;
    uint32_t finishCount, stopCount, fatalCount;
    uint32_t firstStopLine, firstFatalLine, fatalCode;
    uint32_t msgCount;
    uint32_t pad_;
};

struct VlcuContext final {
    // Hot, register-resident part
    uint64_t m_time;
    uint64_t m_randSeed;  // base seed ^ global instance index (SURVEY 8d "Seeds")
    uint32_t m_status;
    uint32_t m_randReset;  // 0 zeros, 1 ones, 2 random (verilated.cpp:577-615)
    uint32_t m_msgCount;
    uint32_t m_plusargs;  // bit n set: design's n-th $test$plusargs string was given
    uint32_t m_errorLimit;
    // Cold part, in HBM
    VlcuCtxRec* m_recp;
    uint64_t* m_rngp;  // this instance's $random / $urandom generator state {s0, s1}
    VlcuMsg* m_msgp;  // this instance's slice of the message log
    uint32_t m_msgCap;

    VLCU_HD uint64_t time() const { return m_time; }
    VLCU_HD void timeInc(uint64_t add) { m_time += add; }
    VLCU_HD bool gotFinish() const { return (m_status & VLCU_ST_FINISH) != 0; }
    VLCU_HD bool gotError() const { return (m_status & VLCU_ST_ERROR) != 0; }
    VLCU_HD bool didInit() const { return (m_status & VLCU_ST_DIDINIT) != 0; }
    VLCU_HD uint64_t randSeed() const { return m_randSeed; }
    VLCU_HD uint32_t randReset() const { return m_randReset; }
    VLCU_HD bool testPlusArg(uint32_t index) const { return (m_plusargs >> index) & 1; }

    // vl_finish (verilated.cpp:166-171): sets gotFinish; a serial model runs it
    // immediately (VerilatedThreadMsgQueue::post with mtaskId()==0)
    VLCU_HD void finish(uint32_t /*line*/) {
        if (!m_recp->finishCount) m_recp->firstFinishTime = m_time;
        ++m_recp->finishCount;
        m_status |= VLCU_ST_FINISH;
    }
    // vl_stop_maybe + vl_stop (verilated.cpp:175-190,218-231) with the default
    // errorLimit: never aborts the batch, records the first occurrence
    VLCU_HD void stop(uint32_t line, bool maybe) {
        const uint32_t count = ++m_recp->errorCount;
        if (maybe && count < m_errorLimit) return;
        if (gotFinish()) return;  // $stop after $finish is ignored
        m_status |= VLCU_ST_ERROR | VLCU_ST_FINISH | VLCU_ST_STOPPED;
        if (!m_recp->stopCount) {
            m_recp->firstStopTime = m_time;
            m_recp->firstStopLine = line;
        }
        ++m_recp->stopCount;
    }
    // vl_fatal (verilated.cpp:194-215) without the process exit
    VLCU_HD void fatal(uint32_t line, uint32_t code) {
        m_status |= VLCU_ST_ERROR | VLCU_ST_FINISH | VLCU_ST_FATAL;
        if (!m_recp->fatalCount) {
            m_recp->firstFatalTime = m_time;
            m_recp->firstFatalLine = line;
            m_recp->fatalCode = code;
        }
        ++m_recp->fatalCount;
    }
    // $display/$write/$error text: log (id, args) for host-side formatting
    VLCU_HD void writef(uint32_t id, uint32_t nargs = 0, uint64_t a0 = 0, uint64_t a1 = 0,
                        uint64_t a2 = 0, uint64_t a3 = 0, uint64_t a4 = 0, uint64_t a5 = 0,
                        uint64_t a6 = 0, uint64_t a7 = 0) {
        if (m_msgCount < m_msgCap) {
            VlcuMsg& msg = m_msgp[m_msgCount];
            msg.time = m_time;
            msg.id = id;
            msg.nargs = nargs;
            msg.args[0] = a0, msg.args[1] = a1, msg.args[2] = a2, msg.args[3] = a3;
            msg.args[4] = a4, msg.args[5] = a5, msg.args[6] = a6, msg.args[7] = a7;
        } else {
            m_status |= VLCU_ST_MSGDROP;
        }
        ++m_msgCount;
    }
};

VLCU_HD uint64_t vlcu_rng_rand64(VlcuContext& ctx) {
    VlRNG rng{ctx.m_rngp[0], ctx.m_rngp[1]};
    const uint64_t result = rng.rand64();
    ctx.m_rngp[0] = rng.s0();
    ctx.m_rngp[1] = rng.s1();
    return result;
}
// (function-like macros: the emitted call has no context argument; every device function of a
// model has `vlCtx` in scope)
#define VL_RANDOM_I() static_cast<IData>(vlcu_rng_rand64(vlCtx))
#define VL_RANDOM_Q() static_cast<QData>(vlcu_rng_rand64(vlCtx))
#define VL_RANDOM_W(obits, outwp) vlcu_random_w(vlCtx, (obits), (outwp))
VLCU_HD WDataOutP vlcu_random_w(VlcuContext& ctx, int obits, WDataOutP outwp) {
    for (int i = 0; i < VL_WORDS_I(obits); ++i) outwp[i] = static_cast<EData>(vlcu_rng_rand64(ctx));
    return outwp;  // (last word is unclean, as on the host)
}

// Older hand-written spellings
#define VLCU_FINISH(ctx, line) (ctx).finish(line)
#define VLCU_STOP(ctx, line, maybe) (ctx).stop(line, maybe)
#define VLCU_FATAL(ctx, line, code) (ctx).fatal(line, code)

//=========================================================================
// What the text printed by EmitCFunc expects from "verilated.h", for device code printed by
// V3EmitCuda.  Every emitted function has `VlcuContext& vlCtx` in scope (parameter) and a
// `typedef <return type> VlcuFuncRet`.

// VL_FINISH_MT / VL_STOP_MT / VL_FATAL_MT (V3EmitCFunc.h:1359-1376; include/verilated.cpp:245-266):
// the file and hierarchy strings are compile-time literals that only the host needs
VLCU_HD void vlcu_finish_mt(VlcuContext& ctx, const char*, int line, const char*) {
    ctx.finish(static_cast<uint32_t>(line));
}
VLCU_HD void vlcu_stop_mt(VlcuContext& ctx, const char*, int line, const char*, bool maybe = true) {
    ctx.stop(static_cast<uint32_t>(line), maybe);
}
VLCU_HD void vlcu_fatal_mt(VlcuContext& ctx, const char*, int line, const char*, const char* msgp) {
    // The scheduler's only fatal is the convergence guard (V3SchedUtil.cpp:80-104)
    const bool converge = msgp[0] == 'D' && msgp[1] == 'I' && msgp[2] == 'D' && msgp[3] == 'N';
    ctx.fatal(static_cast<uint32_t>(line), converge ? VLCU_FATAL_DIDNOTCONVERGE : VLCU_FATAL_USER);
}
#define VL_FINISH_MT(...) vlcu_finish_mt(vlCtx, __VA_ARGS__)
#define VL_STOP_MT(...) vlcu_stop_mt(vlCtx, __VA_ARGS__)
// The C++ model aborts the process here; an instance instead leaves the function it is in
// (the region loops are inside _eval / _eval_settle) with its fatal flag set
#define VL_FATAL_MT(...) \
    do { \
        vlcu_fatal_mt(vlCtx, __VA_ARGS__); \
        return VlcuFuncRet(); \
    } while (0)
// $time scaled to the module's time unit (V3EmitCFunc.h:1404-1412; verilated.h VL_TIME_UNITED_Q)
#define VL_TIME_UNITED_Q(scale) (vlCtx.time() / static_cast<QData>(scale))
#define VL_TIME_Q() (vlCtx.time())

// vlSymsp->_vm_contextp__->...  inside literal C text of the scheduler (assert control,
// debug()) resolves against the instance's context through this stand-in for Vtop__Syms
enum class VerilatedAssertCtlQuery : uint8_t {
    ASSERT_CTL_ON,
    ASSERT_CTL_KILL,
    ASSERT_CTL_PASS_ON_VACUOUS,
    ASSERT_CTL_PASS_ON_NONVACUOUS,
    ASSERT_CTL_FAIL_ON,
};
struct VlcuSymsCtxShim final {
    VlcuContext* m_ctxp;
    // VerilatedContext::assertCtlGet (include/verilated.cpp:3180-3198) with the run-time
    // defaults: everything on, nothing killed
    VLCU_HD uint32_t assertCtlGet(VerilatedAssertCtlQuery query, uint32_t types, uint32_t directives) const {
        if (!types || !directives) return 0;
        return query == VerilatedAssertCtlQuery::ASSERT_CTL_KILL ? 0U : 1U;
    }
    VLCU_HD int debug() const { return 0; }
    VLCU_HD const VlcuSymsCtxShim* operator->() const { return this; }
};
struct VlcuSymsShim final {
    VlcuSymsCtxShim _vm_contextp__;
    VLCU_HD explicit VlcuSymsShim(VlcuContext* ctxp)
        : _vm_contextp__{ctxp} {}
    VLCU_HD const char* name() const { return "TOP"; }
};

// Trace sampling: the value of one traced signal, stored at its trace code (what
// VerilatedTraceBuffer::full{Bit,CData,SData,IData,QData,WData} copy to oldp[code],
// include/verilated_trace_imp.h)
VLCU_HD void vlcu_trace_set(uint32_t* codep, IData value) { codep[0] = value; }
VLCU_HD void vlcu_trace_set(uint32_t* codep, QData value) {
    codep[0] = static_cast<uint32_t>(value);
    codep[1] = static_cast<uint32_t>(value >> 32);
}
VLCU_HD void vlcu_trace_set_w(uint32_t* codep, int words, WDataInP valuep) {
    for (int w = 0; w < words; ++w) codep[w] = valuep[w];
}

// $display argument as logged (low 64 bits of a wide value)
template <typename T>
VLCU_HD uint64_t vlcu_msg_arg(T value) {
    return static_cast<uint64_t>(value);
}
template <size_t N>
VLCU_HD uint64_t vlcu_msg_arg(const VlWide<N>& value) {
    return N > 1 ? (static_cast<uint64_t>(value[1]) << 32) | value[0] : value[0];
}

//=========================================================================
// X-initial reset values (verilated.cpp:577-635); 'ctx' supplies what the
// host code reads from Verilated::threadContextp()

VLCU_HD IData vlcu_scoped_rand_reset_i(const VlcuContext& ctx, int obits, uint64_t scopeHash,
                                     uint64_t salt) {
    if (ctx.randReset() == 0) return 0;
    IData data = ~0U;
    if (ctx.randReset() != 1) {
        VlRNG rng{ctx.randSeed() ^ scopeHash ^ salt};
        data = static_cast<IData>(rng.rand64());
    }
    return data & VL_MASK_I(obits);
}
VLCU_HD QData vlcu_scoped_rand_reset_q(const VlcuContext& ctx, int obits, uint64_t scopeHash,
                                     uint64_t salt) {
    if (ctx.randReset() == 0) return 0;
    QData data = ~0ULL;
    if (ctx.randReset() != 1) {
        VlRNG rng{ctx.randSeed() ^ scopeHash ^ salt};
        data = rng.rand64();
    }
    return data & VL_MASK_Q(obits);
}
VLCU_HD WDataOutP vlcu_scoped_rand_reset_w(const VlcuContext& ctx, int obits, WDataOutP outwp,
                                         uint64_t scopeHash, uint64_t salt) {
    const int words = VL_WORDS_I(obits);
    if (ctx.randReset() == 0) {
        for (int i = 0; i < words; ++i) outwp[i] = 0;
    } else if (ctx.randReset() == 1) {
        for (int i = 0; i < words; ++i) outwp[i] = ~0U;
    } else {
        VlRNG rng{ctx.randSeed() ^ scopeHash ^ salt};
        // One draw per 32-bit limb; the upper half of each draw is discarded
        for (int i = 0; i < words; ++i) outwp[i] = static_cast<EData>(rng.rand64());
    }
    outwp[words - 1] &= VL_MASK_E(obits);
    return outwp;
}
VLCU_HD IData vlcu_scoped_rand_reset_assign_i(const VlcuContext& ctx, int obits, uint64_t scopeHash,
                                            uint64_t salt) {
    VlRNG rng{ctx.randSeed() ^ scopeHash ^ salt};
    return static_cast<IData>(rng.rand64()) & VL_MASK_I(obits);
}
VLCU_HD QData vlcu_scoped_rand_reset_assign_q(const VlcuContext& ctx, int obits, uint64_t scopeHash,
                                            uint64_t salt) {
    VlRNG rng{ctx.randSeed() ^ scopeHash ^ salt};
    return rng.rand64() & VL_MASK_Q(obits);
}
VLCU_HD WDataOutP vlcu_scoped_rand_reset_assign_w(const VlcuContext& ctx, int obits,
                                                WDataOutP outwp, uint64_t scopeHash,
                                                uint64_t salt) {
    VlRNG rng{ctx.randSeed() ^ scopeHash ^ salt};
    const int words = VL_WORDS_I(obits);
    for (int i = 0; i < words - 1; ++i) outwp[i] = static_cast<EData>(rng.rand64());
    outwp[words - 1] = static_cast<EData>(rng.rand64()) & VL_MASK_E(obits);
    return outwp;
}

#ifndef VL_ATTR_UNUSED
#define VL_ATTR_UNUSED __attribute__((unused))
#endif
// verilated_funcs.h VL_ZERO_RESET_W: zero every limb
VLCU_HD WDataOutP VL_ZERO_RESET_W(int obits, WDataOutP outwp) {
    for (int i = 0; i < VL_WORDS_I(obits); ++i) outwp[i] = 0;
    return outwp;
}

// Spellings printed by EmitCFunc::emitVarReset (V3EmitCFunc.cpp:524-666): the context is implicit
#define VL_SCOPED_RAND_RESET_I(obits, scopeHash, salt) vlcu_scoped_rand_reset_i(vlCtx, obits, scopeHash, salt)
#define VL_SCOPED_RAND_RESET_Q(obits, scopeHash, salt) vlcu_scoped_rand_reset_q(vlCtx, obits, scopeHash, salt)
#define VL_SCOPED_RAND_RESET_W(obits, outwp, scopeHash, salt) \
    vlcu_scoped_rand_reset_w(vlCtx, obits, outwp, scopeHash, salt)
#define VL_SCOPED_RAND_RESET_ASSIGN_I(obits, scopeHash, salt) \
    vlcu_scoped_rand_reset_assign_i(vlCtx, obits, scopeHash, salt)
#define VL_SCOPED_RAND_RESET_ASSIGN_Q(obits, scopeHash, salt) \
    vlcu_scoped_rand_reset_assign_q(vlCtx, obits, scopeHash, salt)
#define VL_SCOPED_RAND_RESET_ASSIGN_W(obits, outwp, scopeHash, salt) \
    vlcu_scoped_rand_reset_assign_w(vlCtx, obits, outwp, scopeHash, salt)

//=========================================================================
// Structure-of-arrays state block

struct VlcuState final {
    uint8_t* base;  // start of the block (256-byte aligned)
    uint64_t npad;  // instances padded to a multiple of 128 (1 for a packed record)

    template <typename T>
    VLCU_HD T* ptr(uint32_t soaOffset) const {
        return reinterpret_cast<T*>(base + static_cast<uint64_t>(soaOffset) * npad);
    }
    // Scalar member
    template <typename T>
    VLCU_HD T ld(uint32_t soaOffset, uint64_t inst) const {
        return ptr<T>(soaOffset)[inst];
    }
    template <typename T>
    VLCU_HD void st(uint32_t soaOffset, uint64_t inst, T value) const {
        ptr<T>(soaOffset)[inst] = value;
    }
    // Element 'e' of a multi-element member (VlWide limb / small VlUnpacked entry)
    template <typename T>
    VLCU_HD T ldE(uint32_t soaOffset, uint32_t e, uint64_t inst) const {
        return ptr<T>(soaOffset)[static_cast<uint64_t>(e) * npad + inst];
    }
    template <typename T>
    VLCU_HD void stE(uint32_t soaOffset, uint32_t e, uint64_t inst, T value) const {
        ptr<T>(soaOffset)[static_cast<uint64_t>(e) * npad + inst] = value;
    }
    template <size_t N>
    VLCU_HD void ldW(uint32_t soaOffset, uint64_t inst, VlWide<N>& out) const {
        VLCU_UNROLL
        for (size_t e = 0; e < N; ++e) out[e] = ldE<EData>(soaOffset, static_cast<uint32_t>(e), inst);
    }
    template <size_t N>
    VLCU_HD void stW(uint32_t soaOffset, uint64_t inst, const VlWide<N>& in) const {
        VLCU_UNROLL
        for (size_t e = 0; e < N; ++e) stE<EData>(soaOffset, static_cast<uint32_t>(e), inst, in[e]);
    }
};

// A memory too large for thread-private storage: stays in HBM, one private
// copy per instance, laid out instance-major ("array of instances") so that
// the sequential accesses of one instance (instruction fetch, stack) hit the
// same 32-byte sector instead of gathering across an N-instance stride
// (SURVEY 7.3-4).  operator[] hands out real references, so emitted code uses
// it exactly like a VlUnpacked.
template <typename T_Value, size_t N_Depth>
struct VlcuGlobalArray final {
    T_Value* m_basep;  // this instance's N_Depth elements

    VLCU_HD void bind(const VlcuState& st, uint32_t soaOffset, uint64_t inst) {
        m_basep = st.ptr<T_Value>(soaOffset) + inst * N_Depth;
    }
    VLCU_HD T_Value& operator[](size_t index) { return m_basep[index]; }
    VLCU_HD const T_Value& operator[](size_t index) const { return m_basep[index]; }
    // `a = b` between two arrays of the model copies the ELEMENTS (the handle itself is set by
    // bind() only): V3Sched's `__Vtrigprevexpr = signal` relies on it
    VlcuGlobalArray() = default;
    VlcuGlobalArray(const VlcuGlobalArray&) = default;
    VLCU_HD VlcuGlobalArray& operator=(const VlcuGlobalArray& that) {
        assign(that);
        return *this;
    }
    // whole-array change detection / copy (trigger computation, V3SenExprBuilder.h:253-260)
    template <class T_Other>
    VLCU_HD bool neq(const T_Other& that) const {
        bool differs = false;
        for (size_t i = 0; i < N_Depth; ++i) differs |= ((*this)[i] != that[i]);
        return differs;
    }
    template <class T_Other>
    VLCU_HD void assign(const T_Other& that) {
        for (size_t i = 0; i < N_Depth; ++i) (*this)[i] = that[i];
    }
    template <class T_Other>
    VLCU_HD bool operator==(const T_Other& that) const { return !neq(that); }
    template <class T_Other>
    VLCU_HD bool operator!=(const T_Other& that) const { return neq(that); }
    VLCU_HD constexpr size_t size() const { return N_Depth; }
};

// A large array that stays in the SoA state block as it is at rest, [element][instance]: every
// lane of a warp touching the SAME element (a loop index, a constant) is one fully coalesced
// 32 x sizeof(T) byte run.  V3EmitCuda picks this layout for arrays whose run-time indices do not
// depend on design signals, and VlcuGlobalArray (instance-major) for data-dependent indices, where
// consecutive accesses of one instance should share sectors instead.
template <typename T_Value, size_t N_Depth>
struct VlcuStateArray final {
    T_Value* m_basep;  // element 0 of this instance
    uint32_t m_stride;  // padded instance count (elements between consecutive indices)
    VLCU_HD void bind(const VlcuState& st, uint32_t soaOffset, uint64_t inst) {
        m_basep = st.ptr<T_Value>(soaOffset) + inst;
        m_stride = static_cast<uint32_t>(st.npad);
    }
    VLCU_HD T_Value& operator[](size_t index) {
        return m_basep[static_cast<uint64_t>(m_stride) * static_cast<uint32_t>(index)];
    }
    VLCU_HD const T_Value& operator[](size_t index) const {
        return m_basep[static_cast<uint64_t>(m_stride) * static_cast<uint32_t>(index)];
    }
    // `a = b` between two arrays of the model copies the ELEMENTS (the handle itself is set by
    // bind() only): V3Sched's `__Vtrigprevexpr = signal` relies on it
    VlcuStateArray() = default;
    VlcuStateArray(const VlcuStateArray&) = default;
    VLCU_HD VlcuStateArray& operator=(const VlcuStateArray& that) {
        assign(that);
        return *this;
    }
    // whole-array change detection / copy (trigger computation, V3SenExprBuilder.h:253-260)
    template <class T_Other>
    VLCU_HD bool neq(const T_Other& that) const {
        bool differs = false;
        for (size_t i = 0; i < N_Depth; ++i) differs |= ((*this)[i] != that[i]);
        return differs;
    }
    template <class T_Other>
    VLCU_HD void assign(const T_Other& that) {
        for (size_t i = 0; i < N_Depth; ++i) (*this)[i] = that[i];
    }
    template <class T_Other>
    VLCU_HD bool operator==(const T_Other& that) const { return !neq(that); }
    template <class T_Other>
    VLCU_HD bool operator!=(const T_Other& that) const { return neq(that); }
    VLCU_HD constexpr size_t size() const { return N_Depth; }
};

// The coverage counter array of a model image: see VlcuCovPtr
template <size_t N_Bins>
struct VlcuCovArray final : public VlcuCovPtr {
    VLCU_HD void bind(const VlcuState& st, uint32_t soaOffset, uint64_t inst) {
        m_basep = st.ptr<uint32_t>(soaOffset) + inst;
        m_stride = static_cast<uint32_t>(st.npad);
    }
    VLCU_HD constexpr size_t size() const { return N_Bins; }
};

// A small array that is indexed with run-time values (a register file, coverage counters).
// Such an array cannot live in registers; if it were a plain member of the per-thread model
// struct, nvcc would keep the WHOLE struct in local memory with it.  The emitter therefore
// declares the storage separately (T_Design::Locals, thread-private local memory, L1 resident)
// and gives the model struct only this handle, so every other member is promoted to a register.
template <typename T_Value, size_t N_Depth>
struct VlcuLocalArray final {
    T_Value* m_basep;
    VLCU_HD void bind(T_Value* storagep) { m_basep = storagep; }
    VLCU_HD T_Value& operator[](size_t index) { return m_basep[index]; }
    VLCU_HD const T_Value& operator[](size_t index) const { return m_basep[index]; }
    // `a = b` between two arrays of the model copies the ELEMENTS (the handle itself is set by
    // bind() only): V3Sched's `__Vtrigprevexpr = signal` relies on it
    VlcuLocalArray() = default;
    VlcuLocalArray(const VlcuLocalArray&) = default;
    VLCU_HD VlcuLocalArray& operator=(const VlcuLocalArray& that) {
        assign(that);
        return *this;
    }
    // whole-array change detection / copy (trigger computation, V3SenExprBuilder.h:253-260)
    template <class T_Other>
    VLCU_HD bool neq(const T_Other& that) const {
        bool differs = false;
        for (size_t i = 0; i < N_Depth; ++i) differs |= ((*this)[i] != that[i]);
        return differs;
    }
    template <class T_Other>
    VLCU_HD void assign(const T_Other& that) {
        for (size_t i = 0; i < N_Depth; ++i) (*this)[i] = that[i];
    }
    template <class T_Other>
    VLCU_HD bool operator==(const T_Other& that) const { return !neq(that); }
    template <class T_Other>
    VLCU_HD bool operator!=(const T_Other& that) const { return neq(that); }
    VLCU_HD constexpr size_t size() const { return N_Depth; }
    VLCU_HD operator T_Value*() { return m_basep; }
    VLCU_HD operator const T_Value*() const { return m_basep; }
};

// Threads (= instances) per CTA of every design kernel.  65,536 instances = 1024 CTAs = 6.9 per
// SM, all co-resident (one wave); small CTAs keep the tail short when instances stop early.
#ifndef VLCU_BLOCK
#define VLCU_BLOCK 64
#endif

// The same kind of array (small, run-time indexed, private to an instance) kept in SHARED
// memory instead: the emitter declares `T_Value name[N_Depth][VLCU_BLOCK]` in T_Design::Shared
// (one per CTA) and element i of this thread's instance lives at [i][threadIdx.x].  The bank is
// threadIdx.x % 32 whatever the index, so a warp whose lanes index different elements (32
// instances reading 32 different RISC-V registers) is served in ONE conflict-free pass, where the
// local-memory form needs up to 32 L1 sectors.
template <typename T_Value, size_t N_Depth>
struct VlcuSharedArray final {
    T_Value* m_basep;  // &storage[0][threadIdx.x]
    VLCU_HD void bind(T_Value (*storagep)[VLCU_BLOCK], unsigned lane) { m_basep = &storagep[0][lane]; }
    VLCU_HD T_Value& operator[](size_t index) { return m_basep[index * VLCU_BLOCK]; }
    VLCU_HD const T_Value& operator[](size_t index) const { return m_basep[index * VLCU_BLOCK]; }
    // `a = b` between two arrays of the model copies the ELEMENTS (the handle itself is set by
    // bind() only): V3Sched's `__Vtrigprevexpr = signal` relies on it
    VlcuSharedArray() = default;
    VlcuSharedArray(const VlcuSharedArray&) = default;
    VLCU_HD VlcuSharedArray& operator=(const VlcuSharedArray& that) {
        assign(that);
        return *this;
    }
    // whole-array change detection / copy (trigger computation, V3SenExprBuilder.h:253-260)
    template <class T_Other>
    VLCU_HD bool neq(const T_Other& that) const {
        bool differs = false;
        for (size_t i = 0; i < N_Depth; ++i) differs |= ((*this)[i] != that[i]);
        return differs;
    }
    template <class T_Other>
    VLCU_HD void assign(const T_Other& that) {
        for (size_t i = 0; i < N_Depth; ++i) (*this)[i] = that[i];
    }
    template <class T_Other>
    VLCU_HD bool operator==(const T_Other& that) const { return !neq(that); }
    template <class T_Other>
    VLCU_HD bool operator!=(const T_Other& that) const { return neq(that); }
    VLCU_HD constexpr size_t size() const { return N_Depth; }
};

// Multidimensional unpacked arrays (VlUnpacked<VlUnpacked<T, D1>, D0> in the C++ model): stored
// flat, row-major, in one of the 1-D array classes above; m[i0][i1]... walks down a chain of views
// that folds the indices into the flat element number.
template <class T_Flat, size_t... N_Dims>
struct VlcuDimView;
template <class T_Flat, size_t N_Dim>
struct VlcuDimView<T_Flat, N_Dim> final {  // last dimension: yields the element
    T_Flat* m_flatp;
    size_t m_index;  // flat index of this row's element 0, in units of N_Dim
    VLCU_HD auto operator[](size_t i) const -> decltype((*m_flatp)[0]) { return (*m_flatp)[m_index * N_Dim + i]; }
    VLCU_HD constexpr size_t size() const { return N_Dim; }
};
template <class T_Flat, size_t N_Dim, size_t... N_Rest>
struct VlcuDimView<T_Flat, N_Dim, N_Rest...> final {
    T_Flat* m_flatp;
    size_t m_index;
    VLCU_HD VlcuDimView<T_Flat, N_Rest...> operator[](size_t i) const {
        return VlcuDimView<T_Flat, N_Rest...>{m_flatp, m_index * N_Dim + i};
    }
    VLCU_HD constexpr size_t size() const { return N_Dim; }
};
// T_Flat: the 1-D array class holding all elements; N_Dim0 x N_Rest...: the declared dimensions
template <class T_Flat, size_t N_Dim0, size_t... N_Rest>
struct VlcuMdArray final {
    T_Flat m_flat;
    template <class... T_Args>
    VLCU_HD void bind(T_Args... args) { m_flat.bind(args...); }
    VLCU_HD T_Flat& flat() { return m_flat; }
    VLCU_HD const T_Flat& flat() const { return m_flat; }
    VlcuMdArray() = default;
    VlcuMdArray(const VlcuMdArray&) = default;
    VLCU_HD VlcuMdArray& operator=(const VlcuMdArray& that) {  // element copy, see the 1-D classes
        m_flat.assign(that.m_flat);
        return *this;
    }
    VLCU_HD bool neq(const VlcuMdArray& that) const { return m_flat.neq(that.m_flat); }
    VLCU_HD void assign(const VlcuMdArray& that) { m_flat.assign(that.m_flat); }
    VLCU_HD bool operator==(const VlcuMdArray& that) const { return !neq(that); }
    VLCU_HD bool operator!=(const VlcuMdArray& that) const { return neq(that); }
    VLCU_HD VlcuDimView<T_Flat, N_Rest...> operator[](size_t i) { return VlcuDimView<T_Flat, N_Rest...>{&m_flat, i}; }
    VLCU_HD VlcuDimView<const T_Flat, N_Rest...> operator[](size_t i) const {
        return VlcuDimView<const T_Flat, N_Rest...>{&m_flat, i};
    }
    VLCU_HD constexpr size_t size() const { return N_Dim0; }
};

//=========================================================================
// Tiled model image (designs V3EmitCuda prints in the tiled form: module hierarchy kept, or more
// state than a thread's registers hold).  The state block is an array of TILES; a tile is the
// design's <prefix>__Syms struct -- one struct per scope, as in the C++ model -- in which every
// member is widened to VLCU_TILE slots, one per instance of the tile:
//
//     instance i, member m of scope s  ->  base + (i / VLCU_TILE) * sizeof(Syms)
//                                               + offsetof(Syms, s) + offsetof(Class(s), m)
//                                               + (i % VLCU_TILE) * sizeof(T)
//
// A tile is a warp: the 32 lanes' accesses to a scalar member are one contiguous 32*sizeof(T)-byte
// run at a compile-time offset from the tile's base, so a member access is ONE load or store with
// an immediate offset -- no address arithmetic per access, and all members provably distinct.  A
// warp's whole working set is one contiguous block of HBM (sizeof(Syms) bytes).
#ifndef VLCU_TILE
#define VLCU_TILE 32
#endif
#if defined(__CUDACC__)
#define VLCU_DN __device__ __noinline__
#else
#define VLCU_DN
#endif

// This thread's slot in its tile (VLCU_BLOCK is a multiple of VLCU_TILE)
VLCU_HD uint32_t vlcu_lane() {
#ifdef __CUDA_ARCH__
    return threadIdx.x % VLCU_TILE;
#else
    return 0;
#endif
}

// Ask for the line holding *p ahead of its use (the next leaf's gather, issued by the caller one
// call ahead).  VLCU_PREFETCH_LEVEL: 1 = into L1 (default), 2 = into L2 only, 0 = off.
#ifndef VLCU_PREFETCH_LEVEL
#define VLCU_PREFETCH_LEVEL 1
#endif
template <typename T>
VLCU_HD void vlcu_prefetch(const T* p) {
#if defined(__CUDA_ARCH__) && VLCU_PREFETCH_LEVEL == 1
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#elif defined(__CUDA_ARCH__) && VLCU_PREFETCH_LEVEL == 2
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

// One value per instance of the tile.  Emitted code names the slot explicitly (`sig.v[vlLane]`);
// the conversions are for the few places where EmitCFunc prints a variable without going through
// the AstVarRef visitor (V3EmitCFunc.h:320-345: wide constants assigned in place;
// V3EmitCFunc.cpp:524-666: the literal text of _ctor_var_reset).
template <typename T_Value>
struct VlcuLanes final {
    T_Value v[VLCU_TILE];
    VLCU_HD operator T_Value&() { return v[vlcu_lane()]; }
    VLCU_HD operator const T_Value&() const { return v[vlcu_lane()]; }
    template <typename T_Rhs>
    VLCU_HD void operator=(const T_Rhs& rhs) { v[vlcu_lane()] = static_cast<T_Value>(rhs); }
    template <typename T_Index>
    VLCU_HD auto operator[](T_Index index) -> decltype((v[0][index])) { return v[vlcu_lane()][index]; }
};
template <size_t N_Words>
struct VlcuLanes<VlWide<N_Words>> final {
    VlWide<N_Words> v[VLCU_TILE];
    VLCU_HD operator VlWide<N_Words>&() { return v[vlcu_lane()]; }
    VLCU_HD operator EData*() { return v[vlcu_lane()].data(); }
    VLCU_HD operator const EData*() const { return v[vlcu_lane()].data(); }
    VLCU_HD EData& operator[](size_t index) { return v[vlcu_lane()][index]; }
};

// uint32_t __Vcoverage[N] of a module class in a TILE: one uint64 counter per bin for the tile's
// 32 instances together.  A design of this class has ~10^5 cover points; one counter per instance
// would be 400 KB per instance -- gigabytes of counters bumped at random, every RED a DRAM
// read-modify-write (measured: 0.15 M instance-cycles/s, below one CPU thread) -- while what leaves a
// campaign is the SUM over instances (verilated_cov.cpp:368-438 merges by key).  So the warp
// aggregates: a cover point adds, once, the number of lanes that reached it; a toggle point
// counts, per bit position that changed in any lane, the lanes where it rose and where it fell
// (two ballots).  The counters are the exact sums of the per-instance counters the C++ model keeps.
// The lanes that work on one cover point together.  Where the bump functions are inlined, every
// point has its own code and the lanes executing it together are at that point.  As real functions
// (VLCU_COV_NOINLINE) they have ONE copy of code: lanes that come from different points -- the two
// sides of an if, say -- may be scheduled through it together, and __activemask() alone would mix
// their counters (a narrow signal's lanes would add past its counters).  There the lanes are
// grouped by the counter they were given: one MATCH.ANY.
#ifdef __CUDA_ARCH__
VLCU_D unsigned vlcu_cov_lanes(const unsigned long long* p) {
#if (defined(VLCU_COV_NOINLINE) || defined(VLCU_COV_GROUP_LANES)) && !defined(VLCU_COV_NO_MATCH)  // (NO_MATCH: tests/optest shows the hazard)
    return __match_any_sync(__activemask(), reinterpret_cast<unsigned long long>(p));
#else
    (void)p;
    return __activemask();
#endif
}
#endif
// The lanes that entered a bump function leave it together -- where it is a real function; inlined,
// the code that follows reconverges them anyway, and the barrier would only cost
#ifdef __CUDA_ARCH__
VLCU_D void vlcu_cov_leave(unsigned lanes) {
#ifdef VLCU_COV_NOINLINE
    __syncwarp(lanes);
#else
    (void)lanes;
#endif
}
#endif
// One lane of `lanes` (the lanes executing this together; the lowest) adds `value` to a counter;
// a zero is not sent.  (As the LAST statement of a real function this is a divergent return: the
// other lanes leave while the elected one still adds, and nothing in the caller waits for it --
// the warp would run the rest of the launch in pieces.  The bump functions therefore end with a
// __syncwarp over the lanes that entered.)  (elect.sync with a run-time mask costs a WARPSYNC/ENDCOLLECTIVE region of
// its own; two integer operations find the lowest lane as well.)
#ifdef __CUDA_ARCH__
VLCU_D void vlcu_red_elected(unsigned lanes, unsigned long long* p, uint32_t value) {
    const uint32_t mine = (lanes & (0U - lanes)) == (1U << (threadIdx.x & 31U)) ? value : 0U;
    if (mine) atomicAdd(p, static_cast<unsigned long long>(mine));
}
#endif
struct VlcuTileCovRef final {
    unsigned long long* m_p;
    VLCU_HD void operator++() const {
#ifdef __CUDA_ARCH__
        // the lanes reaching this point together bump the same counter (a line or branch point of
        // the code they all run, for the scope they all are in): one add of their number.  (Any
        // split of the lanes into groups gives the same sum.)
        const unsigned lanes = vlcu_cov_lanes(m_p);
        vlcu_red_elected(lanes, m_p, __popc(lanes));
#else
        ++*m_p;
#endif
    }
};
struct VlcuTileCovPtr final {
    unsigned long long* m_p;
    VLCU_HD VlcuTileCovRef operator[](size_t bin) const { return VlcuTileCovRef{m_p + bin}; }
    VLCU_HD VlcuTileCovPtr operator+(size_t bin) const { return VlcuTileCovPtr{m_p + bin}; }
};
template <size_t N_Bins>
struct VlcuTileCov final {
    unsigned long long c[N_Bins];
    VLCU_HD VlcuTileCovRef operator[](size_t bin) { return VlcuTileCovRef{&c[bin]}; }
    VLCU_HD VlcuTileCovPtr operator+(size_t bin) { return VlcuTileCovPtr{&c[bin]}; }
    // (reset: the lanes that are there -- the last tile of a shard may be partial -- share the work)
    VLCU_HD void zero(uint32_t lane) {
#ifdef __CUDA_ARCH__
        const unsigned lanes = __activemask();
        const uint32_t rank = __popc(lanes & ((1U << lane) - 1U)), step = __popc(lanes);
        for (size_t b = rank; b < N_Bins; b += step) c[b] = 0ULL;
#else
        for (size_t b = lane; b < N_Bins; b += VLCU_TILE) c[b] = 0ULL;
#endif
    }
};
// The wide form (verilated_cov.h:94-105 with its quirk, see VL_COV_TOGGLE_CHG_ST_W above: changed
// bit b of word w bumps the counters of bits w*32+b, w*32+b+32, ... below `width`), warp-cooperative:
// every loop runs the same trips in all lanes, so a 2,048-bit vector does not serialise the warp
// (left per lane, the lanes walk different bits and the warp runs 1.4 lanes wide).
VLCU_COV_FN void vlcu_cov_bump_wide(VlcuTileCovPtr covp, int width, WDataInP newData, WDataInP oldData) {
#ifdef __CUDA_ARCH__
    const unsigned lanes = vlcu_cov_lanes(covp.m_p);
    const bool leader = (threadIdx.x & 31U) == static_cast<unsigned>(__ffs(lanes) - 1);
    for (int w = 0; w < VL_WORDS_I(width); ++w) {
        const EData chg = newData[w] ^ oldData[w];
        const EData rise = chg & newData[w];
        const int limit = width - w * 32;  // j < limit
        uint32_t any = __reduce_or_sync(lanes, chg);
        while (any) {
            const int b = __ffs(static_cast<int>(any)) - 1;
            any &= any - 1U;
            const unsigned changed = __ballot_sync(lanes, (chg >> b) & 1U);
            const unsigned rose = __ballot_sync(lanes, (rise >> b) & 1U);
            if (leader) {
                const unsigned long long up = __popc(rose), down = __popc(changed) - __popc(rose);
                for (int j = b; j < limit; j += 32) {
                    if (up) atomicAdd(covp.m_p + (w * 32 + j) * 2 + 1, up);
                    if (down) atomicAdd(covp.m_p + (w * 32 + j) * 2, down);
                }
            }
        }
    }
    vlcu_cov_leave(lanes);  // leave together, see vlcu_red_elected
#else
    for (int w = 0; w < VL_WORDS_I(width); ++w) {
        EData chg = newData[w] ^ oldData[w];
        const int limit = width - w * 32;
        for (; chg; chg &= chg - 1U) {
            const int b = vlcu_ctz32(chg);
            const uint32_t bit = (newData[w] >> b) & 1U;
            for (int j = b; j < limit; j += 32) ++covp[(w * 32 + j) * 2 + bit];
        }
    }
#endif
}
VLCU_COV_CALLER void VL_COV_TOGGLE_CHG_ST_W(int width, VlcuTileCovPtr covp, WDataInP newData, WDataInP oldData) {
    vlcu_cov_bump_wide(covp, width, newData, oldData);
}
// 32 x 32 bit-matrix transpose across the lanes of a FULL warp: lane l gives row l and receives
// column l (bit k of the result = bit l of lane k's word).  Five butterfly stages of one shuffle,
// one rotate and one LOP3 each.
#ifdef __CUDA_ARCH__
VLCU_D uint32_t vlcu_warp_transpose32(uint32_t x) {
    const uint32_t lane = threadIdx.x & 31U;
#pragma unroll
    for (uint32_t j = 16; j >= 1; j >>= 1) {
        const uint32_t m = j == 16 ? 0xffff0000U : j == 8 ? 0xff00ff00U : j == 4 ? 0xf0f0f0f0U : j == 2 ? 0xccccccccU : 0xaaaaaaaaU;
        const uint32_t y = __shfl_xor_sync(0xffffffffU, x, j);
        const bool upper = lane & j;
        const uint32_t t = __funnelshift_l(y, y, upper ? 32U - j : j);
        const uint32_t keep = upper ? m : ~m;
        x = (x & keep) | (t & ~keep);
    }
    return x;
}
#endif
// The condition of a toggle point, IF(orig ^ change) { INC; change = orig; } (V3Clock.cpp:69-97),
// as the emitter prints it for per-tile counters: the lanes of the warp enter the point together
// when any of them saw a change (an unchanged lane bumps nothing and rewrites its shadow copy with
// the same value), so the bump below can work across the whole warp.
VLCU_HD bool vlcu_cov_any(bool changed) {
#ifdef __CUDA_ARCH__
    return __any_sync(__activemask(), changed);
#else
    return changed;
#endif
}
// Counter 2*i + (bit i of newData) for every bit i of `chg`, summed over the lanes of the warp.
// Full warp (the normal case, see vlcu_cov_any): the changed and the risen masks are transposed
// across the lanes, lane b then holds who changed bit b -- two popcounts give its two counters --
// and the 64 counters go out as two contiguous runs of 32.  The cost does not depend on how many
// bits changed; the per-bit ballot loop is kept for a divergent warp.
VLCU_COV_FN void vlcu_cov_bump_bits(VlcuTileCovPtr covp, uint32_t chg, uint32_t newData) {
#ifdef __CUDA_ARCH__
    const unsigned lanes = vlcu_cov_lanes(covp.m_p);
    const uint32_t rise = chg & newData;
    if (lanes == 0xffffffffU) {
        const uint32_t lane = threadIdx.x & 31U;
        const uint32_t chgT = vlcu_warp_transpose32(chg), riseT = vlcu_warp_transpose32(rise);
        const uint32_t up = __popc(riseT), down = __popc(chgT) - up;
        // counter k (bit k/2, direction k&1) and counter 32+k to lane k
        const uint32_t both = (up << 8) | down;
        const uint32_t lo = __shfl_sync(0xffffffffU, both, lane >> 1);
        const uint32_t hi = __shfl_sync(0xffffffffU, both, 16U + (lane >> 1));
        const uint32_t addLo = (lane & 1U) ? (lo >> 8) : (lo & 0xffU);
        const uint32_t addHi = (lane & 1U) ? (hi >> 8) : (hi & 0xffU);
        if (addLo) atomicAdd(covp.m_p + lane, static_cast<unsigned long long>(addLo));
        if (addHi) atomicAdd(covp.m_p + 32U + lane, static_cast<unsigned long long>(addHi));
        vlcu_cov_leave(0xffffffffU);
        return;
    }
    const bool leader = (threadIdx.x & 31U) == static_cast<unsigned>(__ffs(lanes) - 1);
    uint32_t any = __reduce_or_sync(lanes, chg);
    while (any) {
        uint32_t i;
        asm("bfind.u32 %0, %1;" : "=r"(i) : "r"(any));
        any ^= 1U << i;
        const unsigned changed = __ballot_sync(lanes, (chg >> i) & 1U);
        const unsigned rose = __ballot_sync(lanes, (rise >> i) & 1U);
        if (leader) {
            const unsigned up = __popc(rose), down = __popc(changed) - up;
            if (up) atomicAdd(covp.m_p + 2U * i + 1U, static_cast<unsigned long long>(up));
            if (down) atomicAdd(covp.m_p + 2U * i, static_cast<unsigned long long>(down));
        }
    }
    vlcu_cov_leave(lanes);
#else
    for (; chg; chg &= chg - 1U) {
        const int i = vlcu_ctz32(chg);
        ++covp[2 * i + ((newData >> i) & 1U)];
    }
#endif
}
// Line, branch and expression points of one run of a function: the emitted code sets bit
// (bin % 32) of a local mask per window of 32 bins when a point is passed and hands the mask over
// here when the function ends: transposed across the warp, lane b holds who passed bin b.
VLCU_COV_FN void vlcu_cov_flush_hits(VlcuTileCovPtr covp, uint32_t hits) {
#ifdef __CUDA_ARCH__
    const unsigned lanes = vlcu_cov_lanes(covp.m_p);
    if (lanes == 0xffffffffU) {
        if (!__any_sync(0xffffffffU, hits != 0U)) return;
        const uint32_t passed = __popc(vlcu_warp_transpose32(hits));
        if (passed) atomicAdd(covp.m_p + (threadIdx.x & 31U), static_cast<unsigned long long>(passed));
        vlcu_cov_leave(0xffffffffU);
        return;
    }
    uint32_t any = __reduce_or_sync(lanes, hits);
    while (any) {
        uint32_t i;
        asm("bfind.u32 %0, %1;" : "=r"(i) : "r"(any));
        any ^= 1U << i;
        vlcu_red_elected(lanes, covp.m_p + i, __popc(__ballot_sync(lanes, (hits >> i) & 1U)));
    }
    vlcu_cov_leave(lanes);
#else
    for (; hits; hits &= hits - 1U) ++covp[vlcu_ctz32(hits)];
#endif
}
// One-bit signals (most toggle points of a design): two ballots, no loop
VLCU_COV_FN void vlcu_cov_bump_bit(VlcuTileCovPtr covp, uint32_t chg, uint32_t newData) {
#ifdef __CUDA_ARCH__
    const unsigned lanes = vlcu_cov_lanes(covp.m_p);
    const unsigned changed = __ballot_sync(lanes, chg & 1U);
    const unsigned rose = __ballot_sync(lanes, chg & newData & 1U);
    const unsigned up = __popc(rose);
    vlcu_red_elected(lanes, covp.m_p + 1, up);
    vlcu_red_elected(lanes, covp.m_p, __popc(changed) - up);
    vlcu_cov_leave(lanes);
#else
    if (chg & 1U) ++covp[newData & 1U];
#endif
}
VLCU_COV_CALLER void VL_COV_TOGGLE_CHG_ST_I(int width, VlcuTileCovPtr covp, IData newData, IData oldData) {
    IData chg = newData ^ oldData;
    if (width < 32) chg &= (1U << width) - 1U;
    if (width == 1) {
        vlcu_cov_bump_bit(covp, chg, newData);
    } else {
        vlcu_cov_bump_bits(covp, chg, newData);
    }
}

VLCU_COV_CALLER void VL_COV_TOGGLE_CHG_ST_Q(int width, VlcuTileCovPtr covp, QData newData, QData oldData) {
    QData chg = newData ^ oldData;
    if (width < 64) chg &= (1ULL << width) - 1ULL;
    vlcu_cov_bump_bits(covp, static_cast<uint32_t>(chg), static_cast<uint32_t>(newData));
    vlcu_cov_bump_bits(covp + 64, static_cast<uint32_t>(chg >> 32), static_cast<uint32_t>(newData >> 32));
}

// The counters of a register-form design kept per WARP of 32 instances ([warp][bin] uint64 in the
// block of __Vcoverage; -cuda-cover-per-instance keeps VlcuCovArray instead): `__Vcoverage + n` and
// `++__Vcoverage[n]` of the emitted code land on the per-tile forms above.
template <size_t N_Bins>
struct VlcuWarpCovArray final {
    unsigned long long* m_p;  // counter 0 of this instance's warp
    VLCU_HD void bind(const VlcuState& st, uint32_t soaOffset, uint64_t inst) {
        m_p = reinterpret_cast<unsigned long long*>(st.ptr<uint32_t>(soaOffset)) + (inst / VLCU_TILE) * N_Bins;
    }
    VLCU_HD VlcuTileCovRef operator[](size_t bin) const { return VlcuTileCovRef{m_p + bin}; }
    VLCU_HD VlcuTileCovPtr operator+(size_t bin) const { return VlcuTileCovPtr{m_p + bin}; }
    // (reset: the lanes that are there share the work; `bins` <= N_Bins is the real count)
    VLCU_HD void zero(uint32_t bins) const {
#ifdef __CUDA_ARCH__
        const unsigned lanes = __activemask();
        const uint32_t rank = __popc(lanes & ((1U << (threadIdx.x & 31U)) - 1U)), step = __popc(lanes);
        for (uint32_t b = rank; b < bins; b += step) m_p[b] = 0ULL;
#else
        for (uint32_t b = 0; b < bins; ++b) m_p[b] = 0ULL;
#endif
    }
    VLCU_HD constexpr size_t size() const { return N_Bins; }
};

// A large array kept [element][lane]: an index that is the same in all lanes of a warp (a loop
// counter) touches one contiguous run.  `arr.lane(vlLane)[i]` is element i of this instance.
template <typename T_Value, size_t N_Depth>
struct VlcuLaneArray final {
    T_Value m[N_Depth][VLCU_TILE];
    struct View final {
        T_Value* m_basep;  // &m[0][lane]
        VLCU_HD T_Value& operator[](size_t index) const { return m_basep[index * VLCU_TILE]; }
        VLCU_HD constexpr size_t size() const { return N_Depth; }
        template <class T_Other>
        VLCU_HD bool neq(const T_Other& that) const {
            bool differs = false;
            for (size_t i = 0; i < N_Depth; ++i) differs |= ((*this)[i] != that[i]);
            return differs;
        }
        template <class T_Other>
        VLCU_HD void assign(const T_Other& that) const {
            for (size_t i = 0; i < N_Depth; ++i) (*this)[i] = that[i];
        }
        // `a = b` between two arrays of the model copies the ELEMENTS
        VLCU_HD const View& operator=(const View& that) const {
            assign(that);
            return *this;
        }
        template <class T_Other>
        VLCU_HD bool operator!=(const T_Other& that) const { return neq(that); }
        template <class T_Other>
        VLCU_HD bool operator==(const T_Other& that) const { return !neq(that); }
    };
    VLCU_HD View lane(uint32_t lane) { return View{&m[0][lane]}; }
    VLCU_HD T_Value& operator[](size_t index) { return m[index][vlcu_lane()]; }
};

// A cell of a module class (the C++ model holds `Vtop_sub* cell`, V3EmitCHeaders.cpp:63-72): the
// child scope's struct lies at the same distance from the parent scope's struct in every tile and
// for every instance of the parent class (V3EmitCuda orders Syms depth first), so `cell->sig` is
// pointer arithmetic on `this`.  T_Info (defined once Syms is complete) gives the member's own
// offset in the parent and the distance between the two structs.
template <class T_Child, class T_Info>
struct VlcuCell final {
    VLCU_HD T_Child* operator->() {
        return reinterpret_cast<T_Child*>(reinterpret_cast<char*>(this) - T_Info::kSelf + T_Info::kDelta);
    }
    VLCU_HD operator T_Child*() { return operator->(); }
};

// Shared code naming another scope: the struct of class T_Class that lies T_Info::kDelta bytes from
// this function's own scope struct in every tile (V3EmitCuda "sharing between instances")
#define VLCU_REL(T_Class, T_Info) \
    (*reinterpret_cast<T_Class*>(reinterpret_cast<char*>(vlSelf) + T_Info::kDelta))

// vlSymsp->_vm_contextp__->... in literal C text of a tiled model (cf. VlcuSymsCtxShim): the
// queries have run-time-default answers and need no context
struct VlcuSymsCtxTiled final {
    VLCU_HD uint32_t assertCtlGet(VerilatedAssertCtlQuery query, uint32_t types, uint32_t directives) const {
        if (!types || !directives) return 0;
        return query == VerilatedAssertCtlQuery::ASSERT_CTL_KILL ? 0U : 1U;
    }
    VLCU_HD int debug() const { return 0; }
    VLCU_HD const VlcuSymsCtxTiled* operator->() const { return this; }
};

// Byte offset, inside a tile, of lane 0 of member `member` of scope `scope`
#define VLCU_TILE_OFF(Syms, scope, member) \
    static_cast<uint32_t>(__builtin_offsetof(Syms, scope) + __builtin_offsetof(decltype(Syms::scope), member))

// Everything a kernel needs to find one model's data
struct VlcuModelDev final {
    VlcuState state;
    VlcuCtxRec* ctxRecs;
    VlcuMsg* msgs;
    uint64_t n;  // live instances in this shard
    uint64_t instBase;  // global index of this shard's instance 0 (multi-GPU sharding)
    uint64_t seedBase;
    uint32_t randReset;
    uint32_t msgCap;
    uint32_t plusargs;
    uint32_t errorLimit;
    uint64_t* rng;  // [instance][2]: $random / $urandom generator state
};

// $random / $urandom (VL_RANDOM_I/Q/W, verilated_funcs.h:90-91, verilated.cpp:537-555).  The
// host draws from a thread-local generator that is reseeded from the context's seed whenever a
// seed is set (verilated.cpp:537-549), i.e. with one model per thread each model has its own
// stream seeded with randSeedDefault64() (verilated.cpp:3701-3707).  Here each instance owns its
// generator state (two words in HBM, touched only when the design calls $random).
VLCU_HD void vlcu_rng_seed(uint64_t* statep, uint64_t randSeed) {
    // randSeed() is an int: the 64-bit seed is (seed << 32) ^ seed with sign extension
    const uint64_t s = static_cast<uint64_t>(static_cast<int64_t>(static_cast<int32_t>(randSeed)));
    const VlRNG rng{(s << 32) ^ s};
    statep[0] = rng.s0();
    statep[1] = rng.s1();
}
VLCU_HD uint64_t vlcu_rng_rand64(VlcuContext& ctx);

VLCU_HD void vlcu_ctx_load(VlcuContext& ctx, const VlcuModelDev& m, uint64_t inst) {
    VlcuCtxRec* const recp = &m.ctxRecs[inst];
    ctx.m_recp = recp;
    ctx.m_rngp = m.rng + 2 * inst;
    ctx.m_time = recp->time;
    ctx.m_status = recp->status;
    ctx.m_msgCount = recp->msgCount;
    ctx.m_randSeed = m.seedBase ^ (m.instBase + inst);
    ctx.m_randReset = m.randReset;
    ctx.m_plusargs = m.plusargs;
    ctx.m_errorLimit = m.errorLimit;
    ctx.m_msgp = m.msgs + inst * m.msgCap;
    ctx.m_msgCap = m.msgCap;
}
VLCU_HD void vlcu_ctx_store(const VlcuContext& ctx) {
    ctx.m_recp->time = ctx.m_time;
    ctx.m_recp->status = ctx.m_status;
    ctx.m_recp->msgCount = ctx.m_msgCount;
}

//=========================================================================
// Eval driver: Vtop::eval_step() (V3EmitCModel.cpp:403-437) for one instance

template <class T_Design>
VLCU_D void vlcu_eval_step(typename T_Design::Root* rootp, VlcuContext& ctx) {
    if (VL_UNLIKELY(!ctx.didInit())) {
        ctx.m_status |= VLCU_ST_DIDINIT;
        T_Design::eval_static(rootp, ctx);
        T_Design::eval_initial(rootp, ctx);
        T_Design::eval_settle(rootp, ctx);
    }
    T_Design::eval(rootp, ctx);
}

// Running 64-bit signature of the output ports; lets a K-cycle fused run be checked against the
// oracle without storing every port value of every instance every cycle.  Per eval the output
// words are folded into a 32-bit FNV-1a accumulator (one LOP3 + one IMAD per word), which is then
// mixed into the 64-bit signature once (FNV-1a, 64-bit prime).
#define VLCU_SIG_ACC_INIT 0x811c9dc5U
VLCU_HD uint32_t vlcu_sig_word(uint32_t acc, uint32_t word) { return (acc ^ word) * 0x01000193U; }
VLCU_HD uint64_t vlcu_sig_mix(uint64_t sig, uint32_t acc) {
    return (sig ^ acc) * 0x100000001b3ULL;
}
#define VLCU_SIG_INIT 0xcbf29ce484222325ULL

#endif  // guard
