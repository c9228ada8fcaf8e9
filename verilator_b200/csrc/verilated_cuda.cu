// -*- mode: C++; c-file-style: "cc-mode" -*-
//*************************************************************************
//
// verilated_cuda.cu -- host runtime of the B200 batched evaluator and the
// implementation of the C ABI in include/verilated_cuda_abi.h.
//
// Role in the reference's layering: this is the sibling of
// include/verilated.cpp for models printed by V3EmitCuda -- it owns the
// batched VerilatedContext (one VlcuCtxRec per instance), the structure-of-
// arrays design state in HBM, the port ingest/egress transposes, the
// device-side coverage reduction (+ NCCL sum across shards) and the
// coverage.dat writer (restating include/verilated_cov.cpp:143-171,368-438).
//
// There is deliberately no CPU execution path: every entry point that
// computes launches a kernel, and vlcu_create() fails without a device.
//
//*************************************************************************
#define VLCU_HOST_RUNTIME 1
#include "verilated_cuda_abi.h"
#include "verilated_cuda_host.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>  // types only; the library is dlopen()ed on first use

#include <cctype>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <iterator>
#include <map>
#include <sstream>
#include <string>
#include <map>
#include <vector>

static_assert(sizeof(VlcuCtxRec) == sizeof(vlcu_status_t), "status record layouts must agree");
static_assert(sizeof(VlcuMsg) == sizeof(vlcu_msg_t), "message record layouts must agree");
static_assert(VLCU_RUN_TB_BEGIN == VLCU_RUNF_TB_BEGIN, "flag values must agree");
static_assert(VLCU_RUN_STOP_ON_FINISH == VLCU_RUNF_STOP_ON_FINISH, "flag values must agree");
static_assert(VLCU_TBF_FUZZ_RESET_LOW == 1 && VLCU_TBF_FUZZ_RESET_HIGH == 2 && VLCU_TBF_FUZZ_UNIFORM == 4,
              "fuzz testbench flags (verilated_cuda_launch.h VLCU_FUZZ_*) must agree");
static_assert(sizeof(ncclUniqueId) == VLCU_NCCL_ID_BYTES, "NCCL unique id size");

//=========================================================================
// Errors

static thread_local std::string t_lastError;

static int fail(int code, const std::string& msg) {
    t_lastError = msg;
    return code;
}
static int failCuda(cudaError_t e, const char* what) {
    return fail(VLCU_E_CUDA, std::string{what} + ": " + cudaGetErrorString(e));
}
#define CUDA_TRY(expr) \
    do { \
        const cudaError_t e_ = (expr); \
        if (e_ != cudaSuccess) return failCuda(e_, #expr); \
    } while (0)

//=========================================================================
// Design registry

static std::vector<VlcuDesignOps*>& registry() {
    static std::vector<VlcuDesignOps*> s;
    return s;
}
int vlcu_register_design(VlcuDesignOps* opsp) {
    registry().push_back(opsp);
    return static_cast<int>(registry().size());
}

//=========================================================================
// NCCL, bound lazily so single-GPU use needs no NCCL at all

struct NcclApi final {
    void* libp = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load(std::string& err) {
        if (libp) return true;
        // RTLD_NOLOAD first: reuse a copy the process (e.g. torch) already mapped
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            libp = dlopen(name, RTLD_NOW | RTLD_NOLOAD);
            if (libp) break;
        }
        if (!libp) {
            for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
                libp = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
                if (libp) break;
            }
        }
        if (!libp) {
            err = std::string{"cannot load libnccl: "} + dlerror();
            return false;
        }
#define BIND(field, sym) \
    field = reinterpret_cast<decltype(field)>(dlsym(libp, sym)); \
    if (!field) { \
        err = std::string{"libnccl lacks "} + sym; \
        return false; \
    }
        BIND(GetUniqueId, "ncclGetUniqueId");
        BIND(CommInitRank, "ncclCommInitRank");
        BIND(CommDestroy, "ncclCommDestroy");
        BIND(AllReduce, "ncclAllReduce");
        BIND(GetErrorString, "ncclGetErrorString");
#undef BIND
        return true;
    }
};
static NcclApi& nccl() {
    static NcclApi s;
    return s;
}

//=========================================================================
// The model handle

struct vlcu_model final {
    const VlcuDesignOps* opsp = nullptr;
    const VlcuDesignInfo* infop = nullptr;
    int device = 0;
    uint64_t n = 0, npad = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t evStart = nullptr, evStop = nullptr;
    uint8_t* dState = nullptr;
    VlcuCtxRec* dCtx = nullptr;
    uint64_t* dRng = nullptr;  // [npad][2] $random generator states
    VlcuMsg* dMsgs = nullptr;
    uint64_t* dSigs = nullptr;
    uint64_t* dEvalsDone = nullptr;
    uint64_t* dCovSums = nullptr;
    uint32_t* dCovBinOffset = nullptr;  // tiled designs: byte offset in a tile of every counter's lane 0
    VlcuModelDev dev{};
    uint64_t launches = 0;
    uint64_t evalIndex = 0;  // evals issued so far (VlcuRunDev::evalBase)
    bool haveTiming = false;
    ncclComm_t comm = nullptr;
    int commRank = 0, commWorld = 1;
    std::vector<int> ports;  // member indices of primary IO, in declaration order
    // Reusable device staging
    uint8_t* dStage = nullptr;
    size_t stageBytes = 0;
    uint8_t* dFlush = nullptr;  // L2 flush scratch (vlcu_flush_l2)
    // Asynchronous input staging (vlcu_stage_member / vlcu_commit_staged)
    struct Staged final {
        uint8_t* dBuf = nullptr;  // device copy of the host records, [inst][elems]
        size_t bytes = 0;
        bool pending = false;  // a copy was issued and not yet committed
        bool mask = false;
    };
    size_t pipelineChunkBytes = 64u << 20;  // vlcu_run chunking of host stimulus; 0 = off
    cudaStream_t copyStream = nullptr;
    cudaEvent_t evCopied = nullptr;  // last staged copy finished (copy stream)
    cudaEvent_t evConsumed = nullptr;  // last commit finished reading the staging buffers
    bool haveConsumed = false;
    std::map<int, Staged> staged;  // by member index
    // Asynchronous result read-back (vlcu_results_begin / vlcu_results_wait)
    cudaStream_t resStream = nullptr;
    cudaEvent_t evCaptured = nullptr;  // compute stream: the results were copied aside
    cudaEvent_t evFetched = nullptr;  // result stream: the previous read-back reached the host
    bool fetchPending = false;
    uint8_t* dRes = nullptr;  // device-side copy of the results being read back
    size_t resBytes = 0;
    // Device scratch of vlcu_run (host-side stimulus / observation / trace), kept between calls
    // and grown on demand: a testbench that calls vlcu_run per step pays no cudaMalloc/cudaFree
    enum { RB_STIM, RB_OBS, RB_TRACE, RB_SAMPLES, RB_SCRATCH, RB_TBDATA, RB_COUNT };
    struct RunBuf final {
        void* p = nullptr;
        size_t cap = 0;
    } runBuf[RB_COUNT];
};

namespace {

struct DeviceGuard final {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

cudaError_t runBuffer(vlcu_model* h, int which, size_t bytes, void** out) {
    vlcu_model::RunBuf& b = h->runBuf[which];
    if (bytes > b.cap) {
        // (the stream may still be using the old buffer)
        cudaError_t e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) return e;
        cudaFree(b.p);
        b.p = nullptr;
        b.cap = 0;
        if ((e = cudaMalloc(&b.p, bytes)) != cudaSuccess) return e;
        b.cap = bytes;
    }
    *out = b.p;
    return cudaSuccess;
}

int ensureStage(vlcu_model* h, size_t bytes) {
    if (bytes <= h->stageBytes) return VLCU_OK;
    if (h->dStage) cudaFree(h->dStage);
    h->dStage = nullptr;
    h->stageBytes = 0;
    CUDA_TRY(cudaMalloc(&h->dStage, bytes));
    h->stageBytes = bytes;
    return VLCU_OK;
}

//-------------------------------------------------------------------------
// Port / member transposes between the host's per-instance layout and SoA

// Where the runtime finds a member.  SoA designs: base = state + soaOffset * npad, one tile of
// npad instances.  Tiled designs (VlcuDesignInfo::tile): base = state + soaOffset, `tile`
// instances per tile of tileBytes bytes.  Element e of instance i is then, in units of the
// element type, at  tile(i) + (laneMajor ? lane(i) * elems + e : e * tile + lane(i)).
struct MemberLoc final {
    uint8_t* base;
    uint64_t tile;
    uint64_t tileBytes;
    uint32_t elems;
    uint32_t laneMajor;
};
template <typename T>
__device__ __forceinline__ T* memberAt(const MemberLoc& loc, uint32_t e, uint64_t inst) {
    const uint64_t t = inst / loc.tile, lane = inst - t * loc.tile;
    T* const tilep = reinterpret_cast<T*>(loc.base + t * loc.tileBytes);
    return tilep + (loc.laneMajor ? lane * loc.elems + e : static_cast<uint64_t>(e) * loc.tile + lane);
}

// src: packed records [inst][elems] (one record when broadcast); dst: the member's slots
template <typename T>
__global__ void k_scatter(MemberLoc dst, const T* __restrict__ src, uint64_t n, uint32_t broadcast,
                          T topMask, uint32_t maskEvery) {
    const uint64_t idx = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    const uint32_t elems = dst.elems;
    if (idx >= n * elems) return;
    // instance fastest for [element][instance] members, element fastest for [instance][element]
    const uint64_t inst = dst.laneMajor ? idx / elems : idx % n;
    const uint32_t e = static_cast<uint32_t>(dst.laneMajor ? idx % elems : idx / n);
    T v = src[(broadcast ? 0 : inst * elems) + e];
    // inputs are masked to the declared width on ingest
    if (e % maskEvery == maskEvery - 1) v &= topMask;
    *memberAt<T>(dst, e, inst) = v;
}
template <typename T>
__global__ void k_gather(T* __restrict__ dst, MemberLoc src, uint64_t n) {
    const uint64_t idx = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    const uint32_t elems = src.elems;
    if (idx >= n * elems) return;
    const uint64_t inst = src.laneMajor ? idx / elems : idx % n;
    const uint32_t e = static_cast<uint32_t>(src.laneMajor ? idx % elems : idx / n);
    dst[inst * elems + e] = *memberAt<T>(src, e, inst);
}

// One instance's state as a packed record (what the state block would be with one instance)
struct StateGeom final {
    uint64_t tile, tileBytes;  // as in MemberLoc
    uint64_t offScale;  // member base = state + soaOffset * offScale
    uint32_t recDiv;  // offset of the member in the packed record = soaOffset / recDiv
};
__global__ void k_snapshot(uint8_t* __restrict__ dst, uint8_t* state, StateGeom g, uint64_t inst,
                           const uint32_t* __restrict__ table, uint32_t nMembers) {
    // table: per member {soaOffset, elemBytes, elems, laneMajor}
    for (uint32_t m = blockIdx.x; m < nMembers; m += gridDim.x) {
        const uint32_t off = table[m * 4 + 0], eb = table[m * 4 + 1], elems = table[m * 4 + 2];
        const MemberLoc loc{state + static_cast<uint64_t>(off) * g.offScale, g.tile, g.tileBytes, elems,
                            table[m * 4 + 3] & 1U};
        for (uint64_t b = threadIdx.x; b < static_cast<uint64_t>(eb) * elems; b += blockDim.x) {
            const uint32_t e = static_cast<uint32_t>(b / eb), byte = static_cast<uint32_t>(b % eb);
            const uint8_t* const elemp = eb == 1   ? memberAt<uint8_t>(loc, e, inst)
                                         : eb == 2 ? reinterpret_cast<uint8_t*>(memberAt<uint16_t>(loc, e, inst))
                                         : eb == 4 ? reinterpret_cast<uint8_t*>(memberAt<uint32_t>(loc, e, inst))
                                                   : reinterpret_cast<uint8_t*>(memberAt<uint64_t>(loc, e, inst));
            dst[off / g.recDiv + b] = elemp[byte];
        }
    }
}

// The inverse: one instance's packed record back into the state block
__global__ void k_restore(uint8_t* state, const uint8_t* __restrict__ src, StateGeom g, uint64_t inst,
                          const uint32_t* __restrict__ table, uint32_t nMembers) {
    // table: per member {soaOffset, elemBytes, elems, laneMajor | skip << 1}
    for (uint32_t m = blockIdx.x; m < nMembers; m += gridDim.x) {
        const uint32_t off = table[m * 4 + 0], eb = table[m * 4 + 1], elems = table[m * 4 + 2];
        if (table[m * 4 + 3] & 2U) continue;  // member not part of the image
        const MemberLoc loc{state + static_cast<uint64_t>(off) * g.offScale, g.tile, g.tileBytes, elems,
                            table[m * 4 + 3] & 1U};
        for (uint64_t b = threadIdx.x; b < static_cast<uint64_t>(eb) * elems; b += blockDim.x) {
            const uint32_t e = static_cast<uint32_t>(b / eb), byte = static_cast<uint32_t>(b % eb);
            uint8_t* const elemp = eb == 1   ? memberAt<uint8_t>(loc, e, inst)
                                   : eb == 2 ? reinterpret_cast<uint8_t*>(memberAt<uint16_t>(loc, e, inst))
                                   : eb == 4 ? reinterpret_cast<uint8_t*>(memberAt<uint32_t>(loc, e, inst))
                                             : reinterpret_cast<uint8_t*>(memberAt<uint64_t>(loc, e, inst));
            elemp[byte] = src[off / g.recDiv + b];
        }
    }
}

__global__ void k_time_inc(VlcuCtxRec* ctx, uint64_t n, uint64_t add) {
    const uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i < n) ctx[i].time += add;
}

// Coverage: counters[bin][instance] (uint32, each wrapped like the host's ++) -> uint64 per
// bin.  One CTA per bin streams the bin's row with coalesced 128-bit loads.
__global__ void __launch_bounds__(256) k_cov_reduce(const uint32_t* __restrict__ cov, uint64_t n,
                                                    uint64_t npad, uint32_t bins,
                                                    uint64_t* __restrict__ out) {
    __shared__ uint64_t s_part[8];
    for (uint32_t bin = blockIdx.x; bin < bins; bin += gridDim.x) {
        const uint32_t* const rowp = cov + static_cast<uint64_t>(bin) * npad;
        uint64_t acc = 0;
        const uint64_t nvec = n / 4;
        const uint4* const row4p = reinterpret_cast<const uint4*>(rowp);  // npad % 128 == 0
        for (uint64_t i = threadIdx.x; i < nvec; i += blockDim.x) {
            const uint4 v = row4p[i];
            acc += static_cast<uint64_t>(v.x) + v.y + v.z + v.w;
        }
        for (uint64_t i = nvec * 4 + threadIdx.x; i < n; i += blockDim.x) acc += rowp[i];
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
        if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint64_t total = 0;
            for (int w = 0; w < 8; ++w) total += s_part[w];
            out[bin] = total;
        }
        __syncthreads();
    }
}

// The same for a tiled design: every tile keeps one uint64 counter per bin for its instances
// together (VlcuTileCov), at state + tile * tileBytes + binOffset[bin]: sum them over the tiles.
__global__ void __launch_bounds__(256) k_cov_reduce_tiled(const uint8_t* __restrict__ state, uint64_t tiles,
                                                          uint64_t tileBytes,
                                                          const uint32_t* __restrict__ binOffset, uint32_t bins,
                                                          uint64_t* __restrict__ out) {
    // 32 adjacent bins per CTA (one 256-byte run of a scope's counters per tile), the tiles dealt
    // to the CTA's 8 warps
    __shared__ uint64_t s_part[8][32];
    const uint32_t lane = threadIdx.x & 31U, slice = threadIdx.x >> 5;
    const uint64_t bin = blockIdx.x * 32ULL + lane;
    uint64_t acc = 0;
    if (bin < bins) {
        const uint8_t* const basep = state + binOffset[bin];
#pragma unroll 4
        for (uint64_t t = slice; t < tiles; t += 8) {
            acc += *reinterpret_cast<const unsigned long long*>(basep + t * tileBytes);
        }
    }
    s_part[slice][lane] = acc;
    __syncthreads();
    if (slice == 0 && bin < bins) {
        for (uint32_t k = 1; k < 8; ++k) acc += s_part[k][lane];
        out[bin] = acc;
    }
}

unsigned blocksFor(uint64_t work, unsigned threads) {
    return static_cast<unsigned>((work + threads - 1) / threads);
}

const VlcuMemberDesc* memberAt(const vlcu_model* h, int member) {
    if (!h || member < 0 || static_cast<uint32_t>(member) >= h->infop->nMembers) return nullptr;
    return &h->infop->members[member];
}

bool memberLaneMajor(const VlcuMemberDesc& md) {
    return md.kind == VLCU_MEMBER_MEMORY || (md.flags & VLCU_MEMBERF_LANE_MAJOR);
}
MemberLoc locOf(const vlcu_model* h, const VlcuMemberDesc& md) {
    const uint32_t tile = h->infop->tile;
    MemberLoc loc;
    loc.base = h->dState + (tile ? static_cast<uint64_t>(md.soaOffset) : static_cast<uint64_t>(md.soaOffset) * h->npad);
    loc.tile = tile ? tile : h->npad;
    loc.tileBytes = tile ? static_cast<uint64_t>(h->infop->stateBytes) * tile : 0;
    loc.elems = md.elems;
    loc.laneMajor = memberLaneMajor(md);
    return loc;
}
StateGeom geomOf(const vlcu_model* h) {
    const uint32_t tile = h->infop->tile;
    StateGeom g;
    g.tile = tile ? tile : h->npad;
    g.tileBytes = tile ? static_cast<uint64_t>(h->infop->stateBytes) * tile : 0;
    g.offScale = tile ? 1 : h->npad;
    g.recDiv = tile ? tile : 1;
    return g;
}
// Offset of a member in the packed one-instance record (vlcu_snapshot, save images)
uint32_t recordOffset(const vlcu_model* h, const VlcuMemberDesc& md) {
    return h->infop->tile ? md.soaOffset / h->infop->tile : md.soaOffset;
}

// Width mask of a member's most significant element (inputs are masked on ingest)
uint64_t memberTopMask(const VlcuMemberDesc& md, bool maskWidth) {
    const int topBits = md.width - static_cast<int>(md.elems - 1) * md.elemBytes * 8;
    const bool isWide = md.elems > 1 && md.kind != VLCU_MEMBER_COVERAGE;
    return !maskWidth ? ~0ULL
                      : (isWide || md.elems == 1)
                            ? ((topBits >= md.elemBytes * 8) ? ~0ULL : ((1ULL << topBits) - 1ULL))
                            : ~0ULL;
}

// Device records [inst][elems] (or one record when broadcast) -> the member's slots
int launchScatter(vlcu_model* h, const VlcuMemberDesc& md, const uint8_t* dsrc, bool broadcast,
                  uint64_t topMask) {
    const uint64_t n = h->n;
    const uint64_t work = n * md.elems;
    const unsigned threads = 256, blocks = blocksFor(work, threads);
    const MemberLoc loc = locOf(h, md);
    // the mask applies to the most significant limb of a (wide) value; array elements of a
    // memory are stored as given
    const uint32_t maskEvery = md.elems;
    switch (md.elemBytes) {
    case 1:
        k_scatter<uint8_t><<<blocks, threads, 0, h->stream>>>(loc, dsrc, n, broadcast,
                                                              static_cast<uint8_t>(topMask), maskEvery);
        break;
    case 2:
        k_scatter<uint16_t><<<blocks, threads, 0, h->stream>>>(
            loc, reinterpret_cast<const uint16_t*>(dsrc), n, broadcast, static_cast<uint16_t>(topMask), maskEvery);
        break;
    case 4:
        k_scatter<uint32_t><<<blocks, threads, 0, h->stream>>>(
            loc, reinterpret_cast<const uint32_t*>(dsrc), n, broadcast, static_cast<uint32_t>(topMask), maskEvery);
        break;
    default:
        k_scatter<uint64_t><<<blocks, threads, 0, h->stream>>>(
            loc, reinterpret_cast<const uint64_t*>(dsrc), n, broadcast, topMask, maskEvery);
        break;
    }
    ++h->launches;
    CUDA_TRY(cudaGetLastError());
    return VLCU_OK;
}
// The member's slots -> device records [inst][elems]
int launchGather(vlcu_model* h, const VlcuMemberDesc& md, uint8_t* ddst) {
    const uint64_t n = h->n;
    const unsigned threads = 256, blocks = blocksFor(n * md.elems, threads);
    const MemberLoc loc = locOf(h, md);
    switch (md.elemBytes) {
    case 1: k_gather<uint8_t><<<blocks, threads, 0, h->stream>>>(ddst, loc, n); break;
    case 2: k_gather<uint16_t><<<blocks, threads, 0, h->stream>>>(reinterpret_cast<uint16_t*>(ddst), loc, n); break;
    case 4: k_gather<uint32_t><<<blocks, threads, 0, h->stream>>>(reinterpret_cast<uint32_t*>(ddst), loc, n); break;
    default: k_gather<uint64_t><<<blocks, threads, 0, h->stream>>>(reinterpret_cast<uint64_t*>(ddst), loc, n); break;
    }
    ++h->launches;
    CUDA_TRY(cudaGetLastError());
    return VLCU_OK;
}
// A memory of an SoA design is [instance][element] over the whole shard: plain copies reach it
bool memberIsFlat(const vlcu_model* h, const VlcuMemberDesc& md) {
    return md.kind == VLCU_MEMBER_MEMORY && !h->infop->tile;
}

int transferMember(vlcu_model* h, const VlcuMemberDesc& md, const void* src, void* dst,
                   size_t stride, bool toDevice, bool maskWidth) {
    const size_t recBytes = static_cast<size_t>(md.elemBytes) * md.elems;
    const bool broadcast = toDevice && stride == 0;
    if (!broadcast && stride < recBytes) return fail(VLCU_E_ARG, "stride smaller than one element");
    const uint64_t n = h->n;
    DeviceGuard guard{h->device};
    if (memberIsFlat(h, md) && !broadcast) {  // a strided copy is enough
        uint8_t* const devp = h->dState + static_cast<uint64_t>(md.soaOffset) * h->npad;
        if (toDevice) {
            CUDA_TRY(cudaMemcpy2DAsync(devp, recBytes, src, stride, recBytes, n, cudaMemcpyHostToDevice,
                                       h->stream));
        } else {
            CUDA_TRY(cudaMemcpy2DAsync(dst, stride, devp, recBytes, recBytes, n,
                                       cudaMemcpyDeviceToHost, h->stream));
        }
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        return VLCU_OK;
    }
    const size_t stageBytes = recBytes * (broadcast ? 1 : n);
    if (const int rc = ensureStage(h, stageBytes)) return rc;
    // (the elements of a memory are stored as given: only a value's top limb is masked)
    const uint64_t topMask = memberTopMask(md, maskWidth && md.kind != VLCU_MEMBER_MEMORY);
    if (toDevice) {
        if (broadcast) {
            // one record up, replicated into every instance on the device
            CUDA_TRY(cudaMemcpyAsync(h->dStage, src, recBytes, cudaMemcpyHostToDevice, h->stream));
        } else {
            CUDA_TRY(cudaMemcpy2DAsync(h->dStage, recBytes, src, stride, recBytes, n,
                                       cudaMemcpyHostToDevice, h->stream));
        }
        if (const int rc = launchScatter(h, md, h->dStage, broadcast, topMask)) return rc;
        CUDA_TRY(cudaStreamSynchronize(h->stream));
    } else {
        if (const int rc = launchGather(h, md, h->dStage)) return rc;
        CUDA_TRY(cudaMemcpy2DAsync(dst, stride, h->dStage, recBytes, recBytes, n,
                                   cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
    }
    return VLCU_OK;
}

void fillPortInfo(const VlcuMemberDesc& md, vlcu_port_info_t* out) {
    out->name = md.name;
    out->dir = md.kind == VLCU_MEMBER_IN    ? VLCU_DIR_IN
               : md.kind == VLCU_MEMBER_OUT ? VLCU_DIR_OUT
                                            : VLCU_DIR_INOUT;
    out->width = static_cast<int>(md.width);
    out->elem_bytes = md.elemBytes;
    out->elems = static_cast<int>(md.elems);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (no libcuda link)
typedef CUresult (*VlcuTensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                          const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                          const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
VlcuTensorMapEncodeFn tensorMapEncodeFn() {
    static const VlcuTensorMapEncodeFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess
            || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            p = nullptr;
        }
        return reinterpret_cast<VlcuTensorMapEncodeFn>(p);
    }();
    return fn;
}
// rows [evals * words][npad] of uint32 as a 2-D tensor, box = {VLCU_BLOCK, words}
bool encodeRowMap(uint64_t out[16], const void* basep, uint64_t npad, uint32_t words, uint64_t evals) {
    const VlcuTensorMapEncodeFn fn = tensorMapEncodeFn();
    if (!fn || !basep || !words || words > 256 || !evals || words * evals > 0xffffffffULL) return false;
    if (reinterpret_cast<uintptr_t>(basep) & 15) return false;
    static_assert(sizeof(CUtensorMap) == 16 * sizeof(uint64_t), "CUtensorMap size");
    CUtensorMap map;
    const cuuint64_t dims[2] = {npad, static_cast<cuuint64_t>(words) * evals};
    const cuuint64_t strides[1] = {npad * sizeof(uint32_t)};
    const cuuint32_t box[2] = {VLCU_BLOCK, words};
    const cuuint32_t elemStrides[2] = {1, 1};
    const CUresult rc = fn(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void*>(basep), dims,
                           strides, box, elemStrides, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return false;
    memcpy(out, &map, sizeof(map));
    return true;
}

int launchRun(vlcu_model* h, const char* tb, VlcuRunDev& r, bool markStart = true,
              bool markStop = true) {
    r.npad = h->npad;
    VlcuRowMaps maps{};
    r.rowMaps = nullptr;
    if (!strcmp(tb, "trace") && r.stim && !r.stimBroadcast && !(r.flags & VLCU_RUN_NO_STREAM)
        && r.nEvals) {
        bool ok = encodeRowMap(maps.in, r.stim, h->npad, h->infop->inWords, r.nEvals);
        if (ok && r.obs && r.obsEvals == r.nEvals) {
            ok = encodeRowMap(maps.out, r.obs, h->npad, h->infop->outWords, r.nEvals);
        }
        if (ok) r.rowMaps = &maps;
    }
    r.sigs = h->dSigs;
    r.evalsDone = h->dEvalsDone;
    r.evalBase = h->evalIndex;
    if (markStart) CUDA_TRY(cudaEventRecord(h->evStart, h->stream));
    const cudaError_t e = h->opsp->launchRun(tb, h->dev, r, h->stream);
    if (e == cudaErrorInvalidValue) {
        return fail(VLCU_E_ARG, std::string{"unknown testbench '"} + tb + "' for design "
                                    + h->infop->name);
    }
    if (e != cudaSuccess) return failCuda(e, "kernel launch");
    if (markStop) CUDA_TRY(cudaEventRecord(h->evStop, h->stream));
    ++h->launches;
    h->haveTiming = true;
    h->evalIndex += r.nEvals;
    return VLCU_OK;
}

// "trace" runs whose stimulus (and captured outputs) live in HOST memory and are too large to
// be worth holding on the device: eval-chunks flow host -> device -> host through two slots on
// three streams, so the copy-in of chunk c+1, the kernel of chunk c and the copy-out of chunk
// c-1 overlap (PCIe is full duplex); device memory use is bounded by four chunk buffers.
int runTracePipelined(vlcu_model* h, const vlcu_run_desc_t* d, VlcuRunDev r, size_t chunkEvals) {
    const VlcuDesignInfo& info = *h->infop;
    const size_t inRow = static_cast<size_t>(info.inWords) * h->npad;  // words per eval
    const size_t outRow = static_cast<size_t>(info.outWords) * h->npad;
    const bool wantObs = d->obs && d->obs_evals >= d->n_evals && info.outWords;
    struct Res final {
        uint32_t* dStim[2] = {nullptr, nullptr};
        uint32_t* dObs[2] = {nullptr, nullptr};
        cudaEvent_t evIn[2] = {nullptr, nullptr}, evRun[2] = {nullptr, nullptr},
                    evOut[2] = {nullptr, nullptr};
        cudaStream_t outStream = nullptr;
        ~Res() {
            for (int b = 0; b < 2; ++b) {
                cudaFree(dStim[b]);
                cudaFree(dObs[b]);
                if (evIn[b]) cudaEventDestroy(evIn[b]);
                if (evRun[b]) cudaEventDestroy(evRun[b]);
                if (evOut[b]) cudaEventDestroy(evOut[b]);
            }
            if (outStream) cudaStreamDestroy(outStream);
        }
    } res;
    if (!h->copyStream) {
        CUDA_TRY(cudaStreamCreateWithFlags(&h->copyStream, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&h->evCopied, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&h->evConsumed, cudaEventDisableTiming));
    }
    CUDA_TRY(cudaStreamCreateWithFlags(&res.outStream, cudaStreamNonBlocking));
    for (int b = 0; b < 2; ++b) {
        CUDA_TRY(cudaMalloc(&res.dStim[b], inRow * chunkEvals * sizeof(uint32_t)));
        if (wantObs) CUDA_TRY(cudaMalloc(&res.dObs[b], outRow * chunkEvals * sizeof(uint32_t)));
        CUDA_TRY(cudaEventCreateWithFlags(&res.evIn[b], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&res.evRun[b], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&res.evOut[b], cudaEventDisableTiming));
    }
    const uint64_t total = d->n_evals;
    uint64_t chunk = 0;
    for (uint64_t e0 = 0; e0 < total; e0 += chunkEvals, ++chunk) {
        const int b = static_cast<int>(chunk & 1);
        const uint64_t ne = total - e0 < chunkEvals ? total - e0 : chunkEvals;
        // copy-in: the slot is free once the kernel of chunk c-2 has finished
        if (chunk >= 2) CUDA_TRY(cudaStreamWaitEvent(h->copyStream, res.evRun[b], 0));
        CUDA_TRY(cudaMemcpyAsync(res.dStim[b], d->stim + e0 * inRow, inRow * ne * sizeof(uint32_t),
                                 cudaMemcpyHostToDevice, h->copyStream));
        CUDA_TRY(cudaEventRecord(res.evIn[b], h->copyStream));
        // kernel: needs its stimulus, and the output slot drained by the copy-out of chunk c-2
        CUDA_TRY(cudaStreamWaitEvent(h->stream, res.evIn[b], 0));
        if (wantObs && chunk >= 2) CUDA_TRY(cudaStreamWaitEvent(h->stream, res.evOut[b], 0));
        VlcuRunDev rc = r;
        rc.nEvals = ne;
        rc.stim = res.dStim[b];
        rc.obs = wantObs ? res.dObs[b] : nullptr;
        rc.obsEvals = wantObs ? ne : 0;
        if (chunk) rc.flags &= ~static_cast<uint32_t>(VLCU_RUN_TB_BEGIN);
        if (const int rcode = launchRun(h, "trace", rc, chunk == 0, e0 + ne >= total)) return rcode;
        CUDA_TRY(cudaEventRecord(res.evRun[b], h->stream));
        if (wantObs) {
            CUDA_TRY(cudaStreamWaitEvent(res.outStream, res.evRun[b], 0));
            CUDA_TRY(cudaMemcpyAsync(d->obs + e0 * outRow, res.dObs[b], outRow * ne * sizeof(uint32_t),
                                     cudaMemcpyDeviceToHost, res.outStream));
            CUDA_TRY(cudaEventRecord(res.evOut[b], res.outStream));
        }
    }
    CUDA_TRY(cudaStreamSynchronize(h->copyStream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaStreamSynchronize(res.outStream));
    return VLCU_OK;
}

}  // namespace

//=========================================================================
// C ABI

extern "C" {

int vlcu_abi_version(void) { return VLCU_ABI_VERSION; }
const char* vlcu_last_error(void) { return t_lastError.c_str(); }
int vlcu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}
int vlcu_design_count(void) { return static_cast<int>(registry().size()); }
const char* vlcu_design_name(int index) {
    if (index < 0 || index >= vlcu_design_count()) return nullptr;
    return registry()[index]->info().name;
}

void* vlcu_host_alloc(size_t bytes) {
    void* p = nullptr;
    const cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        failCuda(e, "cudaHostAlloc");
        return nullptr;
    }
    return p;
}
void vlcu_host_free(void* ptr) {
    if (ptr) cudaFreeHost(ptr);
}
void* vlcu_device_alloc(vlcu_model_t* h, size_t bytes) {
    if (!h) return nullptr;
    DeviceGuard guard{h->device};
    void* p = nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) {
        fail(VLCU_E_CUDA, "vlcu_device_alloc: cudaMalloc failed");
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void vlcu_device_free(vlcu_model_t* h, void* ptr) {
    if (!h || !ptr) return;
    DeviceGuard guard{h->device};
    cudaFree(ptr);
}
int vlcu_device_copy(vlcu_model_t* h, void* dst, const void* src, size_t bytes, int to_device) {
    if (!h || !dst || !src) return fail(VLCU_E_ARG, "vlcu_device_copy: null argument");
    DeviceGuard guard{h->device};
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost,
                             h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return VLCU_OK;
}
int vlcu_set_pipeline_chunk(vlcu_model_t* h, size_t bytes) {
    if (!h) return fail(VLCU_E_ARG, "null model");
    h->pipelineChunkBytes = bytes;
    return VLCU_OK;
}

vlcu_model_t* vlcu_create(const char* design, uint64_t n_instances, int device, uint64_t inst_base,
                          uint64_t seed_base, int rand_reset, uint32_t msg_cap,
                          const char* plusargs) {
    if (!design || !n_instances) {
        fail(VLCU_E_ARG, "vlcu_create: design name and a non-zero instance count are required");
        return nullptr;
    }
    const VlcuDesignOps* opsp = nullptr;
    for (VlcuDesignOps* p : registry()) {
        if (!strcmp(p->info().name, design)) opsp = p;
    }
    if (!opsp) {
        fail(VLCU_E_ARG, std::string{"vlcu_create: unknown design '"} + design + "'");
        return nullptr;
    }
    int ndev = 0;
    const cudaError_t de = cudaGetDeviceCount(&ndev);
    if (de != cudaSuccess || ndev == 0) {
        fail(VLCU_E_NODEVICE,
             std::string{"vlcu_create: no CUDA device ("}
                 + (de != cudaSuccess ? cudaGetErrorString(de) : "device count is 0")
                 + "); the batched evaluator has no CPU path");
        return nullptr;
    }
    if (device < 0 || device >= ndev) {
        fail(VLCU_E_ARG, "vlcu_create: device index out of range");
        return nullptr;
    }
    vlcu_model* const h = new vlcu_model;
    h->opsp = opsp;
    h->infop = &opsp->info();
    h->device = device;
    h->n = n_instances;
    h->npad = (n_instances + 127) / 128 * 128;
    for (uint32_t m = 0; m < h->infop->nMembers; ++m) {
        if (h->infop->members[m].kind <= VLCU_MEMBER_INOUT) h->ports.push_back(static_cast<int>(m));
    }
    DeviceGuard guard{device};
    const auto bail = [&](cudaError_t e, const char* what) -> vlcu_model_t* {
        failCuda(e, what);
        vlcu_destroy(h);
        return nullptr;
    };
    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess)
        return bail(e, "cudaStreamCreate");
    if ((e = cudaEventCreate(&h->evStart)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaEventCreate(&h->evStop)) != cudaSuccess) return bail(e, "cudaEventCreate");
    const size_t stateBytes = static_cast<size_t>(h->infop->stateBytes) * h->npad;
    if ((e = cudaMalloc(&h->dState, stateBytes)) != cudaSuccess) return bail(e, "cudaMalloc(state)");
    if ((e = cudaMemsetAsync(h->dState, 0, stateBytes, h->stream)) != cudaSuccess)
        return bail(e, "cudaMemset(state)");
    if ((e = cudaMalloc(&h->dCtx, sizeof(VlcuCtxRec) * h->npad)) != cudaSuccess)
        return bail(e, "cudaMalloc(ctx)");
    if ((e = cudaMalloc(&h->dRng, 2 * sizeof(uint64_t) * h->npad)) != cudaSuccess)
        return bail(e, "cudaMalloc(rng)");
    if (msg_cap) {
        if ((e = cudaMalloc(&h->dMsgs, sizeof(VlcuMsg) * msg_cap * h->n)) != cudaSuccess)
            return bail(e, "cudaMalloc(msgs)");
    }
    if ((e = cudaMalloc(&h->dSigs, sizeof(uint64_t) * h->npad)) != cudaSuccess)
        return bail(e, "cudaMalloc(sigs)");
    if ((e = cudaMalloc(&h->dEvalsDone, sizeof(uint64_t) * h->npad)) != cudaSuccess)
        return bail(e, "cudaMalloc(evalsDone)");
    if (h->infop->coverBins) {
        if ((e = cudaMalloc(&h->dCovSums, sizeof(uint64_t) * h->infop->coverBins))
            != cudaSuccess)
            return bail(e, "cudaMalloc(covSums)");
    }
    // $test$plusargs are resolved here, once, into a bit per string the design asks about
    if (h->infop->nPlusargs > 32) {
        fail(VLCU_E_ARG, "vlcu_create: the design tests more than 32 $test$plusargs strings");
        vlcu_destroy(h);
        return nullptr;
    }
    uint32_t plusBits = 0;
    if (plusargs) {
        std::istringstream is{plusargs};
        std::string tok;
        while (is >> tok) {
            if (!tok.empty() && tok[0] == '+') tok = tok.substr(1);
            for (uint32_t k = 0; k < h->infop->nPlusargs; ++k) {
                // $test$plusargs matches a prefix of the provided argument
                if (tok.compare(0, strlen(h->infop->plusargs[k]), h->infop->plusargs[k]) == 0) {
                    plusBits |= 1u << k;
                }
            }
        }
    }
    h->dev.state = VlcuState{h->dState, h->npad};
    h->dev.ctxRecs = h->dCtx;
    h->dev.rng = h->dRng;
    h->dev.msgs = h->dMsgs;
    h->dev.n = h->n;
    h->dev.instBase = inst_base;
    h->dev.seedBase = seed_base;
    h->dev.randReset = static_cast<uint32_t>(rand_reset);
    h->dev.msgCap = msg_cap;
    h->dev.plusargs = plusBits;
    h->dev.errorLimit = 1;  // VerilatedContext default (include/verilated.h:431)
    if (vlcu_reset(h) != VLCU_OK) {
        vlcu_destroy(h);
        return nullptr;
    }
    return h;
}

void vlcu_destroy(vlcu_model_t* h) {
    if (!h) return;
    DeviceGuard guard{h->device};
    if (h->comm && nccl().CommDestroy) nccl().CommDestroy(h->comm);
    if (h->copyStream) cudaStreamSynchronize(h->copyStream);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (auto& it : h->staged) cudaFree(it.second.dBuf);
    if (h->resStream) cudaStreamSynchronize(h->resStream);
    cudaFree(h->dRes);
    if (h->evCaptured) cudaEventDestroy(h->evCaptured);
    if (h->evFetched) cudaEventDestroy(h->evFetched);
    if (h->resStream) cudaStreamDestroy(h->resStream);
    if (h->evCopied) cudaEventDestroy(h->evCopied);
    if (h->evConsumed) cudaEventDestroy(h->evConsumed);
    if (h->copyStream) cudaStreamDestroy(h->copyStream);
    cudaFree(h->dState);
    cudaFree(h->dCtx);
    cudaFree(h->dRng);
    cudaFree(h->dMsgs);
    cudaFree(h->dSigs);
    cudaFree(h->dEvalsDone);
    cudaFree(h->dCovSums);
    cudaFree(h->dCovBinOffset);
    cudaFree(h->dStage);
    cudaFree(h->dFlush);
    for (vlcu_model::RunBuf& b : h->runBuf) cudaFree(b.p);
    if (h->evStart) cudaEventDestroy(h->evStart);
    if (h->evStop) cudaEventDestroy(h->evStop);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int vlcu_reset(vlcu_model_t* h) {
    if (!h) return fail(VLCU_E_ARG, "null model");
    DeviceGuard guard{h->device};
    CUDA_TRY(cudaMemsetAsync(h->dSigs, 0, sizeof(uint64_t) * h->npad, h->stream));
    CUDA_TRY(cudaMemsetAsync(h->dEvalsDone, 0, sizeof(uint64_t) * h->npad, h->stream));
    CUDA_TRY(cudaMemsetAsync(h->dCtx, 0, sizeof(VlcuCtxRec) * h->npad, h->stream));
    const cudaError_t e = h->opsp->launchReset(h->dev, h->stream);
    if (e != cudaSuccess) return failCuda(e, "reset kernel");
    ++h->launches;
    h->evalIndex = 0;
    if (const int rc = vlcu_reset_signatures(h)) return rc;
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return VLCU_OK;
}

uint64_t vlcu_instances(const vlcu_model_t* h) { return h ? h->n : 0; }
uint64_t vlcu_padded_instances(const vlcu_model_t* h) { return h ? h->npad : 0; }
uint32_t vlcu_state_bytes(const vlcu_model_t* h) { return h ? h->infop->stateBytes : 0; }
uint32_t vlcu_in_words(const vlcu_model_t* h) { return h ? h->infop->inWords : 0; }
uint32_t vlcu_out_words(const vlcu_model_t* h) { return h ? h->infop->outWords : 0; }
uint32_t vlcu_in_bytes(const vlcu_model_t* h) { return h ? h->infop->inBytes : 0; }
uint32_t vlcu_out_bytes(const vlcu_model_t* h) { return h ? h->infop->outBytes : 0; }
uint32_t vlcu_touched_read_bytes(const vlcu_model_t* h) {
    if (!h) return 0;
    return h->infop->touchedReadBytes ? h->infop->touchedReadBytes : h->infop->stateBytes;
}
uint32_t vlcu_touched_write_bytes(const vlcu_model_t* h) {
    if (!h) return 0;
    return h->infop->touchedWriteBytes ? h->infop->touchedWriteBytes : h->infop->stateBytes;
}
uint32_t vlcu_trace_record_bytes(const vlcu_model_t* h) {
    return h ? VLCU_TRACE_HDR + h->infop->stateBytes + 4U * VLCU_TRACE_WORDS_PADDED(h->infop->traceWords) : 0;
}

int vlcu_port_count(const vlcu_model_t* h) { return h ? static_cast<int>(h->ports.size()) : 0; }
int vlcu_port_info(const vlcu_model_t* h, int port, vlcu_port_info_t* out) {
    if (!h || !out || port < 0 || port >= vlcu_port_count(h)) return fail(VLCU_E_ARG, "bad port");
    fillPortInfo(h->infop->members[h->ports[port]], out);
    return VLCU_OK;
}
int vlcu_port_index(const vlcu_model_t* h, const char* name) {
    if (!h || !name) return fail(VLCU_E_ARG, "bad argument");
    for (int p = 0; p < vlcu_port_count(h); ++p) {
        if (!strcmp(h->infop->members[h->ports[p]].name, name)) return p;
    }
    return fail(VLCU_E_ARG, std::string{"no port named '"} + name + "'");
}
int vlcu_set_input(vlcu_model_t* h, int port, const void* src, size_t stride_bytes) {
    if (!h || !src || port < 0 || port >= vlcu_port_count(h)) return fail(VLCU_E_ARG, "bad port");
    const VlcuMemberDesc& md = h->infop->members[h->ports[port]];
    if (md.kind == VLCU_MEMBER_OUT) return fail(VLCU_E_ARG, "port is an output");
    return transferMember(h, md, src, nullptr, stride_bytes, true, true);
}
int vlcu_get_output(vlcu_model_t* h, int port, void* dst, size_t stride_bytes) {
    if (!h || !dst || port < 0 || port >= vlcu_port_count(h)) return fail(VLCU_E_ARG, "bad port");
    return transferMember(h, h->infop->members[h->ports[port]], nullptr, dst, stride_bytes, false,
                          false);
}
int vlcu_member_count(const vlcu_model_t* h) { return h ? static_cast<int>(h->infop->nMembers) : 0; }
const char* vlcu_member_name(const vlcu_model_t* h, int member) {
    const VlcuMemberDesc* const mdp = memberAt(h, member);
    return mdp ? mdp->name : nullptr;
}
int vlcu_member_index(const vlcu_model_t* h, const char* name) {
    if (!h || !name) return fail(VLCU_E_ARG, "bad argument");
    for (uint32_t m = 0; m < h->infop->nMembers; ++m) {
        if (!strcmp(h->infop->members[m].name, name)) return static_cast<int>(m);
    }
    return fail(VLCU_E_ARG, std::string{"no member named '"} + name + "'");
}
int vlcu_member_info(const vlcu_model_t* h, int member, vlcu_port_info_t* out) {
    const VlcuMemberDesc* const mdp = memberAt(h, member);
    if (!mdp || !out) return fail(VLCU_E_ARG, "bad member");
    fillPortInfo(*mdp, out);
    return VLCU_OK;
}
int vlcu_set_member(vlcu_model_t* h, int member, const void* src, size_t stride_bytes) {
    const VlcuMemberDesc* const mdp = memberAt(h, member);
    if (!mdp || !src) return fail(VLCU_E_ARG, "bad member");
    return transferMember(h, *mdp, src, nullptr, stride_bytes, true, false);
}
int vlcu_get_member(vlcu_model_t* h, int member, void* dst, size_t stride_bytes) {
    const VlcuMemberDesc* const mdp = memberAt(h, member);
    if (!mdp || !dst) return fail(VLCU_E_ARG, "bad member");
    return transferMember(h, *mdp, nullptr, dst, stride_bytes, false, false);
}
// Asynchronous double-buffered ingest.  vlcu_stage_member() starts the host->device copy of a
// member's NEXT value on a separate copy stream into a staging buffer (the state is untouched,
// so a vlcu_run in flight keeps computing); vlcu_commit_staged() moves everything staged into
// the state on the model's stream, ordered on the device behind the copies (the host never
// waits).  src must stay valid and unchanged until vlcu_commit_staged + the next sync point.
int vlcu_stage_member(vlcu_model_t* h, int member, const void* src, size_t stride_bytes,
                      int mask_width) {
    const VlcuMemberDesc* const mdp = memberAt(h, member);
    if (!mdp || !src) return fail(VLCU_E_ARG, "bad member");
    const size_t recBytes = static_cast<size_t>(mdp->elemBytes) * mdp->elems;
    if (stride_bytes < recBytes) return fail(VLCU_E_ARG, "stride smaller than one element");
    DeviceGuard guard{h->device};
    if (!h->copyStream) {
        CUDA_TRY(cudaStreamCreateWithFlags(&h->copyStream, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&h->evCopied, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&h->evConsumed, cudaEventDisableTiming));
    }
    vlcu_model::Staged& st = h->staged[member];
    const size_t bytes = recBytes * h->n;
    if (st.bytes < bytes) {
        // a previous commit may still be reading the old buffer
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        if (st.dBuf) cudaFree(st.dBuf);
        st.dBuf = nullptr;
        st.bytes = 0;
        CUDA_TRY(cudaMalloc(&st.dBuf, bytes));
        st.bytes = bytes;
    }
    // the staging buffer is free once the last commit has consumed it
    if (h->haveConsumed) CUDA_TRY(cudaStreamWaitEvent(h->copyStream, h->evConsumed, 0));
    if (stride_bytes == recBytes) {
        CUDA_TRY(cudaMemcpyAsync(st.dBuf, src, bytes, cudaMemcpyHostToDevice, h->copyStream));
    } else {
        CUDA_TRY(cudaMemcpy2DAsync(st.dBuf, recBytes, src, stride_bytes, recBytes, h->n,
                                   cudaMemcpyHostToDevice, h->copyStream));
    }
    st.pending = true;
    st.mask = mask_width != 0;
    return VLCU_OK;
}
int vlcu_commit_staged(vlcu_model_t* h) {
    if (!h) return fail(VLCU_E_ARG, "null model");
    if (!h->copyStream) return VLCU_OK;
    DeviceGuard guard{h->device};
    bool any = false;
    for (auto& it : h->staged) any |= it.second.pending;
    if (!any) return VLCU_OK;
    CUDA_TRY(cudaEventRecord(h->evCopied, h->copyStream));
    CUDA_TRY(cudaStreamWaitEvent(h->stream, h->evCopied, 0));
    for (auto& it : h->staged) {
        vlcu_model::Staged& st = it.second;
        if (!st.pending) continue;
        const VlcuMemberDesc& md = h->infop->members[it.first];
        const size_t recBytes = static_cast<size_t>(md.elemBytes) * md.elems;
        if (memberIsFlat(h, md)) {  // instance-major in HBM already
            uint8_t* const devp = h->dState + static_cast<uint64_t>(md.soaOffset) * h->npad;
            CUDA_TRY(cudaMemcpyAsync(devp, st.dBuf, recBytes * h->n, cudaMemcpyDeviceToDevice,
                                     h->stream));
        } else {
            if (const int rc = launchScatter(h, md, st.dBuf, false, memberTopMask(md, st.mask)))
                return rc;
        }
        st.pending = false;
    }
    CUDA_TRY(cudaEventRecord(h->evConsumed, h->stream));
    h->haveConsumed = true;
    return VLCU_OK;
}
int vlcu_member_device_ptr(vlcu_model_t* h, int member, void** ptr, uint64_t* npad) {
    const VlcuMemberDesc* const mdp = memberAt(h, member);
    if (!mdp || !ptr) return fail(VLCU_E_ARG, "bad member");
    // (a tiled design: lane 0 of tile 0; consecutive tiles are stateBytes * tile bytes apart and
    // *npad is the tile size)
    const MemberLoc loc = locOf(h, *mdp);
    *ptr = loc.base;
    if (npad) *npad = loc.tile;
    return VLCU_OK;
}
uint32_t vlcu_member_offset(const vlcu_model_t* h, int member) {
    const VlcuMemberDesc* const mdp = memberAt(h, member);
    return mdp ? recordOffset(h, *mdp) : 0;
}
uint32_t vlcu_tile(const vlcu_model_t* h) { return h ? h->infop->tile : 0; }

int vlcu_time_inc(vlcu_model_t* h, uint64_t add) {
    if (!h) return fail(VLCU_E_ARG, "null model");
    DeviceGuard guard{h->device};
    k_time_inc<<<blocksFor(h->n, 256), 256, 0, h->stream>>>(h->dCtx, h->n, add);
    ++h->launches;
    CUDA_TRY(cudaGetLastError());
    return VLCU_OK;
}

int vlcu_eval(vlcu_model_t* h) {
    if (!h) return fail(VLCU_E_ARG, "null model");
    DeviceGuard guard{h->device};
    VlcuRunDev r{};
    r.nEvals = 1;
    return launchRun(h, "none", r);
}

int vlcu_run(vlcu_model_t* h, const vlcu_run_desc_t* d) {
    if (!h || !d || !d->tb) return fail(VLCU_E_ARG, "vlcu_run: null argument");
    DeviceGuard guard{h->device};
    const VlcuDesignInfo& info = *h->infop;
    VlcuRunDev r{};
    r.nEvals = d->n_evals;
    r.flags = d->flags & (VLCU_RUNF_TB_BEGIN | VLCU_RUNF_STOP_ON_FINISH | VLCU_RUNF_NO_STREAM
                          | VLCU_RUNF_BULK_1D | VLCU_RUNF_TRACE_CHANGES);
    r.stimBroadcast = (d->flags & VLCU_RUNF_STIM_BROADCAST) ? 1 : 0;
    r.tbSeed = d->tb_seed;
    r.tbFlags = d->tb_flags;
    for (int k = 0; k < 4; ++k) r.tbArg[k] = d->tb_arg[k];
    // Scratch device buffers for host-side stimulus / observation / trace (pooled in the handle)
    uint32_t* dStim = nullptr;
    uint32_t* dObs = nullptr;
    uint8_t* dTrace = nullptr;
    uint64_t* dSamples = nullptr;
    uint32_t* dScratch = nullptr;
    const auto cleanup = [&]() {};
#define RUN_TRY(expr) \
    do { \
        const cudaError_t e_ = (expr); \
        if (e_ != cudaSuccess) { \
            cleanup(); \
            return failCuda(e_, #expr); \
        } \
    } while (0)
#define RUN_BUF(ptr, which, bytes) RUN_TRY(runBuffer(h, vlcu_model::which, (bytes), reinterpret_cast<void**>(&(ptr))))
    if (!strcmp(d->tb, "trace")) {
        if (!d->stim) return fail(VLCU_E_ARG, "vlcu_run: the trace testbench needs stim");
        const size_t rowWords = static_cast<size_t>(info.inWords) * (r.stimBroadcast ? 1 : h->npad);
        // Large host-resident stimulus: stream it in eval-chunks instead of staging it whole
        const bool hostIo = !(d->flags & VLCU_RUNF_STIM_DEVICE) && !(d->obs && (d->flags & VLCU_RUNF_OBS_DEVICE));
        const bool plainObs = !d->obs || !d->obs_evals || d->obs_evals >= d->n_evals;
        const size_t chunkBytes = h->pipelineChunkBytes;
        const size_t rowBytes = (rowWords + (d->obs ? static_cast<size_t>(info.outWords) * h->npad : 0)) * sizeof(uint32_t);
        if (hostIo && plainObs && !r.stimBroadcast && !d->n_samples && info.inWords && chunkBytes
            && !(d->flags & VLCU_RUNF_STOP_ON_FINISH) && rowBytes * d->n_evals > 2 * chunkBytes) {
            size_t chunkEvals = chunkBytes / rowBytes;
            if (chunkEvals < 8) chunkEvals = 8;
            return runTracePipelined(h, d, r, chunkEvals);
        }
        if (d->flags & VLCU_RUNF_STIM_DEVICE) {
            r.stim = d->stim;
        } else {
            const size_t bytes = rowWords * d->n_evals * sizeof(uint32_t);
            RUN_BUF(dStim, RB_STIM, bytes);
            RUN_TRY(cudaMemcpyAsync(dStim, d->stim, bytes, cudaMemcpyHostToDevice, h->stream));
            r.stim = dStim;
        }
    }
    if (d->tb_data && d->tb_data_bytes) {
        if (d->flags & VLCU_RUNF_TBDATA_DEVICE) {
            r.tbPtr[0] = const_cast<void*>(d->tb_data);
        } else {
            uint8_t* dTbData = nullptr;
            RUN_BUF(dTbData, RB_TBDATA, d->tb_data_bytes);
            RUN_TRY(cudaMemcpyAsync(dTbData, d->tb_data, d->tb_data_bytes, cudaMemcpyHostToDevice, h->stream));
            r.tbPtr[0] = dTbData;
        }
    }
    size_t obsBytes = 0;
    if (d->obs && d->obs_evals) {
        r.obsEvals = d->obs_evals < d->n_evals ? d->obs_evals : d->n_evals;
        obsBytes = static_cast<size_t>(info.outWords) * h->npad * r.obsEvals * sizeof(uint32_t);
        if (d->flags & VLCU_RUNF_OBS_DEVICE) {
            r.obs = d->obs;
        } else {
            RUN_BUF(dObs, RB_OBS, obsBytes);
            RUN_TRY(cudaMemsetAsync(dObs, 0, obsBytes, h->stream));
            r.obs = dObs;
        }
    }
    size_t traceBytes = 0;
    if (d->n_samples && d->trace_out && d->trace_evals && d->sample_instances) {
        for (uint32_t s = 0; s < d->n_samples; ++s) {
            if (d->sample_instances[s] >= h->n) {
                cleanup();
                return fail(VLCU_E_ARG, "vlcu_run: sampled instance out of range");
            }
        }
        r.nSamples = d->n_samples;
        r.traceEvals = d->trace_evals < d->n_evals ? d->trace_evals : d->n_evals;
        if (d->flags & VLCU_RUNF_TRACE_CHANGES) {
            if (!info.traceWords) {
                cleanup();
                return fail(VLCU_E_STATE, "vlcu_run: design was not Verilated with --trace");
            }
            if (d->trace_stream_words < VLCU_TRACE_STREAM_HDR + 3U + 2U * info.traceWords) {
                cleanup();
                return fail(VLCU_E_ARG, "vlcu_run: trace_stream_words cannot hold even the first dump");
            }
            r.traceCapWords = d->trace_stream_words;
            traceBytes = sizeof(uint32_t) * r.traceCapWords * r.nSamples;
            const size_t scratchBytes
                = sizeof(uint32_t) * 2U * VLCU_TRACE_WORDS_PADDED(info.traceWords) * r.nSamples;
            RUN_BUF(dScratch, RB_SCRATCH, scratchBytes);
            RUN_TRY(cudaMemsetAsync(dScratch, 0, scratchBytes, h->stream));
            r.traceScratch = dScratch;
        } else {
            traceBytes = static_cast<size_t>(VLCU_TRACE_HDR + info.stateBytes
                                             + 4U * VLCU_TRACE_WORDS_PADDED(info.traceWords))
                         * r.nSamples * r.traceEvals;
        }
        RUN_BUF(dTrace, RB_TRACE, traceBytes);
        RUN_TRY(cudaMemsetAsync(dTrace, 0, traceBytes, h->stream));
        RUN_BUF(dSamples, RB_SAMPLES, sizeof(uint64_t) * r.nSamples);
        RUN_TRY(cudaMemcpyAsync(dSamples, d->sample_instances, sizeof(uint64_t) * r.nSamples,
                                cudaMemcpyHostToDevice, h->stream));
        r.trace = dTrace;
        r.sampleIds = dSamples;
    }
    const int rc = launchRun(h, d->tb, r);
    if (rc != VLCU_OK) {
        cleanup();
        return rc;
    }
    if (dObs) {
        RUN_TRY(cudaMemcpyAsync(d->obs, dObs, obsBytes, cudaMemcpyDeviceToHost, h->stream));
    }
    if (dTrace) {
        RUN_TRY(cudaMemcpyAsync(d->trace_out, dTrace, traceBytes, cudaMemcpyDeviceToHost,
                                h->stream));
    }
    if (dStim || dObs || dTrace) RUN_TRY(cudaStreamSynchronize(h->stream));
    cleanup();
#undef RUN_BUF
#undef RUN_TRY
    return VLCU_OK;
}

int vlcu_final(vlcu_model_t* h) {
    if (!h) return fail(VLCU_E_ARG, "null model");
    DeviceGuard guard{h->device};
    const cudaError_t e = h->opsp->launchFinal(h->dev, h->stream);
    if (e != cudaSuccess) return failCuda(e, "final kernel");
    ++h->launches;
    return VLCU_OK;
}

int vlcu_sync(vlcu_model_t* h) {
    if (!h) return fail(VLCU_E_ARG, "null model");
    DeviceGuard guard{h->device};
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return VLCU_OK;
}

int vlcu_flush_l2(vlcu_model_t* h) {
    if (!h) return fail(VLCU_E_ARG, "null model");
    DeviceGuard guard{h->device};
    const size_t bytes = size_t{256} << 20;  // > 126 MB L2
    if (!h->dFlush) CUDA_TRY(cudaMalloc(&h->dFlush, bytes));
    CUDA_TRY(cudaMemsetAsync(h->dFlush, static_cast<int>(h->launches & 0xff), bytes, h->stream));
    return VLCU_OK;
}

int vlcu_last_kernel_ms(vlcu_model_t* h, float* ms) {
    if (!h || !ms) return fail(VLCU_E_ARG, "null argument");
    if (!h->haveTiming) return fail(VLCU_E_STATE, "no eval/run kernel has been launched yet");
    DeviceGuard guard{h->device};
    CUDA_TRY(cudaEventSynchronize(h->evStop));
    CUDA_TRY(cudaEventElapsedTime(ms, h->evStart, h->evStop));
    return VLCU_OK;
}
uint64_t vlcu_launch_count(const vlcu_model_t* h) { return h ? h->launches : 0; }

int vlcu_status(vlcu_model_t* h, vlcu_status_t* out) {
    if (!h || !out) return fail(VLCU_E_ARG, "null argument");
    DeviceGuard guard{h->device};
    CUDA_TRY(cudaMemcpyAsync(out, h->dCtx, sizeof(VlcuCtxRec) * h->n, cudaMemcpyDeviceToHost,
                             h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return VLCU_OK;
}
int vlcu_signatures(vlcu_model_t* h, uint64_t* out) {
    if (!h || !out) return fail(VLCU_E_ARG, "null argument");
    DeviceGuard guard{h->device};
    CUDA_TRY(cudaMemcpyAsync(out, h->dSigs, sizeof(uint64_t) * h->n, cudaMemcpyDeviceToHost,
                             h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return VLCU_OK;
}
// Asynchronous result read-back for pipelined campaigns.  vlcu_results_begin() copies the
// signatures, the status records and the listed members aside ON THE DEVICE, ordered behind the
// work already issued on the model's stream (a few MB at HBM speed), and starts their
// device->host copies on a separate stream; it returns at once, so the next vlcu_run can be
// launched while the results travel.  vlcu_results_wait() blocks until they have arrived.  One
// read-back can be in flight; destinations should be page-locked (vlcu_host_alloc).
int vlcu_results_begin(vlcu_model_t* h, uint64_t* sigs_out, vlcu_status_t* status_out, int n_members,
                       const int* members, void* const* member_dsts) {
    if (!h || n_members < 0 || (n_members && (!members || !member_dsts))) {
        return fail(VLCU_E_ARG, "vlcu_results_begin: bad argument");
    }
    DeviceGuard guard{h->device};
    if (!h->resStream) {
        CUDA_TRY(cudaStreamCreateWithFlags(&h->resStream, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&h->evCaptured, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&h->evFetched, cudaEventDisableTiming));
    }
    const uint64_t n = h->n;
    size_t need = 0;
    const size_t sigOff = need;
    if (sigs_out) need += sizeof(uint64_t) * n;
    const size_t ctxOff = need;
    if (status_out) need += sizeof(VlcuCtxRec) * n;
    std::vector<size_t> memOff(n_members);
    for (int k = 0; k < n_members; ++k) {
        const VlcuMemberDesc* const mdp = memberAt(h, members[k]);
        if (!mdp || !member_dsts[k]) return fail(VLCU_E_ARG, "vlcu_results_begin: bad member");
        need = (need + 15) & ~size_t{15};
        memOff[k] = need;
        need += static_cast<size_t>(mdp->elemBytes) * mdp->elems * n;
    }
    if (need > h->resBytes) {
        if (h->fetchPending) CUDA_TRY(cudaStreamSynchronize(h->resStream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        cudaFree(h->dRes);
        h->dRes = nullptr;
        h->resBytes = 0;
        CUDA_TRY(cudaMalloc(&h->dRes, need));
        h->resBytes = need;
    }
    // the previous read-back must have left the device buffer before it is overwritten
    if (h->fetchPending) CUDA_TRY(cudaStreamWaitEvent(h->stream, h->evFetched, 0));
    if (sigs_out) {
        CUDA_TRY(cudaMemcpyAsync(h->dRes + sigOff, h->dSigs, sizeof(uint64_t) * n, cudaMemcpyDeviceToDevice,
                                 h->stream));
    }
    if (status_out) {
        CUDA_TRY(cudaMemcpyAsync(h->dRes + ctxOff, h->dCtx, sizeof(VlcuCtxRec) * n, cudaMemcpyDeviceToDevice,
                                 h->stream));
    }
    for (int k = 0; k < n_members; ++k) {
        const VlcuMemberDesc& md = h->infop->members[members[k]];
        const size_t recBytes = static_cast<size_t>(md.elemBytes) * md.elems;
        if (memberIsFlat(h, md)) {  // instance-major already
            const uint8_t* const soap = h->dState + static_cast<uint64_t>(md.soaOffset) * h->npad;
            CUDA_TRY(cudaMemcpyAsync(h->dRes + memOff[k], soap, recBytes * n, cudaMemcpyDeviceToDevice,
                                     h->stream));
            continue;
        }
        if (const int rc = launchGather(h, md, h->dRes + memOff[k])) return rc;
    }
    CUDA_TRY(cudaEventRecord(h->evCaptured, h->stream));
    CUDA_TRY(cudaStreamWaitEvent(h->resStream, h->evCaptured, 0));
    if (sigs_out) {
        CUDA_TRY(cudaMemcpyAsync(sigs_out, h->dRes + sigOff, sizeof(uint64_t) * n, cudaMemcpyDeviceToHost,
                                 h->resStream));
    }
    if (status_out) {
        CUDA_TRY(cudaMemcpyAsync(status_out, h->dRes + ctxOff, sizeof(VlcuCtxRec) * n, cudaMemcpyDeviceToHost,
                                 h->resStream));
    }
    for (int k = 0; k < n_members; ++k) {
        const VlcuMemberDesc& md = h->infop->members[members[k]];
        CUDA_TRY(cudaMemcpyAsync(member_dsts[k], h->dRes + memOff[k],
                                 static_cast<size_t>(md.elemBytes) * md.elems * n, cudaMemcpyDeviceToHost,
                                 h->resStream));
    }
    CUDA_TRY(cudaEventRecord(h->evFetched, h->resStream));
    h->fetchPending = true;
    return VLCU_OK;
}
int vlcu_results_wait(vlcu_model_t* h) {
    if (!h) return fail(VLCU_E_ARG, "null model");
    if (!h->fetchPending) return VLCU_OK;
    DeviceGuard guard{h->device};
    CUDA_TRY(cudaEventSynchronize(h->evFetched));
    h->fetchPending = false;
    return VLCU_OK;
}
int vlcu_reset_signatures(vlcu_model_t* h) {
    if (!h) return fail(VLCU_E_ARG, "null model");
    DeviceGuard guard{h->device};
    std::vector<uint64_t> init(h->npad, VLCU_SIG_INIT);
    CUDA_TRY(cudaMemcpyAsync(h->dSigs, init.data(), sizeof(uint64_t) * h->npad,
                             cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return VLCU_OK;
}
int vlcu_messages(vlcu_model_t* h, uint64_t instance, vlcu_msg_t* out, uint32_t cap,
                  uint32_t* count) {
    if (!h || instance >= h->n || !count) return fail(VLCU_E_ARG, "bad argument");
    DeviceGuard guard{h->device};
    VlcuCtxRec rec;
    CUDA_TRY(cudaMemcpyAsync(&rec, h->dCtx + instance, sizeof(rec), cudaMemcpyDeviceToHost,
                             h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    *count = rec.msgCount;
    uint32_t ncopy = rec.msgCount < h->dev.msgCap ? rec.msgCount : h->dev.msgCap;
    if (ncopy > cap) ncopy = cap;
    if (ncopy && out) {
        CUDA_TRY(cudaMemcpyAsync(out, h->dMsgs + instance * h->dev.msgCap, sizeof(VlcuMsg) * ncopy,
                                 cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
    }
    return VLCU_OK;
}
const char* vlcu_message_format(const vlcu_model_t* h, uint32_t id) {
    if (!h || id >= h->infop->nMessages) return nullptr;
    return h->infop->messages[id];
}

int vlcu_snapshot(vlcu_model_t* h, uint64_t instance, void* out, size_t out_bytes) {
    if (!h || !out || instance >= h->n) return fail(VLCU_E_ARG, "bad argument");
    const VlcuDesignInfo& info = *h->infop;
    if (out_bytes < info.stateBytes) return fail(VLCU_E_ARG, "snapshot buffer too small");
    DeviceGuard guard{h->device};
    std::vector<uint32_t> table(info.nMembers * 4);
    for (uint32_t m = 0; m < info.nMembers; ++m) {
        table[m * 4 + 0] = info.members[m].soaOffset;
        table[m * 4 + 1] = info.members[m].elemBytes;
        table[m * 4 + 2] = info.members[m].elems;
        table[m * 4 + 3] = memberLaneMajor(info.members[m]);
    }
    const size_t tableBytes = table.size() * sizeof(uint32_t);
    const size_t tableOff = (static_cast<size_t>(info.stateBytes) + 15) & ~size_t{15};
    if (const int rc = ensureStage(h, tableOff + tableBytes)) return rc;
    CUDA_TRY(cudaMemsetAsync(h->dStage, 0, info.stateBytes, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->dStage + tableOff, table.data(), tableBytes, cudaMemcpyHostToDevice,
                             h->stream));
    k_snapshot<<<info.nMembers < 64 ? info.nMembers : 64, 128, 0, h->stream>>>(
        h->dStage, h->dState, geomOf(h), instance,
        reinterpret_cast<const uint32_t*>(h->dStage + tableOff), info.nMembers);
    ++h->launches;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out, h->dStage, info.stateBytes, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return VLCU_OK;
}


//-------------------------------------------------------------------------
// verilated_save images
//
// One instance's state in the format VerilatedSave writes for `os << *top` of a --savable build
// of the same design (include/verilated_save.cpp:56-58,87-120,256-270; V3EmitCSyms.cpp:1564-1582;
// V3EmitCImp.cpp:277-350), so that a failing instance of a GPU campaign can be loaded into the
// ordinary C++ model (`VerilatedRestore os; os.open(f); os >> *top;`) and replayed under a
// debugger -- and the other way round.
//
//   "verilatorsave02\n"
//   [bool __Vm_activity]            (--trace builds)
//   bool  __Vm_didInit
//   u64   __Vcheckval               structure checksum of the module's variables
//   VerilatedContext::Serialized    raw struct bytes (layout restated below)
//   u32 len + bytes                 timeFormatSuffix
//   u32 len + bytes                 dumpfile
//   members in __Vserialize order   1/2/4/8-byte scalars, arrays element by element, wide values
//                                   word by word
//   "vltsaved"

// VerilatedContext::Serialized of this reference revision (include/verilated.h:402-441) as the
// x86-64 ABI lays it out: 160 bytes.  kCtxDefault is the image of a freshly constructed context
// (tests/test_save.py compares it, and every offset below, with the genuine runtime).
enum : size_t {
    CTXS_SIZE = 160,
    CTXS_TIME = 0,  // uint64_t m_time
    CTXS_ASSERT_ON = 8,  // uint32 m_assertOn, m_assertLock, m_assertPassOnVacuous, ..NonVacuous, m_assertFailOn
    CTXS_ASSERT_KILL = 28,  // uint32 m_assertKill[24]
    CTXS_CALC_UNUSED = 124,  // bool x6: m_calcUnusedSigs, m_fatalOnError, m_fatalOnVpiError,
    CTXS_FATAL_ON_ERROR = 125,  //          m_gotError, m_gotFinish, m_quiet
    CTXS_FATAL_ON_VPI_ERROR = 126,
    CTXS_GOT_ERROR = 127,
    CTXS_GOT_FINISH = 128,
    CTXS_QUIET = 129,
    CTXS_TIMEUNIT = 130,  // int8 m_timeunit, m_timeprecision
    CTXS_TIMEPRECISION = 131,
    CTXS_ERROR_COUNT = 132,  // int x7: m_errorCount, m_errorLimit, m_randReset, m_randSeed,
    CTXS_ERROR_LIMIT = 136,  //         m_timeFormatUnits, m_timeFormatPrecision, m_timeFormatWidth
    CTXS_RAND_RESET = 140,
    CTXS_RAND_SEED = 144,
    CTXS_TIMEFORMAT_UNITS = 148,
    CTXS_TIMEFORMAT_PRECISION = 152,
    CTXS_TIMEFORMAT_WIDTH = 156,
};
static void ctxDefaultImage(uint8_t out[CTXS_SIZE]) {
    memset(out, 0, CTXS_SIZE);
    const uint32_t ones = 0xffffffffU, zero = 0U;
    memcpy(out + CTXS_ASSERT_ON + 0, &ones, 4);  // m_assertOn: everything enabled
    memcpy(out + CTXS_ASSERT_ON + 4, &zero, 4);  // m_assertLock
    memcpy(out + CTXS_ASSERT_ON + 8, &ones, 4);  // m_assertPassOnVacuous
    memcpy(out + CTXS_ASSERT_ON + 12, &ones, 4);  // m_assertPassOnNonvacuous
    memcpy(out + CTXS_ASSERT_ON + 16, &ones, 4);  // m_assertFailOn
    out[CTXS_FATAL_ON_ERROR] = 1;
    out[CTXS_FATAL_ON_VPI_ERROR] = 1;
    out[CTXS_TIMEUNIT] = static_cast<uint8_t>(-12);  // picoseconds until _Vconfigure says otherwise
    out[CTXS_TIMEPRECISION] = static_cast<uint8_t>(-12);
    const int32_t one = 1, unitsNone = 99, width = 20;
    memcpy(out + CTXS_ERROR_LIMIT, &one, 4);
    memcpy(out + CTXS_RAND_SEED, &one, 4);
    memcpy(out + CTXS_TIMEFORMAT_UNITS, &unitsNone, 4);
    memcpy(out + CTXS_TIMEFORMAT_WIDTH, &width, 4);
}
static const char kSaveHeader[] = "verilatorsave02\n";
static const char kSaveTrailer[] = "vltsaved";

int vlcu_save_context_image(void* out, size_t out_bytes) {
    if (!out || out_bytes < CTXS_SIZE) return fail(VLCU_E_ARG, "buffer too small");
    ctxDefaultImage(static_cast<uint8_t*>(out));
    return static_cast<int>(CTXS_SIZE);
}

static int saveCheck(const vlcu_model_t* h, uint64_t instance) {
    if (!h || instance >= h->n) return fail(VLCU_E_ARG, "bad argument");
    const VlcuDesignInfo& info = *h->infop;
    if (!info.saveOrder || !info.nSaveOrder) return fail(VLCU_E_STATE, "design has no save table");
    if (info.saveUnsupported) {
        return fail(VLCU_E_STATE, std::string{"cannot serialize member '"} + info.saveUnsupported + "'");
    }
    if (info.coverBins) {
        return fail(VLCU_E_STATE, "--savable and --coverage exclude each other (V3Options.cpp:1069)");
    }
    return VLCU_OK;
}

int vlcu_save_instance(vlcu_model_t* h, uint64_t instance, const char* filename) {
    if (const int rc = saveCheck(h, instance)) return rc;
    if (!filename) return fail(VLCU_E_ARG, "null filename");
    const VlcuDesignInfo& info = *h->infop;
    DeviceGuard guard{h->device};
    std::vector<uint8_t> rec(info.stateBytes);
    if (const int rc = vlcu_snapshot(h, instance, rec.data(), rec.size())) return rc;
    VlcuCtxRec ctx;
    CUDA_TRY(cudaMemcpyAsync(&ctx, h->dCtx + instance, sizeof(ctx), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    std::string out;
    const auto put = [&out](const void* p, size_t n) { out.append(static_cast<const char*>(p), n); };
    put(kSaveHeader, sizeof(kSaveHeader) - 1);
    const uint8_t didInit = (ctx.status & VLCU_ST_DIDINIT) ? 1 : 0;
    if (info.saveFlags & VLCU_SAVE_ACTIVITY) put(&didInit, 1);  // eval_step sets __Vm_activity; only a dump clears it
    put(&didInit, 1);
    put(&info.saveCheckval, 8);
    uint8_t cimg[CTXS_SIZE];
    ctxDefaultImage(cimg);
    memcpy(cimg + CTXS_TIME, &ctx.time, 8);
    if (info.saveTimeunit >= 0) cimg[CTXS_TIMEUNIT] = static_cast<uint8_t>(info.saveTimeunit);
    if (info.saveTimeprecision >= 0) cimg[CTXS_TIMEPRECISION] = static_cast<uint8_t>(info.saveTimeprecision);
    cimg[CTXS_GOT_ERROR] = (ctx.status & VLCU_ST_ERROR) ? 1 : 0;
    cimg[CTXS_GOT_FINISH] = (ctx.status & VLCU_ST_FINISH) ? 1 : 0;
    const int32_t errorCount = static_cast<int32_t>(ctx.errorCount);
    const int32_t errorLimit = static_cast<int32_t>(h->dev.errorLimit);
    const int32_t randReset = static_cast<int32_t>(h->dev.randReset);
    const int32_t randSeed = static_cast<int32_t>(h->dev.seedBase ^ (h->dev.instBase + instance));
    memcpy(cimg + CTXS_ERROR_COUNT, &errorCount, 4);
    memcpy(cimg + CTXS_ERROR_LIMIT, &errorLimit, 4);
    memcpy(cimg + CTXS_RAND_RESET, &randReset, 4);
    memcpy(cimg + CTXS_RAND_SEED, &randSeed, 4);
    put(cimg, CTXS_SIZE);
    const uint32_t emptyLen = 0;
    put(&emptyLen, 4);  // timeFormatSuffix ""
    put(&emptyLen, 4);  // dumpfile ""
    for (uint32_t k = 0; k < info.nSaveOrder; ++k) {
        const VlcuMemberDesc& md = info.members[info.saveOrder[k]];
        put(rec.data() + recordOffset(h, md), static_cast<size_t>(md.elemBytes) * md.elems);
    }
    put(kSaveTrailer, sizeof(kSaveTrailer) - 1);
    std::ofstream os{filename, std::ios::binary};
    if (os.fail()) return fail(VLCU_E_IO, std::string{"Can't write '"} + filename + "'");
    os << out;
    return os.good() ? VLCU_OK : fail(VLCU_E_IO, "write failed");
}

int vlcu_restore_instance(vlcu_model_t* h, uint64_t instance, const char* filename) {
    if (const int rc = saveCheck(h, instance)) return rc;
    if (!filename) return fail(VLCU_E_ARG, "null filename");
    const VlcuDesignInfo& info = *h->infop;
    std::ifstream is{filename, std::ios::binary};
    if (is.fail()) return fail(VLCU_E_IO, std::string{"Can't read '"} + filename + "'");
    const std::string in{std::istreambuf_iterator<char>{is}, std::istreambuf_iterator<char>{}};
    size_t pos = 0;
    bool short_ = false;
    const auto get = [&](void* p, size_t n) {
        if (pos + n > in.size()) {
            short_ = true;
            memset(p, 0, n);
            return;
        }
        memcpy(p, in.data() + pos, n);
        pos += n;
    };
    char hdr[sizeof(kSaveHeader) - 1];
    get(hdr, sizeof(hdr));
    if (short_ || memcmp(hdr, kSaveHeader, sizeof(hdr))) {
        return fail(VLCU_E_IO, "Can't deserialize; file has wrong header signature");
    }
    uint8_t activity = 0, didInit = 0;
    if (info.saveFlags & VLCU_SAVE_ACTIVITY) get(&activity, 1);
    get(&didInit, 1);
    uint64_t checkval = 0;
    get(&checkval, 8);
    if (short_ || checkval != info.saveCheckval) {
        return fail(VLCU_E_STATE, "Can't deserialize; the save file is of a different model (checksum)");
    }
    uint8_t cimg[CTXS_SIZE];
    get(cimg, CTXS_SIZE);
    for (int k = 0; k < 2; ++k) {  // timeFormatSuffix, dumpfile
        uint32_t len = 0;
        get(&len, 4);
        if (pos + len > in.size()) short_ = true;
        pos += len;
    }
    std::vector<uint8_t> rec(info.stateBytes, 0);
    std::vector<uint32_t> table(info.nMembers * 4);
    for (uint32_t m = 0; m < info.nMembers; ++m) {
        table[m * 4 + 0] = info.members[m].soaOffset;
        table[m * 4 + 1] = info.members[m].elemBytes;
        table[m * 4 + 2] = info.members[m].elems;
        table[m * 4 + 3] = (info.members[m].kind == VLCU_MEMBER_MEMORY ? 1U : 0U) | 2U;  // skipped unless listed
    }
    for (uint32_t k = 0; k < info.nSaveOrder && !short_; ++k) {
        const uint32_t m = info.saveOrder[k];
        const VlcuMemberDesc& md = info.members[m];
        get(rec.data() + recordOffset(h, md), static_cast<size_t>(md.elemBytes) * md.elems);
        table[m * 4 + 3] &= ~2U;
    }
    char trl[sizeof(kSaveTrailer) - 1];
    get(trl, sizeof(trl));
    if (short_ || memcmp(trl, kSaveTrailer, sizeof(trl)) || pos != in.size()) {
        return fail(VLCU_E_IO, "Can't deserialize; file has wrong end-of-file signature");
    }
    // per-shard settings a single instance cannot change
    int32_t randReset, randSeed;
    memcpy(&randReset, cimg + CTXS_RAND_RESET, 4);
    memcpy(&randSeed, cimg + CTXS_RAND_SEED, 4);
    if (static_cast<uint32_t>(randReset) != h->dev.randReset
        || randSeed != static_cast<int32_t>(h->dev.seedBase ^ (h->dev.instBase + instance))) {
        return fail(VLCU_E_STATE, "the image's randReset/randSeed are not those of this instance "
                                  "(create the model with the seed and instance base of the saved run)");
    }
    DeviceGuard guard{h->device};
    const size_t tableBytes = table.size() * sizeof(uint32_t);
    const size_t tableOff = (static_cast<size_t>(info.stateBytes) + 15) & ~size_t{15};
    if (const int rc = ensureStage(h, tableOff + tableBytes)) return rc;
    CUDA_TRY(cudaMemcpyAsync(h->dStage, rec.data(), info.stateBytes, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->dStage + tableOff, table.data(), tableBytes, cudaMemcpyHostToDevice,
                             h->stream));
    k_restore<<<info.nMembers < 64 ? info.nMembers : 64, 128, 0, h->stream>>>(
        h->dState, h->dStage, geomOf(h), instance, reinterpret_cast<const uint32_t*>(h->dStage + tableOff),
        info.nMembers);
    ++h->launches;
    CUDA_TRY(cudaGetLastError());
    VlcuCtxRec ctx;
    CUDA_TRY(cudaMemcpyAsync(&ctx, h->dCtx + instance, sizeof(ctx), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    memcpy(&ctx.time, cimg + CTXS_TIME, 8);
    int32_t errorCount;
    memcpy(&errorCount, cimg + CTXS_ERROR_COUNT, 4);
    ctx.errorCount = static_cast<uint32_t>(errorCount);
    ctx.status &= ~(VLCU_ST_DIDINIT | VLCU_ST_FINISH | VLCU_ST_ERROR);
    if (didInit) ctx.status |= VLCU_ST_DIDINIT;
    if (cimg[CTXS_GOT_FINISH]) ctx.status |= VLCU_ST_FINISH;
    if (cimg[CTXS_GOT_ERROR]) ctx.status |= VLCU_ST_ERROR;
    CUDA_TRY(cudaMemcpyAsync(h->dCtx + instance, &ctx, sizeof(ctx), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return VLCU_OK;
}

//-------------------------------------------------------------------------
// Coverage

uint32_t vlcu_coverage_bins(const vlcu_model_t* h) { return h ? h->infop->coverBins : 0; }
uint32_t vlcu_coverage_points(const vlcu_model_t* h) { return h ? h->infop->nCoverPoints : 0; }
int vlcu_coverage_point(const vlcu_model_t* h, uint32_t point, uint32_t* bin, const char** filename,
                        int* lineno, int* column, const char** hier, const char** page,
                        const char** comment, const char** linescov) {
    if (!h || point >= h->infop->nCoverPoints) return fail(VLCU_E_ARG, "bad coverage point");
    const VlcuCoverPoint& p = h->infop->coverPoints[point];
    if (bin) *bin = p.bin;
    if (filename) *filename = p.filename;
    if (lineno) *lineno = p.lineno;
    if (column) *column = p.column;
    if (hier) *hier = p.hier;
    if (page) *page = p.page;
    if (comment) *comment = p.comment;
    if (linescov) *linescov = p.linescov;
    return VLCU_OK;
}
int vlcu_coverage_reduce(vlcu_model_t* h, uint64_t* out) {
    if (!h || !out) return fail(VLCU_E_ARG, "null argument");
    const uint32_t bins = h->infop->coverBins;
    if (!bins) return fail(VLCU_E_STATE, "design was not built with coverage");
    DeviceGuard guard{h->device};
    if (h->infop->tile) {
        if (!h->dCovBinOffset) {
            std::vector<uint32_t> offsets;
            for (uint32_t s = 0; s < h->infop->nCoverScopes; ++s) {
                const VlcuCoverScope& cs = h->infop->coverScopes[s];
                for (uint32_t b = 0; b < cs.bins; ++b) offsets.push_back(cs.tileOffset + b * 8U);
            }
            if (offsets.size() != bins) return fail(VLCU_E_STATE, "coverage scope table does not add up to coverBins");
            CUDA_TRY(cudaMalloc(&h->dCovBinOffset, sizeof(uint32_t) * bins));
            CUDA_TRY(cudaMemcpy(h->dCovBinOffset, offsets.data(), sizeof(uint32_t) * bins, cudaMemcpyHostToDevice));
        }
        k_cov_reduce_tiled<<<blocksFor(bins, 32), 256, 0, h->stream>>>(
            h->dState, h->npad / h->infop->tile, static_cast<uint64_t>(h->infop->stateBytes) * h->infop->tile,
            h->dCovBinOffset, bins, h->dCovSums);
    } else if (h->infop->nCoverScopes) {
        // register form, counters per warp of 32 instances: [warp][bin] uint64 in the block of __Vcoverage
        if (!h->dCovBinOffset) {
            std::vector<uint32_t> offsets(bins);
            for (uint32_t b = 0; b < bins; ++b) offsets[b] = b * 8U;
            CUDA_TRY(cudaMalloc(&h->dCovBinOffset, sizeof(uint32_t) * bins));
            CUDA_TRY(cudaMemcpy(h->dCovBinOffset, offsets.data(), sizeof(uint32_t) * bins, cudaMemcpyHostToDevice));
        }
        // (N_Bins of VlcuWarpCovArray: the block's words per instance x 16)
        const uint64_t warpBins = static_cast<uint64_t>((bins + 15U) / 16U) * 16U;
        k_cov_reduce_tiled<<<blocksFor(bins, 32), 256, 0, h->stream>>>(
            h->dState + static_cast<uint64_t>(h->infop->coverageSoaOffset) * h->npad, (h->n + 31U) / 32U,
            warpBins * 8U, h->dCovBinOffset, bins, h->dCovSums);
    } else {
        const uint32_t* const covp = reinterpret_cast<const uint32_t*>(
            h->dState + static_cast<uint64_t>(h->infop->coverageSoaOffset) * h->npad);
        k_cov_reduce<<<bins < 148 * 8 ? bins : 148 * 8, 256, 0, h->stream>>>(covp, h->n, h->npad, bins,
                                                                           h->dCovSums);
    }
    ++h->launches;
    CUDA_TRY(cudaGetLastError());
    if (h->comm) {  // sum the per-shard totals over NVLink
        const ncclResult_t nr = nccl().AllReduce(h->dCovSums, h->dCovSums, bins, ncclUint64,
                                                 ncclSum, h->comm, h->stream);
        if (nr != ncclSuccess) {
            return fail(VLCU_E_NCCL, std::string{"ncclAllReduce: "} + nccl().GetErrorString(nr));
        }
    }
    CUDA_TRY(cudaMemcpyAsync(out, h->dCovSums, sizeof(uint64_t) * bins, cudaMemcpyDeviceToHost,
                             h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return VLCU_OK;
}

namespace {
// Restatement of VerilatedCovImp::dequote / keyValueFormatter / combineHier
// (include/verilated_cov.cpp:143-171,175-220)
std::string covDequote(const std::string& text) {
    std::string out;
    for (const char c : text) {
        if (!std::isprint(static_cast<unsigned char>(c)) || c == '%' || c == '"') {
            char hex[8];
            snprintf(hex, sizeof(hex), "%%%02X", static_cast<unsigned char>(c));
            out += hex;
        } else {
            out += c;
        }
    }
    return out;
}
std::string covKeyValue(const std::string& key, const std::string& value) {
    std::string out = "\001";
    out += (key.length() == 1 && std::isalpha(static_cast<unsigned char>(key[0])))
               ? key
               : covDequote(key);
    out += "\002" + covDequote(value);
    return out;
}
std::string covCombineHier(const std::string& old, const std::string& add) {
    if (old == add) return add;
    if (old.empty()) return add;
    if (add.empty()) return old;
    size_t pre = 0;
    while (old[pre] == add[pre]) ++pre;  // strings differ, so this stops inside both
    size_t apost = old.size() - 1, bpost = add.size() - 1;
    while (old[apost] == add[bpost] && apost > pre && bpost > pre) {
        --apost;
        --bpost;
    }
    const std::string suffix = bpost < add.size() ? add.substr(bpost + 1) : std::string{};
    return old.substr(0, pre) + "*" + suffix;
}
}  // namespace

int vlcu_coverage_write(const vlcu_model_t* h, const uint64_t* counts, const char* filename) {
    if (!h || !counts || !filename) return fail(VLCU_E_ARG, "null argument");
    const VlcuDesignInfo& info = *h->infop;
    std::ofstream os{filename};
    if (os.fail()) return fail(VLCU_E_IO, std::string{"Can't write '"} + filename + "'");
    os << "# SystemC::Coverage-3\n";
    // name -> (hier, count), ordered by name like the host writer's std::map
    std::map<std::string, std::pair<std::string, uint64_t>> events;
    for (uint32_t b = 0; b < info.nCoverPoints; ++b) {
        const VlcuCoverPoint& p = info.coverPoints[b];
        // Key order after VL_COVER_INSERT's de-duplication: filename, lineno, column, hier,
        // page (preceded by the derived type), comment, then linescov when present
        std::string name = covKeyValue("f", p.filename);
        name += covKeyValue("l", std::to_string(p.lineno));
        name += covKeyValue("n", std::to_string(p.column));
        const std::string page = p.page;
        name += covKeyValue("t", page.substr(2, page.find('/') - 2));
        name += covKeyValue("page", page);
        name += covKeyValue("o", p.comment);
        if (p.linescov && p.linescov[0]) name += covKeyValue("S", p.linescov);
        std::string hier = std::string{"TOP"} + p.hier;  // vlNamep + hierp
        if (!hier.empty() && hier[0] == '.') hier = hier.substr(1);
        const auto it = events.find(name);
        if (it != events.end()) {
            it->second.second += counts[p.bin];
            it->second.first = covCombineHier(it->second.first, hier);
        } else {
            events.emplace(name, std::make_pair(hier, counts[p.bin]));
        }
    }
    for (const auto& ev : events) {
        os << "C '" << ev.first;
        if (!ev.second.first.empty()) os << covKeyValue("h", ev.second.first);
        os << "' " << ev.second.second << '\n';
    }
    return os.good() ? VLCU_OK : fail(VLCU_E_IO, "write failed");
}

//-------------------------------------------------------------------------
// Waveforms: VCD text of one sampled instance from the trace records of vlcu_run.
// Restates, for the subset a Verilated model uses (bit / bus / quad / wide / array elements):
//   header + declarations   include/verilated_vcd_c.cpp:117-129, 317-450
//   identifier codes        include/verilated_vcd_c.cpp:400-426  ('!' + code % 94, least digit first)
//   time stamps             include/verilated_vcd_c.cpp:191-203  (an empty one is overwritten)
//   full dump on the first call, changes afterwards, signals in code order
//                           include/verilated_trace_imp.h (dump / runCallbacks, full* / chg*)
//   value text              include/verilated_vcd_c.cpp:620-690  (all `bits` digits, MSB first)

namespace {
struct VcdSig final {
    uint32_t code;  // word index in the trace sample
    uint32_t bits;
    bool bit;  // declared with DECL_BIT
    std::string ident;
};
std::string vcdIdent(uint32_t code) {
    std::string id;
    do {
        id += static_cast<char>('!' + (code % 94));
        code /= 94;
    } while (code--);
    return id;
}
std::string vcdLastWord(const std::string& s) {
    const size_t pos = s.rfind(' ');
    return pos == std::string::npos ? s : s.substr(pos + 1);
}
bool vcdProperScope(int32_t type) {
    return type == VLCU_PREFIX_SCOPE_MODULE || type == VLCU_PREFIX_SCOPE_INTERFACE
           || type == VLCU_PREFIX_STRUCT_PACKED || type == VLCU_PREFIX_STRUCT_UNPACKED
           || type == VLCU_PREFIX_UNION_PACKED;
}
}  // namespace

// One dump(time) call of the sampled instance: the time and the value of every trace word.
// Returns false when there are no more dumps.
typedef std::function<bool(uint64_t& time, const uint32_t*& valsp)> VcdDumpSource;

static int vcdWriteCore(const vlcu_model_t* h, const VcdDumpSource& nextDump, const char* filename) {
    const VlcuDesignInfo& info = *h->infop;
    if (!info.traceDecls || !info.traceWords) {
        return fail(VLCU_E_STATE, "design was not Verilated with --trace");
    }
    std::string out = "$version Generated by VerilatedVcd $end\n$timescale 1ps $end\n";
    int indent = 1;
    const auto printIndent = [&](int change) {
        if (change < 0) indent += change;
        out.append(static_cast<size_t>(indent), ' ');
        if (change > 0) indent += change;
    };
    // VerilatedVcd::pushPrefix / popPrefix / declare (include/verilated_vcd_c.cpp:323-450)
    std::vector<std::pair<std::string, int32_t>> prefixStack{{"", VLCU_PREFIX_SCOPE_MODULE}};
    const auto pushPrefix = [&](const std::string& name, int32_t type) {
        const std::string prevPrefix = prefixStack.back().first;
        if ((name == "$rootio" && !prevPrefix.empty()) || name.empty()) {
            prefixStack.emplace_back(prevPrefix, VLCU_PREFIX_ROOTIO_WRAPPER);
            return;
        }
        const std::string newPrefix = prevPrefix + name;
        const bool proper = vcdProperScope(type);
        if (proper) {
            printIndent(1);
            out += "$scope module " + vcdLastWord(newPrefix) + " $end\n";
        }
        prefixStack.emplace_back(newPrefix + (proper ? " " : ""), type);
    };
    const auto popPrefix = [&]() {
        if (vcdProperScope(prefixStack.back().second)) {
            printIndent(-1);
            out += "$upscope $end\n";
        }
        prefixStack.pop_back();
    };
    std::map<uint32_t, VcdSig> sigs;  // by trace code
    const uint32_t baseCode = 1;  // VerilatedTrace::nextCode() starts at 1
    pushPrefix("TOP", VLCU_PREFIX_SCOPE_MODULE);  // the model's name()
    for (uint32_t d = 0; d < info.nTraceDecls; ++d) {
        const VlcuTraceDecl& td = info.traceDecls[d];
        if (td.op == VLCU_TRACE_PUSH) {
            pushPrefix(td.name, td.arg);
        } else if (td.op == VLCU_TRACE_POP) {
            popPrefix();
        } else {
            const uint32_t bits = static_cast<uint32_t>(td.arg);
            const std::string ident = vcdIdent(baseCode + td.code);
            const std::string hierName = prefixStack.back().first + td.name;
            printIndent(0);
            out += "$var wire " + std::to_string(bits) + ' ' + ident + ' ' + vcdLastWord(hierName);
            if (td.elem >= 0) out += '[' + std::to_string(td.elem) + ']';
            if (td.bussed) out += " [" + std::to_string(td.msb) + ':' + std::to_string(td.lsb) + ']';
            out += " $end\n";
            sigs.emplace(td.code, VcdSig{td.code, bits, !td.bussed, ident});
        }
    }
    popPrefix();
    out += "$enddefinitions $end\n\n\n";

    // values at the last dump that was written (a private copy: the source may reuse its buffer)
    std::vector<uint32_t> prev(info.traceWords);
    const uint32_t* prevp = nullptr;
    uint64_t lastTime = 0;
    size_t timeBegin = std::string::npos, timeEnd = std::string::npos;
    uint64_t time = 0;
    const uint32_t* valsp = nullptr;
    while (nextDump(time, valsp)) {
        if (prevp && time <= lastTime) continue;  // "dump at same time" is skipped by dump()
        // emitTimeChange: a time stamp nothing followed is overwritten
        if (timeBegin != std::string::npos && timeEnd == out.size()) out.resize(timeBegin);
        timeBegin = out.size();
        out += '#' + std::to_string(time) + '\n';
        timeEnd = out.size();
        for (uint32_t k = 0; k < info.nTraceOrder; ++k) {
            const auto it = sigs.find(info.traceOrder[k]);
            if (it == sigs.end()) continue;
            const VcdSig& sg = it->second;
            const uint32_t words = (sg.bits + 31) / 32;
            if (prevp && !memcmp(valsp + sg.code, prevp + sg.code, 4U * words)) continue;
            // emitBit / emitCData..emitWData: all `bits` digits, MSB first
            if (sg.bit) {
                out += static_cast<char>('0' | (valsp[sg.code] & 1U));
            } else {
                out += 'b';
                for (int b = static_cast<int>(sg.bits) - 1; b >= 0; --b) {
                    out += static_cast<char>('0' | ((valsp[sg.code + (b >> 5)] >> (b & 31)) & 1U));
                }
            }
            if (sg.bits != 1) out += ' ';  // 1-bit values have no separator before the identifier
            out += sg.ident;
            out += '\n';
        }
        memcpy(prev.data(), valsp, sizeof(uint32_t) * info.traceWords);
        prevp = prev.data();
        lastTime = time;
    }
    std::ofstream os{filename, std::ios::binary};
    if (os.fail()) return fail(VLCU_E_IO, std::string{"Can't write '"} + filename + "'");
    os << out;
    return os.good() ? VLCU_OK : fail(VLCU_E_IO, "write failed");
}

int vlcu_vcd_write(const vlcu_model_t* h, const void* trace, uint64_t trace_evals,
                   uint32_t n_samples, uint32_t sample, const char* filename) {
    if (!h || !trace || !filename || sample >= n_samples) return fail(VLCU_E_ARG, "bad argument");
    const VlcuDesignInfo& info = *h->infop;
    const size_t recBytes = VLCU_TRACE_HDR + static_cast<size_t>(info.stateBytes)
                            + 4U * VLCU_TRACE_WORDS_PADDED(info.traceWords);
    const uint8_t* const basep = static_cast<const uint8_t*>(trace);
    uint64_t e = 0;
    return vcdWriteCore(
        h,
        [&](uint64_t& time, const uint32_t*& valsp) {
            for (; e < trace_evals; ++e) {
                const uint8_t* const recp = basep + (e * n_samples + sample) * recBytes;
                uint32_t valid;
                memcpy(&valid, recp + 12, sizeof(valid));
                if (!valid) continue;  // the instance had stopped (stop-on-finish)
                memcpy(&time, recp, sizeof(time));
                valsp = reinterpret_cast<const uint32_t*>(recp + VLCU_TRACE_HDR + info.stateBytes);
                ++e;
                return true;
            }
            return false;
        },
        filename);
}

int vlcu_vcd_write_changes(const vlcu_model_t* h, const uint32_t* streams, uint32_t stream_words,
                           uint32_t n_samples, uint32_t sample, const char* filename) {
    if (!h || !streams || !filename || sample >= n_samples || stream_words < VLCU_TRACE_STREAM_HDR) {
        return fail(VLCU_E_ARG, "bad argument");
    }
    const VlcuDesignInfo& info = *h->infop;
    const uint32_t* const sp = streams + static_cast<size_t>(sample) * stream_words;
    const uint32_t used = sp[0] ? sp[0] : VLCU_TRACE_STREAM_HDR;
    if (used > stream_words) return fail(VLCU_E_ARG, "corrupt trace stream: used > capacity");
    if (sp[1]) {
        return fail(VLCU_E_STATE, "trace stream overflowed: " + std::to_string(sp[1])
                                      + " dump(s) dropped; give vlcu_run a larger trace_stream_words");
    }
    std::vector<uint32_t> vals(info.traceWords, 0U);  // running value of every trace word
    uint32_t pos = VLCU_TRACE_STREAM_HDR;
    bool corrupt = false;
    const int rc = vcdWriteCore(
        h,
        [&](uint64_t& time, const uint32_t*& valsp) {
            if (pos >= used) return false;
            const uint32_t head = sp[pos];
            const uint32_t n = head & ~VLCU_TRACE_DUMP_MARK;
            if (!(head & VLCU_TRACE_DUMP_MARK) || pos + 3U + 2U * static_cast<uint64_t>(n) > used) {
                corrupt = true;
                return false;
            }
            time = sp[pos + 1] | (static_cast<uint64_t>(sp[pos + 2]) << 32);
            pos += 3;
            for (uint32_t k = 0; k < n; ++k, pos += 2) {
                if (sp[pos] >= info.traceWords) {
                    corrupt = true;
                    return false;
                }
                vals[sp[pos]] = sp[pos + 1];
            }
            valsp = vals.data();
            return true;
        },
        filename);
    if (corrupt) return fail(VLCU_E_ARG, "corrupt trace stream");
    return rc;
}

//-------------------------------------------------------------------------
// Multi-GPU

int vlcu_comm_unique_id(void* id_out) {
    if (!id_out) return fail(VLCU_E_ARG, "null argument");
    std::string err;
    if (!nccl().load(err)) return fail(VLCU_E_NCCL, err);
    ncclUniqueId id;
    const ncclResult_t nr = nccl().GetUniqueId(&id);
    if (nr != ncclSuccess) return fail(VLCU_E_NCCL, nccl().GetErrorString(nr));
    memcpy(id_out, &id, sizeof(id));
    return VLCU_OK;
}
int vlcu_comm_init(vlcu_model_t* h, const void* id, int rank, int world) {
    if (!h || !id || rank < 0 || rank >= world) return fail(VLCU_E_ARG, "bad argument");
    std::string err;
    if (!nccl().load(err)) return fail(VLCU_E_NCCL, err);
    DeviceGuard guard{h->device};
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    const ncclResult_t nr = nccl().CommInitRank(&h->comm, world, uid, rank);
    if (nr != ncclSuccess) {
        h->comm = nullptr;
        return fail(VLCU_E_NCCL, std::string{"ncclCommInitRank: "} + nccl().GetErrorString(nr));
    }
    h->commRank = rank;
    h->commWorld = world;
    return VLCU_OK;
}
int vlcu_comm_destroy(vlcu_model_t* h) {
    if (!h) return fail(VLCU_E_ARG, "null model");
    if (h->comm) {
        DeviceGuard guard{h->device};
        nccl().CommDestroy(h->comm);
        h->comm = nullptr;
    }
    return VLCU_OK;
}

}  // extern "C"
