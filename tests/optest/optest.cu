// TEST INFRASTRUCTURE (not product code).
// Device-side operator differential test: one test vector per CUDA thread
// through include/verilated_cuda_funcs.h.  The same dispatch table
// (oracle/optest_ops.inc) is compiled against the reference runtime by
// oracle/ref_ops.cpp; tests/test_ops_*.py compares the two bit for bit.
//
// Built twice (tests/optest/Makefile):
//   libvlcu_optest.so       nvcc, sm_100a  -> vlcu_optest_run_cuda   (-m gpu tests)
//   libvlcu_optest_host.so  g++ -x c++     -> vlcu_optest_run_host   (CPU debugging of the
//                           header only; never part of the product path)
#include "verilated_cuda.h"

#include "../../oracle/optest_vec.h"

#define WO(x) (x)
#define WI(x) (x)

enum {
#define OP(name, body) OPID_##name,
#include "../../oracle/optest_ops.inc"
#undef OP
    OPID__COUNT
};

#if defined(__CUDACC__)
#define OPTEST_FN __device__ __noinline__
#else
#define OPTEST_FN static
#endif

OPTEST_FN void optest_one(int op, OpVec* v) {
    switch (op) {
#define OP(name, body) \
    case OPID_##name: { \
        body; \
        break; \
    }
#include "../../oracle/optest_ops.inc"
#undef OP
    default: break;
    }
}

extern "C" int vlcu_optest_count() { return OPID__COUNT; }

// X-initial reset values and the scope hash (device VlRNG / VL_SCOPED_RAND_RESET_* / murmur)
struct RandReq {
    uint64_t seed, scopeHash, salt;
    int32_t obits, randReset;
    uint32_t kind;  // 0 = I, 1 = Q, 2 = W
    uint32_t pad;
    uint32_t out[16];
    char key[64];  // when key[0]: also hash this string
    uint64_t keyHash;
};
OPTEST_FN void optest_rand_one(RandReq* r) {
    VlcuContext ctx{};
    ctx.m_randSeed = r->seed;
    ctx.m_randReset = static_cast<uint32_t>(r->randReset);
    if (r->kind == 0) {
        r->out[0] = vlcu_scoped_rand_reset_i(ctx, r->obits, r->scopeHash, r->salt);
    } else if (r->kind == 1) {
        const QData q = vlcu_scoped_rand_reset_q(ctx, r->obits, r->scopeHash, r->salt);
        r->out[0] = static_cast<uint32_t>(q);
        r->out[1] = static_cast<uint32_t>(q >> 32);
    } else {
        vlcu_scoped_rand_reset_w(ctx, r->obits, r->out, r->scopeHash, r->salt);
    }
    if (r->key[0]) r->keyHash = VL_MURMUR64_HASH(r->key);
}

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#include <stdio.h>
__global__ void vlcu_optest_rand_kernel(RandReq* reqs, uint64_t n) {
    const uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i < n) optest_rand_one(&reqs[i]);
}
extern "C" int vlcu_optest_rand_cuda(RandReq* hostReqs, uint64_t n) {
    RandReq* d = nullptr;
    if (cudaMalloc(&d, n * sizeof(RandReq)) != cudaSuccess) return -2;
    cudaMemcpy(d, hostReqs, n * sizeof(RandReq), cudaMemcpyHostToDevice);
    vlcu_optest_rand_kernel<<<static_cast<unsigned>((n + 63) / 64), 64>>>(d, n);
    const cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        fprintf(stderr, "vlcu_optest_rand: %s\n", cudaGetErrorString(e));
        cudaFree(d);
        return -3;
    }
    cudaMemcpy(hostReqs, d, n * sizeof(RandReq), cudaMemcpyDeviceToHost);
    cudaFree(d);
    return 0;
}
__global__ void vlcu_optest_kernel(int op, OpVec* vecs, uint64_t n) {
    const uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i < n) optest_one(op, &vecs[i]);
}
extern "C" int vlcu_optest_run_cuda(int op, OpVec* hostVecs, uint64_t n) {
    if (op < 0 || op >= OPID__COUNT) return -1;
    OpVec* dvecs = nullptr;
    cudaError_t e = cudaMalloc(&dvecs, n * sizeof(OpVec));
    if (e != cudaSuccess) {
        fprintf(stderr, "vlcu_optest: cudaMalloc: %s\n", cudaGetErrorString(e));
        return -2;
    }
    cudaMemcpy(dvecs, hostVecs, n * sizeof(OpVec), cudaMemcpyHostToDevice);
    const int threads = 128;
    const unsigned blocks = static_cast<unsigned>((n + threads - 1) / threads);
    vlcu_optest_kernel<<<blocks, threads>>>(op, dvecs, n);
    e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
        fprintf(stderr, "vlcu_optest: kernel op %d: %s\n", op, cudaGetErrorString(e));
        cudaFree(dvecs);
        return -3;
    }
    cudaMemcpy(hostVecs, dvecs, n * sizeof(OpVec), cudaMemcpyDeviceToHost);
    cudaFree(dvecs);
    return 0;
}
#else
extern "C" int vlcu_optest_rand_host(RandReq* reqs, uint64_t n) {
    for (uint64_t i = 0; i < n; ++i) optest_rand_one(&reqs[i]);
    return 0;
}
extern "C" int vlcu_optest_run_host(int op, OpVec* vecs, uint64_t n) {
    if (op < 0 || op >= OPID__COUNT) return -1;
    for (uint64_t i = 0; i < n; ++i) optest_one(op, &vecs[i]);
    return 0;
}
#endif
