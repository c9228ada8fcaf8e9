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

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#include <stdio.h>
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
extern "C" int vlcu_optest_run_host(int op, OpVec* vecs, uint64_t n) {
    if (op < 0 || op >= OPID__COUNT) return -1;
    for (uint64_t i = 0; i < n; ++i) optest_one(op, &vecs[i]);
    return 0;
}
#endif
