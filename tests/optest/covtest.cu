// TEST INFRASTRUCTURE (not product code).
// The per-warp coverage primitives of include/verilated_cuda.h (VlcuTileCovPtr forms: toggle points
// of 1..32, 33..64 and >64 bits, line points, hit-mask flushes) driven by one warp of which only the
// first `lanes` threads are there, on a counter array with guard words on both sides.
// tests/test_cov_primitives_gpu.py compares the counters with a scalar model and checks the guards.
#define VLCU_COV_NOINLINE  // as in a tiled model built --coverage
#include "verilated_cuda.h"

#include <cstdio>

struct CovReq {
    uint32_t mode;  // 0 I (per-lane guard), 1 I (warp guard, vlcu_cov_any), 2 Q, 3 W (4 words), 4 line, 5 hits,
                    // 6 two points on the two sides of a per-lane if: `width` bits at counter 0 (even
                    // lanes) and 32 bits at counter 64 (odd lanes)
    uint32_t width, lanes, rounds;
};

__global__ void covtest_kernel(CovReq req, const uint32_t* newp, const uint32_t* oldp, unsigned long long* counters) {
    const uint32_t lane = threadIdx.x;
    if (lane >= req.lanes) return;
    const VlcuTileCovPtr covp{counters};
    for (uint32_t r = 0; r < req.rounds; ++r) {
        // [round][word 0..3][lane]
        const uint32_t* const np = newp + (r * 4) * 32 + lane;
        const uint32_t* const op = oldp + (r * 4) * 32 + lane;
        if (req.mode == 0 || req.mode == 1) {
            const uint32_t mask = req.width < 32 ? (1U << req.width) - 1U : ~0U;
            const bool changed = ((np[0] ^ op[0]) & mask) != 0;
            if (req.mode == 0 ? changed : vlcu_cov_any(changed)) VL_COV_TOGGLE_CHG_ST_I(req.width, covp, np[0], op[0]);
        } else if (req.mode == 2) {
            const QData n = (static_cast<QData>(np[32]) << 32) | np[0], o = (static_cast<QData>(op[32]) << 32) | op[0];
            const QData mask = req.width < 64 ? (1ULL << req.width) - 1ULL : ~0ULL;
            if (vlcu_cov_any(((n ^ o) & mask) != 0)) VL_COV_TOGGLE_CHG_ST_Q(req.width, covp, n, o);
        } else if (req.mode == 3) {
            VlWide<4> n, o;
            for (int w = 0; w < 4; ++w) {
                n[w] = np[w * 32];
                o[w] = op[w * 32];
            }
            const int words = VL_WORDS_I(req.width);
            if (req.width & 31) {
                n[words - 1] &= (1U << (req.width & 31)) - 1U;
                o[words - 1] &= (1U << (req.width & 31)) - 1U;
            }
            bool changed = false;
            for (int w = 0; w < words; ++w) changed |= n[w] != o[w];
            if (changed) VL_COV_TOGGLE_CHG_ST_W(req.width, covp, n.data(), o.data());
        } else if (req.mode == 6) {
            const uint32_t mask = (1U << req.width) - 1U;
            if (lane & 1U) {
                if (np[0] != op[0]) VL_COV_TOGGLE_CHG_ST_I(32, covp + 64, np[0], op[0]);
            } else {
                if ((np[0] ^ op[0]) & mask) VL_COV_TOGGLE_CHG_ST_I(req.width, covp, np[0], op[0]);
            }
        } else if (req.mode == 4) {
            // a line point per set bit of word 0 (bins 0..31), under per-lane conditions
            for (uint32_t b = 0; b < 32; ++b) {
                if ((np[0] >> b) & 1U) ++covp[b];
            }
        } else {
            vlcu_cov_flush_hits(covp, np[0]);
        }
    }
}

extern "C" int vlcu_covtest_run(const CovReq* req, const uint32_t* newp, const uint32_t* oldp, unsigned long long* counters,
                                uint32_t nCounters) {
    const size_t words = static_cast<size_t>(req->rounds) * 4 * 32;
    uint32_t *dn = nullptr, *dold = nullptr;
    unsigned long long* dc = nullptr;
    if (cudaMalloc(&dn, words * 4) != cudaSuccess || cudaMalloc(&dold, words * 4) != cudaSuccess
        || cudaMalloc(&dc, nCounters * 8ULL) != cudaSuccess) {
        return -2;
    }
    cudaMemcpy(dn, newp, words * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dold, oldp, words * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dc, counters, nCounters * 8ULL, cudaMemcpyHostToDevice);
    covtest_kernel<<<1, 32>>>(*req, dn, dold, dc + nCounters / 4);  // counters start after the first guard quarter
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) fprintf(stderr, "vlcu_covtest: %s\n", cudaGetErrorString(e));
    cudaMemcpy(counters, dc, nCounters * 8ULL, cudaMemcpyDeviceToHost);
    cudaFree(dn);
    cudaFree(dold);
    cudaFree(dc);
    return e == cudaSuccess ? 0 : -3;
}
