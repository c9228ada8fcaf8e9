// EXPERIMENT (test infrastructure, not product code): north_star's parenthesis "lanes of a warp
// for VlWide >64-bit words", measured instead of argued (round-1 verdict, item 7).
//
// The per-cycle work of wide_bus.v (256-bit add / subtract / variable shifts / 64x64 multiply /
// half-swap mix / select / 70-bit part-select / unsigned and signed compare / $countones / a
// 128-bit rotating hash) hand-written twice for N instances x K cycles:
//
//   form A  one instance per thread, the eight 32-bit limbs of a value in eight registers
//           (what V3EmitCuda's models do, with the VL_*_W operators of verilated_cuda_funcs.h);
//   form B  one instance per 8 LANES, lane l holding limb l: carries by ballot + a carry
//           lookahead on 8-bit masks, shifts by two shuffles, compares by ballot + find-first.
//
// Both forms produce every instance's final acc / hash / compare-and-count signature; the program
// checks that they agree bit for bit and prints one JSON line with both timings.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o widelanes widelanes.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#define CHECK(x) \
    do { \
        cudaError_t e_ = (x); \
        if (e_ != cudaSuccess) { \
            fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); \
            exit(1); \
        } \
    } while (0)

struct Out {
    uint32_t acc[8];
    uint32_t hash[4];
    uint32_t sig;  // folds out_lt, out_lts, out_ones and the slice of every cycle
};

__host__ __device__ inline uint64_t mix64(uint64_t x) {
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
    return x ^ (x >> 31);
}
// stimulus word w of instance i in cycle c (in_data limbs 0..7, word 8 = {op, shift})
__device__ inline uint32_t stim(uint64_t inst, uint32_t cycle, uint32_t w) {
    return static_cast<uint32_t>(mix64(inst * 0x9e3779b97f4a7c15ULL + cycle * 0xd1342543de82ef95ULL + w) >> 32);
}

//============================================================================ form A
__device__ inline void shl256(uint32_t* o, const uint32_t* a, uint32_t sh) {
    const uint32_t w = sh >> 5, b = sh & 31;
#pragma unroll
    for (int l = 0; l < 8; ++l) {
        const int s = l - static_cast<int>(w);
        const uint32_t lo = s >= 0 ? a[s] : 0, lo1 = s - 1 >= 0 ? a[s - 1] : 0;
        o[l] = b ? ((lo << b) | (lo1 >> (32 - b))) : lo;
    }
}
__device__ inline void shr256(uint32_t* o, const uint32_t* a, uint32_t sh, uint32_t fill) {
    const uint32_t w = sh >> 5, b = sh & 31;
#pragma unroll
    for (int l = 0; l < 8; ++l) {
        const uint32_t s = l + w;
        const uint32_t lo = s < 8 ? a[s] : fill, hi = s + 1 < 8 ? a[s + 1] : fill;
        o[l] = b ? ((lo >> b) | (hi << (32 - b))) : lo;
    }
}
__global__ void __launch_bounds__(64) k_form_a(Out* out, uint32_t n, uint32_t cycles) {
    const uint32_t inst = blockIdx.x * blockDim.x + threadIdx.x;
    if (inst >= n) return;
    uint32_t acc[8] = {1, 0, 0, 0, 0, 0, 0, 0}, stage[8] = {0}, hash[4] = {0};
    uint32_t sh = 0, op = 0, sig = 0x811c9dc5U;
    for (uint32_t c = 0; c < cycles; ++c) {
        uint32_t sum[8], dif[8], shl[8], shr[8], sra[8], mixv[8], res[8];
        uint32_t carry = 0, borrow = 0;
#pragma unroll
        for (int l = 0; l < 8; ++l) {
            const uint64_t s = static_cast<uint64_t>(acc[l]) + stage[l] + carry;
            sum[l] = static_cast<uint32_t>(s);
            carry = static_cast<uint32_t>(s >> 32);
            const uint64_t d = static_cast<uint64_t>(acc[l]) - stage[l] - borrow;
            dif[l] = static_cast<uint32_t>(d);
            borrow = static_cast<uint32_t>(d >> 63);
        }
        shl256(shl, stage, sh);
        shr256(shr, stage, sh, 0);
        shr256(sra, stage, sh, (stage[7] >> 31) ? 0xffffffffU : 0U);
        const uint64_t x = (static_cast<uint64_t>(stage[1]) << 32) | stage[0];
        const uint64_t y = (static_cast<uint64_t>(acc[3]) << 32) | acc[2];
        const uint64_t plo = x * y, phi = __umul64hi(x, y);
        const uint32_t prod[4] = {static_cast<uint32_t>(plo), static_cast<uint32_t>(plo >> 32),
                                  static_cast<uint32_t>(phi), static_cast<uint32_t>(phi >> 32)};
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            mixv[l] = acc[l + 4] ^ ~prod[l];
            mixv[l + 4] = acc[l] ^ prod[l];
        }
        uint32_t par = 0, ones = 0;
#pragma unroll
        for (int l = 0; l < 8; ++l) {
            par ^= acc[l];
            ones += __popc(acc[l]);
        }
        par = __popc(par) & 1U;
#pragma unroll
        for (int l = 0; l < 8; ++l) {
            const uint32_t rot = (acc[l] << 1) | (l ? (acc[l - 1] >> 31) : par);
            const uint32_t sel = (acc[l] & stage[l]) | (~acc[l] & shl[l]);
            res[l] = op == 0 ? sum[l] : op == 1 ? dif[l] : op == 2 ? shl[l] : op == 3 ? shr[l] : op == 4 ? sra[l]
                     : op == 5 ? mixv[l] : op == 6 ? sel : rot;
        }
        uint32_t sl[8];
        shr256(sl, res, sh & 127U, 0);
        sl[2] &= 0x3fU;
        // unsigned / signed compare acc < stage
        bool lt = false;
#pragma unroll
        for (int l = 0; l < 8; ++l) {
            if (acc[l] != stage[l]) lt = acc[l] < stage[l];
        }
        const bool sa = acc[7] >> 31, sb = stage[7] >> 31;
        const bool lts = sa != sb ? sa : lt;
        sig = (sig ^ (ones | (lt << 16) | (lts << 17))) * 0x01000193U;
        sig = (sig ^ sl[0] ^ sl[1] ^ sl[2]) * 0x01000193U;
        // registers
        const uint32_t h3 = hash[3];
        const uint32_t nh0 = ((hash[0] << 1) | (h3 >> 31)) ^ prod[0] ^ sl[0];
        const uint32_t nh1 = ((hash[1] << 1) | (hash[0] >> 31)) ^ prod[1] ^ sl[1];
        const uint32_t nh2 = ((hash[2] << 1) | (hash[1] >> 31)) ^ prod[2] ^ sl[2];
        const uint32_t nh3 = ((hash[3] << 1) | (hash[2] >> 31)) ^ prod[3];
        hash[0] = nh0, hash[1] = nh1, hash[2] = nh2, hash[3] = nh3;
#pragma unroll
        for (int l = 0; l < 8; ++l) {
            acc[l] = res[l];
            stage[l] = stim(inst, c, l);
        }
        const uint32_t ctl = stim(inst, c, 8);
        sh = ctl & 0xffU;
        op = (ctl >> 8) & 7U;
    }
    Out o;
    for (int l = 0; l < 8; ++l) o.acc[l] = acc[l];
    for (int l = 0; l < 4; ++l) o.hash[l] = hash[l];
    o.sig = sig;
    out[inst] = o;
}

//============================================================================ form B
// 8 lanes per instance; G = this group's lanes within the warp
__device__ inline uint32_t grp_shfl(uint32_t v, int srcLimb, uint32_t fill) {
    const uint32_t got = __shfl_sync(0xffffffffU, v, srcLimb & 7, 8);
    return (srcLimb >= 0 && srcLimb < 8) ? got : fill;
}
// carries INTO each limb from per-limb generate / propagate flags (8-bit masks of the group)
__device__ inline uint32_t carries_in(uint32_t g, uint32_t p) {
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) c = ((g | (p & c)) << 1) & 0xffU;
    return c;
}
__global__ void __launch_bounds__(64) k_form_b(Out* out, uint32_t n, uint32_t cycles) {
    const uint32_t gid = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;  // instance
    const int l = threadIdx.x & 7;  // limb
    const uint32_t gshift = (threadIdx.x & 31) & ~7;  // first lane of the group within the warp
    const uint32_t inst = gid < n ? gid : n - 1;  // (padding groups shadow the last instance)
    uint32_t acc = l == 0 ? 1U : 0U, stage = 0, hash = 0, sh = 0, op = 0, sig = 0x811c9dc5U;
    for (uint32_t c = 0; c < cycles; ++c) {
        // add / subtract: per-limb sums, then the carries by lookahead over the group's masks
        const uint32_t s0 = acc + stage, d0 = acc - stage;
        const uint32_t g_add = (__ballot_sync(0xffffffffU, s0 < acc) >> gshift) & 0xffU;
        const uint32_t p_add = (__ballot_sync(0xffffffffU, s0 == 0xffffffffU) >> gshift) & 0xffU;
        const uint32_t g_sub = (__ballot_sync(0xffffffffU, acc < stage) >> gshift) & 0xffU;
        const uint32_t p_sub = (__ballot_sync(0xffffffffU, d0 == 0U) >> gshift) & 0xffU;
        const uint32_t sum = s0 + ((carries_in(g_add, p_add) >> l) & 1U);
        const uint32_t dif = d0 - ((carries_in(g_sub, p_sub) >> l) & 1U);
        // shifts: limb l takes from limbs l -/+ w and the neighbour
        const int w = sh >> 5;
        const uint32_t b = sh & 31;
        const uint32_t fillS = (grp_shfl(stage, 7, 0) >> 31) ? 0xffffffffU : 0U;
        const uint32_t l_lo = grp_shfl(stage, l - w, 0), l_lo1 = grp_shfl(stage, l - w - 1, 0);
        const uint32_t shl = b ? ((l_lo << b) | (l_lo1 >> (32 - b))) : l_lo;
        const uint32_t r_lo = __shfl_sync(0xffffffffU, stage, (l + w) & 7, 8), r_hi = __shfl_sync(0xffffffffU, stage, (l + w + 1) & 7, 8);
        const bool in0 = l + w < 8, in1 = l + w + 1 < 8;
        const uint32_t shr = b ? (((in0 ? r_lo : 0U) >> b) | ((in1 ? r_hi : 0U) << (32 - b))) : (in0 ? r_lo : 0U);
        const uint32_t sra = b ? (((in0 ? r_lo : fillS) >> b) | ((in1 ? r_hi : fillS) << (32 - b))) : (in0 ? r_lo : fillS);
        // 64 x 64 -> 128 product: every lane forms it from four shuffled limbs (cheaper than
        // distributing nine partial products), limb l of the result in lanes 0..3
        const uint64_t x = (static_cast<uint64_t>(grp_shfl(stage, 1, 0)) << 32) | grp_shfl(stage, 0, 0);
        const uint64_t y = (static_cast<uint64_t>(grp_shfl(acc, 3, 0)) << 32) | grp_shfl(acc, 2, 0);
        const uint64_t plo = x * y, phi = __umul64hi(x, y);
        const uint32_t prodl = (l & 3) == 0 ? static_cast<uint32_t>(plo) : (l & 3) == 1 ? static_cast<uint32_t>(plo >> 32)
                               : (l & 3) == 2 ? static_cast<uint32_t>(phi) : static_cast<uint32_t>(phi >> 32);
        const uint32_t swapped = __shfl_xor_sync(0xffffffffU, acc, 4, 8);
        const uint32_t mixv = swapped ^ (l < 4 ? ~prodl : prodl);
        // parity and population count of acc over the group
        uint32_t par = acc, ones = __popc(acc);
#pragma unroll
        for (int k = 1; k < 8; k <<= 1) {
            par ^= __shfl_xor_sync(0xffffffffU, par, k, 8);
            ones += __shfl_xor_sync(0xffffffffU, ones, k, 8);
        }
        par = __popc(par) & 1U;
        const uint32_t below = grp_shfl(acc, l - 1, 0);
        const uint32_t rot = (acc << 1) | (l ? (below >> 31) : par);
        const uint32_t sel = (acc & stage) | (~acc & shl);
        const uint32_t res = op == 0 ? sum : op == 1 ? dif : op == 2 ? shl : op == 3 ? shr : op == 4 ? sra
                             : op == 5 ? mixv : op == 6 ? sel : rot;
        // 70-bit part-select of res at sh[6:0]
        const int w2 = (sh & 127U) >> 5;
        const uint32_t b2 = sh & 31;
        const uint32_t q_lo = __shfl_sync(0xffffffffU, res, (l + w2) & 7, 8), q_hi = __shfl_sync(0xffffffffU, res, (l + w2 + 1) & 7, 8);
        const bool qi0 = l + w2 < 8, qi1 = l + w2 + 1 < 8;
        uint32_t sl = b2 ? (((qi0 ? q_lo : 0U) >> b2) | ((qi1 ? q_hi : 0U) << (32 - b2))) : (qi0 ? q_lo : 0U);
        if (l == 2) sl &= 0x3fU;
        if (l > 2) sl = 0;
        // compares: the most significant limb that differs decides
        const uint32_t ne = (__ballot_sync(0xffffffffU, acc != stage) >> gshift) & 0xffU;
        const uint32_t lo = (__ballot_sync(0xffffffffU, acc < stage) >> gshift) & 0xffU;
        const bool lt = ne ? ((lo >> (31 - __clz(ne))) & 1U) : false;
        const bool sa = grp_shfl(acc, 7, 0) >> 31, sb = fillS != 0;
        const bool lts = sa != sb ? sa : lt;
        const uint32_t slx = sl ^ __shfl_xor_sync(0xffffffffU, sl, 1, 8);
        const uint32_t slAll = slx ^ __shfl_xor_sync(0xffffffffU, slx, 2, 8);  // lanes 0..3: sl0^sl1^sl2
        sig = (sig ^ (ones | (lt << 16) | (lts << 17))) * 0x01000193U;
        sig = (sig ^ grp_shfl(slAll, 0, 0)) * 0x01000193U;
        // registers: hash in lanes 0..3
        const uint32_t hbelow = grp_shfl(hash, (l + 3) & 3, 0);
        if (l < 4) hash = ((hash << 1) | (hbelow >> 31)) ^ prodl ^ sl;
        acc = res;
        stage = stim(inst, c, l);
        const uint32_t ctl = stim(inst, c, 8);
        sh = ctl & 0xffU;
        op = (ctl >> 8) & 7U;
    }
    if (gid < n) {
        out[gid].acc[l] = acc;
        if (l < 4) out[gid].hash[l] = hash;
        if (l == 0) out[gid].sig = sig;
    }
}

int main(int argc, char** argv) {
    const uint32_t n = argc > 1 ? atoi(argv[1]) : 4096;
    const uint32_t cycles = argc > 2 ? atoi(argv[2]) : 8192;
    Out *dA, *dB;
    CHECK(cudaMalloc(&dA, sizeof(Out) * n));
    CHECK(cudaMalloc(&dB, sizeof(Out) * n));
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0));
    CHECK(cudaEventCreate(&e1));
    float msA = 1e30f, msB = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        float ms;
        CHECK(cudaEventRecord(e0));
        k_form_a<<<(n + 63) / 64, 64>>>(dA, n, cycles);
        CHECK(cudaEventRecord(e1));
        CHECK(cudaEventSynchronize(e1));
        CHECK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep && ms < msA) msA = ms;
        CHECK(cudaEventRecord(e0));
        k_form_b<<<(n * 8 + 63) / 64, 64>>>(dB, n, cycles);
        CHECK(cudaEventRecord(e1));
        CHECK(cudaEventSynchronize(e1));
        CHECK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep && ms < msB) msB = ms;
    }
    CHECK(cudaGetLastError());
    std::vector<Out> a(n), b(n);
    CHECK(cudaMemcpy(a.data(), dA, sizeof(Out) * n, cudaMemcpyDeviceToHost));
    CHECK(cudaMemcpy(b.data(), dB, sizeof(Out) * n, cudaMemcpyDeviceToHost));
    uint32_t bad = 0, distinct = 0;
    for (uint32_t i = 0; i < n; ++i) {
        bool same = a[i].sig == b[i].sig;
        for (int l = 0; l < 8; ++l) same &= a[i].acc[l] == b[i].acc[l];
        for (int l = 0; l < 4; ++l) same &= a[i].hash[l] == b[i].hash[l];
        bad += !same;
        distinct += i && a[i].sig != a[i - 1].sig;
    }
    printf("{\"instances\": %u, \"cycles\": %u, \"mismatches\": %u, \"distinct_neighbours\": %u, "
           "\"thread_per_instance_ms\": %.4f, \"eight_lanes_per_instance_ms\": %.4f, "
           "\"thread_per_instance_Gcps\": %.4f, \"eight_lanes_per_instance_Gcps\": %.4f}\n",
           n, cycles, bad, distinct, msA, msB, n * (double)cycles / msA / 1e6, n * (double)cycles / msB / 1e6);
    return bad ? 1 : 0;
}
