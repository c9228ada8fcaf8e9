// -*- mode: C++; c-file-style: "cc-mode" -*-
// Stimulus of the wide_bus workload.  ONE source compiled twice (device testbench wide_tb.h and
// oracle harness oracle/models/wide_bus/harness.cpp) so both sides drive bit-identical inputs.
// Stateless: the inputs applied before eval `e` of an instance are a pure function of
// (seed, e), drawn from the reference's PRNG (VlRNG = splitmix64 -> xoroshiro128**).
#ifndef WIDE_STIM_H_
#define WIDE_STIM_H_
#ifndef WIDE_STIM_FN
#define WIDE_STIM_FN inline
#endif

template <class T_Top>
WIDE_STIM_FN void wide_drive_inputs(T_Top& top, uint64_t seed, uint64_t e) {
    VlRNG rng{seed + 0x9e3779b97f4a7c15ULL * (e + 1)};
    for (int w = 0; w < 8; w += 2) {
        const uint64_t r = rng.rand64();
        top.in_data[w] = static_cast<uint32_t>(r);
        top.in_data[w + 1] = static_cast<uint32_t>(r >> 32);
    }
    const uint64_t c = rng.rand64();
    // mostly short shifts, sometimes the whole 8-bit range (shifts >= 256 cannot occur)
    top.in_shift = ((c >> 16) & 3U) ? static_cast<uint32_t>(c & 0x3fU) : static_cast<uint32_t>(c & 0xffU);
    top.in_op = static_cast<uint32_t>((c >> 8) & 7U);
    top.in_valid = ((c >> 11) & 7U) != 0U;
}

#endif  // guard
