// -*- mode: C++; c-file-style: "cc-mode" -*-
//*************************************************************************
//
// verilated_cuda_funcs.h -- the VL_* operator library as device functions.
//
// One CUDA thread evaluates one model instance, so every operator here works
// on thread-private registers / local arrays.  Names, argument order and
// bit-level results (including which results are left "dirty" above their
// width) follow the host runtime so that V3EmitCuda can print the same calls
// V3EmitCFunc prints (src/V3EmitCFunc.cpp:39-195):
//
//   reference                                    here
//   include/verilated_funcs.h:436-505            clean / assign / assign-bit
//   include/verilated_funcs.h:621-693            extend / sign-extend
//   include/verilated_funcs.h:696-923            reductions, countones, onehot, clog2
//   include/verilated_funcs.h:926-1113           wide logic and compares
//   include/verilated_funcs.h:1117-1386          negate/add/sub/mul/div/mod/pow
//   include/verilated.cpp:663-870                wide divide (Knuth D), wide pow
//   include/verilated_funcs.h:1391-1585          insert helpers, replicate
//   include/verilated_funcs.h:1589-1658,1953,2071-2100  stream-left
//   include/verilated_funcs.h:2392-2471          concat
//   include/verilated_funcs.h:2474-2733          shifts
//   include/verilated_funcs.h:2736-2856          bit / part select
//   include/verilated_funcs.h:3346-3461          assign-select / select-assign
//   include/verilated_funcs.h:3464-3598          cond, wide constants
//
// Word loops are written so that, once the (always constant) width arguments
// are propagated by inlining, nvcc fully unrolls them and keeps the limbs in
// registers.
//
//*************************************************************************
#ifndef VERILATOR_VERILATED_CUDA_FUNCS_H_
#define VERILATOR_VERILATED_CUDA_FUNCS_H_

#include "verilated_cuda_types.h"

//=========================================================================
// Bit-count primitives: PTX popc/clz on the device, builtins in host test builds

VLCU_HD int vlcu_popc32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
VLCU_HD int vlcu_popc64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}
// Bit position of the most significant one, plus one; 0 for x == 0
VLCU_HD int vlcu_msbp1_32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return 32 - __clz(static_cast<int>(x));
#else
    return x ? 32 - __builtin_clz(x) : 0;
#endif
}
VLCU_HD int vlcu_msbp1_64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return 64 - __clzll(static_cast<long long>(x));
#else
    return x ? 64 - __builtin_clzll(x) : 0;
#endif
}

//=========================================================================
// Quad <-> wide packing (verilated_funcs.h:207-220)

VLCU_HD WDataOutP VL_SET_WQ(WDataOutP owp, QData data) {
    owp[0] = static_cast<EData>(data);
    owp[1] = static_cast<EData>(data >> 32);
    return owp;
}
VLCU_HD WDataOutP VL_SET_WI(WDataOutP owp, IData data) {
    owp[0] = data;
    owp[1] = 0;
    return owp;
}
VLCU_HD QData VL_SET_QW(WDataInP lwp) {
    return (static_cast<QData>(lwp[1]) << 32) | static_cast<QData>(lwp[0]);
}
#define VL_SET_QII(ld, rd) ((static_cast<QData>(ld) << 32ULL) | static_cast<QData>(rd))

//=========================================================================
// Sign helpers (verilated_funcs.h:328-344)

#define VL_SIGN_I(nbits, lhs) ((lhs) >> VL_BITBIT_I((nbits) - VL_UL(1)))
#define VL_SIGN_Q(nbits, lhs) ((lhs) >> VL_BITBIT_Q((nbits) - 1ULL))
#define VL_SIGN_E(nbits, lhs) ((lhs) >> VL_BITBIT_E((nbits) - VL_EUL(1)))
#define VL_SIGN_W(nbits, rwp) \
    ((rwp)[VL_BITWORD_E((nbits) - VL_EUL(1))] >> VL_BITBIT_E((nbits) - VL_EUL(1)))
#define VL_SIGNONES_E(nbits, lhs) (-(VL_SIGN_E(nbits, lhs)))

VLCU_HD IData VL_EXTENDSIGN_I(int lbits, IData lhs) { return 0U - (lhs & (1U << (lbits - 1))); }
VLCU_HD QData VL_EXTENDSIGN_Q(int lbits, QData lhs) {
    return 0ULL - (lhs & (1ULL << (lbits - 1)));
}

//=========================================================================
// Word-array fill/copy

VLCU_HD WDataOutP VL_MEMSET_ZERO_W(WDataOutP owp, int words) {
    for (int i = 0; i < words; ++i) owp[i] = 0;
    return owp;
}
VLCU_HD WDataOutP VL_MEMSET_ONES_W(WDataOutP owp, int words) {
    for (int i = 0; i < words; ++i) owp[i] = ~0U;
    return owp;
}
VLCU_HD WDataOutP VL_MEMCPY_W(WDataOutP owp, WDataInP iwp, int words) {
    for (int i = 0; i < words; ++i) owp[i] = iwp[i];
    return owp;
}

//=========================================================================
// Clean / assign (verilated_funcs.h:436-469)

#define VL_CLEAN_II(obits, lbits, lhs) ((lhs) & (VL_MASK_I(obits)))
#define VL_CLEAN_QQ(obits, lbits, lhs) ((lhs) & (VL_MASK_Q(obits)))
#define VL_ASSIGNCLEAN_W(obits, owp, lwp) VL_CLEAN_WW((obits), (owp), (lwp))

VLCU_HD WDataOutP _vl_clean_inplace_w(int obits, WDataOutP owp) {
    owp[VL_WORDS_I(obits) - 1] &= VL_MASK_E(obits);
    return owp;
}
VLCU_HD WDataOutP VL_CLEAN_WW(int obits, WDataOutP owp, WDataInP lwp) {
    const int top = VL_WORDS_I(obits) - 1;
    for (int i = 0; i < top; ++i) owp[i] = lwp[i];
    owp[top] = lwp[top] & VL_MASK_E(obits);
    return owp;
}
VLCU_HD WDataOutP VL_ZERO_W(int obits, WDataOutP owp) {
    return VL_MEMSET_ZERO_W(owp, VL_WORDS_I(obits));
}
VLCU_HD WDataOutP VL_ALLONES_W(int obits, WDataOutP owp) {
    const int top = VL_WORDS_I(obits) - 1;
    for (int i = 0; i < top; ++i) owp[i] = ~0U;
    owp[top] = VL_MASK_E(obits);
    return owp;
}
VLCU_HD WDataOutP VL_ASSIGN_W(int obits, WDataOutP owp, WDataInP lwp) {
    return VL_MEMCPY_W(owp, lwp, VL_WORDS_I(obits));
}

// Single-bit assignment; rhs must be clean (verilated_funcs.h:471-505)
VLCU_HD void VL_ASSIGNBIT_II(int bit, VlcuMem8& lhsr, IData rhs) {
    const int b = VL_BITBIT_I(bit);
    lhsr = static_cast<VlcuMem8>((lhsr & ~(1U << b)) | (rhs << b));
}
VLCU_HD void VL_ASSIGNBIT_II(int bit, VlcuMem16& lhsr, IData rhs) {
    const int b = VL_BITBIT_I(bit);
    lhsr = static_cast<VlcuMem16>((lhsr & ~(1U << b)) | (rhs << b));
}
VLCU_HD void VL_ASSIGNBIT_II(int bit, IData& lhsr, IData rhs) {
    const int b = VL_BITBIT_I(bit);
    lhsr = (lhsr & ~(1U << b)) | (rhs << b);
}
VLCU_HD void VL_ASSIGNBIT_QI(int bit, QData& lhsr, QData rhs) {
    const int b = VL_BITBIT_Q(bit);
    lhsr = (lhsr & ~(1ULL << b)) | (rhs << b);
}
VLCU_HD void VL_ASSIGNBIT_WI(int bit, WDataOutP owp, IData rhs) {
    const int b = VL_BITBIT_E(bit);
    EData& w = owp[VL_BITWORD_E(bit)];
    w = (w & ~(1U << b)) | (rhs << b);
}
VLCU_HD void VL_ASSIGNBIT_IO(int bit, VlcuMem8& lhsr) {
    lhsr = static_cast<VlcuMem8>(lhsr | (1U << VL_BITBIT_I(bit)));
}
VLCU_HD void VL_ASSIGNBIT_IO(int bit, VlcuMem16& lhsr) {
    lhsr = static_cast<VlcuMem16>(lhsr | (1U << VL_BITBIT_I(bit)));
}
VLCU_HD void VL_ASSIGNBIT_IO(int bit, IData& lhsr) { lhsr |= (1U << VL_BITBIT_I(bit)); }
VLCU_HD void VL_ASSIGNBIT_QO(int bit, QData& lhsr) { lhsr |= (1ULL << VL_BITBIT_Q(bit)); }
VLCU_HD void VL_ASSIGNBIT_WO(int bit, WDataOutP owp) {
    owp[VL_BITWORD_E(bit)] |= (1U << VL_BITBIT_E(bit));
}

//=========================================================================
// Width extension (verilated_funcs.h:621-693).  Zero extension needs clean
// input; sign extension returns a dirty value for the <=64 bit forms.

#define VL_EXTEND_II(obits, lbits, lhs) ((lhs))
#define VL_EXTEND_QI(obits, lbits, lhs) (static_cast<QData>(lhs))
#define VL_EXTEND_QQ(obits, lbits, lhs) ((lhs))

VLCU_HD WDataOutP VL_EXTEND_WI(int obits, int, WDataOutP owp, IData ld) {
    owp[0] = ld;
    for (int i = 1; i < VL_WORDS_I(obits); ++i) owp[i] = 0;
    return owp;
}
VLCU_HD WDataOutP VL_EXTEND_WQ(int obits, int, WDataOutP owp, QData ld) {
    VL_SET_WQ(owp, ld);
    for (int i = 2; i < VL_WORDS_I(obits); ++i) owp[i] = 0;
    return owp;
}
VLCU_HD WDataOutP VL_EXTEND_WW(int obits, int lbits, WDataOutP owp, WDataInP lwp) {
    const int lwords = VL_WORDS_I(lbits);
    for (int i = lwords; i < VL_WORDS_I(obits); ++i) owp[i] = 0;
    for (int i = 0; i < lwords; ++i) owp[i] = lwp[i];
    return owp;
}

VLCU_HD IData VL_EXTENDS_II(int, int lbits, IData lhs) { return VL_EXTENDSIGN_I(lbits, lhs) | lhs; }
VLCU_HD QData VL_EXTENDS_QI(int, int lbits, QData lhs) { return VL_EXTENDSIGN_Q(lbits, lhs) | lhs; }
VLCU_HD QData VL_EXTENDS_QQ(int, int lbits, QData lhs) { return VL_EXTENDSIGN_Q(lbits, lhs) | lhs; }

VLCU_HD WDataOutP VL_EXTENDS_WI(int obits, int lbits, WDataOutP owp, IData ld) {
    const bool neg = VL_SIGN_E(lbits, ld) != 0;  // (bit above the sign is never set: clean input)
    const EData fill = neg ? ~0U : 0U;
    owp[0] = neg ? (ld | ~VL_MASK_E(lbits)) : ld;
    for (int i = 1; i < VL_WORDS_I(obits); ++i) owp[i] = fill;
    return owp;
}
VLCU_HD WDataOutP VL_EXTENDS_WQ(int obits, int lbits, WDataOutP owp, QData ld) {
    VL_SET_WQ(owp, ld);
    const bool neg = VL_SIGN_E(lbits, owp[1]) != 0;
    const EData fill = neg ? ~0U : 0U;
    if (neg) owp[1] |= ~VL_MASK_E(lbits);
    for (int i = 2; i < VL_WORDS_I(obits); ++i) owp[i] = fill;
    return owp;
}
VLCU_HD WDataOutP VL_EXTENDS_WW(int obits, int lbits, WDataOutP owp, WDataInP lwp) {
    const int lwords = VL_WORDS_I(lbits);
    const EData msw = lwp[lwords - 1];
    const bool neg = VL_SIGN_E(lbits, msw) != 0;
    const EData fill = neg ? ~0U : 0U;
    for (int i = 0; i < lwords - 1; ++i) owp[i] = lwp[i];
    owp[lwords - 1] = neg ? (msw | ~VL_MASK_E(lbits)) : msw;
    for (int i = lwords; i < VL_WORDS_I(obits); ++i) owp[i] = fill;
    return owp;
}

//=========================================================================
// Reductions (verilated_funcs.h:696-785)

#define VL_REDAND_II(lbits, lhs) ((lhs) == VL_MASK_I(lbits))
#define VL_REDAND_IQ(lbits, lhs) ((lhs) == VL_MASK_Q(lbits))
VLCU_HD IData VL_REDAND_IW(int lbits, WDataInP lwp) {
    const int words = VL_WORDS_I(lbits);
    EData acc = lwp[0];
    for (int i = 1; i < words - 1; ++i) acc &= lwp[i];
    acc &= ~VL_MASK_E(lbits) | lwp[words - 1];
    return acc == ~0U;
}

#define VL_REDOR_I(lhs) ((lhs) != 0)
#define VL_REDOR_Q(lhs) ((lhs) != 0)
VLCU_HD IData VL_REDOR_W(int words, WDataInP lwp) {
    EData acc = 0;
    for (int i = 0; i < words; ++i) acc |= lwp[i];
    return acc != 0;
}

// The host runtime (built with GCC) returns a clean parity for all but the
// 2-bit form, whose result is only valid in bit 0.
VLCU_HD IData VL_REDXOR_2(IData r) { return r ^ (r >> 1); }
VLCU_HD IData VL_REDXOR_4(IData r) { return vlcu_popc32(r) & 1; }
VLCU_HD IData VL_REDXOR_8(IData r) { return vlcu_popc32(r) & 1; }
VLCU_HD IData VL_REDXOR_16(IData r) { return vlcu_popc32(r) & 1; }
VLCU_HD IData VL_REDXOR_32(IData r) { return vlcu_popc32(r) & 1; }
VLCU_HD IData VL_REDXOR_64(QData r) { return vlcu_popc64(r) & 1; }
VLCU_HD IData VL_REDXOR_W(int words, WDataInP lwp) {
    EData acc = lwp[0];
    for (int i = 1; i < words; ++i) acc ^= lwp[i];
    return VL_REDXOR_32(acc);
}

//=========================================================================
// Population counts and friends (verilated_funcs.h:786-923)

VLCU_HD IData VL_COUNTONES_I(IData lhs) { return vlcu_popc32(lhs); }
VLCU_HD IData VL_COUNTONES_Q(QData lhs) { return vlcu_popc64(lhs); }
#define VL_COUNTONES_E VL_COUNTONES_I
VLCU_HD IData VL_COUNTONES_W(int words, WDataInP lwp) {
    EData n = 0;
    for (int i = 0; i < words; ++i) n += vlcu_popc32(lwp[i]);
    return n;
}

// $countbits with three control values, each 0 or 1 (X/Z never match in 2-state)
VLCU_HD IData VL_COUNTBITS_I(int lbits, IData lhs, IData ctrl0, IData ctrl1, IData ctrl2) {
    const int nOnes = (ctrl0 & 1) + (ctrl1 & 1) + (ctrl2 & 1);
    if (nOnes == 3) return vlcu_popc32(lhs);
    if (nOnes == 0) {
        const IData mask = (lbits == 32) ? ~0U : ((1U << lbits) - 1U);
        return vlcu_popc32(~lhs & mask);
    }
    return (lbits == 32) ? 32 : lbits;
}
VLCU_HD IData VL_COUNTBITS_Q(int lbits, QData lhs, IData ctrl0, IData ctrl1, IData ctrl2) {
    return VL_COUNTBITS_I(32, static_cast<IData>(lhs), ctrl0, ctrl1, ctrl2)
           + VL_COUNTBITS_I(lbits - 32, static_cast<IData>(lhs >> 32), ctrl0, ctrl1, ctrl2);
}
#define VL_COUNTBITS_E VL_COUNTBITS_I
VLCU_HD IData VL_COUNTBITS_W(int lbits, int words, WDataInP lwp, IData ctrl0, IData ctrl1,
                             IData ctrl2) {
    EData n = 0;
    for (int i = 0; i < words; ++i) {
        const int wbits = (i == words - 1) ? VL_BITBIT_I(lbits) : 32;
        n += VL_COUNTBITS_I(wbits, lwp[i], ctrl0, ctrl1, ctrl2);
    }
    return n;
}

VLCU_HD IData VL_ONEHOT_I(IData lhs) { return vlcu_popc32(lhs) == 1; }
VLCU_HD IData VL_ONEHOT_Q(QData lhs) { return vlcu_popc64(lhs) == 1; }
VLCU_HD IData VL_ONEHOT_W(int words, WDataInP lwp) { return VL_COUNTONES_W(words, lwp) == 1; }
VLCU_HD IData VL_ONEHOT0_I(IData lhs) { return vlcu_popc32(lhs) <= 1; }
VLCU_HD IData VL_ONEHOT0_Q(QData lhs) { return vlcu_popc64(lhs) <= 1; }
VLCU_HD IData VL_ONEHOT0_W(int words, WDataInP lwp) { return VL_COUNTONES_W(words, lwp) <= 1; }

VLCU_HD IData VL_MOSTSETBITP1_I(IData lhs) { return vlcu_msbp1_32(lhs); }
VLCU_HD IData VL_MOSTSETBITP1_Q(QData lhs) { return vlcu_msbp1_64(lhs); }
VLCU_HD IData VL_MOSTSETBITP1_W(int words, WDataInP lwp) {
    for (int i = words - 1; i >= 0; --i) {
        if (lwp[i]) return i * 32 + vlcu_msbp1_32(lwp[i]);
    }
    return 0;
}

// ceil(log2(x)); 0 for x == 0 and x == 1
VLCU_HD IData VL_CLOG2_I(IData lhs) { return lhs ? vlcu_msbp1_32(lhs - 1) : 0; }
VLCU_HD IData VL_CLOG2_Q(QData lhs) { return lhs ? vlcu_msbp1_64(lhs - 1) : 0; }
VLCU_HD IData VL_CLOG2_W(int words, WDataInP lwp) {
    const IData top = VL_MOSTSETBITP1_W(words, lwp);
    if (!top) return 0;
    // Exact power of two -> index of that bit, otherwise one more
    return (top - 1) + ((VL_COUNTONES_W(words, lwp) == 1) ? 0 : 1);
}

//=========================================================================
// Wide bitwise logic (verilated_funcs.h:926-957)

VLCU_HD WDataOutP VL_AND_W(int words, WDataOutP owp, WDataInP lwp, WDataInP rwp) {
    for (int i = 0; i < words; ++i) owp[i] = lwp[i] & rwp[i];
    return owp;
}
VLCU_HD WDataOutP VL_OR_W(int words, WDataOutP owp, WDataInP lwp, WDataInP rwp) {
    for (int i = 0; i < words; ++i) owp[i] = lwp[i] | rwp[i];
    return owp;
}
VLCU_HD WDataOutP VL_XOR_W(int words, WDataOutP owp, WDataInP lwp, WDataInP rwp) {
    for (int i = 0; i < words; ++i) owp[i] = lwp[i] ^ rwp[i];
    return owp;
}
VLCU_HD WDataOutP VL_NOT_W(int words, WDataOutP owp, WDataInP lwp) {
    for (int i = 0; i < words; ++i) owp[i] = ~lwp[i];
    return owp;
}
VLCU_HD IData VL_CHANGEXOR_W(int words, WDataInP lwp, WDataInP rwp) {
    IData od = 0;
    for (int i = 0; i < words; ++i) od |= lwp[i] ^ rwp[i];
    return od;
}

//=========================================================================
// Comparisons (verilated_funcs.h:959-1114); inputs clean, result 0/1

VLCU_HD IData VL_EQ_W(int words, WDataInP lwp, WDataInP rwp) {
    return VL_CHANGEXOR_W(words, lwp, rwp) == 0;
}
#define VL_NEQ_W(words, lwp, rwp) (!VL_EQ_W(words, lwp, rwp))

// -1 / 0 / +1 ordering of two unsigned wides, most significant limb first
VLCU_HD int _vl_cmp_w(int words, WDataInP lwp, WDataInP rwp) {
    int result = 0;  // Decided by the highest differing limb; scan upward so it wins
    for (int i = 0; i < words; ++i) {
        if (lwp[i] != rwp[i]) result = (lwp[i] > rwp[i]) ? 1 : -1;
    }
    return result;
}
#define VL_LT_W(words, lwp, rwp) (_vl_cmp_w(words, lwp, rwp) < 0)
#define VL_LTE_W(words, lwp, rwp) (_vl_cmp_w(words, lwp, rwp) <= 0)
#define VL_GT_W(words, lwp, rwp) (_vl_cmp_w(words, lwp, rwp) > 0)
#define VL_GTE_W(words, lwp, rwp) (_vl_cmp_w(words, lwp, rwp) >= 0)

VLCU_HD int _vl_cmps_w(int lbits, WDataInP lwp, WDataInP rwp) {
    const int words = VL_WORDS_I(lbits);
    const EData lneg = VL_SIGN_E(lbits, lwp[words - 1]);
    const EData rneg = VL_SIGN_E(lbits, rwp[words - 1]);
    if (!lneg && rneg) return 1;
    if (lneg && !rneg) return -1;
    return _vl_cmp_w(words, lwp, rwp);  // Same sign: two's complement orders like unsigned
}
#define VL_LTS_IWW(lbits, lwp, rwp) (_vl_cmps_w(lbits, lwp, rwp) < 0)
#define VL_LTES_IWW(lbits, lwp, rwp) (_vl_cmps_w(lbits, lwp, rwp) <= 0)
#define VL_GTS_IWW(lbits, lwp, rwp) (_vl_cmps_w(lbits, lwp, rwp) > 0)
#define VL_GTES_IWW(lbits, lwp, rwp) (_vl_cmps_w(lbits, lwp, rwp) >= 0)

// Signed compares (verilated_funcs.h VL_GTS_III etc. sign-extend both sides to 64 bits): values
// of up to 32 bits are moved to the top of a 32-bit register instead (a shift by a constant, none
// at all for 32-bit operands) and compared as int32 -- the same order, half the instructions.
VLCU_HD int32_t vlcu_stop32(int lbits, IData v) { return static_cast<int32_t>(v << (32 - lbits)); }
VLCU_HD int64_t vlcu_sext64(int lbits, QData v) {
    return static_cast<int64_t>(v << (64 - lbits)) >> (64 - lbits);
}
VLCU_HD IData VL_GTS_III(int lbits, IData lhs, IData rhs) {
    return vlcu_stop32(lbits, lhs) > vlcu_stop32(lbits, rhs);
}
VLCU_HD IData VL_GTS_IQQ(int lbits, QData lhs, QData rhs) {
    return vlcu_sext64(lbits, lhs) > vlcu_sext64(lbits, rhs);
}
VLCU_HD IData VL_GTES_III(int lbits, IData lhs, IData rhs) {
    return vlcu_stop32(lbits, lhs) >= vlcu_stop32(lbits, rhs);
}
VLCU_HD IData VL_GTES_IQQ(int lbits, QData lhs, QData rhs) {
    return vlcu_sext64(lbits, lhs) >= vlcu_sext64(lbits, rhs);
}
VLCU_HD IData VL_LTS_III(int lbits, IData lhs, IData rhs) {
    return vlcu_stop32(lbits, lhs) < vlcu_stop32(lbits, rhs);
}
VLCU_HD IData VL_LTS_IQQ(int lbits, QData lhs, QData rhs) {
    return vlcu_sext64(lbits, lhs) < vlcu_sext64(lbits, rhs);
}
VLCU_HD IData VL_LTES_III(int lbits, IData lhs, IData rhs) {
    return vlcu_stop32(lbits, lhs) <= vlcu_stop32(lbits, rhs);
}
VLCU_HD IData VL_LTES_IQQ(int lbits, QData lhs, QData rhs) {
    return vlcu_sext64(lbits, lhs) <= vlcu_sext64(lbits, rhs);
}

//=========================================================================
// Arithmetic (verilated_funcs.h:1117-1303).  Wide results leave the top limb
// dirty; division by zero and signed MIN / -1 yield 0.

VLCU_HD WDataOutP VL_NEGATE_W(int words, WDataOutP owp, WDataInP lwp) {
    EData carry = 1;
    for (int i = 0; i < words; ++i) {
        const EData inv = ~lwp[i];
        const EData sum = inv + carry;
        carry = (sum < inv);
        owp[i] = sum;
    }
    return owp;
}
VLCU_HD void VL_NEGATE_INPLACE_W(int words, WDataOutP owp_lwp) {
    VL_NEGATE_W(words, owp_lwp, owp_lwp);
}

VLCU_HD IData VL_DIV_III(int, IData lhs, IData rhs) { return rhs ? lhs / rhs : 0; }
VLCU_HD QData VL_DIV_QQQ(int, QData lhs, QData rhs) { return rhs ? lhs / rhs : 0; }
VLCU_HD IData VL_MODDIV_III(int, IData lhs, IData rhs) { return rhs ? lhs % rhs : 0; }
VLCU_HD QData VL_MODDIV_QQQ(int, QData lhs, QData rhs) { return rhs ? lhs % rhs : 0; }

VLCU_HD WDataOutP VL_ADD_W(int words, WDataOutP owp, WDataInP lwp, WDataInP rwp) {
    QData acc = 0;
    for (int i = 0; i < words; ++i) {
        acc += static_cast<QData>(lwp[i]) + static_cast<QData>(rwp[i]);
        owp[i] = static_cast<EData>(acc);
        acc >>= 32;
    }
    return owp;
}
VLCU_HD WDataOutP VL_SUB_W(int words, WDataOutP owp, WDataInP lwp, WDataInP rwp) {
    QData acc = 1;  // l - r == l + ~r + 1
    for (int i = 0; i < words; ++i) {
        acc += static_cast<QData>(lwp[i]) + static_cast<QData>(static_cast<EData>(~rwp[i]));
        owp[i] = static_cast<EData>(acc);
        acc >>= 32;
    }
    return owp;
}
// Truncating schoolbook multiply.  owp must not alias the inputs.
VLCU_HD WDataOutP VL_MUL_W(int words, WDataOutP owp, WDataInP lwp, WDataInP rwp) {
    for (int i = 0; i < words; ++i) owp[i] = 0;
    for (int l = 0; l < words; ++l) {
        QData carry = 0;
        const QData lw = lwp[l];
        for (int r = 0; l + r < words; ++r) {
            const QData t = lw * static_cast<QData>(rwp[r]) + owp[l + r] + carry;
            owp[l + r] = static_cast<EData>(t);
            carry = t >> 32;
        }
    }
    return owp;
}

VLCU_HD IData VL_MULS_III(int lbits, IData lhs, IData rhs) {
    return VL_EXTENDS_II(32, lbits, lhs) * VL_EXTENDS_II(32, lbits, rhs);
}
VLCU_HD QData VL_MULS_QQQ(int lbits, QData lhs, QData rhs) {
    return VL_EXTENDS_QQ(64, lbits, lhs) * VL_EXTENDS_QQ(64, lbits, rhs);
}

// Magnitude of a signed wide in 'storep' when negative, else the input itself
VLCU_HD WDataInP vlcu_abs_w(int lbits, int words, WDataOutP storep, WDataInP lwp, bool neg) {
    if (!neg) return lwp;
    VL_NEGATE_W(words, storep, lwp);
    storep[words - 1] &= VL_MASK_E(lbits);
    return storep;
}

template <int T_MaxWords = VL_MULS_MAX_WORDS>
VLCU_HD WDataOutP VL_MULS_WWW(int lbits, WDataOutP owp, WDataInP lwp, WDataInP rwp) {
    const int words = VL_WORDS_I(lbits);
    EData lstore[T_MaxWords];
    EData rstore[T_MaxWords];
    const bool lneg = VL_SIGN_E(lbits, lwp[words - 1]) & 1;
    const bool rneg = VL_SIGN_E(lbits, rwp[words - 1]) & 1;
    lwp = vlcu_abs_w(lbits, words, lstore, lwp, lneg);
    rwp = vlcu_abs_w(lbits, words, rstore, rwp, rneg);
    VL_MUL_W(words, owp, lwp, rwp);
    owp[words - 1] &= VL_MASK_E(lbits);
    if (lneg != rneg) VL_NEGATE_W(words, owp, owp);
    return owp;  // Top limb dirty
}

VLCU_HD IData VL_DIVS_III(int lbits, IData lhs, IData rhs) {
    if (VLCU_UNLIKELY(rhs == 0)) return 0;
    if (VLCU_UNLIKELY(lhs == 0x80000000U && rhs == 0xffffffffU)) return 0;
    const int32_t l = static_cast<int32_t>(VL_EXTENDS_II(32, lbits, lhs));
    const int32_t r = static_cast<int32_t>(VL_EXTENDS_II(32, lbits, rhs));
    return static_cast<IData>(l / r);
}
VLCU_HD QData VL_DIVS_QQQ(int lbits, QData lhs, QData rhs) {
    if (VLCU_UNLIKELY(rhs == 0)) return 0;
    if (VLCU_UNLIKELY(lhs == 0x8000000000000000ULL && rhs == 0xffffffffffffffffULL)) return 0;
    const int64_t l = vlcu_sext64(lbits, lhs);
    const int64_t r = vlcu_sext64(lbits, rhs);
    return static_cast<QData>(l / r);
}
VLCU_HD IData VL_MODDIVS_III(int lbits, IData lhs, IData rhs) {
    if (VLCU_UNLIKELY(rhs == 0)) return 0;
    if (VLCU_UNLIKELY(lhs == 0x80000000U && rhs == 0xffffffffU)) return 0;
    const int32_t l = static_cast<int32_t>(VL_EXTENDS_II(32, lbits, lhs));
    const int32_t r = static_cast<int32_t>(VL_EXTENDS_II(32, lbits, rhs));
    return static_cast<IData>(l % r);
}
VLCU_HD QData VL_MODDIVS_QQQ(int lbits, QData lhs, QData rhs) {
    if (VLCU_UNLIKELY(rhs == 0)) return 0;
    if (VLCU_UNLIKELY(lhs == 0x8000000000000000ULL && rhs == 0xffffffffffffffffULL)) return 0;
    const int64_t l = vlcu_sext64(lbits, lhs);
    const int64_t r = vlcu_sext64(lbits, rhs);
    return static_cast<QData>(l % r);
}

// Unsigned wide divide / modulus, Knuth's algorithm D on 32-bit digits
// (include/verilated.cpp:663-776).  T_MaxWords bounds the scratch arrays; the
// emitter instantiates it with the operand word count so they stay small.
template <int T_MaxWords = VL_MULS_MAX_WORDS>
VLCU_HD WDataOutP _vl_moddiv_w(int lbits, WDataOutP owp, WDataInP lwp, WDataInP rwp,
                               bool is_modulus) {
    const int words = VL_WORDS_I(lbits);
    for (int i = 0; i < words; ++i) owp[i] = 0;
    const int umsbp1 = VL_MOSTSETBITP1_W(words, lwp);
    const int vmsbp1 = VL_MOSTSETBITP1_W(words, rwp);
    if (VLCU_UNLIKELY(vmsbp1 == 0) || VLCU_UNLIKELY(umsbp1 == 0)) return owp;  // x/0, 0/x -> 0
    const int uw = VL_WORDS_I(umsbp1);  // Digits in dividend ("m + n")
    const int vw = VL_WORDS_I(vmsbp1);  // Digits in divisor ("n")

    if (vw == 1) {  // Short division
        const uint64_t d = rwp[0];
        uint64_t rem = 0;
        for (int j = uw - 1; j >= 0; --j) {
            const uint64_t cur = (rem << 32) | lwp[j];
            const uint64_t q = cur / d;
            owp[j] = static_cast<EData>(q);
            rem = cur - static_cast<uint64_t>(static_cast<EData>(q)) * d;
        }
        if (is_modulus) {
            owp[0] = static_cast<EData>(rem);
            for (int i = 1; i < words; ++i) owp[i] = 0;
        }
        return owp;
    }

    uint32_t un[T_MaxWords + 1];  // Normalised dividend, one extra digit
    uint32_t vn[T_MaxWords + 1];  // Normalised divisor
    for (int i = 0; i < words + 1; ++i) un[i] = vn[i] = 0;

    // Normalise so the divisor's top digit has its MSB set
    const int s = 31 - VL_BITBIT_I(vmsbp1 - 1);
    if (s) {
        for (int i = vw - 1; i > 0; --i) vn[i] = (rwp[i] << s) | (rwp[i - 1] >> (32 - s));
        vn[0] = rwp[0] << s;
        un[uw] = lwp[uw - 1] >> (32 - s);
        for (int i = uw - 1; i > 0; --i) un[i] = (lwp[i] << s) | (lwp[i - 1] >> (32 - s));
        un[0] = lwp[0] << s;
    } else {
        for (int i = 0; i < vw; ++i) vn[i] = rwp[i];
        for (int i = 0; i < uw; ++i) un[i] = lwp[i];
        un[uw] = 0;
    }

    const uint64_t vtop = vn[vw - 1];
    const uint64_t vnext = vn[vw - 2];
    for (int j = uw - vw; j >= 0; --j) {
        // Estimate the quotient digit from the top two dividend digits
        const uint64_t num = (static_cast<uint64_t>(un[j + vw]) << 32) | un[j + vw - 1];
        uint64_t qhat = num / vtop;
        uint64_t rhat = num - qhat * vtop;
        while (qhat >= 0x100000000ULL || (qhat * vnext) > ((rhat << 32) + un[j + vw - 2])) {
            --qhat;
            rhat += vtop;
            if (rhat >= 0x100000000ULL) break;
        }
        // Multiply and subtract
        int64_t t = 0;
        uint64_t k = 0;
        for (int i = 0; i < vw; ++i) {
            const uint64_t p = qhat * vn[i];
            t = static_cast<int64_t>(static_cast<uint64_t>(un[i + j]) - k - (p & 0xFFFFFFFFULL));
            un[i + j] = static_cast<uint32_t>(t);
            k = (p >> 32) - static_cast<uint64_t>(t >> 32);
        }
        t = static_cast<int64_t>(static_cast<uint64_t>(un[j + vw]) - k);
        un[j + vw] = static_cast<uint32_t>(t);
        owp[j] = static_cast<EData>(qhat);
        if (t < 0) {  // Estimate was one too large: add the divisor back
            --owp[j];
            uint64_t c = 0;
            for (int i = 0; i < vw; ++i) {
                const uint64_t sum
                    = static_cast<uint64_t>(un[i + j]) + static_cast<uint64_t>(vn[i]) + c;
                un[i + j] = static_cast<uint32_t>(sum);
                c = sum >> 32;
            }
            un[j + vw] += static_cast<uint32_t>(c);
        }
    }

    if (is_modulus) {  // Remainder is the de-normalised low digits of un
        for (int i = 0; i < vw; ++i) {
            owp[i] = s ? ((un[i] >> s) | (un[i + 1] << (32 - s))) : un[i];
        }
        for (int i = vw; i < words; ++i) owp[i] = 0;
    }
    return owp;
}
#define VL_DIV_WWW(lbits, owp, lwp, rwp) (_vl_moddiv_w(lbits, owp, lwp, rwp, 0))
#define VL_MODDIV_WWW(lbits, owp, lwp, rwp) (_vl_moddiv_w(lbits, owp, lwp, rwp, 1))

// Signed wide divide / modulus by sign-magnitude (verilated_funcs.h:1270-1303)
template <int T_MaxWords = VL_MULS_MAX_WORDS>
VLCU_HD WDataOutP VL_DIVS_WWW(int lbits, WDataOutP owp, WDataInP lwp, WDataInP rwp) {
    const int words = VL_WORDS_I(lbits);
    const bool lneg = VL_SIGN_E(lbits, lwp[words - 1]) & 1;
    const bool rneg = VL_SIGN_E(lbits, rwp[words - 1]) & 1;
    EData lstore[T_MaxWords];
    EData rstore[T_MaxWords];
    lwp = vlcu_abs_w(lbits, words, lstore, lwp, lneg);
    rwp = vlcu_abs_w(lbits, words, rstore, rwp, rneg);
    if (lneg != rneg) {
        EData quot[T_MaxWords];
        _vl_moddiv_w<T_MaxWords>(lbits, quot, lwp, rwp, false);
        VL_NEGATE_W(words, owp, quot);
        return _vl_clean_inplace_w(lbits, owp);
    }
    return _vl_moddiv_w<T_MaxWords>(lbits, owp, lwp, rwp, false);
}
template <int T_MaxWords = VL_MULS_MAX_WORDS>
VLCU_HD WDataOutP VL_MODDIVS_WWW(int lbits, WDataOutP owp, WDataInP lwp, WDataInP rwp) {
    const int words = VL_WORDS_I(lbits);
    const bool lneg = VL_SIGN_E(lbits, lwp[words - 1]) & 1;
    const bool rneg = VL_SIGN_E(lbits, rwp[words - 1]) & 1;
    EData lstore[T_MaxWords];
    EData rstore[T_MaxWords];
    lwp = vlcu_abs_w(lbits, words, lstore, lwp, lneg);
    rwp = vlcu_abs_w(lbits, words, rstore, rwp, rneg);
    if (lneg) {  // Remainder takes the dividend's sign
        EData rem[T_MaxWords];
        _vl_moddiv_w<T_MaxWords>(lbits, rem, lwp, rwp, true);
        VL_NEGATE_W(words, owp, rem);
        return _vl_clean_inplace_w(lbits, owp);
    }
    return _vl_moddiv_w<T_MaxWords>(lbits, owp, lwp, rwp, true);
}

//=========================================================================
// Power (verilated_funcs.h:1305-1386, verilated.cpp:778-869): square and
// multiply over the rbits exponent bits, result dirty.

VLCU_HD IData VL_POW_III(int, int, int rbits, IData lhs, IData rhs) {
    if (VLCU_UNLIKELY(rhs == 0)) return 1;
    if (VLCU_UNLIKELY(lhs == 0)) return 0;
    IData power = lhs;
    IData out = 1;
    for (int i = 0; i < rbits; ++i) {
        if (i > 0) power *= power;
        if (rhs & (1ULL << i)) out *= power;
    }
    return out;
}
VLCU_HD QData VL_POW_QQQ(int, int, int rbits, QData lhs, QData rhs) {
    if (VLCU_UNLIKELY(rhs == 0)) return 1;
    if (VLCU_UNLIKELY(lhs == 0)) return 0;
    QData power = lhs;
    QData out = 1;
    for (int i = 0; i < rbits; ++i) {
        if (i > 0) power *= power;
        if (rhs & (1ULL << i)) out *= power;
    }
    return out;
}
template <int T_MaxWords = VL_MULS_MAX_WORDS>
VLCU_HD WDataOutP VL_POW_WWW(int obits, int, int rbits, WDataOutP owp, WDataInP lwp,
                             WDataInP rwp) {
    const int owords = VL_WORDS_I(obits);
    EData power[T_MaxWords];
    EData scratch[T_MaxWords];
    owp[0] = 1;
    for (int i = 1; i < owords; ++i) owp[i] = 0;
    for (int i = 0; i < owords; ++i) power[i] = lwp[i];
    for (int bit = 0; bit < rbits; ++bit) {
        if (bit > 0) {
            for (int i = 0; i < owords; ++i) scratch[i] = power[i];
            VL_MUL_W(owords, power, scratch, scratch);
        }
        if (VL_BITISSET_W(rwp, bit)) {
            for (int i = 0; i < owords; ++i) scratch[i] = owp[i];
            VL_MUL_W(owords, owp, scratch, power);
        }
    }
    return owp;
}
template <int T_MaxWords = VL_MULS_MAX_WORDS>
VLCU_HD WDataOutP VL_POW_WWQ(int obits, int lbits, int rbits, WDataOutP owp, WDataInP lwp,
                             QData rhs) {
    EData rhsw[2];
    VL_SET_WQ(rhsw, rhs);
    return VL_POW_WWW<T_MaxWords>(obits, lbits, rbits, owp, lwp, rhsw);
}
VLCU_HD QData VL_POW_QQW(int, int, int rbits, QData lhs, WDataInP rwp) {
    if (!VL_REDOR_W(VL_WORDS_I(rbits), rwp)) return 1;
    if (VLCU_UNLIKELY(lhs == 0)) return 0;
    QData power = lhs;
    QData out = 1;
    for (int bit = 0; bit < rbits; ++bit) {
        if (bit > 0) power *= power;
        if (VL_BITISSET_W(rwp, bit)) out *= power;
    }
    return out;
}
#define VL_POW_IIQ(obits, lbits, rbits, lhs, rhs) VL_POW_QQQ(obits, lbits, rbits, lhs, rhs)
#define VL_POW_IIW(obits, lbits, rbits, lhs, rwp) VL_POW_QQW(obits, lbits, rbits, lhs, rwp)
#define VL_POW_QQI(obits, lbits, rbits, lhs, rhs) VL_POW_QQQ(obits, lbits, rbits, lhs, rhs)
#define VL_POW_WWI(obits, lbits, rbits, owp, lwp, rhs) \
    VL_POW_WWQ(obits, lbits, rbits, owp, lwp, rhs)

// Signed power: a negative exponent gives 0, except bases 1 and -1 (and 0 -> 'X' -> 0)
VLCU_HD IData VL_POWSS_III(int obits, int, int rbits, IData lhs, IData rhs, bool lsign,
                           bool rsign) {
    if (VLCU_UNLIKELY(rhs == 0)) return 1;
    if (rsign && VL_SIGN_I(rbits, rhs)) {
        if (lhs == 1) return 1;
        if (lsign && lhs == VL_MASK_I(obits)) return (rhs & 1) ? VL_MASK_I(obits) : 1;
        return 0;
    }
    return VL_POW_III(obits, rbits, rbits, lhs, rhs);
}
VLCU_HD QData VL_POWSS_QQQ(int obits, int, int rbits, QData lhs, QData rhs, bool lsign,
                           bool rsign) {
    if (VLCU_UNLIKELY(rhs == 0)) return 1;
    if (rsign && VL_SIGN_Q(rbits, rhs)) {
        if (lhs == 1) return 1;
        if (lsign && lhs == VL_MASK_Q(obits)) return (rhs & 1) ? VL_MASK_Q(obits) : 1;
        return 0;
    }
    return VL_POW_QQQ(obits, rbits, rbits, lhs, rhs);
}
template <int T_MaxWords = VL_MULS_MAX_WORDS>
VLCU_HD WDataOutP VL_POWSS_WWW(int obits, int, int rbits, WDataOutP owp, WDataInP lwp,
                               WDataInP rwp, bool lsign, bool rsign) {
    if (rsign && VL_SIGN_W(rbits, rwp)) {
        const int words = VL_WORDS_I(obits);
        VL_ZERO_W(obits, owp);
        // Classify the base: 0 = middle limbs all zero, ~0 = top limb is the all-ones pattern
        EData lor = 0;
        for (int i = 1; i < words - 1; ++i) lor |= lwp[i];
        lor |= (lwp[words - 1] == VL_MASK_E(rbits)) ? ~0U : 0U;
        if (lor == 0 && lwp[0] == 0) return owp;  // 0 ** negative is 'X' -> 0
        if (lor == 0 && lwp[0] == 1) {
            owp[0] = 1;
            return owp;
        }
        if (lsign && lor == ~0U && lwp[0] == ~0U) {  // -1 ** n
            if (rwp[0] & 1) return VL_ALLONES_W(obits, owp);
            owp[0] = 1;
            return owp;
        }
        return owp;
    }
    return VL_POW_WWW<T_MaxWords>(obits, rbits, rbits, owp, lwp, rwp);
}
template <int T_MaxWords = VL_MULS_MAX_WORDS>
VLCU_HD WDataOutP VL_POWSS_WWQ(int obits, int lbits, int rbits, WDataOutP owp, WDataInP lwp,
                               QData rhs, bool lsign, bool rsign) {
    EData rhsw[2];
    VL_SET_WQ(rhsw, rhs);
    return VL_POWSS_WWW<T_MaxWords>(obits, lbits, rbits, owp, lwp, rhsw, lsign, rsign);
}
VLCU_HD QData VL_POWSS_QQW(int obits, int, int rbits, QData lhs, WDataInP rwp, bool lsign,
                           bool rsign) {
    if (rsign && VL_SIGN_W(rbits, rwp)) {
        if (lhs == 1) return 1;
        if (lsign && lhs == VL_MASK_Q(obits)) return (rwp[0] & 1) ? VL_MASK_Q(obits) : 1;
        return 0;
    }
    return VL_POW_QQW(obits, rbits, rbits, lhs, rwp);
}
#define VL_POWSS_IIQ(obits, lbits, rbits, lhs, rhs, lsign, rsign) \
    VL_POWSS_QQQ(obits, lbits, rbits, lhs, rhs, lsign, rsign)
#define VL_POWSS_IIW(obits, lbits, rbits, lhs, rwp, lsign, rsign) \
    VL_POWSS_QQW(obits, lbits, rbits, lhs, rwp, lsign, rsign)
#define VL_POWSS_QQI(obits, lbits, rbits, lhs, rhs, lsign, rsign) \
    VL_POWSS_QQQ(obits, lbits, rbits, lhs, rhs, lsign, rsign)
#define VL_POWSS_WWI(obits, lbits, rbits, owp, lwp, rhs, lsign, rsign) \
    VL_POWSS_WWQ(obits, lbits, rbits, owp, lwp, rhs, lsign, rsign)

//=========================================================================
// Insert helpers (verilated_funcs.h:1391-1530): place the low bits of the
// source into destination bits [hbit:lbit]; 'rbits' is the destination's
// declared width, used to keep its top word clean.

VLCU_HD IData vlcu_insert32(IData dst, IData ld, int hbit, int lbit, int rbits) {
    const IData field = VL_MASK_I(hbit - lbit + 1) << lbit;
    return (dst & ~field) | ((ld << lbit) & field & VL_MASK_I(rbits));
}
VLCU_HD void _vl_insert_II(VlcuMem8& lhsr, IData ld, int hbit, int lbit, int rbits) {
    lhsr = static_cast<VlcuMem8>(vlcu_insert32(lhsr, ld, hbit, lbit, rbits));
}
VLCU_HD void _vl_insert_II(VlcuMem16& lhsr, IData ld, int hbit, int lbit, int rbits) {
    lhsr = static_cast<VlcuMem16>(vlcu_insert32(lhsr, ld, hbit, lbit, rbits));
}
VLCU_HD void _vl_insert_II(IData& lhsr, IData ld, int hbit, int lbit, int rbits) {
    lhsr = vlcu_insert32(lhsr, ld, hbit, lbit, rbits);
}
VLCU_HD void _vl_insert_QQ(QData& lhsr, QData ld, int hbit, int lbit, int rbits) {
    const QData field = VL_MASK_Q(hbit - lbit + 1) << lbit;
    lhsr = (lhsr & ~field) | ((ld << lbit) & field & VL_MASK_Q(rbits));
}

VLCU_HD void _vl_insert_WI(WDataOutP iowp, IData ld, int hbit, int lbit, int rbits = 0) {
    const int hoff = VL_BITBIT_E(hbit);
    const int loff = VL_BITBIT_E(lbit);
    const int roff = VL_BITBIT_E(rbits);
    const int hword = VL_BITWORD_E(hbit);
    const int lword = VL_BITWORD_E(lbit);
    const int rword = VL_BITWORD_E(rbits);
    // Only the destination's top word needs cleaning
    const EData clean = (hword == rword) ? VL_MASK_E(roff) : ~0U;
    if (hoff == 31 && loff == 0) {  // Whole aligned word
        iowp[lword] = ld & clean;
    } else if (hword == lword) {  // Field inside one word
        const EData field = VL_MASK_E(hoff - loff + 1) << loff;
        iowp[lword] = (iowp[lword] & ~field) | ((ld << loff) & field & clean);
    } else {  // Field straddles two words
        const EData lowfield = ~0U << loff;
        iowp[lword] = (iowp[lword] & ~lowfield) | ((ld << loff) & lowfield);
        // The upper word does not exist when the destination ends on a word boundary
        if (VLCU_LIKELY(!(hword == rword && roff == 0))) {
            const EData highfield = VL_MASK_E(hoff + 1);
            iowp[hword]
                = (iowp[hword] & ~highfield) | ((ld >> (32 - loff)) & highfield & clean);
        }
    }
}

// Extract lwp[hbit:lbit] into the low bits of a 32-bit destination of width rbits
VLCU_HD void _vl_insert_IW(IData& lhsr, WDataInP lwp, int hbit, int lbit, int rbits = 0) {
    const int hoff = VL_BITBIT_E(hbit);
    const int loff = VL_BITBIT_E(lbit);
    const int hword = VL_BITWORD_E(hbit);
    const int lword = VL_BITWORD_E(lbit);
    const IData clean = VL_MASK_I(rbits);
    if (hword == lword) {
        const IData field = VL_MASK_I(hoff - loff + 1);
        lhsr = (lhsr & ~field) | ((lwp[lword] >> loff) & field & clean);
    } else {
        const int nlow = 32 - loff;  // Bits supplied by the low source word
        const IData lowfield = VL_MASK_E(32 - loff);
        const IData highfield = VL_MASK_E(hoff + 1) << nlow;
        lhsr = (lhsr & ~lowfield) | ((lwp[lword] >> loff) & lowfield & clean);
        lhsr = (lhsr & ~highfield) | ((lwp[hword] << nlow) & highfield & clean);
    }
}

VLCU_HD void _vl_insert_WW(WDataOutP iowp, WDataInP lwp, int hbit, int lbit, int rbits = 0) {
    const int hoff = VL_BITBIT_E(hbit);
    const int loff = VL_BITBIT_E(lbit);
    const int roff = VL_BITBIT_E(rbits);
    const int lword = VL_BITWORD_E(lbit);
    const int hword = VL_BITWORD_E(hbit);
    const int rword = VL_BITWORD_E(rbits);
    const int words = VL_WORDS_I(hbit - lbit + 1);  // Source words
    const EData clean = (hword == rword) ? VL_MASK_E(roff) : ~0U;
    const EData highfield = VL_MASK_E(hoff + 1);  // Bits of the top destination word in range

    if (loff == 0) {  // Source limbs line up with destination limbs
        for (int i = 0; i < words - 1; ++i) iowp[lword + i] = lwp[i];
        if (hoff == 31) {
            iowp[hword] = lwp[words - 1] & clean;
        } else {
            iowp[hword] = (iowp[hword] & ~highfield) | (lwp[words - 1] & highfield & clean);
        }
        return;
    }
    // Unaligned: every destination word takes the high part of one source
    // limb shifted up and the low part of the previous one shifted down.
    const EData upfield = ~0U << loff;  // Destination bits fed by "limb << loff"
    const int nup = 32 - loff;
    for (int ow = lword; ow <= hword; ++ow) {
        const int i = ow - lword;
        const EData old = iowp[ow];
        // Bits below loff: previous source limb, or preserved in the first word
        const EData below = (i > 0) ? (lwp[i - 1] >> nup) : (old & ~upfield);
        // Bits from loff up: this source limb, or preserved when the source has run out
        const EData above = (i < words) ? (lwp[i] << loff) : (old & upfield);
        const EData merged = below | above;
        if (ow == hword) {
            iowp[ow] = (old & ~highfield) | (merged & highfield & clean);
        } else {
            iowp[ow] = merged;
        }
    }
}
VLCU_HD void _vl_insert_WQ(WDataOutP iowp, QData ld, int hbit, int lbit, int rbits = 0) {
    EData lw[2];
    VL_SET_WQ(lw, ld);
    _vl_insert_WW(iowp, lw, hbit, lbit, rbits);
}

//=========================================================================
// Replication (verilated_funcs.h:1532-1585); rep is a clean constant

#define VL_REPLICATE_IOI(lbits, ld, rep) (-(ld))  // Iff lbits==1
#define VL_REPLICATE_QOI(lbits, ld, rep) (-(static_cast<QData>(ld)))  // Iff lbits==1

VLCU_HD IData VL_REPLICATE_III(int lbits, IData ld, IData rep) {
    IData out = ld;
    for (unsigned i = 1; i < rep; ++i) out = (out << lbits) | ld;
    return out;
}
VLCU_HD QData VL_REPLICATE_QII(int lbits, IData ld, IData rep) {
    QData out = ld;
    for (unsigned i = 1; i < rep; ++i) out = (out << lbits) | static_cast<QData>(ld);
    return out;
}
VLCU_HD WDataOutP VL_REPLICATE_WII(int lbits, WDataOutP owp, IData ld, IData rep) {
    owp[0] = ld;
    for (unsigned i = 1; i < VL_WORDS_I(static_cast<unsigned>(lbits) * rep); ++i) owp[i] = 0;
    for (unsigned i = 1; i < rep; ++i) _vl_insert_WI(owp, ld, i * lbits + lbits - 1, i * lbits);
    return owp;
}
VLCU_HD WDataOutP VL_REPLICATE_WQI(int lbits, WDataOutP owp, QData ld, IData rep) {
    VL_SET_WQ(owp, ld);
    for (unsigned i = 2; i < VL_WORDS_I(static_cast<unsigned>(lbits) * rep); ++i) owp[i] = 0;
    for (unsigned i = 1; i < rep; ++i) _vl_insert_WQ(owp, ld, i * lbits + lbits - 1, i * lbits);
    return owp;
}
VLCU_HD WDataOutP VL_REPLICATE_WWI(int lbits, WDataOutP owp, WDataInP lwp, IData rep) {
    const unsigned lwords = VL_WORDS_I(static_cast<unsigned>(lbits));
    for (unsigned i = 0; i < lwords; ++i) owp[i] = lwp[i];
    for (unsigned i = lwords; i < VL_WORDS_I(static_cast<unsigned>(lbits * rep)); ++i) owp[i] = 0;
    for (unsigned i = 1; i < rep; ++i) _vl_insert_WW(owp, lwp, i * lbits + lbits - 1, i * lbits);
    return owp;
}

//=========================================================================
// Streaming {<<{}} (verilated_funcs.h:1587-1658, 1953-1963, 2071-2100)

// Power-of-two slice sizes: butterfly network.  Bits of a partial most
// significant slice are first moved up so the reversal leaves no gap.
VLCU_HD IData VL_STREAML_FAST_III(int lbits, IData ld, IData rd_log2) {
    IData ret = ld;
    if (rd_log2) {
        const uint32_t whole = lbits & ~VL_MASK_I(rd_log2);  // Largest slice multiple <= lbits
        const uint32_t rem = lbits - whole;  // Bits in the partial top slice
        const IData topmask = (whole == 32) ? 0U : (VL_MASK_I(rem) << whole);
        ret = (ret & ~topmask) | ((ret & topmask) << ((1U << rd_log2) - rem));
    }
    if (rd_log2 <= 0) ret = ((ret >> 1) & 0x55555555U) | ((ret & 0x55555555U) << 1);
    if (rd_log2 <= 1) ret = ((ret >> 2) & 0x33333333U) | ((ret & 0x33333333U) << 2);
    if (rd_log2 <= 2) ret = ((ret >> 4) & 0x0f0f0f0fU) | ((ret & 0x0f0f0f0fU) << 4);
    if (rd_log2 <= 3) ret = ((ret >> 8) & 0x00ff00ffU) | ((ret & 0x00ff00ffU) << 8);
    if (rd_log2 <= 4) ret = (ret >> 16) | (ret << 16);
    return ret >> (32 - lbits);
}
VLCU_HD QData VL_STREAML_FAST_QQI(int lbits, QData ld, IData rd_log2) {
    QData ret = ld;
    if (rd_log2) {
        const uint32_t whole = lbits & ~VL_MASK_I(rd_log2);
        const uint32_t rem = lbits - whole;
        const QData topmask = (whole == 64) ? 0ULL : (VL_MASK_Q(rem) << whole);
        ret = (ret & ~topmask) | ((ret & topmask) << ((1ULL << rd_log2) - rem));
    }
    if (rd_log2 <= 0) {
        ret = ((ret >> 1) & 0x5555555555555555ULL) | ((ret & 0x5555555555555555ULL) << 1);
    }
    if (rd_log2 <= 1) {
        ret = ((ret >> 2) & 0x3333333333333333ULL) | ((ret & 0x3333333333333333ULL) << 2);
    }
    if (rd_log2 <= 2) {
        ret = ((ret >> 4) & 0x0f0f0f0f0f0f0f0fULL) | ((ret & 0x0f0f0f0f0f0f0f0fULL) << 4);
    }
    if (rd_log2 <= 3) {
        ret = ((ret >> 8) & 0x00ff00ff00ff00ffULL) | ((ret & 0x00ff00ff00ff00ffULL) << 8);
    }
    if (rd_log2 <= 4) {
        ret = ((ret >> 16) & 0x0000ffff0000ffffULL) | ((ret & 0x0000ffff0000ffffULL) << 16);
    }
    if (rd_log2 <= 5) ret = (ret >> 32) | (ret << 32);
    return ret >> (64 - lbits);
}
// General slice sizes: move slice [istart +: rd] to the mirrored position
VLCU_HD IData VL_STREAML_III(int lbits, IData ld, IData rd) {
    IData ret = 0;
    const IData mask = VL_MASK_I(rd);
    for (int istart = 0; istart < lbits; istart += rd) {
        int ostart = lbits - static_cast<int>(rd) - istart;
        if (ostart < 0) ostart = 0;
        ret |= ((ld >> istart) & mask) << ostart;
    }
    return ret;
}
VLCU_HD QData VL_STREAML_QQI(int lbits, QData ld, IData rd) {
    QData ret = 0;
    const QData mask = VL_MASK_Q(rd);
    for (int istart = 0; istart < lbits; istart += rd) {
        int ostart = lbits - static_cast<int>(rd) - istart;
        if (ostart < 0) ostart = 0;
        ret |= ((ld >> istart) & mask) << ostart;
    }
    return ret;
}
VLCU_HD WDataOutP VL_STREAML_WWI(int lbits, WDataOutP owp, WDataInP lwp, IData rd) {
    VL_ZERO_W(lbits, owp);
    const int ssize = (rd < static_cast<IData>(lbits)) ? rd : lbits;
    for (int istart = 0; istart < lbits; istart += rd) {
        int ostart = lbits - static_cast<int>(rd) - istart;
        if (ostart < 0) ostart = 0;
        for (int sbit = 0; sbit < ssize && sbit < lbits - istart; ++sbit) {
            const EData bit = (VL_BITRSHIFT_W(lwp, (istart + sbit)) & 1U)
                              << VL_BITBIT_E(ostart + sbit);
            owp[VL_BITWORD_E(ostart + sbit)] |= bit;
        }
    }
    return owp;
}

//=========================================================================
// Concatenation (verilated_funcs.h:2390-2471): inputs clean, output clean

#define VL_CONCAT_III(obits, lbits, rbits, ld, rd) \
    (static_cast<IData>(ld) << (rbits) | static_cast<IData>(rd))
#define VL_CONCAT_QII(obits, lbits, rbits, ld, rd) \
    (static_cast<QData>(ld) << (rbits) | static_cast<QData>(rd))
#define VL_CONCAT_QIQ(obits, lbits, rbits, ld, rd) \
    (static_cast<QData>(ld) << (rbits) | static_cast<QData>(rd))
#define VL_CONCAT_QQI(obits, lbits, rbits, ld, rd) \
    (static_cast<QData>(ld) << (rbits) | static_cast<QData>(rd))
#define VL_CONCAT_QQQ(obits, lbits, rbits, ld, rd) \
    (static_cast<QData>(ld) << (rbits) | static_cast<QData>(rd))

// Lay down the right operand in the low words and zero the rest
VLCU_HD void vlcu_concat_base_i(int obits, WDataOutP owp, IData rd) {
    owp[0] = rd;
    for (int i = 1; i < VL_WORDS_I(obits); ++i) owp[i] = 0;
}
VLCU_HD void vlcu_concat_base_q(int obits, WDataOutP owp, QData rd) {
    VL_SET_WQ(owp, rd);
    for (int i = 2; i < VL_WORDS_I(obits); ++i) owp[i] = 0;
}
VLCU_HD void vlcu_concat_base_w(int obits, int rbits, WDataOutP owp, WDataInP rwp) {
    const int rwords = VL_WORDS_I(rbits);
    for (int i = 0; i < rwords; ++i) owp[i] = rwp[i];
    for (int i = rwords; i < VL_WORDS_I(obits); ++i) owp[i] = 0;
}
VLCU_HD WDataOutP VL_CONCAT_WII(int obits, int lbits, int rbits, WDataOutP owp, IData ld,
                                IData rd) {
    vlcu_concat_base_i(obits, owp, rd);
    _vl_insert_WI(owp, ld, rbits + lbits - 1, rbits);
    return owp;
}
VLCU_HD WDataOutP VL_CONCAT_WWI(int obits, int lbits, int rbits, WDataOutP owp, WDataInP lwp,
                                IData rd) {
    vlcu_concat_base_i(obits, owp, rd);
    _vl_insert_WW(owp, lwp, rbits + lbits - 1, rbits);
    return owp;
}
VLCU_HD WDataOutP VL_CONCAT_WIW(int obits, int lbits, int rbits, WDataOutP owp, IData ld,
                                WDataInP rwp) {
    vlcu_concat_base_w(obits, rbits, owp, rwp);
    _vl_insert_WI(owp, ld, rbits + lbits - 1, rbits);
    return owp;
}
VLCU_HD WDataOutP VL_CONCAT_WIQ(int obits, int lbits, int rbits, WDataOutP owp, IData ld,
                                QData rd) {
    vlcu_concat_base_q(obits, owp, rd);
    _vl_insert_WI(owp, ld, rbits + lbits - 1, rbits);
    return owp;
}
VLCU_HD WDataOutP VL_CONCAT_WQI(int obits, int lbits, int rbits, WDataOutP owp, QData ld,
                                IData rd) {
    vlcu_concat_base_i(obits, owp, rd);
    _vl_insert_WQ(owp, ld, rbits + lbits - 1, rbits);
    return owp;
}
VLCU_HD WDataOutP VL_CONCAT_WQQ(int obits, int lbits, int rbits, WDataOutP owp, QData ld,
                                QData rd) {
    vlcu_concat_base_q(obits, owp, rd);
    _vl_insert_WQ(owp, ld, rbits + lbits - 1, rbits);
    return owp;
}
VLCU_HD WDataOutP VL_CONCAT_WWQ(int obits, int lbits, int rbits, WDataOutP owp, WDataInP lwp,
                                QData rd) {
    vlcu_concat_base_q(obits, owp, rd);
    _vl_insert_WW(owp, lwp, rbits + lbits - 1, rbits);
    return owp;
}
VLCU_HD WDataOutP VL_CONCAT_WQW(int obits, int lbits, int rbits, WDataOutP owp, QData ld,
                                WDataInP rwp) {
    vlcu_concat_base_w(obits, rbits, owp, rwp);
    _vl_insert_WQ(owp, ld, rbits + lbits - 1, rbits);
    return owp;
}
VLCU_HD WDataOutP VL_CONCAT_WWW(int obits, int lbits, int rbits, WDataOutP owp, WDataInP lwp,
                                WDataInP rwp) {
    vlcu_concat_base_w(obits, rbits, owp, rwp);
    _vl_insert_WW(owp, lwp, rbits + lbits - 1, rbits);
    return owp;
}

//=========================================================================
// Shifts (verilated_funcs.h:2474-2733).  A shift amount at or beyond the
// word size gives 0 (sign fill for >>>); a wide amount with any high limb
// set likewise.

VLCU_HD void _vl_shiftl_inplace_w(int obits, WDataOutP iowp, IData rd /*1 or 4*/) {
    const int words = VL_WORDS_I(obits);
    for (int i = words - 1; i >= 1; --i) iowp[i] = (iowp[i] << rd) | (iowp[i - 1] >> (32 - rd));
    iowp[0] <<= rd;
    iowp[words - 1] &= VL_MASK_E(obits);
}

VLCU_HD IData VL_SHIFTL_III(int, int, int, IData lhs, IData rhs) {
    return (rhs >= 32) ? 0 : (lhs << rhs);  // Dirty
}
VLCU_HD IData VL_SHIFTL_IIQ(int obits, int, int, IData lhs, QData rhs) {
    return (rhs >= 32) ? 0 : ((lhs << rhs) & VL_MASK_I(obits));
}
VLCU_HD QData VL_SHIFTL_QQI(int, int, int, QData lhs, IData rhs) {
    return (rhs >= 64) ? 0 : (lhs << rhs);  // Dirty
}
VLCU_HD QData VL_SHIFTL_QQQ(int obits, int, int, QData lhs, QData rhs) {
    return (rhs >= 64) ? 0 : ((lhs << rhs) & VL_MASK_Q(obits));
}
VLCU_HD WDataOutP VL_SHIFTL_WWI(int obits, int, int, WDataOutP owp, WDataInP lwp, IData rd) {
    const int owords = VL_WORDS_I(obits);
    const int wshift = VL_BITWORD_E(rd);
    const int bshift = VL_BITBIT_E(rd);
    if (rd >= static_cast<IData>(obits)) {
        for (int i = 0; i < owords; ++i) owp[i] = 0;
    } else if (bshift == 0) {
        for (int i = 0; i < owords; ++i) owp[i] = (i >= wshift) ? lwp[i - wshift] : 0;
    } else {
        for (int i = 0; i < owords; ++i) owp[i] = 0;
        _vl_insert_WW(owp, lwp, obits - 1, rd);
    }
    return owp;
}
// True when any limb above the first is non-zero (shift amount >= 2**32)
VLCU_HD bool vlcu_shift_huge(int rbits, WDataInP rwp) {
    EData hi = 0;
    for (int i = 1; i < VL_WORDS_I(rbits); ++i) hi |= rwp[i];
    return hi != 0;
}
VLCU_HD WDataOutP VL_SHIFTL_WWW(int obits, int lbits, int rbits, WDataOutP owp, WDataInP lwp,
                                WDataInP rwp) {
    if (VLCU_UNLIKELY(vlcu_shift_huge(rbits, rwp))) return VL_ZERO_W(obits, owp);
    return VL_SHIFTL_WWI(obits, lbits, 32, owp, lwp, rwp[0]);
}
VLCU_HD WDataOutP VL_SHIFTL_WWQ(int obits, int lbits, int rbits, WDataOutP owp, WDataInP lwp,
                                QData rd) {
    EData rw[2];
    VL_SET_WQ(rw, rd);
    return VL_SHIFTL_WWW(obits, lbits, rbits, owp, lwp, rw);
}
VLCU_HD IData VL_SHIFTL_IIW(int obits, int, int rbits, IData lhs, WDataInP rwp) {
    if (VLCU_UNLIKELY(vlcu_shift_huge(rbits, rwp))) return 0;
    return VL_SHIFTL_III(obits, obits, 32, lhs, rwp[0]);
}
VLCU_HD QData VL_SHIFTL_QQW(int obits, int, int rbits, QData lhs, WDataInP rwp) {
    if (VLCU_UNLIKELY(vlcu_shift_huge(rbits, rwp))) return 0;
    return VL_SHIFTL_QQI(obits, obits, 32, lhs, rwp[0]);
}

VLCU_HD IData VL_SHIFTR_III(int, int, int, IData lhs, IData rhs) {
    return (rhs >= 32) ? 0 : (lhs >> rhs);
}
VLCU_HD IData VL_SHIFTR_IIQ(int, int, int, IData lhs, QData rhs) {
    return (rhs >= 32) ? 0 : (lhs >> rhs);
}
VLCU_HD QData VL_SHIFTR_QQI(int, int, int, QData lhs, IData rhs) {
    return (rhs >= 64) ? 0 : (lhs >> rhs);
}
VLCU_HD QData VL_SHIFTR_QQQ(int, int, int, QData lhs, QData rhs) {
    return (rhs >= 64) ? 0 : (lhs >> rhs);
}
VLCU_HD WDataOutP VL_SHIFTR_WWI(int obits, int, int, WDataOutP owp, WDataInP lwp, IData rd) {
    const int owords = VL_WORDS_I(obits);
    const int wshift = VL_BITWORD_E(rd);
    const int bshift = VL_BITBIT_E(rd);
    if (rd >= static_cast<IData>(obits)) {
        for (int i = 0; i < owords; ++i) owp[i] = 0;
    } else if (bshift == 0) {
        const int ncopy = owords - wshift;
        for (int i = 0; i < owords; ++i) owp[i] = (i < ncopy) ? lwp[i + wshift] : 0;
    } else {
        const int nkeep = VL_WORDS_I(obits - rd);  // Result words that receive data
        for (int i = 0; i < owords; ++i) {
            EData w = 0;
            if (i < nkeep) {
                w = lwp[i + wshift] >> bshift;
                if (i + wshift + 1 < owords) w |= lwp[i + wshift + 1] << (32 - bshift);
            }
            owp[i] = w;
        }
    }
    return owp;
}
VLCU_HD WDataOutP VL_SHIFTR_WWW(int obits, int lbits, int rbits, WDataOutP owp, WDataInP lwp,
                                WDataInP rwp) {
    if (VLCU_UNLIKELY(vlcu_shift_huge(rbits, rwp))) return VL_ZERO_W(obits, owp);
    return VL_SHIFTR_WWI(obits, lbits, 32, owp, lwp, rwp[0]);
}
VLCU_HD WDataOutP VL_SHIFTR_WWQ(int obits, int lbits, int rbits, WDataOutP owp, WDataInP lwp,
                                QData rd) {
    EData rw[2];
    VL_SET_WQ(rw, rd);
    return VL_SHIFTR_WWW(obits, lbits, rbits, owp, lwp, rw);
}
VLCU_HD IData VL_SHIFTR_IIW(int obits, int, int rbits, IData lhs, WDataInP rwp) {
    if (VLCU_UNLIKELY(vlcu_shift_huge(rbits, rwp))) return 0;
    return VL_SHIFTR_III(obits, obits, 32, lhs, rwp[0]);
}
VLCU_HD QData VL_SHIFTR_QQW(int obits, int, int rbits, QData lhs, WDataInP rwp) {
    if (VLCU_UNLIKELY(vlcu_shift_huge(rbits, rwp))) return 0;
    return VL_SHIFTR_QQI(obits, obits, 32, lhs, rwp[0]);
}

// Arithmetic right shift: the sign is bit lbits-1 of the input (lbits may be
// narrower than obits), result not clean above obits only via the sign fill.
VLCU_HD IData VL_SHIFTRS_III(int obits, int lbits, int, IData lhs, IData rhs) {
    if (obits == 32 && lbits == 32) {  // the machine's arithmetic shift (clamps the count itself)
        return static_cast<IData>(static_cast<int32_t>(lhs) >> (rhs < 31U ? rhs : 31U));
    }
    const IData sign = 0U - (lhs >> (lbits - 1));
    if (VLCU_UNLIKELY(rhs >= 32)) return sign & VL_MASK_I(obits);
    const IData vacated = ~(VL_MASK_I(lbits) >> rhs);
    return (lhs >> rhs) | (sign & vacated & VL_MASK_I(obits));
}
VLCU_HD QData VL_SHIFTRS_QQI(int obits, int lbits, int, QData lhs, IData rhs) {
    const QData sign = 0ULL - (lhs >> (lbits - 1));
    if (VLCU_UNLIKELY(rhs >= 64)) return sign & VL_MASK_Q(obits);
    const QData vacated = ~(VL_MASK_Q(lbits) >> rhs);
    return (lhs >> rhs) | (sign & vacated & VL_MASK_Q(obits));
}
VLCU_HD IData VL_SHIFTRS_IQI(int obits, int lbits, int rbits, QData lhs, IData rhs) {
    return static_cast<IData>(VL_SHIFTRS_QQI(obits, lbits, rbits, lhs, rhs));
}
VLCU_HD WDataOutP VL_SHIFTRS_WWI(int obits, int lbits, int, WDataOutP owp, WDataInP lwp,
                                 IData rd) {
    const int owords = VL_WORDS_I(obits);
    const int wshift = VL_BITWORD_E(rd);
    const int bshift = VL_BITBIT_E(rd);
    const int msw = owords - 1;
    const EData sign = VL_SIGNONES_E(lbits, lwp[msw]);
    if (rd >= static_cast<IData>(obits)) {
        for (int i = 0; i <= msw; ++i) owp[i] = sign;
    } else if (bshift == 0) {
        const int ncopy = owords - wshift;
        for (int i = 0; i < ncopy; ++i) owp[i] = lwp[i + wshift];
        owp[ncopy - 1] |= ~VL_MASK_E(obits) & sign;
        for (int i = ncopy; i < owords; ++i) owp[i] = sign;
    } else {
        const int nkeep = VL_WORDS_I(obits - rd);
        for (int i = 0; i < nkeep; ++i) {
            EData w = lwp[i + wshift] >> bshift;
            if (i + wshift + 1 < owords) w |= lwp[i + wshift + 1] << (32 - bshift);
            owp[i] = w;
        }
        if (nkeep) owp[nkeep - 1] |= sign & ~VL_MASK_E(obits - bshift);
        for (int i = nkeep; i < owords; ++i) owp[i] = sign;
    }
    owp[msw] &= VL_MASK_E(lbits);
    return owp;
}
VLCU_HD WDataOutP VL_SHIFTRS_WWW(int obits, int lbits, int rbits, WDataOutP owp, WDataInP lwp,
                                 WDataInP rwp) {
    if (VLCU_UNLIKELY(vlcu_shift_huge(rbits, rwp) || rwp[0] >= static_cast<IData>(obits))) {
        const int owords = VL_WORDS_I(obits);
        if (VL_SIGN_E(lbits, lwp[owords - 1])) {
            VL_MEMSET_ONES_W(owp, owords);
            owp[owords - 1] &= VL_MASK_E(lbits);
        } else {
            VL_MEMSET_ZERO_W(owp, owords);
        }
        return owp;
    }
    return VL_SHIFTRS_WWI(obits, lbits, 32, owp, lwp, rwp[0]);
}
VLCU_HD WDataOutP VL_SHIFTRS_WWQ(int obits, int lbits, int rbits, WDataOutP owp, WDataInP lwp,
                                 QData rd) {
    EData rw[2];
    VL_SET_WQ(rw, rd);
    return VL_SHIFTRS_WWW(obits, lbits, rbits, owp, lwp, rw);
}
VLCU_HD IData VL_SHIFTRS_IIW(int obits, int lbits, int rbits, IData lhs, WDataInP rwp) {
    if (VLCU_UNLIKELY(vlcu_shift_huge(rbits, rwp) || rwp[0] >= static_cast<IData>(obits))) {
        const IData sign = 0U - (lhs >> (lbits - 1));
        return sign & VL_MASK_I(obits);
    }
    return VL_SHIFTRS_III(obits, lbits, 32, lhs, rwp[0]);
}
VLCU_HD QData VL_SHIFTRS_QQW(int obits, int lbits, int rbits, QData lhs, WDataInP rwp) {
    if (VLCU_UNLIKELY(vlcu_shift_huge(rbits, rwp) || rwp[0] >= static_cast<IData>(obits))) {
        const QData sign = 0ULL - (lhs >> (lbits - 1));
        return sign & VL_MASK_Q(obits);
    }
    return VL_SHIFTRS_QQI(obits, lbits, 32, lhs, rwp[0]);
}
VLCU_HD IData VL_SHIFTRS_IIQ(int obits, int lbits, int rbits, IData lhs, QData rhs) {
    EData rw[2];
    VL_SET_WQ(rw, rhs);
    return VL_SHIFTRS_IIW(obits, lbits, rbits, lhs, rw);
}
VLCU_HD QData VL_SHIFTRS_QQQ(int obits, int lbits, int rbits, QData lhs, QData rhs) {
    EData rw[2];
    VL_SET_WQ(rw, rhs);
    return VL_SHIFTRS_QQW(obits, lbits, rbits, lhs, rw);
}

//=========================================================================
// Bit and part selects (verilated_funcs.h:2736-2856); results are dirty and
// out-of-range selects read as all ones.  The three range tests below differ
// on purpose: they are what the host runtime does.

#define VL_BITSEL_IIII(lbits, lhs, rhs) ((lhs) >> (rhs))
#define VL_BITSEL_QIII(lbits, lhs, rhs) ((lhs) >> (rhs))
#define VL_BITSEL_QQII(lbits, lhs, rhs) ((lhs) >> (rhs))
#define VL_BITSEL_IQII(lbits, lhs, rhs) (static_cast<IData>((lhs) >> (rhs)))

VLCU_HD IData VL_BITSEL_IWII(int lbits, WDataInP lwp, IData rd) {
    if (VLCU_UNLIKELY(rd > static_cast<IData>(lbits))) return ~0U;
    return lwp[VL_BITWORD_E(rd)] >> VL_BITBIT_E(rd);
}

#define VL_SEL_IIII(lbits, lhs, lsb, width) ((lhs) >> (lsb))
#define VL_SEL_QQII(lbits, lhs, lsb, width) ((lhs) >> (lsb))
#define VL_SEL_IQII(lbits, lhs, lsb, width) (static_cast<IData>((lhs) >> (lsb)))

VLCU_HD IData VL_SEL_IWII(int lbits, WDataInP lwp, IData lsb, IData width) {
    const int msb = lsb + width - 1;
    if (VLCU_UNLIKELY(msb >= lbits)) return ~0U;
    const int lw = VL_BITWORD_E(static_cast<int>(lsb));
    const int hw = VL_BITWORD_E(msb);
    const EData low = lwp[lw] >> VL_BITBIT_E(lsb);
    if (hw == lw) return low;
    return (lwp[hw] << (32 - VL_BITBIT_E(lsb))) | low;
}
VLCU_HD QData VL_SEL_QWII(int lbits, WDataInP lwp, IData lsb, IData width) {
    const int msb = lsb + width - 1;
    if (VLCU_UNLIKELY(msb > lbits)) return ~0ULL;
    const int lw = VL_BITWORD_E(static_cast<int>(lsb));
    const int hw = VL_BITWORD_E(msb);
    const int nlow = 32 - VL_BITBIT_E(lsb);  // Bits taken from the lowest word
    const QData low = lwp[lw] >> VL_BITBIT_E(lsb);
    if (hw == lw) return low;
    if (hw == lw + 1) return (static_cast<QData>(lwp[hw]) << nlow) | low;
    return (static_cast<QData>(lwp[hw]) << (nlow + 32)) | (static_cast<QData>(lwp[lw + 1]) << nlow)
           | low;
}
VLCU_HD WDataOutP VL_SEL_WWII(int obits, int lbits, WDataOutP owp, WDataInP lwp, IData lsb,
                              IData width) {
    const int owords = VL_WORDS_I(obits);
    const int msb = lsb + width - 1;
    const int wshift = VL_BITWORD_E(lsb);
    const int bshift = VL_BITBIT_E(lsb);
    if (VLCU_UNLIKELY(msb > lbits)) {
        for (int i = 0; i < owords - 1; ++i) owp[i] = ~0U;
        owp[owords - 1] = VL_MASK_E(obits);
    } else if (bshift == 0) {
        for (int i = 0; i < owords; ++i) owp[i] = lwp[i + wshift];
    } else {
        const int nsrc = VL_WORDS_I(msb - static_cast<int>(lsb) + 1);
        const int topword = VL_BITWORD_E(msb);
        for (int i = 0; i < owords; ++i) {
            EData w = 0;
            if (i < nsrc) {
                w = lwp[i + wshift] >> bshift;
                if (i + wshift + 1 <= topword) w |= lwp[i + wshift + 1] << (32 - bshift);
            }
            owp[i] = w;
        }
    }
    return owp;
}

//=========================================================================
// Range assignment (verilated_funcs.h:3346-3461); rhs may be dirty

VLCU_HD void VL_ASSIGNSEL_II(int rbits, int obits, int lsb, VlcuMem8& lhsr, IData rhs) {
    _vl_insert_II(lhsr, rhs, lsb + obits - 1, lsb, rbits);
}
VLCU_HD void VL_ASSIGNSEL_II(int rbits, int obits, int lsb, VlcuMem16& lhsr, IData rhs) {
    _vl_insert_II(lhsr, rhs, lsb + obits - 1, lsb, rbits);
}
VLCU_HD void VL_ASSIGNSEL_II(int rbits, int obits, int lsb, IData& lhsr, IData rhs) {
    _vl_insert_II(lhsr, rhs, lsb + obits - 1, lsb, rbits);
}
VLCU_HD void VL_ASSIGNSEL_QI(int rbits, int obits, int lsb, QData& lhsr, IData rhs) {
    _vl_insert_QQ(lhsr, rhs, lsb + obits - 1, lsb, rbits);
}
VLCU_HD void VL_ASSIGNSEL_QQ(int rbits, int obits, int lsb, QData& lhsr, QData rhs) {
    _vl_insert_QQ(lhsr, rhs, lsb + obits - 1, lsb, rbits);
}
VLCU_HD void VL_ASSIGNSEL_WI(int rbits, int obits, int lsb, WDataOutP iowp, IData rhs) {
    _vl_insert_WI(iowp, rhs, lsb + obits - 1, lsb, rbits);
}
VLCU_HD void VL_ASSIGNSEL_WQ(int rbits, int obits, int lsb, WDataOutP iowp, QData rhs) {
    _vl_insert_WQ(iowp, rhs, lsb + obits - 1, lsb, rbits);
}
VLCU_HD void VL_ASSIGNSEL_WW(int rbits, int obits, int lsb, WDataOutP iowp, WDataInP rwp) {
    _vl_insert_WW(iowp, rwp, lsb + obits - 1, lsb, rbits);
}

// Select-then-assign: copy rhs[roffset +: obits] to the low bits of lhs
VLCU_HD void VL_SELASSIGN_II(int rbits, int obits, VlcuMem8& lhsr, IData rhs, int roffset) {
    _vl_insert_II(lhsr, rhs >> roffset, obits - 1, 0, rbits);
}
VLCU_HD void VL_SELASSIGN_II(int rbits, int obits, VlcuMem16& lhsr, IData rhs, int roffset) {
    _vl_insert_II(lhsr, rhs >> roffset, obits - 1, 0, rbits);
}
VLCU_HD void VL_SELASSIGN_II(int rbits, int obits, IData& lhsr, IData rhs, int roffset) {
    _vl_insert_II(lhsr, rhs >> roffset, obits - 1, 0, rbits);
}
template <typename T_Narrow>
VLCU_HD void vlcu_selassign_iq(int rbits, int obits, T_Narrow& lhsr, QData rhs, int roffset) {
    // Masks are truncated to the destination type before use
    const T_Narrow clean = static_cast<T_Narrow>(VL_MASK_I(rbits));
    const T_Narrow field = static_cast<T_Narrow>(VL_MASK_I(obits));
    lhsr = static_cast<T_Narrow>((lhsr & ~field)
                                 | (static_cast<T_Narrow>(rhs >> roffset) & (field & clean)));
}
VLCU_HD void VL_SELASSIGN_IQ(int rbits, int obits, VlcuMem8& lhsr, QData rhs, int roffset) {
    vlcu_selassign_iq(rbits, obits, lhsr, rhs, roffset);
}
VLCU_HD void VL_SELASSIGN_IQ(int rbits, int obits, VlcuMem16& lhsr, QData rhs, int roffset) {
    vlcu_selassign_iq(rbits, obits, lhsr, rhs, roffset);
}
VLCU_HD void VL_SELASSIGN_IQ(int rbits, int obits, IData& lhsr, QData rhs, int roffset) {
    vlcu_selassign_iq(rbits, obits, lhsr, rhs, roffset);
}
VLCU_HD void VL_SELASSIGN_QQ(int rbits, int obits, QData& lhsr, QData rhs, int roffset) {
    _vl_insert_QQ(lhsr, rhs >> roffset, obits - 1, 0, rbits);
}
VLCU_HD void VL_SELASSIGN_IW(int rbits, int obits, VlcuMem8& lhsr, WDataInP rhs, int roffset) {
    IData l = lhsr;
    _vl_insert_IW(l, rhs, roffset + obits - 1, roffset, rbits);
    lhsr = static_cast<VlcuMem8>(l);
}
VLCU_HD void VL_SELASSIGN_IW(int rbits, int obits, VlcuMem16& lhsr, WDataInP rhs, int roffset) {
    IData l = lhsr;
    _vl_insert_IW(l, rhs, roffset + obits - 1, roffset, rbits);
    lhsr = static_cast<VlcuMem16>(l);
}
VLCU_HD void VL_SELASSIGN_IW(int rbits, int obits, IData& lhsr, WDataInP rhs, int roffset) {
    _vl_insert_IW(lhsr, rhs, roffset + obits - 1, roffset, rbits);
}
VLCU_HD void VL_SELASSIGN_QW(int rbits, int obits, QData& lhsr, WDataInP rhs, int roffset) {
    IData low = static_cast<IData>(lhsr);
    IData high = static_cast<IData>(lhsr >> 32);
    if (obits <= 32) {
        _vl_insert_IW(low, rhs, obits + roffset - 1, roffset, 32);
    } else {
        _vl_insert_IW(low, rhs, roffset + 32 - 1, roffset, 32);
        _vl_insert_IW(high, rhs, roffset + obits - 1, roffset + 32, rbits - 32);
    }
    lhsr = (static_cast<QData>(high) << 32) | low;
}
VLCU_HD void VL_SELASSIGN_WW(int rbits, int obits, WDataOutP iowp, WDataInP rwp, int roffset) {
    const int wordoff = roffset / 32;
    const int lsb = roffset & 31;
    const int upperbits = lsb == 0 ? 0 : 32 - lsb;
    if (lsb != 0) {  // Copy enough bits to reach a source word boundary
        const int w = obits < upperbits ? obits : upperbits;
        const EData field = VL_MASK_E(w);
        iowp[0] = (iowp[0] & ~field) | ((rwp[wordoff] >> lsb) & field);
        if (w == obits) return;
        obits -= w;
    }
    _vl_insert_WW(iowp, rwp + (wordoff + (lsb != 0)), upperbits + obits - 1, upperbits, rbits);
}

//=========================================================================
// Conditional (verilated_funcs.h:3464-3468): a macro so only one arm evaluates

#define VL_COND_WIWW(obits, owp, cond, w1p, w2p) \
    VL_MEMCPY_W(owp, (cond) ? static_cast<WDataInP>(w1p) : static_cast<WDataInP>(w2p), \
                VL_WORDS_I(obits))

//=========================================================================
// Wide constants (verilated_funcs.h:3470-3597).  Words are given most
// significant first; all destination words above those given are zeroed.

VLCU_HD WDataOutP vlcu_const_finish(int obits, WDataOutP o, int wordsSet) {
    for (int i = wordsSet; i < VL_WORDS_I(obits); ++i) o[i] = 0;
    return o;
}
VLCU_HD WDataOutP VL_CONST_W_1X(int obits, WDataOutP o, EData d0) {
    o[0] = d0;
    return vlcu_const_finish(obits, o, 1);
}
VLCU_HD WDataOutP VL_CONST_W_2X(int obits, WDataOutP o, EData d1, EData d0) {
    o[0] = d0, o[1] = d1;
    return vlcu_const_finish(obits, o, 2);
}
VLCU_HD WDataOutP VL_CONST_W_3X(int obits, WDataOutP o, EData d2, EData d1, EData d0) {
    o[0] = d0, o[1] = d1, o[2] = d2;
    return vlcu_const_finish(obits, o, 3);
}
VLCU_HD WDataOutP VL_CONST_W_4X(int obits, WDataOutP o, EData d3, EData d2, EData d1, EData d0) {
    o[0] = d0, o[1] = d1, o[2] = d2, o[3] = d3;
    return vlcu_const_finish(obits, o, 4);
}
VLCU_HD WDataOutP VL_CONST_W_5X(int obits, WDataOutP o, EData d4, EData d3, EData d2, EData d1,
                                EData d0) {
    o[0] = d0, o[1] = d1, o[2] = d2, o[3] = d3, o[4] = d4;
    return vlcu_const_finish(obits, o, 5);
}
VLCU_HD WDataOutP VL_CONST_W_6X(int obits, WDataOutP o, EData d5, EData d4, EData d3, EData d2,
                                EData d1, EData d0) {
    o[0] = d0, o[1] = d1, o[2] = d2, o[3] = d3, o[4] = d4, o[5] = d5;
    return vlcu_const_finish(obits, o, 6);
}
VLCU_HD WDataOutP VL_CONST_W_7X(int obits, WDataOutP o, EData d6, EData d5, EData d4, EData d3,
                                EData d2, EData d1, EData d0) {
    o[0] = d0, o[1] = d1, o[2] = d2, o[3] = d3, o[4] = d4, o[5] = d5, o[6] = d6;
    return vlcu_const_finish(obits, o, 7);
}
VLCU_HD WDataOutP VL_CONST_W_8X(int obits, WDataOutP o, EData d7, EData d6, EData d5, EData d4,
                                EData d3, EData d2, EData d1, EData d0) {
    o[0] = d0, o[1] = d1, o[2] = d2, o[3] = d3, o[4] = d4, o[5] = d5, o[6] = d6, o[7] = d7;
    return vlcu_const_finish(obits, o, 8);
}
// High part: words start at word VL_WORDS_I(lsb)
VLCU_HD WDataOutP VL_CONSTHI_W_1X(int obits, int lsb, WDataOutP o, EData d0) {
    const int b = VL_WORDS_I(lsb);
    o[b] = d0;
    return vlcu_const_finish(obits, o, b + 1);
}
VLCU_HD WDataOutP VL_CONSTHI_W_2X(int obits, int lsb, WDataOutP o, EData d1, EData d0) {
    const int b = VL_WORDS_I(lsb);
    o[b] = d0, o[b + 1] = d1;
    return vlcu_const_finish(obits, o, b + 2);
}
VLCU_HD WDataOutP VL_CONSTHI_W_3X(int obits, int lsb, WDataOutP o, EData d2, EData d1, EData d0) {
    const int b = VL_WORDS_I(lsb);
    o[b] = d0, o[b + 1] = d1, o[b + 2] = d2;
    return vlcu_const_finish(obits, o, b + 3);
}
VLCU_HD WDataOutP VL_CONSTHI_W_4X(int obits, int lsb, WDataOutP o, EData d3, EData d2, EData d1,
                                  EData d0) {
    const int b = VL_WORDS_I(lsb);
    o[b] = d0, o[b + 1] = d1, o[b + 2] = d2, o[b + 3] = d3;
    return vlcu_const_finish(obits, o, b + 4);
}
VLCU_HD WDataOutP VL_CONSTHI_W_5X(int obits, int lsb, WDataOutP o, EData d4, EData d3, EData d2,
                                  EData d1, EData d0) {
    const int b = VL_WORDS_I(lsb);
    o[b] = d0, o[b + 1] = d1, o[b + 2] = d2, o[b + 3] = d3, o[b + 4] = d4;
    return vlcu_const_finish(obits, o, b + 5);
}
VLCU_HD WDataOutP VL_CONSTHI_W_6X(int obits, int lsb, WDataOutP o, EData d5, EData d4, EData d3,
                                  EData d2, EData d1, EData d0) {
    const int b = VL_WORDS_I(lsb);
    o[b] = d0, o[b + 1] = d1, o[b + 2] = d2, o[b + 3] = d3, o[b + 4] = d4, o[b + 5] = d5;
    return vlcu_const_finish(obits, o, b + 6);
}
VLCU_HD WDataOutP VL_CONSTHI_W_7X(int obits, int lsb, WDataOutP o, EData d6, EData d5, EData d4,
                                  EData d3, EData d2, EData d1, EData d0) {
    const int b = VL_WORDS_I(lsb);
    o[b] = d0, o[b + 1] = d1, o[b + 2] = d2, o[b + 3] = d3, o[b + 4] = d4, o[b + 5] = d5;
    o[b + 6] = d6;
    return vlcu_const_finish(obits, o, b + 7);
}
VLCU_HD WDataOutP VL_CONSTHI_W_8X(int obits, int lsb, WDataOutP o, EData d7, EData d6, EData d5,
                                  EData d4, EData d3, EData d2, EData d1, EData d0) {
    const int b = VL_WORDS_I(lsb);
    o[b] = d0, o[b + 1] = d1, o[b + 2] = d2, o[b + 3] = d3, o[b + 4] = d4, o[b + 5] = d5;
    o[b + 6] = d6, o[b + 7] = d7;
    return vlcu_const_finish(obits, o, b + 8);
}
// Low part of a constant wider than 8 words: no zero fill
VLCU_HD void VL_CONSTLO_W_8X(int lsb, WDataOutP obase, EData d7, EData d6, EData d5, EData d4,
                             EData d3, EData d2, EData d1, EData d0) {
    WDataOutP o = obase + VL_WORDS_I(lsb);
    o[0] = d0, o[1] = d1, o[2] = d2, o[3] = d3, o[4] = d4, o[5] = d5, o[6] = d6, o[7] = d7;
}

//=========================================================================
// Masked match for casez tables (verilated_funcs.h:271-302)

VLCU_HD IData VL_MATCHMASKED_I(int, IData lhs, WDataInP matchp) {
    IData i = 0;
    while ((matchp[i * 2] & lhs) != matchp[i * 2 + 1]) ++i;
    return i;
}
VLCU_HD IData VL_MATCHMASKED_Q(int, QData lhs, WDataInP matchp) {
    IData i = 0;
    while ((VL_SET_QW(matchp + i * 4) & lhs) != VL_SET_QW(matchp + i * 4 + 2)) ++i;
    return i;
}
VLCU_HD IData VL_MATCHMASKED_W(int lbits, WDataInP lhsp, WDataInP matchp) {
    const int iwords = VL_WORDS_I(lbits);
    IData i = 0;
    while (true) {
        WDataInP maskp = matchp + (i * iwords * 2);
        WDataInP bitsp = maskp + iwords;
        EData diff = 0;
        for (int j = 0; j < iwords; ++j) diff |= (maskp[j] & lhsp[j]) ^ bitsp[j];
        if (!diff) return i;
        ++i;
    }
}

#endif  // guard
