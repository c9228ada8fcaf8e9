// TEST INFRASTRUCTURE (not product code).
// One operator test vector, shared by the reference-side and device-side
// dispatchers (oracle/ref_ops.cpp, tests/optest/optest.cu) and mirrored by
// the numpy dtype in tests/optest_util.py.
#ifndef VLCU_OPTEST_VEC_H_
#define VLCU_OPTEST_VEC_H_
#include <stdint.h>
#define OPVEC_NP 8
#define OPVEC_IN_WORDS 16
#define OPVEC_OUT_WORDS 160
struct OpVec {
    int32_t p[OPVEC_NP];  // widths / word counts / bit positions
    uint32_t a[OPVEC_IN_WORDS];  // left operand limbs
    uint32_t b[OPVEC_IN_WORDS];  // right operand limbs
    uint32_t c[OPVEC_IN_WORDS];  // third operand limbs
    uint32_t out[OPVEC_OUT_WORDS];  // result limbs (preloaded for read-modify-write ops)
};
#endif
