// TEST INFRASTRUCTURE (not product code) -- the operator-level ORACLE.
//
// Thin extern "C" wrappers around the REFERENCE's own runtime, compiled from
// the sources where they lie under /root/reference/include (see
// oracle/Makefile); nothing from the reference is copied into this repo.
// Built into oracle/_ref/libvlref.so, which travels to the GPU box.
//
//   vlref_optest_run      every VL_* operator of include/verilated_funcs.h /
//                         verilated.cpp listed in optest_ops.inc
//   vlref_murmur64        VL_MURMUR64_HASH          (verilated.cpp:2173)
//   vlref_rand_reset_*    VL_SCOPED_RAND_RESET_I/Q/W (verilated.cpp:577-615)
//   vlref_rng_*           VlRNG(seed).rand64()       (verilated.cpp:480-487)
#include "verilated.h"
#include "verilated_cov.h"

#include "optest_vec.h"

#define WO(x) WDataOutP::external(x)
#define WI(x) WDataInP::external(x)

enum {
#define OP(name, body) OPID_##name,
#include "optest_ops.inc"
#undef OP
    OPID__COUNT
};

static void run_one(int op, OpVec* v) {
    switch (op) {
#define OP(name, body) \
    case OPID_##name: { \
        body; \
        break; \
    }
#include "optest_ops.inc"
#undef OP
    default: break;
    }
}

extern "C" {
int vlref_optest_count() { return OPID__COUNT; }
int vlref_optest_run(int op, OpVec* vecs, uint64_t n) {
    if (op < 0 || op >= OPID__COUNT) return -1;
    for (uint64_t i = 0; i < n; ++i) run_one(op, &vecs[i]);
    return 0;
}

uint64_t vlref_murmur64(const char* key) { return VL_MURMUR64_HASH(key); }

static VerilatedContext* ctx_for(uint64_t seed, int randReset) {
    static thread_local VerilatedContext* ctxp = nullptr;
    if (!ctxp) ctxp = new VerilatedContext;
    ctxp->randReset(randReset);
    ctxp->randSeed(static_cast<int>(seed));
    Verilated::threadContextp(ctxp);
    return ctxp;
}
uint32_t vlref_rand_reset_i(int obits, uint64_t scopeHash, uint64_t salt, uint64_t seed,
                            int randReset) {
    ctx_for(seed, randReset);
    return VL_SCOPED_RAND_RESET_I(obits, scopeHash, salt);
}
uint64_t vlref_rand_reset_q(int obits, uint64_t scopeHash, uint64_t salt, uint64_t seed,
                            int randReset) {
    ctx_for(seed, randReset);
    return VL_SCOPED_RAND_RESET_Q(obits, scopeHash, salt);
}
void vlref_rand_reset_w(int obits, uint32_t* outp, uint64_t scopeHash, uint64_t salt,
                        uint64_t seed, int randReset) {
    ctx_for(seed, randReset);
    VL_SCOPED_RAND_RESET_W(obits, WDataOutP::external(outp), scopeHash, salt);
}
void vlref_rng_rand64(uint64_t seed, uint64_t* outp, int n) {
    VlRNG rng{seed};
    for (int i = 0; i < n; ++i) outp[i] = rng.rand64();
}
}  // extern "C"
