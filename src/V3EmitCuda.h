// -*- mode: C++; c-file-style: "cc-mode" -*-
//*************************************************************************
// DESCRIPTION: Verilator: Emit the batched CUDA (sm_100a) model of the design
//
// New emitter, sibling of V3EmitC::emitcImp/emitcModel/emitcSyms (V3EmitC.h:25-35).
// It runs on the same AST state they do (after V3Sched/V3Order/... at
// Verilator.cpp:647) and writes <prefix>__cuda.cu: the design as device code for
// include/verilated_cuda*.h, one model instance per CUDA thread.
//
// Hooked into verilator_bin by integration/verilator-cuda.patch: the --cuda option
// (V3Options.cpp) and one call after V3EmitC::emitcImp() (Verilator.cpp:662); see INTEGRATION.md.
// Also writes <prefix>__cuda.mk (nvcc rules, cf. V3EmitMk) and <prefix>__ports.inc.
//*************************************************************************

#ifndef VERILATOR_V3EMITCUDA_H_
#define VERILATOR_V3EMITCUDA_H_

#include "config_build.h"
#include "verilatedos.h"

class V3EmitCuda final {
public:
    static void emitCuda() VL_MT_DISABLED;
    // --cuda / --no-cuda
    static void enable(bool flag) { s_enabled = flag; }
    static bool enabled() { return s_enabled; }
    // --cuda-array-layout / --no-cuda-array-layout: per-array choice between [element][instance]
    // and instance-major placement of large arrays (off: always instance-major)
    static void layoutSelection(bool flag) { s_layoutSelection = flag; }
    static bool layoutSelection() { return s_layoutSelection; }
    // --cuda-array-interleave: the index-uniformity hint.  Large arrays indexed by design signals
    // are kept [element][instance] too -- right when the instances index in lock-step (one
    // program for all, BASELINE config 3: a warp's fetch is then one 128-byte line instead of 32
    // sectors), wrong when every instance goes its own way
    static void interleaveAll(bool flag) { s_interleaveAll = flag; }
    static bool interleaveAll() { return s_interleaveAll; }
    // --cuda-stream / --no-cuda-stream: force the tiled (model image stays in HBM, module
    // hierarchy kept) or the register form (single $root class held by the thread); default: the
    // emitter chooses by hierarchy and size
    static void streamMode(bool flag) { s_streamMode = flag ? 1 : 0; }
    static int streamMode() { return s_streamMode; }
    // --cuda-leaf-gather / --no-cuda-leaf-gather (tiled form): leaf functions gather the scalar
    // members they touch into locals at entry and scatter the written ones at exit
    static void leafCaching(bool flag) { s_leafCaching = flag; }
    static bool leafCaching() { return s_leafCaching; }
    // --cuda-gather-words <n>: 32-bit words of members one leaf may hold in locals
    static void gatherWords(int words) { s_gatherWords = words; }
    static size_t gatherWords() { return static_cast<size_t>(s_gatherWords); }
    // --cuda-leaf-prefetch / --no-cuda-leaf-prefetch: ... and the caller issues the next leaf's
    // gather as prefetches one call ahead
    // --coverage counters of a register-form design: one set per instance, cov[bin][instance]
    // (on), or one uint64 set per warp of 32 instances bumped warp-wide (off, the default: the
    // product of a campaign is the sum over instances)
    static void coverPerInstance(bool flag) { s_coverPerInstance = flag; }
    static bool coverPerInstance() { return s_coverPerInstance; }
    static void leafPrefetch(bool flag) { s_leafPrefetch = flag; }
    static bool leafPrefetch() { return s_leafPrefetch; }
    // --cuda-chunk-nodes <n> (tiled form): AST nodes of leaf code inlined into one chunk function;
    // 0 (default): every big leaf is a function of its own
    static void chunkNodes(int nodes) { s_chunkNodes = nodes; }
    static size_t chunkNodes() { return static_cast<size_t>(s_chunkNodes); }
    // --cuda-share-code / --no-cuda-share-code (tiled form): print one copy of the functions of a
    // module class for all its instances where they differ only in the scopes they name
    static void shareCode(bool flag) { s_shareCode = flag; }
    static bool shareCode() { return s_shareCode; }

private:
    static bool s_enabled;
    static bool s_layoutSelection;
    static bool s_interleaveAll;
    static int s_streamMode;  // -1 automatic, 0 register form, 1 tiled form
    static bool s_leafCaching;
    static int s_gatherWords;
    static bool s_leafPrefetch;
    static bool s_coverPerInstance;
    static int s_chunkNodes;
    static bool s_shareCode;
};

#endif  // Guard
