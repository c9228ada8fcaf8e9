// -*- mode: C++; c-file-style: "cc-mode" -*-
//*************************************************************************
// DESCRIPTION: Verilator: Emit the batched CUDA (sm_100a) model of the design
//
// New emitter, sibling of V3EmitC::emitcImp/emitcModel/emitcSyms (V3EmitC.h:25-35).
// It runs on the same AST state they do (after V3Sched/V3Order/... at
// Verilator.cpp:647) and writes <prefix>__cuda.cu: the design as device code for
// include/verilated_cuda*.h, one model instance per CUDA thread.
//
// Hooking it into verilator_bin is one call beside V3EmitC::emitcFiles()
// (Verilator.cpp:679); see INTEGRATION.md.
//*************************************************************************

#ifndef VERILATOR_V3EMITCUDA_H_
#define VERILATOR_V3EMITCUDA_H_

#include "config_build.h"
#include "verilatedos.h"

class V3EmitCuda final {
public:
    static void emitCuda() VL_MT_DISABLED;
};

#endif  // Guard
