// TEST/BUILD GLUE: calls V3EmitCuda::emitCuda() from an unmodified Verilator.cpp.
// A maintainer adds one line after V3EmitC::emitcImp() (src/Verilator.cpp:662); this build
// cannot edit the reference, so the same call is injected with `ld --wrap` on that symbol.
#include "V3EmitCuda.h"
#include <cstdlib>
#include <cstring>
extern "C" void __real__ZN7V3EmitC8emitcImpEv();
extern "C" void __wrap__ZN7V3EmitC8emitcImpEv() {
    __real__ZN7V3EmitC8emitcImpEv();
    const char* const envp = std::getenv("VERILATOR_CUDA");
    if (!envp || std::strcmp(envp, "0") != 0) V3EmitCuda::emitCuda();
}
