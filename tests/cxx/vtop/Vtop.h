// TEST INFRASTRUCTURE: the "Vtop.h" examples/make_tracing_c/sim_main.cpp includes, bound to the
// batched CUDA evaluator (one instance) through include/verilated_cuda_vtop.h
#define VLCU_VTOP Vtop
#define VLCU_VTOP_DESIGN "make_tracing_c"
#define VLCU_VTOP_PORTS "Vtop_nocov__ports.inc"
#include "verilated_cuda_vtop.h"
