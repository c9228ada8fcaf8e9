# -*- Makefile -*-
######################################################################
# DESCRIPTION: Makefile commands for the batched CUDA (sm_100a) model of a Verilated design
#
# Included by the <prefix>__cuda.mk that `verilator --cuda` writes (V3EmitCuda::emitMk), the way
# <prefix>.mk includes verilated.mk (include/verilated.mk.in:142-230) for the C++ model.
#
# Inputs (set by <prefix>__cuda.mk, overridable on the make command line):
#   VM_PREFIX         --prefix of the design; <prefix>__cuda.cu is the device model
#   VLCU_ROOT         kit root: include/verilated_cuda*.h and the runtime verilated_cuda.cu
#   VLCU_DESIGN       name the design registers under (vlcu_create(design, ...))
#   VLCU_TESTBENCH    optional header with the design's device testbenches
# Output: lib<prefix>__cuda.so = device model + batch runtime behind verilated_cuda_abi.h
######################################################################

NVCC ?= nvcc
# Blackwell B200 only: no fallback architectures, no PTX JIT
VLCU_ARCH ?= -gencode arch=compute_100a,code=sm_100a
VLCU_OPT ?= -O3
VLCU_NVCCFLAGS ?= -std=c++17 $(VLCU_OPT) -lineinfo -Xcompiler -fPIC -diag-suppress 177,186,550
VLCU_INCLUDE ?= $(VLCU_ROOT)/include
# the runtime source: beside the headers in an installed kit, under verilator_b200/csrc in the
# development tree
VLCU_RUNTIME ?= $(firstword $(wildcard $(VLCU_INCLUDE)/verilated_cuda.cu \
                                        $(VLCU_ROOT)/verilator_b200/csrc/verilated_cuda.cu))
VLCU_LDLIBS ?= -lcudart -ldl

VLCU_DEFS := -DVLCU_DESIGN_NAME='"$(VLCU_DESIGN)"'
ifneq ($(VLCU_TESTBENCH),)
  VLCU_DEFS += -DVLCU_TESTBENCH_INCLUDE='"$(notdir $(VLCU_TESTBENCH))"' -I$(dir $(abspath $(VLCU_TESTBENCH)))
endif

VLCU_LIB := lib$(VM_PREFIX)__cuda.so

default: $(VLCU_LIB)

$(VM_PREFIX)__cuda.o: $(VM_PREFIX)__cuda.cu $(VLCU_TESTBENCH) $(wildcard $(VLCU_INCLUDE)/verilated_cuda*.h)
	$(NVCC) $(VLCU_ARCH) $(VLCU_NVCCFLAGS) -I$(VLCU_INCLUDE) -I. $(VLCU_DEFS) -c $< -o $@

verilated_cuda.o: $(VLCU_RUNTIME) $(wildcard $(VLCU_INCLUDE)/verilated_cuda*.h)
	$(NVCC) $(VLCU_ARCH) $(VLCU_NVCCFLAGS) -I$(VLCU_INCLUDE) -c $< -o $@

$(VLCU_LIB): $(VM_PREFIX)__cuda.o verilated_cuda.o
	$(NVCC) $(VLCU_ARCH) -shared -o $@ $^ $(VLCU_LDLIBS)

clean::
	rm -f $(VM_PREFIX)__cuda.o verilated_cuda.o $(VLCU_LIB)

.PHONY: default clean
