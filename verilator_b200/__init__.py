"""verilator_b200 -- B200-native batched evaluator for Verilated models.

The product is ``libverilated_cuda.so`` (hand-written CUDA for sm_100a + C++ host
runtime behind the C ABI of include/verilated_cuda_abi.h).  Python here is test /
bench glue only: :mod:`verilator_b200.abi` is a ctypes binding of that ABI.
"""
__all__ = ["abi"]
