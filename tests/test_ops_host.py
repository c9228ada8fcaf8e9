"""CPU-side differential test of the device operator library.

include/verilated_cuda_funcs.h is compiled as HOST code (tests/optest/libvlcu_optest_host.so,
debug/test build only -- the product never runs on the CPU) and every VL_* operator is compared
bit for bit, dirty bits included, with the ORACLE: the reference's own
include/verilated_funcs.h + verilated.cpp compiled into oracle/_ref/libvlref.so.
The same vectors run on the B200 in tests/test_ops_gpu.py.
"""
import numpy as np
import pytest

import optest_util as U

N_VECTORS = 3000


@pytest.fixture(scope="module")
def libs():
    return U.load_ref()[1], U.load_host()[1]


def test_op_table_is_consistent():
    ref_lib, _ = U.load_ref()
    host_lib, _ = U.load_host()
    assert ref_lib.vlref_optest_count() == len(U.OPS) == host_lib.vlcu_optest_count()
    assert len(U.OPS) > 200


@pytest.mark.parametrize("op", U.OPS)
def test_operator_matches_reference(libs, op):
    ref_fn, host_fn = libs
    vecs = U.gen(op, N_VECTORS, seed=20240229)
    want = U.run(ref_fn, op, vecs)
    got = U.run(host_fn, op, vecs)
    assert U.first_mismatch(op, want, got) is None, U.first_mismatch(op, want, got)
    # the test must actually exercise the operator
    assert (want["out"] != vecs["out"]).any() or op in ("VL_ZERO_W",), f"{op}: vectors had no effect"


def test_harness_detects_a_wrong_result(libs):
    """Sanity of the harness itself: a corrupted result is reported."""
    ref_fn, _ = libs
    vecs = U.gen("VL_ADD_W", 64, seed=1)
    want = U.run(ref_fn, "VL_ADD_W", vecs)
    got = want.copy()
    got["out"][7, 0] ^= 1
    assert U.first_mismatch("VL_ADD_W", want, got) is not None
