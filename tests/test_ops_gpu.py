"""B200 differential test of the device operator library (include/verilated_cuda_funcs.h):
one test vector per CUDA thread (tests/optest/optest.cu) against the ORACLE, the
reference's own include/verilated_funcs.h compiled into oracle/_ref/libvlref.so.
Bit-exact, dirty bits included."""
import ctypes
import json
import os

import numpy as np
import pytest

import optest_util as U
import oracle_util as O

pytestmark = pytest.mark.gpu
N_VECTORS = 20000


@pytest.fixture(scope="module")
def libs():
    return U.load_ref()[1], U.load_cuda()[1]


@pytest.mark.parametrize("op", U.OPS)
def test_operator_matches_reference_on_gpu(libs, op):
    ref_fn, cuda_fn = libs
    vecs = U.gen(op, N_VECTORS, seed=977)
    want = U.run(ref_fn, op, vecs)
    got = U.run(cuda_fn, op, vecs)
    assert U.first_mismatch(op, want, got) is None, U.first_mismatch(op, want, got)


RAND_REQ = np.dtype([("seed", "<u8"), ("scopeHash", "<u8"), ("salt", "<u8"), ("obits", "<i4"),
                     ("randReset", "<i4"), ("kind", "<u4"), ("pad", "<u4"), ("out", "<u4", 16),
                     ("key", "S64"), ("keyHash", "<u8")])


def _rand_fn():
    lib, _ = U.load_cuda()
    fn = lib.vlcu_optest_rand_cuda
    fn.argtypes = [ctypes.c_void_p, ctypes.c_uint64]
    fn.restype = ctypes.c_int
    return fn


def test_x_initial_known_answers_on_gpu():
    """t_x_rand_stability.out values (reference regression golden) from the device PRNG."""
    gold = json.load(open(os.path.join(U.ROOT, "tests", "golden", "x_rand_stability.json")))
    reqs = np.zeros(len(gold["cases"]), dtype=RAND_REQ)
    for r, c in zip(reqs, gold["cases"]):
        r["seed"] = gold["seed"]
        r["randReset"] = gold["rand_reset"]
        r["scopeHash"] = O.murmur64(c["scope"])
        r["salt"] = O.murmur64(c["salt"])
        r["obits"] = c["bits"]
        r["kind"] = "IQW".index(c["kind"])
        r["key"] = c["scope"].encode()
    assert _rand_fn()(reqs.ctypes.data, len(reqs)) == 0
    for r, c in zip(reqs, gold["cases"]):
        words = (c["bits"] + 31) // 32
        got = sum(int(w) << (32 * i) for i, w in enumerate(r["out"][:words]))
        assert got == int(c["hex"], 16), c
        assert int(r["keyHash"]) == O.murmur64(c["scope"])


def test_scoped_rand_reset_random_on_gpu():
    rng = np.random.default_rng(5)
    n = 4096
    reqs = np.zeros(n, dtype=RAND_REQ)
    reqs["seed"] = rng.integers(0, 1 << 63, n, dtype=np.uint64)
    reqs["scopeHash"] = rng.integers(0, 1 << 63, n, dtype=np.uint64)
    reqs["salt"] = rng.integers(0, 1 << 63, n, dtype=np.uint64)
    reqs["randReset"] = rng.integers(0, 3, n)
    reqs["kind"] = 2
    reqs["obits"] = rng.integers(65, 512, n)
    want = [O.rand_reset_w(int(r["obits"]), int(r["scopeHash"]), int(r["salt"]),
                           int(r["seed"]) & 0x7fffffff, int(r["randReset"])) for r in reqs]
    reqs["seed"] &= np.uint64(0x7fffffff)  # VerilatedContext::randSeed is an int
    assert _rand_fn()(reqs.ctypes.data, n) == 0
    for r, w in zip(reqs, want):
        assert (r["out"][:len(w)] == w).all()
