"""Per-warp coverage primitives (include/verilated_cuda.h, VlcuTileCovPtr forms) against a scalar
model of verilated_cov.h:78-105: toggle points of every width class, line points and hit-mask
flushes, with full and partial warps, and guard words around the counters."""
import ctypes
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "optest", "libvlcu_covtest.so")


class CovReq(ctypes.Structure):
    _fields_ = [("mode", ctypes.c_uint32), ("width", ctypes.c_uint32), ("lanes", ctypes.c_uint32),
                ("rounds", ctypes.c_uint32)]


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        pytest.fail("tests/optest/libvlcu_covtest.so not built (run __graft_entry__.build())")
    return ctypes.CDLL(LIB)


def model(mode, width, lanes, new, old):
    """counters[...] as the host macros count them, summed over the lanes"""
    want = np.zeros(512, dtype=np.uint64)
    rounds = new.shape[0]
    for r in range(rounds):
        for lane in range(lanes):
            if mode == 6:
                n0, o0 = int(new[r, 0, lane]), int(old[r, 0, lane])
                base, w6 = (64, 32) if lane & 1 else (0, width)
                for b in range(w6):
                    if ((n0 ^ o0) >> b) & 1:
                        want[base + 2 * b + ((n0 >> b) & 1)] += 1
            elif mode in (0, 1, 2, 3):
                words = (width + 31) // 32
                n = [int(new[r, w, lane]) for w in range(4)]
                o = [int(old[r, w, lane]) for w in range(4)]
                if mode == 3:
                    if width & 31:
                        n[words - 1] &= (1 << (width & 31)) - 1
                        o[words - 1] &= (1 << (width & 31)) - 1
                    # VL_COV_TOGGLE_CHG_ST_W with its quirk (verilated_cov.h:94-105 as built by the
                    # reference's flags): changed bit b of word w bumps bits w*32+b, +32, ... < width
                    for w in range(words):
                        chg = n[w] ^ o[w]
                        for b in range(32):
                            if (chg >> b) & 1:
                                bit = (n[w] >> b) & 1
                                j = b
                                while j < width - w * 32:
                                    want[(w * 32 + j) * 2 + bit] += 1
                                    j += 32
                else:
                    nv = n[0] | (n[1] << 32 if mode == 2 else 0)
                    ov = o[0] | (o[1] << 32 if mode == 2 else 0)
                    for b in range(width):
                        if ((nv ^ ov) >> b) & 1:
                            want[2 * b + ((nv >> b) & 1)] += 1
            else:  # line points / hit masks: bin b once per set bit
                for b in range(32):
                    if (int(new[r, 0, lane]) >> b) & 1:
                        want[b] += 1
    return want


CASES = [(0, 1), (0, 5), (0, 16), (0, 32), (1, 1), (1, 2), (1, 7), (1, 16), (1, 17), (1, 31), (1, 32),
         (2, 33), (2, 48), (2, 64), (3, 65), (3, 100), (3, 128), (4, 32), (5, 32), (6, 1), (6, 3), (6, 4)]


@pytest.mark.parametrize("mode,width", CASES)
@pytest.mark.parametrize("lanes", [32, 8, 1, 31])
@pytest.mark.parametrize("density", [0.5, 0.03])
def test_per_warp_counters_equal_the_scalar_model(lib, mode, width, lanes, density):
    rounds = 24
    rng = np.random.default_rng(1000 * mode + 10 * width + lanes)
    old = rng.integers(0, 1 << 32, size=(rounds, 4, 32), dtype=np.uint64).astype(np.uint32)
    flip = (rng.random((rounds, 4, 32, 32)) < density)
    flipw = (flip * (1 << np.arange(32, dtype=np.uint64))).sum(axis=3).astype(np.uint32)
    new = old ^ flipw
    if mode >= 4:
        new = flipw
    total = 2048  # [512 guard][512 counters][1024 guard]
    counters = np.full(total, 0x5a5a5a5a00000000, dtype=np.uint64)
    counters[512:1024] = rng.integers(0, 1000, size=512).astype(np.uint64)
    start = counters.copy()
    req = CovReq(mode, width, lanes, rounds)
    rc = lib.vlcu_covtest_run(ctypes.byref(req), new.ctypes.data_as(ctypes.c_void_p), old.ctypes.data_as(ctypes.c_void_p),
                              counters.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(total))
    assert rc == 0
    assert (counters[:512] == start[:512]).all() and (counters[1024:] == start[1024:]).all(), "wrote outside the counters"
    got = counters[512:1024] - start[512:1024]
    want = model(mode, width, lanes, new, old)
    assert (got == want).all(), np.argwhere(got != want)[:8]
