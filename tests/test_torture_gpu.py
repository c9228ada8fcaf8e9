"""Breadth of the emitter + operator library: seventeen hand-written designs and four random ones, each a different corner of
what V3Sched/V3Order/V3Expand/V3Table/V3Delayed/V3Unroll produce (scalar signed/unsigned
mul/div/mod and shifts, case FSM + constant-pool table + casez, byte-enable memories with
run-time indices, two clocks + negedge + asynchronous reset, 96/130-bit multiply/divide/shift,
unrolled loops, data-dependent while/for loops + $countones/$onehot/power, arrays of wide and
64-bit elements with two write ports, multidimensional arrays in every storage class, a 4 KB
state array streamed by a loop, NBA commit queues, $urandom / $error / $fatal / inout / initial
ROM / final, derived and gated clocks + latch + signed wide arithmetic, a parameterised module hierarchy, functions and tasks, generate regions, a byte table too big for shared memory; tests/gen_random_design.py's expression
soups), taken through the reference's own compiler and V3EmitCuda with no hand-written
code per design, and compared bit for bit with the genuine Verilated C++ model under random
stimulus on every port at every eval."""
import zlib

import numpy as np
import pytest

import oracle_util as O

pytestmark = pytest.mark.gpu

DESIGNS = ["t01_alu", "t02_fsm", "t03_mem", "t04_clocks", "t05_wide", "t06_loops", "t07_dyn", "t08_widemem", "t09_mdarray", "t10_bigstate", "t11_nbaq", "t12_sys", "t13_mix", "t14_hier", "t15_func", "t16_gen", "t17_lut",
           "rnd01", "rnd02", "rnd03", "rnd04"]


@pytest.fixture(scope="module")
def abi():
    from verilator_b200 import abi
    assert abi.lib.vlcu_device_count() > 0, "no CUDA device: GPU tests need a B200"
    return abi


def random_stimulus(m, abi, n, evals, rng, hold=None):
    """Rows [eval][in_word][npad]: `clk` toggles every eval, every other input word is random and
    changes on falling edges only (except designs that want glitchy inputs); biased values
    (0, all ones, small) are mixed in so corner cases (division by zero, MIN / -1, shift >= width)
    occur."""
    stim = np.zeros((evals, m.in_words, m.npad), dtype=np.uint32)
    word = 0
    names = []
    for name, (p, direction, width, eb, elems) in sorted(m.ports.items(), key=lambda kv: kv[1][0]):
        if direction not in (abi.DIR_IN, abi.DIR_INOUT):
            continue
        words = elems if elems > 1 else (2 if eb == 8 else 1)
        for w in range(words):
            names.append((name, w))
        word += words
    assert word == m.in_words
    for k, (name, w) in enumerate(names):
        if name == "clk":
            stim[:, k, :] = (np.arange(evals) % 2 == 0)[:, None]  # 1, 0, 1, 0, ...
            continue
        if name == "clk2":
            stim[:, k, :] = ((np.arange(evals) // 3) % 2)[:, None]
            continue
        vals = rng.integers(0, 1 << 32, size=(evals, n), dtype=np.uint64).astype(np.uint32)
        kind = rng.integers(0, 8, size=(evals, n))
        vals = np.where(kind == 0, 0, vals)
        vals = np.where(kind == 1, 0xffffffff, vals)
        vals = np.where(kind == 2, vals & 0x3f, vals)
        vals = np.where(kind == 3, 0x80000000, vals)
        if name in ("arst_n", "rst"):
            vals = (rng.integers(0, 16, size=(evals, n)) != 0).astype(np.uint32) if name == "arst_n" else \
                (rng.integers(0, 32, size=(evals, n)) == 0).astype(np.uint32)
            vals[:4] = 0 if name == "arst_n" else 1
        # inputs change on the evals where clk falls (odd evals) and are held over the next rising
        # edge, like sim_main.cpp pokes its inputs
        vals[2::2] = vals[1:-1:2] if evals % 2 == 0 else vals[1::2]
        stim[:, k, :n] = vals
    return stim


# $urandom draws from a thread-local generator that the host runtime reseeds from the thread's
# context whenever ANY context sets its seed (a process-wide epoch, verilated.cpp:537-549): with one
# model at a time per thread that is one fresh stream per model, which is what the device does; with
# several harness threads the epoch bumps of the others would restart a model's stream mid-run
ORACLE_THREADS = {"t12_sys": 1}

# CUDA-only variants: (registered design, the design whose genuine model is the reference)
VARIANTS = {"t10_bigstate_im": "t10_bigstate"}  # --no-cuda-array-layout: instance-major array


@pytest.mark.parametrize("design", DESIGNS + sorted(VARIANTS))
@pytest.mark.parametrize("rand_reset", [2, 0])
def test_every_port_every_eval_matches_the_genuine_model(abi, design, rand_reset):
    n, evals = 96, 600
    ref_design = VARIANTS.get(design, design)
    rng = np.random.default_rng(zlib.crc32(ref_design.encode()))  # (str hashes are salted per process)
    with abi.Model(design, n, seed=1000, rand_reset=rand_reset) as m:
        stim = random_stimulus(m, abi, n, evals, rng)
        want_obs, want_sigs = O.generic_run(ref_design, stim[:, :, :n], seed=1000, rand_reset=rand_reset,
                                            threads=ORACLE_THREADS.get(ref_design, 4))
        assert (m.in_words, m.out_words) == O.generic_words(ref_design)
        obs, _ = m.run("trace", evals, stim=stim, obs_evals=evals)
        bad = np.argwhere(obs[:, :, :n] != want_obs)
        assert len(bad) == 0, (design, bad[:5].tolist())
        assert (m.signatures() == want_sigs).all()
        # the stimulus really exercised the design: outputs vary over time and across instances
        assert len(np.unique(want_sigs)) > n // 2
