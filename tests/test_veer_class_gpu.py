"""The VeeR-class design (BASELINE configs 3-5 class: ~37,000 bits of scalar flops in 34 scopes of
9 module classes, non-inlined hierarchy, tiled model image in HBM) against the genuine Verilated
C++ model of the same source: every output port after every eval."""
import os

import numpy as np
import pytest

import oracle_util as O
from verilator_b200 import abi

pytestmark = pytest.mark.gpu

DESIGN = "veer_class"
SPARSE = 0b11100000  # irq_src[63:0] and ctl: non-zero in 1 period of 32
VC = dict(tb_flags=abi.FUZZ_RESET_LOW, tb_arg=(4, SPARSE, 0, 0))
VC_O = dict(tb_flags=O.FUZZ_RESET_LOW, reset_periods=4, sparse_mask=SPARSE)


def test_design_is_tiled_and_hierarchical():
    with abi.Model(DESIGN, 40, seed=3, rand_reset=2) as m:
        assert abi.lib.vlcu_tile(m.h) == 32
        assert (m.in_words, m.out_words) == (8, 13)
        scopes = {name.rsplit(".", 1)[0] for name in m.members if "." in name}
        assert len(scopes) == 32  # 4 x (hart, btb, csr, lsu, muldiv, pic, timer, bridge) + the root
        assert m.state_bytes > 8000  # bytes of persistent state per instance
        scalar_bits = sum(w for (_, _, w, _, elems, _) in m.members.values() if elems * 8 <= 64 or w > 64)
        assert scalar_bits >= 20000


@pytest.mark.parametrize("n,evals,rand_reset", [(200, 400, 2), (33, 300, 0), (1, 200, 1)])
def test_ports_every_eval_fuzz(n, evals, rand_reset):
    want_obs, want_sigs = O.generic_fuzz(DESIGN, n, evals, inst_base=5, seed=9, rand_reset=rand_reset, tb_seed=21,
                                         obs_evals=evals, threads=8, **VC_O)
    with abi.Model(DESIGN, n, inst_base=5, seed=9, rand_reset=rand_reset) as m:
        obs, _ = m.run("fuzz", evals, obs_evals=evals, tb_seed=21, **VC)
        bad = np.argwhere(obs != want_obs)
        assert bad.size == 0, f"first mismatch (eval, word, instance) = {bad[0]}"
        assert (m.signatures() == want_sigs).all()
        assert len(np.unique(want_sigs)) == n


def test_ports_every_eval_stimulus_rows():
    """The same through the "trace" testbench: stimulus rows in, observation rows out."""
    n, evals = 96, 256
    iw, _ = O.generic_words(DESIGN)
    rows = O.fuzz_rows(n, evals, iw, seed=4, tb_seed=8, **VC_O)
    want_obs, want_sigs = O.generic_run(DESIGN, rows, seed=4, rand_reset=2, threads=8)
    with abi.Model(DESIGN, n, seed=4, rand_reset=2) as m:
        obs, _ = m.run("trace", evals, stim=rows, obs_evals=evals)
        assert (obs == want_obs).all() and (m.signatures() == want_sigs).all()
        # ... and in two launches (state, context and signature carried in HBM)
        m.reset()
        o1, _ = m.run("trace", 100, stim=rows[:100], obs_evals=100)
        o2, _ = m.run("trace", evals - 100, stim=rows[100:], obs_evals=evals - 100)
        assert (np.concatenate([o1, o2]) == want_obs).all() and (m.signatures() == want_sigs).all()


def test_uniform_stimulus_diverges_by_x_initial_only():
    """BASELINE config 3: one stimulus for all instances, per-instance --x-initial unique seeds."""
    n, evals = 64, 500
    flags = abi.FUZZ_RESET_LOW | abi.FUZZ_UNIFORM
    _, want = O.generic_fuzz(DESIGN, n, evals, seed=2, rand_reset=2, tb_seed=6, tb_flags=flags, reset_periods=4,
                             sparse_mask=SPARSE, threads=8)
    with abi.Model(DESIGN, n, seed=2, rand_reset=2) as m:
        m.run("fuzz", evals, tb_seed=6, tb_flags=flags, tb_arg=VC["tb_arg"])
        assert (m.signatures() == want).all() and len(np.unique(want)) == n
    with abi.Model(DESIGN, n, seed=2, rand_reset=0) as m:
        m.run("fuzz", evals, tb_seed=6, tb_flags=flags, tb_arg=VC["tb_arg"])
        assert len(np.unique(m.signatures())) == 1


def test_members_of_every_scope_round_trip():
    """set_member / get_member / snapshot on the tiled layout: scalars, wide values, arrays."""
    n = 70
    rng = np.random.default_rng(12)
    with abi.Model(DESIGN, n, seed=1, rand_reset=2) as m:
        before = {i: m.snapshot(i) for i in (0, 31, 32, 69)}
        picks = [name for name in m.members if name.endswith((".__PVT__pc_f", ".__PVT__l2_data", ".__PVT__dccm", ".__PVT__rf", ".__PVT__mcycle", ".__PVT__sb0_be"))]
        assert len(picks) >= 6 * 4
        written = {}
        for name in picks[::3]:
            _, _, width, eb, elems, _ = m.members[name]
            vals = rng.integers(0, 1 << 32, size=(n, elems), dtype=np.uint64).astype(abi._ELEM_DTYPE[eb])
            if elems * eb * 8 <= 64 and width < eb * 8:
                vals &= (1 << width) - 1
            m.set_member(name, vals)
            written[name] = vals
        for name, vals in written.items():
            got = m.get_member(name)
            assert (got.reshape(n, -1) == vals).all(), name
        for i, rec in before.items():
            now = m.snapshot(i)
            changed = np.zeros(rec.size, dtype=bool)
            for name, vals in written.items():
                _, _, _, eb, elems, off = m.members[name]
                changed[off:off + eb * elems] = True
                assert (m.record_member(now, name) == vals[i]).all()
            assert (now[~changed] == rec[~changed]).all()
        # a broadcast write reaches every instance
        name = picks[1]
        _, _, _, eb, elems, _ = m.members[name]
        one = rng.integers(0, 200, size=elems).astype(abi._ELEM_DTYPE[eb])
        m.set_member(name, one)
        assert (m.get_member(name).reshape(n, -1) == one).all()


def test_shards_are_slices_of_one_run():
    n, evals = 160, 200
    with abi.Model(DESIGN, n, seed=5, rand_reset=2) as m:
        m.run("fuzz", evals, tb_seed=1, **VC)
        whole = m.signatures()
    parts = []
    for base, cnt in ((0, 50), (50, 110)):
        with abi.Model(DESIGN, cnt, inst_base=base, seed=5, rand_reset=2) as m:
            m.run("fuzz", evals, tb_seed=1, **VC)
            parts.append(m.signatures())
    assert (np.concatenate(parts) == whole).all()


def test_config3_size_sampled_against_the_oracle():
    """16,384 instances (BASELINE config 3): 48 of them checked against the genuine model."""
    n, evals = 16384, 128
    with abi.Model(DESIGN, n, seed=1, rand_reset=2) as m:
        m.run("fuzz", evals, tb_seed=3, **VC)
        sigs = m.signatures()
        assert len(np.unique(sigs)) == n
    for i in list(range(0, n, 512)) + list(range(n - 16, n)):
        _, want = O.generic_fuzz(DESIGN, 1, evals, inst_base=i, seed=1, rand_reset=2, tb_seed=3, **VC_O)
        assert sigs[i] == want[0], i


VEERCOV_LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "veercov", "libvlcu_veercov.so")


def test_coverage_campaign_merged_file_equals_the_reference_writer(tmp_path):
    """BASELINE config 5 on one GPU: the --coverage build of the VeeR-class design (line + toggle
    points of every scope, ~80,000 counters; 16-entry branch target buffers, see oracle/Makefile),
    a fuzz campaign, the counters -- one per point per TILE, the warp adds up its lanes -- reduced
    on the device, and ONE merged coverage.dat: byte for byte the file the reference's
    VerilatedCov writes for the same instances."""
    if not os.path.exists(VEERCOV_LIB):
        pytest.skip("tests/veercov/libvlcu_veercov.so not built")
    abi.load_design_library(VEERCOV_LIB)
    n, evals = 40, 260
    ref_path, gpu_path = str(tmp_path / "ref.dat"), str(tmp_path / "gpu.dat")
    want_sigs = O.generic_cov("veer_class_cov", n, evals, ref_path, inst_base=9, seed=3, rand_reset=2, tb_seed=12, **VC_O)
    with abi.Model("veer_class_cov", n, inst_base=9, seed=3, rand_reset=2) as m:
        assert abi.lib.vlcu_tile(m.h) == 32 and m.coverage_bins() > 50000
        m.run("fuzz", evals, tb_seed=12, **VC)
        assert (m.signatures() == want_sigs).all()
        sums = m.coverage_reduce()
        m.coverage_write(sums, gpu_path)
        assert (sums > 0).sum() > 10000
    ours, ref = open(gpu_path).read(), open(ref_path).read()
    assert ref.count("\n") > 5000 and "v_toggle" in ref and "v_line" in ref
    assert ours == ref
