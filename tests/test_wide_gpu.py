"""Parity of the emitted CUDA model of wide_bus.v (BASELINE.json config 4 workload class: a
VlWide-heavy wide-bus path) with the genuine Verilated C++ model, through the C ABI.  The design
keeps VL_ADD_W / VL_SUB_W / VL_MUL_W / VL_SHIFTL_WWI / VL_SHIFTR_WWI / VL_SHIFTRS_WWI / VL_LT_W /
VL_LTS_IWW / VL_COUNTONES_W and a variable 70-bit part-select of a 256-bit value on the hot path,
so this is the operator library exercised end to end through V3EmitCuda."""
import numpy as np
import pytest

import oracle_util as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def abi():
    from verilator_b200 import abi
    assert abi.lib.vlcu_device_count() > 0, "no CUDA device: GPU tests need a B200"
    assert "wide_bus" in abi.designs()
    return abi


@pytest.mark.parametrize("rand_reset,reset_time", [(2, 0), (0, 0), (1, 7)])
def test_ports_every_eval_match_the_genuine_model(abi, rand_reset, reset_time):
    n, evals = 150, 1000
    want = O.wide_run(n, evals, seed=21, rand_reset=rand_reset, tb_seed=99, reset_time=reset_time,
                      obs_evals=evals)
    with abi.Model("wide_bus", n, seed=21, rand_reset=rand_reset) as m:
        assert (m.out_words, m.in_words) == (19, 13)
        obs, _ = m.run("wide_rand", evals, flags=abi.RUNF_TB_BEGIN, obs_evals=evals, tb_seed=99,
                       tb_arg=(reset_time, 0, 0, 0))
        assert (obs[:, :, :n] == want.obs).all()
        assert (m.signatures() == want.sigs).all()
        assert (m.get_member("wide_bus__DOT__acc") == want.acc).all()
        # the accumulators really are wide and different per instance
        assert len(np.unique(want.acc[:, 7])) > n // 2


def test_split_runs_equal_one_run(abi):
    n = 70
    with abi.Model("wide_bus", n, seed=3, rand_reset=2) as a, abi.Model("wide_bus", n, seed=3, rand_reset=2) as b:
        a.run("wide_rand", 600, flags=abi.RUNF_TB_BEGIN, tb_seed=1)
        b.run("wide_rand", 1, flags=abi.RUNF_TB_BEGIN, tb_seed=1)
        b.run("wide_rand", 332, tb_seed=1)
        b.run("wide_rand", 267, tb_seed=1)
        assert (a.signatures() == b.signatures()).all()
        assert (a.get_member("wide_bus__DOT__hash") == b.get_member("wide_bus__DOT__hash")).all()


def test_vcd_equals_the_reference_writer(abi, tmp_path):
    n, evals, samples = 24, 200, [0, 11, 23]
    with abi.Model("wide_bus", n, seed=4, rand_reset=2) as m:
        _, trace = m.run("wide_rand", evals, flags=abi.RUNF_TB_BEGIN, tb_seed=8, samples=samples,
                         trace_evals=evals)
        for k, inst in enumerate(samples):
            ref_path, our_path = str(tmp_path / f"ref{inst}.vcd"), str(tmp_path / f"our{inst}.vcd")
            O.wide_run(n, evals, seed=4, rand_reset=2, tb_seed=8, vcd_path=ref_path, vcd_instance=inst)
            m.vcd_write(trace, k, our_path)
            ours, want = open(our_path).read(), open(ref_path).read()
            assert "$var wire 256" in want
            assert ours == want, (inst, ours[:600], want[:600])
        m.reset()
        _, streams = m.run("wide_rand", evals, flags=abi.RUNF_TB_BEGIN, tb_seed=8, samples=samples,
                           trace_evals=evals, trace_stream_words=4 + evals * (3 + 2 * 102))
        assert (streams[:, 1] == 0).all()
        for k, inst in enumerate(samples):
            chg_path = str(tmp_path / f"chg{inst}.vcd")
            m.vcd_write_changes(streams, k, chg_path)
            assert open(chg_path).read() == open(str(tmp_path / f"ref{inst}.vcd")).read(), inst


def test_full_batch_sampled_against_the_genuine_model(abi):
    n, evals = 16384, 512
    with abi.Model("wide_bus", n, seed=1, rand_reset=2) as m:
        m.run("wide_rand", evals, flags=abi.RUNF_TB_BEGIN, tb_seed=5)
        sig = m.signatures()
        assert len(np.unique(sig)) == n
        for base in (0, 8000, n - 40):
            want = O.wide_run(40, evals, inst_base=base, seed=1, rand_reset=2, tb_seed=5, threads=4)
            assert (sig[base:base + 40] == want.sigs).all()
