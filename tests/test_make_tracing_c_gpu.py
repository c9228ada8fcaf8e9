"""Parity of the batched CUDA evaluator with the oracle for examples/make_tracing_c
(BASELINE.json config 1), through the C ABI (include/verilated_cuda_abi.h).

Oracle = the genuine Verilated C++ model (written by the reference's own compiler) against the
genuine runtime, driven by oracle/models/make_tracing_c/harness.cpp; pinned by tests/test_oracle.py.  Bit-exact: every output port of every instance after every
eval, context flags, $display records, coverage counters, coverage.dat text.
"""
import json
import os

import numpy as np
import pytest

import oracle_util as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STATUS_FIELDS = ["time", "first_finish_time", "first_stop_time", "status", "error_count",
                 "finish_count", "stop_count", "fatal_count", "first_stop_line", "msg_count"]


@pytest.fixture(scope="module")
def abi():
    from verilator_b200 import abi
    assert abi.lib.vlcu_device_count() > 0, "no CUDA device: GPU tests need a B200"
    return abi


def check_status(got, want):
    for f in STATUS_FIELDS:
        g = got[f]
        if f == "status":  # the oracle's console is unbounded: it has no "log overflowed" flag
            g = g & ~np.uint32(32)
        assert (g == want[f]).all(), (f, g[:8], want[f][:8])


@pytest.mark.parametrize("design", ["make_tracing_c", "make_tracing_c_cov"])
@pytest.mark.parametrize("per_instance,period,rand_reset", [(False, 0, 2), (True, 0, 2),
                                                            (True, 128, 2), (True, 128, 0)])
def test_ports_every_eval_match_oracle(abi, design, per_instance, period, rand_reset):
    n, evals = 300, 700
    want = O.mtc_run(n, evals, seed=7, rand_reset=rand_reset, tb_seed=99, per_instance=per_instance,
                     reset_period=period, obs_evals=evals)
    with abi.Model(design, n, seed=7, rand_reset=rand_reset, msg_cap=8) as m:
        obs, _ = m.run("sim_main", evals, flags=abi.RUNF_TB_BEGIN, obs_evals=evals, tb_seed=99,
                       tb_flags=1 if per_instance else 0, tb_arg=(period, 0, 0, 0))
        assert (obs == want.obs).all()
        assert (m.signatures() == want.sigs).all()
        check_status(m.status(), want.status)
        assert (m.get_member("top__DOT__sub__DOT__count_c") == want.final[:, 0]).all()
        assert (m.get_output("out_quad") == (want.obs[-1, 1].astype(np.uint64)
                                             | (want.obs[-1, 2].astype(np.uint64) << 32))).all()


def test_golden_behavioural_vectors(abi):
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "make_tracing_c_ports.json")))["default"]
    with abi.Model("make_tracing_c", 1, rand_reset=2) as m:
        obs, _ = m.run("sim_main", gold["n_evals"], flags=abi.RUNF_TB_BEGIN, obs_evals=gold["n_evals"])
        rows = np.array(gold["rows"], dtype=np.uint64)
        assert (obs[:, :, 0] == rows[:, :6]).all()
        st = m.status()
        assert int(st["first_finish_time"][0]) == 17


def test_stop_on_finish_like_sim_main_and_console(abi):
    want = O.mtc_run(64, 1000, rand_reset=2, stop_on_finish=True, trace_plusarg=True)
    with abi.Model("make_tracing_c", 64, rand_reset=2, msg_cap=8, plusargs="+trace") as m:
        m.run("sim_main", 1000, flags=abi.RUNF_TB_BEGIN | abi.RUNF_STOP_ON_FINISH)
        st = m.status()
        check_status(st, want.status)
        assert (st["time"] == 17).all()
        msgs, count = m.messages(0)
        text = "".join(m.format_message(x) for x in msgs)
        assert count == 3 and text == want.console
        assert text.endswith("*-* All Finished *-*\n")


def test_split_runs_equal_one_run(abi):
    """State and context persist across vlcu_run calls (K-cycle fusion is transparent)."""
    n = 257
    with abi.Model("make_tracing_c", n, seed=3, rand_reset=2) as a, \
            abi.Model("make_tracing_c", n, seed=3, rand_reset=2) as b:
        a.run("sim_main", 1000, flags=abi.RUNF_TB_BEGIN, tb_seed=5, tb_flags=1, tb_arg=(128, 0, 0, 0))
        b.run("sim_main", 1, flags=abi.RUNF_TB_BEGIN, tb_seed=5, tb_flags=1, tb_arg=(128, 0, 0, 0))
        b.run("sim_main", 332, tb_seed=5, tb_flags=1, tb_arg=(128, 0, 0, 0))
        b.run("sim_main", 667, tb_seed=5, tb_flags=1, tb_arg=(128, 0, 0, 0))
        assert (a.signatures() == b.signatures()).all()
        for f in STATUS_FIELDS:
            assert (a.status()[f] == b.status()[f]).all()
        assert (a.snapshot(100) == b.snapshot(100)).all()


def test_host_driven_eval_like_a_cpp_testbench(abi):
    """Drop-in use: poke ports from the host, vlcu_eval(), read ports -- one eval at a time,
    exactly what sim_main.cpp does with Vtop, for all instances at once."""
    n, evals = 33, 60
    want = O.mtc_run(n, evals, rand_reset=2, tb_seed=11, per_instance=True, obs_evals=evals)
    rng_in = O.mtc_run  # (stimulus is regenerated below the same way the testbench does)
    with abi.Model("make_tracing_c", n, rand_reset=2) as m:
        # initial values as the device testbench would set them: read them back after begin()
        m.run("sim_main", 0, flags=abi.RUNF_TB_BEGIN, tb_seed=11, tb_flags=1)
        in_quad = m.get_member("in_quad").astype(np.uint64)
        clk, t = 0, 0
        for e in range(evals):
            t += 1
            clk ^= 1
            m.time_inc(1)
            m.set_input("clk", np.uint8(clk))
            if not clk:
                m.set_input("reset_l", np.uint8(0 if 1 < t < 10 else 1))
                in_quad = (in_quad + np.uint64(0x12)) & np.uint64((1 << 40) - 1)
                m.set_input("in_quad", in_quad)
            m.eval()
            assert (m.get_output("out_small") == want.obs[e, 0]).all()
            oq = m.get_output("out_quad")
            assert (oq == (want.obs[e, 1].astype(np.uint64) | (want.obs[e, 2].astype(np.uint64) << 32))).all()
            assert (m.get_output("out_wide") == want.obs[e, 3:6].T).all()
        check_status(m.status(), want.status)


def test_cxx_testbench_through_the_model_binding(abi, tmp_path):
    """sim_main.cpp's loop written in C++ against include/verilated_cuda_model.h (the Vtop-shaped
    binding over the C ABI): every port value of every eval equals the oracle's."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cxx", "sim_main_batch")
    n, evals = 48, 60
    out = str(tmp_path / "ports.bin")
    res = subprocess.run([exe, str(n), str(evals), out], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    assert res.stdout.strip() == f"finished=1 time={evals} first_finish=17"
    rec = np.dtype([("small", "u1"), ("quad", "<u8"), ("wide", "<u4", 3)])
    got = np.fromfile(out, dtype=rec).reshape(evals, n)
    want = O.mtc_run(n, evals, rand_reset=2, trace_plusarg=True, obs_evals=evals)
    assert (got["small"] == want.obs[:, 0, :]).all()
    assert (got["quad"] == (want.obs[:, 1].astype(np.uint64) | (want.obs[:, 2].astype(np.uint64) << 32))).all()
    assert (got["wide"] == want.obs[:, 3:6].transpose(0, 2, 1)).all()


def test_trace_testbench_with_host_stimulus_rows(abi):
    """Generic 'trace' testbench: stimulus rows [eval][in_word][instance] from the host."""
    n, evals = 130, 80
    want = O.mtc_run(n, evals, rand_reset=2, tb_seed=4, per_instance=True, obs_evals=evals)
    with abi.Model("make_tracing_c", n, rand_reset=2) as m:
        m.run("sim_main", 0, flags=abi.RUNF_TB_BEGIN, tb_seed=4, tb_flags=1)
        in_small = m.get_member("in_small").astype(np.uint32)
        in_quad = m.get_member("in_quad").astype(np.uint64)
        in_wide = m.get_member("in_wide").astype(np.uint32)
        stim = np.zeros((evals, m.in_words, m.npad), dtype=np.uint32)
        clk, t, reset_l = 0, 0, 1
        for e in range(evals):
            t += 1
            clk ^= 1
            if not clk:
                reset_l = 0 if 1 < t < 10 else 1
                in_quad = (in_quad + np.uint64(0x12)) & np.uint64((1 << 40) - 1)
            stim[e, 0, :n] = clk
            stim[e, 1, :n] = reset_l
            stim[e, 2, :n] = in_small
            stim[e, 3, :n] = (in_quad & np.uint64(0xffffffff)).astype(np.uint32)
            stim[e, 4, :n] = (in_quad >> np.uint64(32)).astype(np.uint32)
            stim[e, 5:8, :n] = in_wide.T
        obs, _ = m.run("trace", evals, stim=stim, obs_evals=evals)
        assert (obs == want.obs).all()
        check_status(m.status(), want.status)


def random_rows(m, evals, seed):
    """Random stimulus rows for make_tracing_c (clk toggling, a reset pulse, random data ports)."""
    rng = np.random.default_rng(seed)
    stim = rng.integers(0, 1 << 32, size=(evals, m.in_words, m.npad), dtype=np.uint64).astype(np.uint32)
    stim[:, 0, :] = (np.arange(evals, dtype=np.uint32) & 1)[:, None] ^ 1
    stim[:, 1, :] = (rng.random((evals, m.npad)) < 0.97)
    return stim


def test_port_streaming_kernel_and_host_pipeline_equal_the_direct_kernel(abi):
    """The same rows through (a) the direct-access kernel, (b) the streaming kernel with 2-D tensor
    TMA tiles and with 1-D bulk copies, (c) the streaming kernels without output capture, (d) vlcu_run's chunked host pipeline and
    (e) device-resident row buffers give the same ports after every eval, state and signatures;
    (a) is the form the other tests check against the reference model."""
    n, evals = 1000, 333  # not multiples of the CTA size / pipeline depth / chunk size
    with abi.Model("make_tracing_c", n, rand_reset=2, seed=9) as m:
        stim = random_rows(m, evals, 21)

        def fresh():
            m.reset()
            m.run("sim_main", 0, flags=abi.RUNF_TB_BEGIN, tb_seed=4, tb_flags=1)

        fresh()
        obs_a, _ = m.run("trace", evals, stim=stim, obs_evals=evals, flags=abi.RUNF_NO_STREAM)
        sig_a, st_a, rec_a = m.signatures(), m.status(), m.snapshot(n - 1)
        fresh()
        obs_b, _ = m.run("trace", evals, stim=stim, obs_evals=evals)
        assert (obs_b == obs_a).all()
        assert (m.signatures() == sig_a).all() and (m.status() == st_a).all()
        assert (m.snapshot(n - 1) == rec_a).all()
        fresh()
        obs_c, _ = m.run("trace", evals, stim=stim, obs_evals=evals, flags=abi.RUNF_BULK_1D)
        assert (obs_c == obs_a).all() and (m.signatures() == sig_a).all()
        for flags in (0, abi.RUNF_BULK_1D):
            fresh()
            m.run("trace", evals, stim=stim, flags=flags)
            assert (m.signatures() == sig_a).all() and (m.snapshot(n - 1) == rec_a).all()
        # two consecutive runs: rows are relative to each run
        fresh()
        o1, _ = m.run("trace", 100, stim=stim[:100], obs_evals=100)
        o2, _ = m.run("trace", evals - 100, stim=stim[100:], obs_evals=evals - 100)
        assert (np.concatenate([o1, o2]) == obs_a).all() and (m.signatures() == sig_a).all()
        # chunked host pipeline: 8 evals per chunk
        fresh()
        m.set_pipeline_chunk(1)
        obs_d, _ = m.run("trace", evals, stim=stim, obs_evals=evals)
        assert (obs_d == obs_a).all() and (m.signatures() == sig_a).all()
        fresh()
        m.run("trace", evals, stim=stim)
        assert (m.signatures() == sig_a).all() and (m.snapshot(n - 1) == rec_a).all()
        m.set_pipeline_chunk(64 << 20)
        # device-resident rows
        fresh()
        d_stim = abi.lib.vlcu_device_alloc(m.h, stim.nbytes)
        obs_e = np.zeros((evals, m.out_words, m.npad), dtype=np.uint32)
        d_obs = abi.lib.vlcu_device_alloc(m.h, obs_e.nbytes)
        assert d_stim and d_obs
        abi._check(abi.lib.vlcu_device_copy(m.h, d_stim, stim.ctypes.data, stim.nbytes, 1))
        m.run_rows(evals, d_stim, d_obs, flags=abi.RUNF_STIM_DEVICE | abi.RUNF_OBS_DEVICE)
        abi._check(abi.lib.vlcu_device_copy(m.h, obs_e.ctypes.data, d_obs, obs_e.nbytes, 0))
        assert (obs_e[:, :, :n] == obs_a).all() and (m.signatures() == sig_a).all()
        abi.lib.vlcu_device_free(m.h, d_stim)
        abi.lib.vlcu_device_free(m.h, d_obs)


@pytest.mark.parametrize("evals", [1, 2, 7, 8, 9, 17])
@pytest.mark.parametrize("n", [1, 64, 65, 257])
def test_port_streaming_pipeline_edges(abi, n, evals):
    """Fewer evals than pipeline stages, exactly one pipeline's worth, one more; a single
    instance, exactly one CTA, one instance more than a CTA."""
    with abi.Model("make_tracing_c", n, rand_reset=2, seed=3) as m:
        stim = random_rows(m, evals, 100 * n + evals)
        m.run("sim_main", 0, flags=abi.RUNF_TB_BEGIN, tb_seed=4, tb_flags=1)
        obs_a, _ = m.run("trace", evals, stim=stim, obs_evals=evals, flags=abi.RUNF_NO_STREAM)
        sig_a = m.signatures()
        for flags in (0, abi.RUNF_BULK_1D):
            m.reset()
            m.run("sim_main", 0, flags=abi.RUNF_TB_BEGIN, tb_seed=4, tb_flags=1)
            obs_b, _ = m.run("trace", evals, stim=stim, obs_evals=evals, flags=flags)
            assert (obs_b == obs_a).all() and (m.signatures() == sig_a).all()


@pytest.mark.parametrize("stop_on_finish,period", [(False, 0), (False, 40), (True, 0)])
def test_vcd_of_sampled_instances_equals_the_reference_writer(abi, tmp_path, stop_on_finish, period):
    """R8: waveforms of sampled instances.  The device samples full state after every eval;
    vlcu_vcd_write turns the records into VCD text, which must be byte-identical to what the
    reference's own VerilatedVcdC writes for the same instance (dump(time) after every eval)."""
    n, evals = 96, 400
    samples = [0, 41, 95]
    flags = abi.RUNF_TB_BEGIN | (abi.RUNF_STOP_ON_FINISH if stop_on_finish else 0)
    with abi.Model("make_tracing_c", n, seed=9, rand_reset=2, msg_cap=4, plusargs="+trace") as m:
        _, trace = m.run("sim_main", evals, flags=flags, tb_seed=77, tb_flags=1, tb_arg=(period, 0, 0, 0),
                         samples=samples, trace_evals=evals)
        for k, inst in enumerate(samples):
            ref_path, our_path = str(tmp_path / f"ref{inst}.vcd"), str(tmp_path / f"our{inst}.vcd")
            O.mtc_run(n, evals, seed=9, rand_reset=2, stop_on_finish=stop_on_finish, tb_seed=77,
                      per_instance=True, reset_period=period, trace_plusarg=True, vcd_path=ref_path,
                      vcd_instance=inst)
            m.vcd_write(trace, k, our_path)
            ours, want = open(our_path).read(), open(ref_path).read()
            assert want.count("\n#") > 10 and "$var wire 70" in want
            assert ours == want, (inst, ours[:600], want[:600])
        # N2: the same waveforms from on-device change detection (only changes leave the device)
        m.reset()
        cap = 4 + evals * (3 + 2 * 15)
        _, streams = m.run("sim_main", evals, flags=flags, tb_seed=77, tb_flags=1, tb_arg=(period, 0, 0, 0),
                           samples=samples, trace_evals=evals, trace_stream_words=cap)
        assert (streams[:, 1] == 0).all() and (streams[:, 0] * 4 < trace[:, 0].nbytes // 3).all()
        for k, inst in enumerate(samples):
            chg_path = str(tmp_path / f"chg{inst}.vcd")
            m.vcd_write_changes(streams, k, chg_path)
            assert open(chg_path).read() == open(str(tmp_path / f"ref{inst}.vcd")).read(), inst


def test_x_initial_unique_reset_values(abi):
    """ctor_var_reset: per-instance seeds feed VL_SCOPED_RAND_RESET_* (config 3 mechanism)."""
    n = 1000
    want = O.mtc_run(n, 0, inst_base=5000, seed=0x1234, rand_reset=2)
    with abi.Model("make_tracing_c", n, inst_base=5000, seed=0x1234, rand_reset=2) as m:
        got = m.get_member("top__DOT__sub__DOT__count_c")
        assert len(set(got.tolist())) > n * 0.99
    # the oracle's final count_c after zero evals is its reset value; sim_main then overwrites
    # the input ports, which the GPU model has not done yet: compare count_c only
    assert (got == want.final[:, 0]).all()
    want1 = O.mtc_run(8, 0, seed=1, rand_reset=1)
    with abi.Model("make_tracing_c", 8, seed=1, rand_reset=1) as m:
        assert (m.get_member("top__DOT__sub__DOT__count_c") == want1.final[:, 0]).all()
        assert (m.get_member("top__DOT__sub__DOT__count_c") == 0xffffffff).all()


def test_coverage_counters_and_file(abi, tmp_path):
    n, evals = 200, 900
    want = O.mtc_run(n, evals, seed=2, rand_reset=2, tb_seed=8, per_instance=True, reset_period=128)
    with abi.Model("make_tracing_c_cov", n, seed=2, rand_reset=2) as m:
        m.run("sim_main", evals, flags=abi.RUNF_TB_BEGIN, tb_seed=8, tb_flags=1, tb_arg=(128, 0, 0, 0))
        sums = m.coverage_reduce()
        assert m.coverage_bins() == len(want.cov) == 543
        assert (sums == want.cov).all()
        per_inst = m.get_member("__Vcoverage")
        assert (per_inst.astype(np.uint64).sum(axis=0) == want.cov).all()
    # single instance: the file must be byte-identical to the reference writer's
    ref_path, gpu_path = str(tmp_path / "ref.dat"), str(tmp_path / "gpu.dat")
    O.mtc_run(1, evals, seed=2, rand_reset=2, reset_period=128, cov_path=ref_path, cov_instance=0)
    with abi.Model("make_tracing_c_cov", 1, seed=2, rand_reset=2) as m:
        m.run("sim_main", evals, flags=abi.RUNF_TB_BEGIN, tb_arg=(128, 0, 0, 0))
        m.coverage_write(m.coverage_reduce(), gpu_path)
    assert open(gpu_path, "rb").read() == open(ref_path, "rb").read()


def test_assertion_failure_is_recorded_not_fatal(abi):
    """Original reset rule on a long run: count_c passes 100, the design's assertion fires.
    $finish came first, so (as in the reference's vl_stop) the stop is ignored."""
    n, evals = 16, 400
    want = O.mtc_run(n, evals, rand_reset=2)
    with abi.Model("make_tracing_c", n, rand_reset=2, msg_cap=4) as m:
        m.run("sim_main", evals, flags=abi.RUNF_TB_BEGIN)
        check_status(m.status(), want.status)
        assert (m.status()["status"] & abi.ST_MSGDROP).all()


def test_large_batch_signature_properties(abi):
    """BASELINE size (65,536 instances): instances with identical stimulus have identical
    signatures, different stimulus gives different signatures, sample checked against oracle."""
    n, evals = 65536, 256
    with abi.Model("make_tracing_c", n, seed=1, rand_reset=2) as m:
        m.run("sim_main", evals, flags=abi.RUNF_TB_BEGIN, tb_seed=3, tb_flags=1, tb_arg=(128, 0, 0, 0))
        sig = m.signatures()
        ms = m.last_kernel_ms()
        assert ms > 0
    assert len(np.unique(sig)) > n * 0.99
    want = O.mtc_run(64, evals, inst_base=0, seed=1, rand_reset=2, tb_seed=3, per_instance=True,
                     reset_period=128)
    assert (sig[:64] == want.sigs).all()
    want_hi = O.mtc_run(8, evals, inst_base=65000, seed=1, rand_reset=2, tb_seed=3,
                        per_instance=True, reset_period=128)
    assert (sig[65000:65008] == want_hi.sigs).all()
    with abi.Model("make_tracing_c", 4096, seed=1, rand_reset=2) as m:
        m.run("sim_main", evals, flags=abi.RUNF_TB_BEGIN, tb_arg=(128, 0, 0, 0))
        assert len(np.unique(m.signatures())) == 1


def test_huge_batches_index_arithmetic(abi):
    """2^25 + 5 instances: element offsets beyond 2^32 bytes in the state block, samples at both
    ends against the oracle.  2^22 + 5 instances: port rows streamed for a batch whose row stride
    exceeds 2^22 elements.  Neither is a multiple of the CTA or padding size."""
    n, evals = (1 << 25) + 5, 32
    with abi.Model("make_tracing_c", n, seed=1, rand_reset=2, msg_cap=0) as m:
        assert m.npad * lib_state_bytes(abi, m) > (1 << 32)
        m.run("sim_main", evals, flags=abi.RUNF_TB_BEGIN, tb_seed=3, tb_flags=1, tb_arg=(16, 0, 0, 0))
        sig = m.signatures()
        for base in (0, (1 << 24) + 77, n - 8):
            want = O.mtc_run(8, evals, inst_base=base, seed=1, rand_reset=2, tb_seed=3, per_instance=True,
                             reset_period=16)
            assert (sig[base:base + 8] == want.sigs).all(), base
    n = (1 << 22) + 5
    with abi.Model("make_tracing_c", n, seed=1, rand_reset=2, msg_cap=0) as m:
        stim = np.zeros((8, m.in_words, m.npad), dtype=np.uint32)
        stim[:, 0, :] = ((np.arange(8) & 1) ^ 1)[:, None]
        stim[:, 1, :] = 1
        stim[:, 3, :] = (np.arange(m.npad, dtype=np.uint64) * 2654435761 & 0xffffffff).astype(np.uint32)[None, :]
        obs_a, _ = m.run("trace", 8, stim=stim, obs_evals=8, flags=abi.RUNF_NO_STREAM)
        sig_a = m.signatures()
        m.reset()
        obs_b, _ = m.run("trace", 8, stim=stim, obs_evals=8)
        assert (obs_a == obs_b).all() and (m.signatures() == sig_a).all()
        assert len(np.unique(obs_a[7, 1, :])) > 1000


def lib_state_bytes(abi, m):
    return abi.lib.vlcu_state_bytes(m.h)


def _sha_file(path):
    import hashlib
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 22), b""):
            h.update(chunk)
    return h.hexdigest(), os.path.getsize(path)


def test_config1_one_million_cycles_vcd_bit_exact(abi, tmp_path):
    """BASELINE.json configs[0] as written: examples/make_tracing_c for 1,000,000 clock cycles
    (2,000,000 evals), the VCD of an instance compared bit for bit with what the reference's
    VerilatedVcdC writes, and the port signatures of every instance with the genuine model's.
    The shipped design $finishes once count_c reaches 3 (sub.v:24-29) and asserts further on; for
    a run this long the testbench keeps going after gotFinish and pulses the reset every 128 time
    units (the device SURVEY 8d documents for config 1, the same on both sides), so the counter
    stays below the assertion's threshold.  The waveform leaves the device as a change stream
    (N2): ~150 MB for the 2,000,000 dumps."""
    n, evals = 48, 2_000_000
    sample = 17
    ref_path, our_path = str(tmp_path / "ref.vcd"), str(tmp_path / "our.vcd")
    want = O.mtc_run(n, evals, seed=9, rand_reset=2, tb_seed=77, per_instance=True, reset_period=128,
                     trace_plusarg=True, vcd_path=ref_path, vcd_instance=sample, threads=8)
    with abi.Model("make_tracing_c", n, seed=9, rand_reset=2, msg_cap=4, plusargs="+trace") as m:
        cap = 4 + evals * (3 + 2 * 8)
        _, streams = m.run("sim_main", evals, flags=abi.RUNF_TB_BEGIN, tb_seed=77, tb_flags=1,
                           tb_arg=(128, 0, 0, 0), samples=[sample], trace_evals=evals, trace_stream_words=cap)
        assert streams[0, 1] == 0 and streams[0, 2] == evals  # nothing dropped, one dump per eval
        m.vcd_write_changes(streams, 0, our_path)
        assert (m.signatures() == want.sigs).all()
        st = m.status()
        assert (st["time"] == evals).all() and (st["stop_count"] == 0).all() and (st["fatal_count"] == 0).all()
    ours, ref = _sha_file(our_path), _sha_file(ref_path)
    assert ref[1] > 50_000_000  # a real waveform: tens of MB of value changes
    assert ours == ref


def test_unmodified_sim_main_cpp_on_the_vtop_shim(abi, tmp_path):
    """examples/make_tracing_c/sim_main.cpp, UNMODIFIED, compiled against a three-line
    Vtop.h over include/verilated_cuda_vtop.h -- the generated model class with its public port
    members (V3EmitCModel.cpp:108-117), one instance on the GPU.  The program pokes top->clk,
    top->in_quad ..., calls top->eval(), polls contextp->gotFinish(), prints the design's
    $display text and the example's own trace lines: the console must be the genuine model's."""
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cxx", "sim_main_vtop")
    if not os.path.exists(exe):
        pytest.skip("tests/cxx/sim_main_vtop is built where /root/reference is")
    res = subprocess.run([exe], capture_output=True, text=True, cwd=str(tmp_path), timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    want = O.mtc_run(1, 1000, rand_reset=2, stop_on_finish=True)
    assert "*-* All Finished *-*" in res.stdout
    for line in want.console.splitlines():  # the design's $display / $write text
        assert line in res.stdout, line
    # the testbench's own per-eval report of the ports (sim_main.cpp:109-112) stops where the
    # genuine run does ($finish at time 17) and shows the genuine model's outputs
    evals = [ln for ln in res.stdout.splitlines() if ln.startswith("[") and "clk=" in ln]
    assert len(evals) == 17 and evals[-1].startswith("[17] ")
    obs = O.mtc_run(1, 17, rand_reset=2, obs_evals=17).obs[:, :, 0]
    for e, line in enumerate(evals):
        quad = int(obs[e, 1]) | (int(obs[e, 2]) << 32)
        assert line.endswith(f"-> oquad={quad:x} owide={int(obs[e, 5]):x}_{int(obs[e, 4]):08x}_{int(obs[e, 3]):08x}"), (e, line)
