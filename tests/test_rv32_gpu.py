"""Parity of the batched CUDA evaluator with the oracle for the RV32I core (BASELINE.json
config 2 workload: 65,536 instances each running a random-seeded synthetic program), through
the C ABI.  Bit-exact: all top-level ports after every eval, register file, memory, context."""
import numpy as np
import pytest

import oracle_util as O
import rv32_iss as ISS

pytestmark = pytest.mark.gpu
SCALARS = ["rv32_soc__DOT__pc", "rv32_soc__DOT__insn", "rv32_soc__DOT__ld_addr",
           "rv32_soc__DOT__instret", "rv32_soc__DOT__state", "trap", "retired", "clk"]


@pytest.fixture(scope="module")
def abi():
    from verilator_b200 import abi
    assert abi.lib.vlcu_device_count() > 0, "no CUDA device: GPU tests need a B200"
    assert "rv32i" in abi.designs()
    return abi


def check_final_state(m, want):
    assert (m.get_member("rv32_soc__DOT__regs") == want.regs).all()
    for k, name in enumerate(SCALARS):
        assert (m.get_member(name).astype(np.uint32) == want.scalars[:, k]).all(), name
    if want.mem is not None:
        assert (m.get_member("rv32_soc__DOT__mem") == want.mem).all()


@pytest.mark.parametrize("rand_reset,reset_time", [(2, 0), (0, 0), (2, 5)])
def test_ports_every_eval_match_oracle(abi, rand_reset, reset_time):
    n, evals = 200, 1200
    want = O.rv32_run(n, evals, seed=11, rand_reset=rand_reset, tb_seed=2024, reset_time=reset_time,
                      obs_evals=evals, want_mem=True)
    with abi.Model("rv32i", n, seed=11, rand_reset=rand_reset) as m:
        obs, _ = m.run("rv32_rand", evals, flags=abi.RUNF_TB_BEGIN, obs_evals=evals, tb_seed=2024,
                       tb_arg=(reset_time, 0, 0, 0))
        assert (obs == want.obs).all()
        assert (m.signatures() == want.sigs).all()
        assert (m.status()["time"] == want.times).all()
        check_final_state(m, want)


def test_same_program_x_initial_unique_divergence_only(abi):
    """BASELINE config 3's setup: every instance runs the same program; they differ only in the
    x-initial unique reset values (randSeed = seed ^ global index)."""
    n, evals = 200, 700
    want = O.rv32_run(n, evals, inst_base=300, seed=12, rand_reset=2, tb_seed=40, same_program=True,
                      obs_evals=evals, want_mem=True)
    with abi.Model("rv32i", n, inst_base=300, seed=12, rand_reset=2) as m:
        obs, _ = m.run("rv32_rand", evals, flags=abi.RUNF_TB_BEGIN, tb_seed=40, tb_flags=2, obs_evals=evals)
        assert (obs == want.obs).all()
        check_final_state(m, want)
        assert (m.signatures() == want.sigs).all()
        mem0 = O.rv32_gen_program(40)
        assert len(np.unique(want.sigs)) > 1  # the reset values did make instances differ


def test_campaign_rounds_fresh_program_and_reset_per_run(abi):
    """Fuzz-campaign form used by bench.py's e2e leg: every vlcu_run starts a fresh program and
    holds reset for the run's first evals (tb_flags bit 0), state otherwise carried over."""
    n, evals, rounds = 160, 700, 3
    want = O.rv32_run(n, evals, seed=5, rand_reset=2, tb_seed=900, reset_time=6, rounds=rounds,
                      obs_evals=evals * rounds, want_mem=True)
    with abi.Model("rv32i", n, seed=5, rand_reset=2) as m:
        for k in range(rounds):
            obs, _ = m.run("rv32_rand", evals, flags=abi.RUNF_TB_BEGIN, obs_evals=evals,
                           tb_seed=900 + k, tb_flags=1, tb_arg=(6, 0, 0, 0))
            assert (obs == want.obs[k * evals:(k + 1) * evals]).all(), k
        assert (m.signatures() == want.sigs).all()
        check_final_state(m, want)


def test_coverage_fuzz_campaign_counters_and_file(abi, tmp_path):
    """BASELINE.json config 5 in miniature: the --coverage build of the core (line, toggle,
    branch and expression points: 1,664 counters per instance) run as a fuzz campaign -- fresh
    random program and reset every round -- with the counters summed on the device.  The sums
    must equal the sum over the genuine C++ model's instances, and one instance's coverage.dat
    must be byte-identical to what the reference's VerilatedCov writes."""
    n, evals, rounds = 120, 500, 3
    want = O.rv32_run(n, evals, seed=5, rand_reset=2, tb_seed=40, reset_time=6, rounds=rounds, cov=True)
    with abi.Model("rv32i_cov", n, seed=5, rand_reset=2) as m:
        for k in range(rounds):
            m.run("rv32_rand", evals, flags=abi.RUNF_TB_BEGIN, tb_seed=40 + k, tb_flags=1, tb_arg=(6, 0, 0, 0))
        sums = m.coverage_reduce()
        assert m.coverage_bins() == len(want.cov) == 1664
        assert (sums == want.cov).all()
        assert (m.signatures() == want.sigs).all()
        assert (sums > 0).sum() > 800  # the campaign really toggles most of the design
    ref_path, gpu_path = str(tmp_path / "ref.dat"), str(tmp_path / "gpu.dat")
    O.rv32_run(1, evals, seed=5, rand_reset=2, tb_seed=40, reset_time=6, rounds=1, cov=True,
               cov_path=ref_path, cov_instance=0)
    with abi.Model("rv32i_cov", 1, seed=5, rand_reset=2) as m:
        m.run("rv32_rand", evals, flags=abi.RUNF_TB_BEGIN, tb_seed=40, tb_flags=1, tb_arg=(6, 0, 0, 0))
        m.coverage_write(m.coverage_reduce(), gpu_path)
    assert open(gpu_path, "rb").read() == open(ref_path, "rb").read()


def test_x_initial_reset_of_arrays_and_scalars(abi):
    want = O.rv32_run(300, 0, inst_base=77, seed=123, rand_reset=2, want_mem=True)
    with abi.Model("rv32i", 300, inst_base=77, seed=123, rand_reset=2) as m:
        assert (m.get_member("rv32_soc__DOT__regs") == want.regs).all()
        assert (m.get_member("rv32_soc__DOT__pc") == want.scalars[:, 0]).all()
        # before the testbench loads a program the memory holds its x-initial value
        mem = m.get_member("rv32_soc__DOT__mem")
        assert (mem == mem[:, :1]).all() and len(np.unique(mem[:, 0])) == 300


def test_host_loaded_memory_images_and_iss(abi):
    """Programs supplied by the host (vlcu_set_member on the instance-major memory), including
    illegal instructions / misaligned / out-of-range accesses that must trap."""
    rng = np.random.default_rng(8)
    n, evals = 96, 500
    images = rng.integers(0, 1 << 32, size=(n, 1024), dtype=np.uint64).astype(np.uint32)
    for i in range(n):
        good = O.rv32_gen_program(900 + i)
        sel = rng.random(1024) < 0.85
        images[i, sel] = good[sel]
    want = O.rv32_run(n, evals, seed=4, rand_reset=2, mem_images=images, obs_evals=evals, want_mem=True)
    with abi.Model("rv32i", n, seed=4, rand_reset=2) as m:
        regs0 = m.get_member("rv32_soc__DOT__regs")
        m.set_member("rv32_soc__DOT__mem", images)
        assert (m.get_member("rv32_soc__DOT__mem") == images).all()
        obs, _ = m.run("rv32_host", evals, flags=abi.RUNF_TB_BEGIN, obs_evals=evals)
        assert (obs == want.obs).all()
        check_final_state(m, want)
        assert m.get_output("trap").sum() > 0
        # and straight against the instruction-set simulator
        regs, mem, pc = m.get_member("rv32_soc__DOT__regs"), m.get_member("rv32_soc__DOT__mem"), \
            m.get_member("rv32_soc__DOT__pc")
        for i in range(0, n, 7):
            sim = ISS.Rv32(images[i], regs0[i])
            sim.run(evals)
            assert sim.regs[1:] == regs[i, 1:].tolist() and sim.mem == mem[i].tolist()
            assert sim.pc == int(pc[i])


def pinned(abi, shape, dtype):
    import ctypes
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    ptr = abi.lib.vlcu_host_alloc(nbytes)
    assert ptr
    return np.frombuffer((ctypes.c_uint8 * nbytes).from_address(ptr), dtype=dtype).reshape(shape)


def test_pipelined_rounds_staged_while_the_previous_round_runs(abi):
    """vlcu_stage_member / vlcu_commit_staged: the images of round r+1 are copied while round r
    computes and must neither disturb round r nor differ from a synchronous vlcu_set_member."""
    n, evals, rounds = 192, 400, 3
    sets = [np.stack([O.rv32_gen_program(5000 + 97 * r + i) for i in range(n)]) for r in range(rounds)]
    wants = [O.rv32_run(n, evals, seed=6, rand_reset=2, mem_images=img, want_mem=True) for img in sets]
    bufs = [pinned(abi, (n, 1024), np.uint32) for _ in range(2)]
    with abi.Model("rv32i", n, seed=6, rand_reset=2) as m:
        bufs[0][:] = sets[0]
        m.stage_member("rv32_soc__DOT__mem", bufs[0])
        for r in range(rounds):
            m.reset()  # every round starts from a freshly constructed model, as the oracle's does
            m.commit_staged()
            m.run("rv32_host", evals, flags=abi.RUNF_TB_BEGIN)
            if r + 1 < rounds:  # in flight together with the kernel above
                bufs[(r + 1) & 1][:] = sets[r + 1]
                m.stage_member("rv32_soc__DOT__mem", bufs[(r + 1) & 1])
            assert (m.signatures() == wants[r].sigs).all(), r
            check_final_state(m, wants[r])
    # asynchronous read-back: results of round r are fetched while round r+1 computes
    with abi.Model("rv32i", n, seed=6, rand_reset=2) as m:
        res = [(pinned(abi, (n,), np.uint64), pinned(abi, (n,), abi.STATUS_DTYPE), pinned(abi, (n, 32), np.uint32),
                pinned(abi, (n, 1024), np.uint32)) for _ in range(2)]
        for r in range(rounds):
            m.reset()
            m.set_member("rv32_soc__DOT__mem", sets[r])
            m.run("rv32_host", evals, flags=abi.RUNF_TB_BEGIN)
            if r > 0:
                m.results_wait()
                sg, st, rg, mm = res[(r - 1) & 1]
                assert (sg == wants[r - 1].sigs).all() and (rg == wants[r - 1].regs).all()
                assert (mm == wants[r - 1].mem).all() and (st["time"] == evals).all()
            sg, st, rg, mm = res[r & 1]
            m.results_begin(sg, st, [("rv32_soc__DOT__regs", rg), ("rv32_soc__DOT__mem", mm)])
        m.results_wait()
        sg, st, rg, mm = res[(rounds - 1) & 1]
        assert (sg == wants[-1].sigs).all() and (rg == wants[-1].regs).all() and (mm == wants[-1].mem).all()
    # ports take the scatter path, masked to their width like vlcu_set_input
    with abi.Model("make_tracing_c", n) as m:
        q = pinned(abi, (n, 1), np.uint64)
        q[:, 0] = np.arange(n, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
        w = pinned(abi, (n, 3), np.uint32)
        w[:] = np.arange(3 * n, dtype=np.uint32).reshape(n, 3) * np.uint32(0x01000193) + np.uint32(0xFFFFFF00)
        m.stage_member("in_quad", q, mask_width=True)
        m.stage_member("in_wide", w, mask_width=True)
        m.commit_staged()
        assert (m.get_member("in_quad") == (q[:, 0] & np.uint64((1 << 40) - 1))).all()
        got = m.get_member("in_wide")
        assert (got[:, :2] == w[:, :2]).all() and (got[:, 2] == (w[:, 2] & 0x3F)).all()


def test_split_runs_and_snapshot(abi):
    n = 130
    with abi.Model("rv32i", n, seed=6, rand_reset=2) as a, abi.Model("rv32i", n, seed=6, rand_reset=2) as b:
        a.run("rv32_rand", 900, flags=abi.RUNF_TB_BEGIN, tb_seed=31)
        b.run("rv32_rand", 7, flags=abi.RUNF_TB_BEGIN, tb_seed=31)
        b.run("rv32_rand", 500, tb_seed=31)
        b.run("rv32_rand", 393, tb_seed=31)
        assert (a.signatures() == b.signatures()).all()
        assert (a.get_member("rv32_soc__DOT__mem") == b.get_member("rv32_soc__DOT__mem")).all()
        snap = a.snapshot(77)
        assert (snap == b.snapshot(77)).all()
        assert (a.record_member(snap, "rv32_soc__DOT__mem") == a.get_member("rv32_soc__DOT__mem")[77]).all()
        assert (a.record_member(snap, "rv32_soc__DOT__regs") == a.get_member("rv32_soc__DOT__regs")[77]).all()


def test_trace_samples_of_selected_instances(abi):
    """Waveform sampling: full-state records of a few instances after every eval."""
    n, evals = 300, 64
    samples = [3, 150, 299]
    want = O.rv32_run(n, evals, seed=2, rand_reset=2, tb_seed=5, obs_evals=evals)
    with abi.Model("rv32i", n, seed=2, rand_reset=2) as m:
        _, trace = m.run("rv32_rand", evals, flags=abi.RUNF_TB_BEGIN, tb_seed=5, samples=samples,
                         trace_evals=evals)
        for e in range(evals):
            for k, inst in enumerate(samples):
                rec = trace[e, k]
                t = int(np.frombuffer(rec[:8].tobytes(), dtype=np.uint64)[0])
                assert t == e + 1
                body = rec[abi.TRACE_HDR:]
                assert int(m.record_member(body, "mem_addr")) == int(want.obs[e, 3, inst])
                assert int(m.record_member(body, "retired")) == int(want.obs[e, 7, inst])


def test_vcd_of_sampled_cores_equals_the_reference_writer(abi, tmp_path):
    """Waveforms (R8): the device evaluates the model's own trace functions for the sampled
    instances after every eval (every traced expression, the register file element by element);
    vlcu_vcd_write must produce byte for byte what the reference's VerilatedVcdC writes for the
    genuine C++ model of the same instance."""
    n, evals = 40, 300
    samples = [0, 23, 39]
    with abi.Model("rv32i", n, seed=2, rand_reset=2) as m:
        _, trace = m.run("rv32_rand", evals, flags=abi.RUNF_TB_BEGIN, tb_seed=5, samples=samples,
                         trace_evals=evals)
        for k, inst in enumerate(samples):
            ref_path, our_path = str(tmp_path / f"ref{inst}.vcd"), str(tmp_path / f"our{inst}.vcd")
            O.rv32_run(n, evals, seed=2, rand_reset=2, tb_seed=5, vcd_path=ref_path, vcd_instance=inst)
            m.vcd_write(trace, k, our_path)
            ours, want = open(our_path).read(), open(ref_path).read()
            assert "regs[31] [31:0]" in want and want.count("\n#") > evals // 2
            assert ours == want, (inst, ours[:800], want[:800])
        # N2: on-device change detection.  Full records carry 4.3 KB of state + 88 trace words per
        # eval; the change stream a few words per eval
        m.reset()
        cap = 4 + evals * (3 + 2 * 88)
        _, streams = m.run("rv32_rand", evals, flags=abi.RUNF_TB_BEGIN, tb_seed=5, samples=samples,
                           trace_evals=evals, trace_stream_words=cap)
        assert (streams[:, 1] == 0).all() and (streams[:, 2] == evals).all()
        assert streams[:, 0].max() * 4 * 8 < trace[:, 0].nbytes
        for k, inst in enumerate(samples):
            chg_path = str(tmp_path / f"chg{inst}.vcd")
            m.vcd_write_changes(streams, k, chg_path)
            assert open(chg_path).read() == open(str(tmp_path / f"ref{inst}.vcd")).read(), inst
        # a stream too small for the run: dumps are dropped, counted, and the writer refuses
        m.reset()
        _, small = m.run("rv32_rand", evals, flags=abi.RUNF_TB_BEGIN, tb_seed=5, samples=samples,
                         trace_evals=evals, trace_stream_words=600)
        assert (small[:, 1] > 0).all() and (small[:, 0] <= 600).all()
        with pytest.raises(RuntimeError, match="overflow"):
            m.vcd_write_changes(small, 0, str(tmp_path / "small.vcd"))


def test_full_batch_signatures_sample_against_oracle(abi):
    """BASELINE config-2 size.  Every instance runs a different program: signatures are all
    distinct; contiguous samples from both ends of the batch are checked against the oracle."""
    n, evals = 65536, 1024
    with abi.Model("rv32i", n, seed=1, rand_reset=2) as m:
        m.run("rv32_rand", evals, flags=abi.RUNF_TB_BEGIN, tb_seed=9)
        sig = m.signatures()
        retired = m.get_output("retired")
        assert m.get_output("trap").sum() == 0
    assert len(np.unique(sig)) == n
    assert retired.min() > 100
    for base in (0, 32768, 65536 - 48):
        want = O.rv32_run(48, evals, inst_base=base, seed=1, rand_reset=2, tb_seed=9, threads=4)
        assert (sig[base:base + 48] == want.sigs).all()
        assert (retired[base:base + 48] == want.scalars[:, 6]).all()


def test_campaign_with_host_supplied_seeds(abi):
    """bench.py's end-to-end leg: the host hands one 64-bit program seed per instance to every
    round (vlcu_run_desc_t.tb_data); the images are made on the device.  Same as loading the
    images rv32_gen_program makes from those seeds on the host."""
    n, evals = 130, 500
    rng = np.random.default_rng(77)
    with abi.Model("rv32i", n, seed=5, rand_reset=2) as m:
        sigs = []
        for k in range(2):
            seeds = rng.integers(0, 1 << 63, size=n, dtype=np.uint64)
            images = np.stack([O.rv32_gen_program(int(s)) for s in seeds])
            want = O.rv32_run(n, evals, seed=5, rand_reset=2, reset_time=6, rounds=1, mem_images=images,
                              obs_evals=evals)
            if k:  # (a second round on the same models: compare what a fresh oracle can)
                m.reset()
            obs, _ = m.run("rv32_seeds", evals, flags=abi.RUNF_TB_BEGIN, obs_evals=evals, tb_flags=1,
                           tb_arg=(6, 0, 0, 0), tb_data=seeds)
            assert (obs == want.obs).all() and (m.signatures() == want.sigs).all(), k
            assert (m.get_member("rv32_soc__DOT__mem")[:, :512] == images[:, :512]).sum() > n * 500
            sigs.append(m.signatures())
        assert (sigs[0] != sigs[1]).any()



@pytest.mark.parametrize("same_program", [False, True])
def test_interleaved_memory_layout_variant(abi, same_program):
    """the same core emitted with --cuda-array-interleave (memory kept
    [word][instance] although it is indexed by design signals) is the same machine."""
    n, evals = 200, 700
    want = O.rv32_run(n, evals, seed=3, rand_reset=2, tb_seed=21, obs_evals=evals, want_mem=True,
                      same_program=same_program)
    with abi.Model("rv32i_il", n, seed=3, rand_reset=2) as m:
        obs, _ = m.run("rv32_rand", evals, flags=abi.RUNF_TB_BEGIN, tb_seed=21, tb_flags=2 if same_program else 0,
                       obs_evals=evals)
        assert (obs == want.obs).all() and (m.signatures() == want.sigs).all()
        assert (m.get_member("rv32_soc__DOT__mem") == want.mem).all()
        assert (m.get_member("rv32_soc__DOT__regs") == want.regs).all()
