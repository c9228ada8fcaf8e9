"""rv32_soc: the hand-scheduled model (oracle build, genuine runtime) against an independent
instruction-set simulator, and the program generator's guarantees (CPU only)."""
import numpy as np
import pytest

import oracle_util as O
import rv32_iss as ISS


def check_against_iss(res, i, mem0, regs0, n_evals, reset_time=16):
    sim = ISS.Rv32(mem0, regs0, pc=int(res.init.scalars[i, 0]), state=int(res.init.scalars[i, 4]),
                   instret=int(res.init.scalars[i, 3]), insn=int(res.init.scalars[i, 1]),
                   ld_addr=int(res.init.scalars[i, 2]))
    sim.run(n_evals, reset_time)
    assert sim.regs[1:] == res.regs[i, 1:].tolist()
    assert sim.mem == res.mem[i].tolist()
    pc, insn, ld_addr, instret, state, trap, retired, clk = [int(x) for x in res.scalars[i]]
    assert (sim.pc, sim.instret, sim.state) == (pc, instret, state)
    assert trap == int(sim.state == ISS.S_TRAP) and retired == instret
    return sim


@pytest.mark.parametrize("n_evals", [40, 333, 3000])
def test_scheduled_model_matches_iss_on_generated_programs(n_evals):
    n = 24
    res = O.rv32_run(n, n_evals, seed=5, rand_reset=2, tb_seed=1234, want_mem=True)
    res.init = O.rv32_run(n, 0, seed=5, rand_reset=2, tb_seed=1234, want_mem=True)
    retired = 0
    for i in range(n):
        mem0 = O.rv32_gen_program(1234 ^ i)
        assert (mem0 == res.init.mem[i]).all()
        sim = check_against_iss(res, i, mem0, res.init.regs[i], n_evals)
        retired += sim.instret
    if n_evals >= 333:
        assert retired > n * 20  # programs really execute


def test_generated_programs_never_trap():
    n, n_evals = 64, 6000
    res = O.rv32_run(n, n_evals, seed=9, rand_reset=2, tb_seed=77)
    assert (res.scalars[:, 5] == 0).all(), "a generated program trapped"
    assert (res.scalars[:, 3] > 500).all()


def test_arbitrary_memory_images_including_illegal_instructions():
    """Random bit patterns as programs: traps, misaligned and out-of-range accesses."""
    rng = np.random.default_rng(3)
    n, n_evals = 48, 400
    images = rng.integers(0, 1 << 32, size=(n, 1024), dtype=np.uint64).astype(np.uint32)
    # make a fair share of words look like real instructions so execution gets somewhere
    for i in range(n):
        good = O.rv32_gen_program(500 + i)
        sel = rng.random(1024) < 0.8
        images[i, sel] = good[sel]
    res = O.rv32_run(n, n_evals, seed=2, rand_reset=2, mem_images=images, want_mem=True)
    res.init = O.rv32_run(n, 0, seed=2, rand_reset=2, mem_images=images, want_mem=True)
    for i in range(n):
        check_against_iss(res, i, images[i], res.init.regs[i], n_evals)
    assert res.scalars[:, 5].sum() > 0  # some of them trapped


def test_x_initial_makes_instances_differ():
    res = O.rv32_run(16, 0, seed=1, rand_reset=2)
    assert len({tuple(r) for r in res.regs.tolist()}) == 16
    # every element of an unpacked array draws the same reset value (reference quirk)
    assert (res.regs == res.regs[:, :1]).all()
