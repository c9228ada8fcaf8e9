"""CPU side of the VeeR-class design (verilator_b200/csrc/designs/veer_class): the genuine
Verilated model behaves like a core, and the two restatements of the in-kernel "fuzz" testbench
(numpy rows fed to the generic harness; the harness's own row generator) agree."""
import numpy as np

import oracle_util as O

VC = dict(tb_flags=O.FUZZ_RESET_LOW, reset_periods=4, sparse_mask=0b11100000)


def test_fuzz_restatements_agree():
    n, evals = 5, 120
    iw, ow = O.generic_words("veer_class")
    assert (iw, ow) == (8, 13)  # clk rst_l ifu_rdata bus_rdata[2] irq_src[2] ctl | 13 output words
    rows = O.fuzz_rows(n, evals, iw, inst_base=3, seed=7, tb_seed=11, **VC)
    assert (rows[:, 0, 0] == (np.arange(evals) & 1) ^ 1).all()  # rising edge first
    assert (rows[:8, 1, :] == 0).all() and (rows[8:, 1, :] == 1).all()  # reset for 4 periods
    assert (rows[0::2] == rows[1::2])[:, 1:].all()  # inputs change once per clock period
    want_obs, want_sigs = O.generic_run("veer_class", rows, inst_base=3, seed=7, rand_reset=2)
    obs, sigs = O.generic_fuzz("veer_class", n, evals, inst_base=3, seed=7, rand_reset=2, tb_seed=11,
                               obs_evals=evals, **VC)
    assert (obs == want_obs).all() and (sigs == want_sigs).all()


def test_harts_retire_and_instances_diverge():
    n, evals = 6, 600
    obs, sigs = O.generic_fuzz("veer_class", n, evals, seed=1, rand_reset=2, tb_seed=5, obs_evals=evals, **VC)
    status = obs[:, 10, :]  # {retired[15:0], pic_pri, tmr_irq, pic_irq, trapped}
    assert ((status[-1] >> 16) > 50).all()  # instructions retire on every instance
    assert len(np.unique(sigs)) == n
    # uniform stimulus: the instances still differ, through their --x-initial values only
    _, sigs_u = O.generic_fuzz("veer_class", n, evals, seed=1, rand_reset=2, tb_seed=5,
                               tb_flags=VC["tb_flags"] | O.FUZZ_UNIFORM, reset_periods=4, sparse_mask=VC["sparse_mask"])
    assert len(np.unique(sigs_u)) == n
    # ... and with zeroed initial values they are identical
    _, sigs_z = O.generic_fuzz("veer_class", n, evals, seed=1, rand_reset=0, tb_seed=5,
                               tb_flags=VC["tb_flags"] | O.FUZZ_UNIFORM, reset_periods=4, sparse_mask=VC["sparse_mask"])
    assert len(np.unique(sigs_z)) == 1
