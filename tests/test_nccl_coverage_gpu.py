"""The coverage all-reduce (the only collective north_star allows): vlcu_comm_init +
vlcu_coverage_reduce with a real NCCL communicator.  World size 1 runs wherever there is one GPU;
the N-rank form (tests/nccl_cov_worker.py under torchrun) runs when the box has N GPUs."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import oracle_util as O
from verilator_b200 import abi

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_world_size_one_communicator():
    """ncclCommInitRank + ncclAllReduce(u64, sum) really run (a one-rank communicator): the sums
    equal the device-local ones and the oracle's, and coverage.dat is the reference writer's."""
    n, evals, rounds = 96, 400, 2
    want = O.rv32_run(n, evals, seed=5, rand_reset=2, tb_seed=40, reset_time=6, rounds=rounds, cov=True)
    with abi.Model("rv32i_cov", n, seed=5, rand_reset=2) as m:
        for k in range(rounds):
            m.run("rv32_rand", evals, flags=abi.RUNF_TB_BEGIN, tb_seed=40 + k, tb_flags=1, tb_arg=(6, 0, 0, 0))
        local = m.coverage_reduce()
        m.comm_init(abi.Model.comm_unique_id(), 0, 1)
        merged = m.coverage_reduce()
        assert (merged == local).all() and (merged == want.cov).all()
        abi._check(abi.lib.vlcu_comm_destroy(m.h))
        assert (m.coverage_reduce() == want.cov).all()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2])
def test_torchrun_ranks_merge_to_the_oracle_sum(tmp_path, world):
    if abi.lib.vlcu_device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    out = str(tmp_path / "nccl.json")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "nccl_cov_worker.py"), "200", "400", "2", out]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    got = json.load(open(out))
    assert got["world"] == world and got["all_ranks_ok"]
    assert got["nccl_equals_oracle_sum_over_all_instances"] and got["nccl_equals_torch"]
    assert got["coverage_dat_keys_equal_reference_writer"]
