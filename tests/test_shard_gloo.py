"""Multi-GPU host logic on CPU: world_size-2 and -3 gloo jobs.

What shards: instances, by contiguous blocks of the global index.  What is exchanged: the
coverage sums (one all-reduce) and, for the rendezvous, the NCCL unique id.  On the GPU the
all-reduce is ncclAllReduce inside vlcu_coverage_reduce; here gloo carries the same data so the
partition, seeding-by-global-index and merge logic are checked without a device."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import oracle_util as O
from verilator_b200 import shard as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_block_partition_tiles_the_batch():
    for total in (0, 1, 7, 64, 65536, 100003):
        for world in (1, 2, 3, 8):
            blocks = [S.shard(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and sum(n for _, n in blocks) == total
            for (b0, n0), (b1, _) in zip(blocks, blocks[1:]):
                assert b0 + n0 == b1
            sizes = [n for _, n in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        S.shard(10, 2, 2)


def test_single_process_helpers_are_identity():
    assert S.allreduce_max(None, 3.5) == 3.5
    assert (S.allreduce_sum_u64(None, [1, 2]) == np.array([1, 2], dtype=np.uint64)).all()
    assert S.broadcast_bytes(None, b"abc", 3) == b"abc"


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_campaign_equals_one_big_run(tmp_path, world):
    total, evals = 50, 300
    out = str(tmp_path / "rank0.json")
    port = _free_port()
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "shard_worker.py"),
                                       str(total), str(evals), out], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        text, _ = p.communicate(timeout=300)
        assert p.returncode == 0, text
    got = json.load(open(out))
    want = O.mtc_run(total, evals, seed=7, rand_reset=2, tb_seed=11, per_instance=True, reset_period=64)
    assert got["world"] == world and got["uid_ok"] and got["tmax"] == float(world)
    assert got["cov"] == want.cov.tolist(), "merged coverage differs from the unsharded run"
    assert got["sigs"] == want.sigs.tolist(), "results depend on the sharding"
    assert sum(got["cov"]) > 0
