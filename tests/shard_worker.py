"""Worker of tests/test_shard_gloo.py: one rank of a world_size-N gloo job (CPU only).

Each rank plays one GPU shard of a coverage campaign, with the ORACLE standing in for the
device (no GPU here): it runs its block of global instances, then the ranks merge coverage
sums / exchange a fake unique id / take the max of their step times exactly as bench.py does
on the NCCL backend.  Rank 0 writes what it saw to the result file."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import oracle_util as O  # noqa: E402
from verilator_b200 import shard as S  # noqa: E402


def main():
    total, evals, out_path = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    rank, world, local = S.env_rank_world()
    dist = S.init_process_group("gloo")
    base, n = S.shard(total, rank, world)
    r = O.mtc_run(n, evals, inst_base=base, seed=7, rand_reset=2, tb_seed=11, per_instance=True,
                  reset_period=64)
    cov = S.allreduce_sum_u64(dist, r.cov)
    uid = S.broadcast_bytes(dist, bytes(range(128)) if rank == 0 else b"", 128)
    tmax = S.allreduce_max(dist, 1.0 + rank)
    # gather the per-shard signatures on rank 0 in rank order (only to check shard invariance)
    import torch
    sigs = torch.zeros(total, dtype=torch.int64)
    sigs[base:base + n] = torch.from_numpy(r.sigs.view(np.int64).copy())
    dist.all_reduce(sigs)
    if rank == 0:
        json.dump({"cov": cov.tolist(), "uid_ok": uid == bytes(range(128)), "tmax": tmax,
                   "sigs": sigs.numpy().view(np.uint64).tolist(), "world": world}, open(out_path, "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
