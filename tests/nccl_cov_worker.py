"""One rank of an N-GPU coverage campaign (launched by torchrun, NCCL backend): the ONE collective
of the system -- ncclAllReduce(u64, sum) of the per-shard coverage sums inside
vlcu_coverage_reduce -- checked against the reference semantics
(include/verilated_cov.cpp:368-438: uint32 items summed into uint64 per key): every rank must end
up with the sum over ALL instances of the genuine C++ model's counters, and rank 0's merged
coverage.dat must be the file the reference writer produces for those sums.

usage: torchrun --nproc-per-node N tests/nccl_cov_worker.py <instances per rank> <evals> <rounds> <out.json>"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import oracle_util as O  # noqa: E402
from verilator_b200 import abi, shard as S  # noqa: E402


def main():
    n, evals, rounds, out_path = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    rank, world, local = S.env_rank_world()
    dist = S.init_process_group("nccl", local)
    with abi.Model("rv32i_cov", n, device=local, inst_base=rank * n, seed=5, rand_reset=2) as m:
        S.attach_coverage_comm(m, dist, local)  # vlcu_comm_unique_id -> broadcast -> vlcu_comm_init
        for k in range(rounds):
            m.run("rv32_rand", evals, flags=abi.RUNF_TB_BEGIN, tb_seed=40 + k, tb_flags=1, tb_arg=(6, 0, 0, 0))
        m.sync()
        t0 = time.perf_counter()
        sums = m.coverage_reduce()  # k_cov_reduce + ncclAllReduce + copy back
        reduce_ms = 1e3 * (time.perf_counter() - t0)
        # the same merge through torch.distributed, from the local sums of a second model without
        # a communicator: cross-checks NCCL's result with an independent path
        with abi.Model("rv32i_cov", n, device=local, inst_base=rank * n, seed=5, rand_reset=2) as m2:
            for k in range(rounds):
                m2.run("rv32_rand", evals, flags=abi.RUNF_TB_BEGIN, tb_seed=40 + k, tb_flags=1, tb_arg=(6, 0, 0, 0))
            local_sums = m2.coverage_reduce()
        torch_sums = S.allreduce_sum_u64(dist, local_sums, local)
        ok_torch = bool((torch_sums == sums).all())
        # every rank checks the merged sums against the oracle's sum over ALL instances
        want = O.rv32_run(n * world, evals, seed=5, rand_reset=2, tb_seed=40, reset_time=6, rounds=rounds, cov=True,
                          threads=os.cpu_count() // max(world, 1) or 1)
        ok_oracle = bool((sums == want.cov).all())
        dat_ok = None
        if rank == 0:
            gpu_path = out_path + ".gpu.dat"
            m.coverage_write(sums, gpu_path)
            # the reference writer, fed the same merged counts: rank 0 re-runs ONE instance to get the
            # file layout, then compares every point line's count with the merged sums
            ref_path = out_path + ".ref.dat"
            O.rv32_run(1, evals, seed=5, rand_reset=2, tb_seed=40, reset_time=6, rounds=rounds, cov=True,
                       cov_path=ref_path, cov_instance=0)
            ref_lines = open(ref_path).read().splitlines()
            gpu_lines = open(gpu_path).read().splitlines()
            # same keys in the same order; counts are the merged ones (>= the single instance's)
            keys = lambda lines: [ln.rsplit(" ", 1)[0] for ln in lines]
            dat_ok = keys(ref_lines) == keys(gpu_lines) and sum(int(ln.rsplit(" ", 1)[1]) for ln in gpu_lines[1:]) > 0
    flags = S.allreduce_max(dist, 0.0 if (ok_torch and ok_oracle) else 1.0, local)
    if rank == 0:
        json.dump({"world": world, "instances_per_rank": n, "bins": int(len(sums)), "nccl_equals_torch": ok_torch,
                   "nccl_equals_oracle_sum_over_all_instances": ok_oracle, "all_ranks_ok": flags == 0.0,
                   "coverage_dat_keys_equal_reference_writer": dat_ok, "reduce_ms": reduce_ms,
                   "total_count": int(sums.sum())}, open(out_path, "w"))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if flags != 0.0:
        raise SystemExit(1)


if __name__ == "__main__":
    main()
