"""Multi-GPU glue (Python is CI/benchmark glue only; the data path is in the C ABI).

Instances of a batch never communicate, so a job over G GPUs is G independent shards:
rank g owns the contiguous block of global instance indices shard(total, g, G).  Because
every per-instance seed derives from the GLOBAL index (vlcu_create's inst_base), results do
not depend on G.  The single collective of the whole system is the coverage-counter sum
(SURVEY.md 8e): each shard reduces its uint32 counters on the device to uint64 per bin, then
ncclAllReduce(sum) over NVLink inside vlcu_coverage_reduce.  This module holds the rendezvous
pieces around that: rank/world discovery, the block partition, the broadcast of the NCCL
unique id, and max/sum reductions of host scalars over torch.distributed (nccl on GPUs, gloo
in the CPU tests).
"""
import os

import numpy as np


def env_rank_world():
    """(rank, world, local_rank) from the torchrun environment; (0, 1, 0) when not launched."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def bind_to_gpu_numa_node(local_rank):
    """Pin this process to the host cores of the NUMA node GPU `local_rank` hangs off, BEFORE any
    page-locked buffer is allocated: cudaHostAlloc then places the staging buffers in that node's
    memory (first touch), and the host->device copies of an 8-GPU job stop crossing the socket
    interconnect.  Returns the node number, or None when the topology cannot be read (then
    nothing is changed)."""
    import subprocess
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i",
                              str(local_rank)], capture_output=True, text=True, timeout=20).stdout.strip()
        if not bus:
            return None
        dom, rest = bus.split(":", 1)  # nvidia-smi prints an 8-digit domain, sysfs uses 4
        path = f"/sys/bus/pci/devices/{dom[-4:].lower()}:{rest.lower()}/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def shard(total, rank, world):
    """Contiguous block partition: returns (inst_base, n) for `rank`; sizes differ by at most
    one and the blocks tile [0, total) in rank order."""
    if not (0 <= rank < world) or total < 0:
        raise ValueError("bad shard arguments")
    q, r = divmod(total, world)
    n = q + (1 if rank < r else 0)
    base = rank * q + min(rank, r)
    return base, n


def init_process_group(backend=None, local_rank=0):
    """torch.distributed for WORLD_SIZE > 1 (None otherwise).  backend: 'nccl' with a GPU per
    rank, 'gloo' on CPU."""
    rank, world, _ = env_rank_world()
    if world <= 1:
        return None
    import torch
    import torch.distributed as dist
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    kw = {}
    if backend == "nccl":
        torch.cuda.set_device(local_rank)
        kw["device_id"] = torch.device("cuda", local_rank)
    if not dist.is_initialized():
        dist.init_process_group(backend=backend, **kw)
    return dist


def _device(dist, local_rank):
    import torch
    if dist is not None and dist.get_backend() == "nccl":
        return torch.device("cuda", local_rank)
    return torch.device("cpu")


def broadcast_bytes(dist, payload, nbytes, local_rank=0, src=0):
    """Rank `src` supplies `payload` (bytes of length nbytes); every rank gets a copy.  Used for
    the 128-byte ncclUniqueId handed to vlcu_comm_init."""
    if dist is None:
        return bytes(payload)
    import torch
    buf = torch.zeros(nbytes, dtype=torch.uint8, device=_device(dist, local_rank))
    if dist.get_rank() == src:
        if len(payload) != nbytes:
            raise ValueError("payload length")
        buf.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(buf, src=src)
    return bytes(buf.cpu().numpy().tobytes())


def allreduce_max(dist, value, local_rank=0):
    if dist is None:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=_device(dist, local_rank))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def allreduce_sum_u64(dist, values, local_rank=0):
    """Sum of uint64 vectors across ranks through torch.distributed (wrap-around arithmetic, the
    same as ncclUint64/ncclSum).  The product path does this inside vlcu_coverage_reduce with
    NCCL; this host version exists for the gloo tests and for cross-checking."""
    values = np.ascontiguousarray(values, dtype=np.uint64)
    if dist is None:
        return values.copy()
    import torch
    t = torch.from_numpy(values.view(np.int64).copy()).to(_device(dist, local_rank))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy().view(np.uint64)


def attach_coverage_comm(model, dist, local_rank=0):
    """Create the NCCL communicator the model's coverage reduction uses: rank 0 makes the unique
    id, everybody receives it and calls vlcu_comm_init."""
    if dist is None:
        return
    rank, world = dist.get_rank(), dist.get_world_size()
    uid = model.comm_unique_id() if rank == 0 else b""
    uid = broadcast_bytes(dist, uid, 128, local_rank)
    model.comm_init(uid, rank, world)
