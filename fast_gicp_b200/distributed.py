"""Multi-GPU plumbing for replica-style runs (17k-pt registrations do not shard; one rank per GPU, each rank
registers its own pairs, no data-path collective).  torch.distributed is used for the barrier and the max-over-ranks of the
device time only.  Backend-agnostic so the logic is testable on CPU with gloo."""
import os

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend, device=None):
    """Initialise the default process group from the torchrun environment (MASTER_ADDR/PORT, RANK, WORLD_SIZE)."""
    rank, world, _ = env_rank_world()
    if world > 1 and not dist.is_initialized():
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kw)
    return rank, world


def barrier(cuda=False):
    if cuda:
        torch.cuda.synchronize()
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
    if cuda:
        torch.cuda.synchronize()


def max_over_ranks(values, device="cpu"):
    """Element-wise MAX of a list of floats over all ranks (device time of a multi-GPU run is the slowest rank's)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def sum_over_ranks(values, device="cpu"):
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t]


def partition(n_items, rank, world):
    """Contiguous block partition of n_items over world ranks (first n_items % world ranks get one more)."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def aggregate_throughput(local_units, local_ms, device="cpu"):
    """Whole-job throughput [units/s] = units of all ranks / slowest rank's device time."""
    total = sum_over_ranks([local_units], device)[0]
    worst = max_over_ranks([local_ms], device)[0]
    return total / (worst * 1e-3), worst


def finalize():
    if dist.is_initialized():
        dist.destroy_process_group()
