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


def _all_gather_bytes(blob):
    """All-gather of one equal-length byte string per rank (any backend: the tensor lives where the backend wants it)."""
    world = dist.get_world_size()
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    mine = torch.tensor(list(blob), dtype=torch.uint8, device=dev)
    out = [torch.zeros(len(blob), dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(out, mine)
    return [bytes(t.cpu().tolist()) for t in out]


def setup_source_sharding(core, n_source, max_points=0):
    """Wire one handle per rank for a sharded registration: exchange the IPC handles of the result mailboxes (and of
    the covariance arena when max_points > 0 -> stage 1 sharded too), set this rank's slice of the source, barrier.
    Returns (begin, end) of the slice.  The process group must be initialised (one rank per GPU)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    core.comm_init(rank, world, _all_gather_bytes(core.comm_export()))
    if max_points > 0:
        core.comm_init_arena(_all_gather_bytes(core.comm_export_arena(max_points)))
    lo, hi = partition(n_source, rank, world)
    core.set_source_shard(lo, hi)
    dist.barrier()  # every mailbox / arena is cleared and mapped before anyone launches into it
    if max_points > 0:
        core.set_stage1_sharding(True)
    return lo, hi


def finalize():
    if dist.is_initialized():
        dist.destroy_process_group()
