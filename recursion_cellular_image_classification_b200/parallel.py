"""Data parallelism of the hot path: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch) as
plumbing.  Replaces torch.nn.DataParallel (reference main.py:94): no per-step parameter broadcast, no GPU-0
gather/loss/optimizer hot spot; every rank holds the full replicated model and optimizer state and the only
exchange per step is the gradient all-reduce, issued in a few large slices while backward is still running.

The same helpers run under the gloo backend on CPU tensors (tests/test_parallel_cpu.py).
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """torchrun-style environment (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, local_rank, world


def default_device():
    """Where a model built without an explicit device lives: the rank's own GPU under torchrun (also made the current
    device, so that main.py's `model.to('cuda')` and `device = 'cuda'` mean that GPU on every rank), plain "cuda" in a
    single process, "cpu" when there is no GPU (host-side surface only: compute still fails loudly)."""
    if not torch.cuda.is_available():
        return "cpu"
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local_rank)
        return "cuda:%d" % local_rank
    return "cuda"


def world_size():
    return dist.get_world_size() if dist.is_initialized() else 1


def rank_world():
    """(rank, world size) of the initialised process group — (0, 1) in a single process."""
    if dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def max_int(value, device):
    """Largest `value` over the ranks (an agreement step for sizes only some ranks know)."""
    if world_size() == 1:
        return int(value)
    t = torch.tensor([int(value)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(t.item())


def shard_range(n_items, rank, world):
    """Contiguous, balanced split of n_items work units (experiments, wells, batches) over ranks."""
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


class PhasedGradAllReduce:
    """Gradient all-reduce interleaved with the executor's backward phases (rxb_dn121_train_step phase p leaves
    the flat-gradient slice `ranges[p]` final).  Slices are few and large (5): on NVSwitch the cost is launch
    latency, not link count."""

    def __init__(self, flat_grad, ranges):
        self.flat_grad = flat_grad
        self.ranges = ranges
        self.works = []

    def after_phase(self, p):
        if world_size() > 1:
            b, e = self.ranges[p]
            if e > b:
                self.works.append(dist.all_reduce(self.flat_grad[b:e], op=dist.ReduceOp.SUM, async_op=True))

    def wait(self):
        for w in self.works:
            w.wait()
        self.works = []


def allreduce_stats(acc):
    """Exact int64 (sum x, sum x^2, pixel count) accumulators of per-experiment statistics, summed over ranks
    when the images of an experiment are split across GPUs."""
    if world_size() > 1:
        for t in acc:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return acc


def allgather_rows(local_rows, counts):
    """Test-time: every rank computed probabilities for its shard of wells; gather [N, C] on all ranks.
    counts[r] = rows held by rank r."""
    if world_size() == 1:
        return local_rows
    C = local_rows.shape[1]
    mx = max(counts)
    pad = torch.zeros(mx, C, dtype=local_rows.dtype, device=local_rows.device)
    pad[:local_rows.shape[0]] = local_rows
    out = [torch.empty_like(pad) for _ in counts]
    dist.all_gather(out, pad)
    return torch.cat([o[:n] for o, n in zip(out, counts)], dim=0)
