"""Multi-GPU plumbing of the sampling path: one process per GPU, utterances sharded contiguously, NO collective
on the data path.  The only collectives are the optional final gather of the mels and the
max-over-ranks reduction of a device-measured time; both work on any ``torch.distributed`` backend
(``nccl`` on the GPUs, ``gloo`` in the CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def contiguous_shard(n_items: int, world_size: int, rank: int) -> tuple[int, int]:
    """[lo, hi) of rank's contiguous shard; the first ``n_items % world_size`` ranks take one extra item."""
    if world_size <= 0 or not 0 <= rank < world_size:
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    base, extra = divmod(int(n_items), world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(rank: int, world_size: int, *tensors, dim: int = 0):
    """Slice every tensor (or None) along ``dim`` to this rank's shard of utterances."""
    n = next(t.shape[dim] for t in tensors if t is not None)
    lo, hi = contiguous_shard(n, world_size, rank)
    return tuple(None if t is None else t.narrow(dim, lo, hi - lo).contiguous() for t in tensors)


def max_over_ranks(value: float, device=None, group=None) -> float:
    """Largest ``value`` over the ranks of ``group`` (every multi-GPU time is reported this way)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def gather_shards(local: torch.Tensor, n_items: int, group=None) -> torch.Tensor:
    """All-gather ragged contiguous shards (dim 0) back into the full batch, in rank order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [hi - lo for lo, hi in (contiguous_shard(n_items, world, r) for r in range(world))]
    pad = max(sizes)
    buf = local.new_zeros((pad,) + tuple(local.shape[1:]))
    buf[: local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)
