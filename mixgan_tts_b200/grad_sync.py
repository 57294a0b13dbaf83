"""Data-parallel gradient all-reduce for the Denoiser's training path (BASELINE configs[4]; the reference
wraps the whole model in ``nn.DataParallel``, ``train.py:43-44`` — here one process per GPU).

The library's backward finishes the flat gradient in contiguous slices (tail, residual blocks from the
last to the first, head).  ``plan_buckets`` groups consecutive slices into buckets of about ``bucket_bytes``;
``GradSync.reduce_async`` starts the all-reduce of a finished bucket on the process group's communication
stream (NCCL over NVLink) while the compute stream continues with the next bucket, and ``finish`` makes the
compute stream wait for all of them and turns sums into means.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def plan_buckets(ranges, bucket_bytes=None, elem_bytes: int = 4):
    """``ranges[s] = (flat_begin, flat_end)`` of backward segment ``s`` (execution order).  Returns a list of
    ``(seg_begin, seg_end, flat_begin, flat_end)``: consecutive segments whose slices are adjacent, merged until a
    bucket holds at least ``bucket_bytes`` (``None`` = everything in one bucket)."""
    if not ranges:
        return []
    if bucket_bytes is None:
        lo = min(b for b, _ in ranges)
        hi = max(e for _, e in ranges)
        return [(0, len(ranges), lo, hi)]
    out = []
    sb, lo, hi = 0, ranges[0][0], ranges[0][1]
    for s in range(1, len(ranges)):
        b, e = ranges[s]
        adjacent = (e == lo) or (b == hi)
        if adjacent and (hi - lo) * elem_bytes < bucket_bytes:
            lo, hi = min(lo, b), max(hi, e)
            continue
        out.append((sb, s, lo, hi))
        sb, lo, hi = s, b, e
    out.append((sb, len(ranges), lo, hi))
    return out


class GradSync:
    """Averages gradient buckets over a process group, overlapped with the rest of the backward."""

    def __init__(self, process_group=None, bucket_bytes: int = 16 << 20):
        if not dist.is_initialized():
            raise RuntimeError("GradSync needs an initialised torch.distributed process group")
        self.group = process_group
        self.world = dist.get_world_size(process_group)
        self.bucket_bytes = int(bucket_bytes)
        self._pending = []
        self.bytes_reduced = 0

    def reduce_async(self, bucket: torch.Tensor):
        """Start ``sum`` over ranks of ``bucket`` (a contiguous slice of the flat gradient), in place."""
        if self.world == 1:
            return
        work = dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self._pending.append((work, bucket))
        self.bytes_reduced += bucket.numel() * bucket.element_size()

    def finish(self):
        """Wait (stream-ordered on CUDA) for every started bucket and scale sums to means."""
        for work, bucket in self._pending:
            work.wait()
            bucket.mul_(1.0 / self.world)
        self._pending = []
