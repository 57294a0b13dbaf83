"""Host logic of the data-parallel training path on CPU: gradient-bucket planning over the library's backward segments
and a world_size-2 ``gloo`` run of ``GradSync`` that averages buckets exactly as one all-reduce of the whole vector."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mixgan_tts_b200 import _lib
from mixgan_tts_b200.grad_sync import GradSync, plan_buckets


def segment_ranges(multi=0, layers=20):
    lib = _lib.load()
    d = _lib.ModelDims(80, 256, 256, layers, multi)
    out = []
    for s in range(lib.mgb_train_segments(C.byref(d))):
        b, e = C.c_size_t(0), C.c_size_t(0)
        assert lib.mgb_train_segment_range(C.byref(d), s, C.byref(b), C.byref(e)) == 0
        out.append((b.value, e.value))
    return out, lib.mgb_flat_weight_count(C.byref(d))


@pytest.mark.parametrize("multi,layers", [(0, 20), (1, 20), (0, 1), (1, 3)])
def test_segments_tile_the_flat_gradient_back_to_front(multi, layers):
    ranges, total = segment_ranges(multi, layers)
    assert len(ranges) == layers + 2
    assert ranges[0][1] == total and ranges[-1][0] == 0
    assert all(a[0] == b[1] for a, b in zip(ranges, ranges[1:]))       # each segment ends where the previous began
    assert sum(e - b for b, e in ranges) == total


@pytest.mark.parametrize("bucket_bytes", [None, 1, 4 << 20, 16 << 20, 1 << 30])
def test_bucket_plan_covers_every_segment_once(bucket_bytes):
    ranges, total = segment_ranges()
    plan = plan_buckets(ranges, bucket_bytes)
    assert plan[0][0] == 0 and plan[-1][1] == len(ranges)
    assert all(a[1] == b[0] for a, b in zip(plan, plan[1:]))
    assert all(a[2] == b[3] for a, b in zip(plan, plan[1:]))           # flat slices are adjacent, back to front
    assert sum(fe - fb for _, _, fb, fe in plan) == total
    for sb, se, fb, fe in plan:
        assert fb == min(ranges[s][0] for s in range(sb, se)) and fe == max(ranges[s][1] for s in range(sb, se))
    if bucket_bytes == 1:
        assert len(plan) == len(ranges)
    if bucket_bytes in (None, 1 << 30):
        assert len(plan) == 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ranges, total = segment_ranges(0, 2)
        g = torch.Generator().manual_seed(100 + rank)
        flat = torch.randn(total, generator=g)
        mine = flat.clone()
        sync = GradSync(bucket_bytes=1 << 20)
        for sb, se, fb, fe in plan_buckets(ranges, sync.bucket_bytes):
            sync.reduce_async(mine[fb:fe])         # in the product this follows mgb_denoiser_backward(seg sb..se)
        sync.finish()
        ref = flat.clone()
        dist.all_reduce(ref)
        ref /= world
        if rank == 0:
            np.save(os.path.join(out_dir, "ok.npy"),
                    np.array([float(torch.allclose(mine, ref, rtol=0, atol=1e-7)), sync.bytes_reduced, total * 4], dtype=np.float64))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_bucketed_average_equals_one_allreduce(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    ok, reduced, total_bytes = np.load(tmp_path / "ok.npy")
    assert ok == 1.0
    assert reduced == total_bytes
