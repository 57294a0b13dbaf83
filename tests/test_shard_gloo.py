"""N>1 host logic on CPU: world_size-2 ``gloo`` processes shard a batch by utterance, run the path's CPU checker on
their shard with NO collective in between, and the gathered result equals the unsharded run bit for bit (utterances
never interact).  Also covers the shard arithmetic and the max-over-ranks timing reduction bench.py uses."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mixgan_tts_b200 import shard

from helpers import Case


def test_contiguous_shard_partitions_exactly():
    for n in (0, 1, 7, 64, 513):
        for w in (1, 2, 3, 8):
            spans = [shard.contiguous_shard(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.contiguous_shard(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    try:
        c = Case("LJSpeech", "naive", False, 5, 24, wseed=3, iseed=11, layers=2)      # 5 utterances: ragged shards 3 + 2
        cond, pad, x_T = c.t("cond"), c.t("pad_mask"), c.t("x_T")
        noises = c.t("noises")
        lo, hi = shard.contiguous_shard(c.B, world, rank)
        scond, spad, sx = shard.shard_batch(rank, world, cond, pad, x_T)
        (snoise,) = shard.shard_batch(rank, world, noises, dim=1)
        assert scond.shape[0] == hi - lo
        mel = c.oracle.forward_inference(scond, None, spad, x_T=sx, noises=snoise)[0]
        full = shard.gather_shards(mel, c.B)
        t = shard.max_over_ranks(10.0 + rank)
        if rank == 0:
            ref = c.oracle_forward()[0]
            np.save(os.path.join(out_dir, "ok.npy"),
                    np.array([float(torch.equal(full, ref)), t, full.shape[0]], dtype=np.float64))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_shards_reproduce_the_full_batch(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    ok, tmax, n = np.load(tmp_path / "ok.npy")
    assert ok == 1.0, "gathered shards differ from the unsharded run"
    assert tmax == 11.0, "max over ranks"
    assert n == 5
