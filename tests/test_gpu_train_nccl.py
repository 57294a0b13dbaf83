"""Data-parallel training on two GPUs (skipped on a one-GPU box): every rank runs the Denoiser's forward + library backward
on its own micro-batch with a ``GradSync`` attached; the parameters' gradients must equal the mean of the two ranks'
unsynchronised gradients, and the bucketed all-reduce must not change a bit of the local part of the computation."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir, precision):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from helpers import Case
        from mixgan_tts_b200 import GaussianDiffusion
        from mixgan_tts_b200.grad_sync import GradSync
        c = Case("LJSpeech", "naive", False, 2, 96, wseed=3, iseed=50 + rank, layers=4)      # different data per rank
        gd = GaussianDiffusion(c.args, c.pc, c.mc, c.tc, precision=precision)
        gd.denoise_fn.load_state_dict({k: torch.from_numpy(v) for k, v in c.W.items()})
        gd = gd.cuda().train()
        den = gd.denoise_fn
        x, cond = c.t("x_T").cuda(), c.t("cond").transpose(1, 2).contiguous().cuda()
        t = torch.tensor([3, 1]).cuda()
        r = torch.randn(2, 1, 80, 96, generator=torch.Generator().manual_seed(7)).cuda()

        def grads(sync):
            den.grad_sync = sync
            den.zero_grad(set_to_none=True)
            cr = cond.clone().requires_grad_(True)
            (den(x, t, cr, None) * r).sum().backward()
            torch.cuda.synchronize()
            return torch.cat([p.grad.reshape(-1) for p in den.parameters()]), cr.grad.clone()

        local, lcond = grads(None)
        synced, scond = grads(GradSync(bucket_bytes=1 << 20))                # several buckets for the 4-layer model
        mean = local.clone()
        dist.all_reduce(mean)
        mean /= world
        ok_mean = torch.allclose(synced, mean, rtol=1e-6, atol=1e-7 * float(mean.abs().max()))
        ok_cond = torch.equal(scond, lcond)                                   # input gradients stay local
        flag = torch.tensor([float(ok_mean), float(ok_cond)], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            np.save(os.path.join(out_dir, "ok.npy"), flag.cpu().numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_two_gpu_gradient_average(tmp_path, precision):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path), precision), nprocs=world, join=True)
    ok = np.load(tmp_path / "ok.npy")
    assert ok[0] == 1.0, "synchronised gradients are not the mean over ranks"
    assert ok[1] == 1.0, "the all-reduce changed a local input gradient"


@pytest.mark.parametrize("precision", ["bf16", "fp16", "fp32"])
def test_one_process_drives_two_devices(precision):
    """The > 48 KB shared-memory opt-in of every kernel is a PER-DEVICE attribute and the kernel-layout weight caches are
    per device: the same process must be able to run the module on cuda:0 and on cuda:1 (and nn.DataParallel, which the
    reference wraps the model in, train.py:43-44, must give the single-device result)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import copy
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import Case
    from mixgan_tts_b200 import GaussianDiffusion
    c = Case("LJSpeech", "naive", False, 4, 160, wseed=3, iseed=70, layers=20)
    gd0 = GaussianDiffusion(c.args, c.pc, c.mc, c.tc, precision=precision)
    gd0.denoise_fn.load_state_dict({k: torch.from_numpy(v) for k, v in c.W.items()})
    gd0 = gd0.eval()
    outs = []
    for d in (0, 1):
        gd = copy.deepcopy(gd0).to(f"cuda:{d}")
        to = lambda k: c.t(k).to(f"cuda:{d}")
        with torch.no_grad():
            outs.append(gd(None, to("cond"), None, to("pad_mask"), x_T=to("x_T"), noises=to("noises"))[0].cpu())
    assert torch.equal(outs[0], outs[1])
    # one module object moved from device to device: the caches must follow
    gd = copy.deepcopy(gd0).to("cuda:0")
    with torch.no_grad():
        a = gd(None, c.t("cond").cuda(0), None, c.t("pad_mask").cuda(0), x_T=c.t("x_T").cuda(0), noises=c.t("noises").cuda(0))[0].cpu()
        gd = gd.to("cuda:1")
        b = gd(None, c.t("cond").cuda(1), None, c.t("pad_mask").cuda(1), x_T=c.t("x_T").cuda(1), noises=c.t("noises").cuda(1))[0].cpu()
    assert torch.equal(a, outs[0]) and torch.equal(b, outs[0])
    # nn.DataParallel over the Denoiser (inference): replicas share the module's __dict__ and run in parallel threads
    den = copy.deepcopy(gd0.denoise_fn).to("cuda:0")
    dp = torch.nn.DataParallel(den, device_ids=[0, 1])
    x, cond = c.t("x_T").cuda(0), c.t("cond").transpose(1, 2).contiguous().cuda(0)
    t = torch.tensor([3, 2, 1, 0], device="cuda:0")
    with torch.no_grad():
        single = den(x, t, cond, None)
        both = dp(x, t, cond, None)
    assert torch.equal(single.cpu(), both.cpu())
