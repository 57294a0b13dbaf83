"""CPU-only checks: the C-ABI library loads and exports every symbol the header declares, the
drop-in modules mirror the reference's interface (state_dict keys, buffers, signatures), and the
product path refuses to run without a GPU instead of falling back."""
import ctypes as C
import inspect
import os
import re

import numpy as np
import pytest
import torch

from mixgan_tts_b200 import GaussianDiffusion, Denoiser, _lib, configs, synth
from mixgan_tts_b200.schedule import noise_schedule_list, posterior_buffers
from oracle import ref_loader, schedule as oracle_schedule

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "mixgan_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mgb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 13
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/mixgan_b200.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes prototype"
    assert lib.mgb_abi_version() == 2


def test_sizes_are_computable_without_gpu():
    lib = _lib.load()
    for multi in (0, 1):
        dims = _lib.ModelDims(80, 256, 256, 20, multi)
        n = lib.mgb_flat_weight_count(C.byref(dims))
        assert n == (13764176 if not multi else 15074896)      # SURVEY.md §8a parameter counts
        assert lib.mgb_packed_bytes(C.byref(dims), 0) >= n * 4
        assert lib.mgb_workspace_bytes(C.byref(dims), 0, 64, 800, 4) > 0
    assert lib.mgb_workspace_bytes(C.byref(dims), 0, 0, 800, 4) == 0


def test_no_cpu_fallback():
    lib = _lib.load()
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert lib.mgb_device_check(0) == _lib.E_ARCH
    args, pc, mc, tc = configs.make_configs()
    gd = GaussianDiffusion(args, pc, mc, tc)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        gd(None, torch.zeros(1, 8, 256), None, torch.zeros(1, 8, dtype=torch.bool))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        gd.denoise_fn(torch.zeros(1, 1, 80, 8), torch.zeros(1, dtype=torch.long), torch.zeros(1, 256, 8), None)


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "mixgan_tts_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn


@pytest.mark.parametrize("model,K", [("naive", 4), ("shallow", 1)])
def test_buffers_match_oracle_schedule_bitwise(model, K):
    args, pc, mc, tc = configs.make_configs("LJSpeech", model)
    gd = GaussianDiffusion(args, pc, mc, tc)
    assert gd.num_timesteps == K
    ref = oracle_schedule.diffusion_buffers(oracle_schedule.noise_schedule("vpsde", K, 0.1, 40, 0.008))
    for k, v in ref.items():
        assert np.array_equal(getattr(gd, k).numpy().view(np.uint32), v.view(np.uint32)), k
    for mode in ("linear", "cosine", "vpsde"):
        a = noise_schedule_list(mode, 6, 0.1, 0.5, 0.008)
        b = oracle_schedule.noise_schedule(mode, 6, 0.1, 0.5, 0.008)
        assert np.array_equal(a, b)
        pa, pb = posterior_buffers(a), oracle_schedule.diffusion_buffers(b)
        assert all(np.array_equal(pa[k].view(np.uint32), pb[k].view(np.uint32)) for k in pa)
    with pytest.raises(NotImplementedError):
        noise_schedule_list("nope", 4)


@pytest.mark.parametrize("multi", [False, True])
def test_state_dict_keys_and_shapes(multi):
    args, pc, mc, tc = configs.make_configs("LJSpeech", "naive", multi)
    gd = GaussianDiffusion(args, pc, mc, tc)
    W = synth.make_denoiser_weights(0, multi_speaker=multi)
    sd = gd.denoise_fn.state_dict()
    assert set(sd) == set(W)
    for k in W:
        assert tuple(sd[k].shape) == W[k].shape, k
    assert float(gd.denoise_fn.output_projection.conv.weight.abs().max()) == 0.0   # modules.py:418
    gd.denoise_fn.load_state_dict({k: torch.from_numpy(v) for k, v in W.items()}, strict=True)
    if ref_loader.available():
        ref = ref_loader.load().GaussianDiffusion(args, pc, mc, tc)
        rsd = ref.state_dict()
        assert set(rsd) == set(gd.state_dict())
        for k in rsd:
            assert rsd[k].shape == gd.state_dict()[k].shape, k
            assert rsd[k].dtype == gd.state_dict()[k].dtype, k
        gd.load_state_dict(rsd, strict=True)          # a reference checkpoint drops in


def test_method_signatures_cover_the_reference():
    if not ref_loader.available():
        pytest.skip("reference not present")
    ref = ref_loader.load().GaussianDiffusion
    for name in ("forward", "sampling", "p_sample", "q_posterior", "q_posterior_sample", "q_sample",
                 "diffuse_fn", "diffuse_trace", "interpolate", "norm_spec", "denorm_spec", "out2mel",
                 "q_mean_variance", "predict_start_from_noise"):
        rp = list(inspect.signature(getattr(ref, name)).parameters)
        mp = list(inspect.signature(getattr(GaussianDiffusion, name)).parameters)
        assert mp[:len(rp)] == rp, (name, rp, mp)
    rp = list(inspect.signature(ref_loader.load().Denoiser.forward).parameters)
    assert list(inspect.signature(Denoiser.forward).parameters)[:len(rp)] == rp


def test_light_helpers_match_oracle_on_cpu():
    """q_sample / q_posterior / norm / denorm are plain torch and run anywhere."""
    from helpers import Case
    c = Case("LJSpeech", "naive", False, 2, 16, 0, 3, layers=1)
    gd = GaussianDiffusion(c.args, c.pc, c.mc, c.tc)
    x0, xt, nz = c.t("x_T"), c.t("noises")[0], c.t("noises")[1]
    t = torch.tensor([3, 0])
    a = gd.q_posterior_sample(x0, xt, t, noise=nz)
    b = c.oracle.q_posterior_sample(x0, xt, t, nz)
    assert torch.equal(a, b)
    assert torch.equal(gd.q_sample(x0, t, nz), c.oracle.q_sample(x0, t, nz))
    m = torch.randn(2, 16, 80)
    assert torch.equal(gd.norm_spec(m), c.oracle.norm_spec(m))
    assert torch.equal(gd.denorm_spec(m), c.oracle.denorm_spec(m))
    tt = torch.tensor([2, -1])
    assert torch.equal(gd.diffuse_fn(m, tt.clone(), noise=nz), c.oracle.diffuse_fn(m, tt.clone(), nz))


def test_weight_cache_fingerprint_sees_fused_optimizer_steps():
    """torch's fused optimizers update parameters in place WITHOUT bumping Tensor._version; the kernel-layout weight
    caches must still be invalidated (an optimizer post-step hook advances an epoch that is part of the key) - but only
    by optimizers that own a Denoiser parameter (the discriminator's optimizer must not force a repack)."""
    from mixgan_tts_b200 import Denoiser, configs
    from mixgan_tts_b200.modules import _param_fingerprint
    _, pc, mc, _ = configs.make_configs("LJSpeech", "naive", residual_layers=1)
    den = Denoiser(pc, mc)
    ps = den._ordered_params()
    for p in ps:
        p.grad = torch.randn_like(p)
    fp0 = _param_fingerprint(ps)
    assert _param_fingerprint(ps) == fp0
    other = [torch.nn.Parameter(torch.randn(8))]
    other[0].grad = torch.randn(8)
    torch.optim.SGD(other, lr=1e-2).step()
    assert _param_fingerprint(ps) == fp0              # an unrelated optimizer leaves the caches valid
    try:
        opt = torch.optim.Adam(ps, lr=1e-2, fused=True)
    except (RuntimeError, ValueError):
        opt = torch.optim.Adam(ps, lr=1e-2)
    opt.step()
    assert _param_fingerprint(ps) != fp0
    # writes through .data bump nothing: invalidate_packed() is the documented escape hatch
    e0 = den._pack_epoch
    den.invalidate_packed()
    assert den._pack_epoch == e0 + 1 and not den._packed and not den._flat
