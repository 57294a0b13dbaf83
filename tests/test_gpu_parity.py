"""Parity of the CUDA path (through the C ABI / drop-in modules) against the CPU oracle and the
reference-made golden vectors, on identical inputs and identical injected noise.

Tolerances (relative L2, checked on the NORMALISED x0 as well as on the returned mel, because the
denormalised mel has a large DC offset that flatters the ratio — SURVEY.md §7):
  fp32 mode (CUDA cores, exact fp32 operands): north_star bar 1e-3; this suite asserts 1e-4.
  fp16 mode (the reference-precision mode ON THE TENSOR CORES: fp16 operands = TF32's significand, fp32 accumulate and
             streams): asserts the north_star fp32 bar, 1e-3 on normalised x0 (measured 2.5-4.5e-4), 1e-4 on the mel.
  bf16 mode: stated tolerance 8e-3 on normalised x0 (measured 3.4e-3), 1e-3 on the denormalised mel (measured 2.8e-4).
"""
import ctypes as C

import numpy as np
import pytest
import torch

from mixgan_tts_b200 import GaussianDiffusion, _lib
from mixgan_tts_b200.length_regulator import LengthRegulator

from helpers import GOLDEN_CASES, TRAIN_CASES, TRAIN_KEYS, Case, golden_case, load_golden, rel_l2, train_case

pytestmark = pytest.mark.gpu

TOL = {"fp32": dict(norm=1e-4, mel=1e-4), "fp16": dict(norm=1e-3, mel=1e-4), "bf16": dict(norm=8e-3, mel=1e-3)}


def available_precisions():
    lib = _lib.load()
    dims = _lib.ModelDims(80, 256, 256, 20, 0)
    return [p for p, code in (("fp32", 0), ("bf16", 1), ("fp16", 3)) if lib.mgb_packed_bytes(C.byref(dims), code) > 0]


PRECS = available_precisions()


def build(case: Case, precision: str) -> GaussianDiffusion:
    gd = GaussianDiffusion(case.args, case.pc, case.mc, case.tc, precision=precision)
    gd.denoise_fn.load_state_dict({k: torch.from_numpy(v) for k, v in case.W.items()}, strict=True)
    return gd.cuda().eval()


def cu(t):
    return None if t is None else t.cuda()


@pytest.mark.parametrize("precision", PRECS)
@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_denoiser_forward_vs_golden_and_oracle(name, precision):
    g, c = load_golden(name), golden_case(name)
    gd = build(c, precision)
    t = torch.from_numpy(g["denoiser_t"]).cuda()
    with torch.no_grad():            # the inference kernels of this precision
        out = gd.denoise_fn(cu(c.t("x_T")), t, cu(c.t("cond")).transpose(1, 2), cu(c.t("spk")))
    assert out.shape == g["denoiser_out"].shape
    assert rel_l2(out, g["denoiser_out"]) < TOL[precision]["norm"]
    # with autograd on (parameters require grad) the same call is the training forward (stash-keeping kernels)
    out_t = gd.denoise_fn(cu(c.t("x_T")), t, cu(c.t("cond")).transpose(1, 2), cu(c.t("spk")))
    assert out_t.requires_grad
    assert rel_l2(out_t, g["denoiser_out"]) < TOL[gd.denoise_fn.train_precision]["norm"]


@pytest.mark.parametrize("precision", PRECS)
@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_p_sample_vs_golden(name, precision):
    g, c = load_golden(name), golden_case(name)
    gd = build(c, precision)
    K = gd.num_timesteps
    t = torch.full((c.B,), K - 1, dtype=torch.long, device="cuda")
    out, x0 = gd.p_sample(cu(c.t("x_T")), t, cu(c.t("cond")).transpose(1, 2), cu(c.t("spk")),
                          noise=cu(c.t("noises"))[K - 1], return_x0=True)
    assert rel_l2(out, g["p_sample_out"]) < TOL[precision]["norm"]
    assert float(x0.abs().max()) <= 1.0


@pytest.mark.parametrize("precision", PRECS)
@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_full_inference_vs_golden_and_oracle(name, precision):
    g, c = load_golden(name), golden_case(name)
    gd = build(c, precision)
    res = gd(None, cu(c.t("cond")), cu(c.t("spk")), cu(c.t("pad_mask")), coarse_mel=cu(c.t("coarse_mel")),
             x_T=cu(c.t("x_T")), noises=cu(c.t("noises")), start_noise=cu(c.t("start_noise")))
    mel = res[0]
    assert res[1] is None and res[2] is None and res[3] is None
    assert rel_l2(mel, g["final_mel"]) < TOL[precision]["mel"]
    final, states, x0s, start = c.oracle_forward()
    assert rel_l2(mel, final) < TOL[precision]["mel"]
    # normalised-domain check on the valid frames
    valid = ~c.t("pad_mask")
    x0_ref = x0s[-1][:, 0].transpose(1, 2)[valid]
    x0_gpu = gd.norm_spec(mel).cpu()[valid]
    assert rel_l2(x0_gpu, x0_ref) < TOL[precision]["norm"] * 2
    pad = c.t("pad_mask")
    if pad.any():
        assert float(mel.cpu()[pad].abs().max()) == 0.0


@pytest.mark.parametrize("precision", PRECS)
def test_sampling_states_match_reference_structure(precision):
    c = golden_case("naive_lj_B2_T64")
    g = load_golden("naive_lj_B2_T64")
    gd = build(c, precision)
    gd.cond, gd.spk_emb = cu(c.t("cond")).transpose(1, 2), None
    states = gd.sampling(noise=cu(c.t("x_T")), noises=cu(c.t("noises")))
    assert len(states) == gd.num_timesteps + 1 and tuple(states[0].shape) == (2, 64, 80)
    assert rel_l2(states[1], g["state_after_first_step"]) < TOL[precision]["mel"]
    # state 0 is denorm(x_T) exactly
    ref0 = c.oracle.denorm_spec(c.t("x_T")[:, 0].transpose(1, 2))
    assert rel_l2(states[0], ref0) < 1e-6


@pytest.mark.parametrize("precision", PRECS)
@pytest.mark.parametrize("B,T", [(1, 1), (1, 2), (2, 3), (1, 127), (1, 128), (2, 129), (1, 257)])
def test_edge_shapes_vs_oracle(B, T, precision):
    c = Case("LJSpeech", "naive", False, B, T, wseed=2, iseed=100 + T, layers=20)
    gd = build(c, precision)
    mel = gd(None, cu(c.t("cond")), None, cu(c.t("pad_mask")), x_T=cu(c.t("x_T")), noises=cu(c.t("noises")))[0]
    final, _, _, _ = c.oracle_forward()
    assert rel_l2(mel, final) < TOL[precision]["mel"]


@pytest.mark.parametrize("precision", PRECS)
def test_multi_speaker_requires_embedding(precision):
    c = Case("AISHELL3", "shallow", True, 1, 16, wseed=1, iseed=2)
    gd = build(c, precision)
    with pytest.raises(TypeError):
        gd.denoise_fn(cu(c.t("x_T")), torch.zeros(1, dtype=torch.long, device="cuda"),
                      cu(c.t("cond")).transpose(1, 2), None)


@pytest.mark.parametrize("precision", PRECS)
def test_shard_equivalence_and_determinism_full_size(precision):
    """BASELINE config-2 shape (B=64 would take the oracle minutes): size-independent properties.
    (a) two runs are bit-identical; (b) utterance shards reproduce the full batch bit for bit (the
    multi-GPU sharding contract); (c) output is finite and padded frames are zero."""
    c = Case("LJSpeech", "naive", False, 16, 800, wseed=0, iseed=1234)
    gd = build(c, precision)
    args = dict(x_T=cu(c.t("x_T")), noises=cu(c.t("noises")))
    cond, pad = cu(c.t("cond")), cu(c.t("pad_mask"))
    a = gd(None, cond, None, pad, **args)[0]
    b = gd(None, cond, None, pad, **args)[0]
    assert torch.equal(a, b)
    halves = []
    for sl in (slice(0, 8), slice(8, 16)):
        halves.append(gd(None, cond[sl], None, pad[sl], x_T=args["x_T"][sl].contiguous(),
                         noises=args["noises"][:, sl].contiguous())[0])
    assert torch.equal(torch.cat(halves), a)
    assert torch.isfinite(a).all()
    assert float(a[pad].abs().max()) == 0.0
    # one utterance against the oracle at full length
    c1 = Case("LJSpeech", "naive", False, 1, 800, wseed=0, iseed=77)
    m1 = gd(None, cu(c1.t("cond")), None, cu(c1.t("pad_mask")), x_T=cu(c1.t("x_T")), noises=cu(c1.t("noises")))[0]
    assert rel_l2(m1, c1.oracle_forward()[0]) < TOL[precision]["mel"]


@pytest.mark.parametrize("precision", PRECS)
def test_batch_of_full_length_utterances_vs_oracle(precision):
    """B=8 x T=800 against the oracle in one piece: 8 x 801 rows are 28 tiles of 236 output rows, most of which span two
    utterances (the shard test above only compares the GPU with itself at this size)."""
    c = Case("LJSpeech", "naive", False, 8, 800, wseed=0, iseed=31)
    gd = build(c, precision)
    mel = gd(None, cu(c.t("cond")), None, cu(c.t("pad_mask")), x_T=cu(c.t("x_T")), noises=cu(c.t("noises")))[0]
    final, _, x0s, _ = c.oracle_forward()
    assert rel_l2(mel, final) < TOL[precision]["mel"]
    valid = ~c.t("pad_mask")
    # per utterance, so that one bad tile cannot hide in the batch norm
    for b in range(c.B):
        x0_ref = x0s[-1][b, 0].transpose(0, 1)[valid[b]]
        x0_gpu = gd.norm_spec(mel[b:b + 1]).cpu()[0][valid[b]]
        assert rel_l2(x0_gpu, x0_ref) < TOL[precision]["norm"] * 2, b


# Dynamic range: a trained checkpoint is not a random-init one.  Weights, biases and the conditioner are scaled up so that
# the residual stream, the skip sum and the pre-activations are several times larger than at initialisation (saturated
# gates, |u| of a few tens), which exercises the 16-bit group spills and the constants folded into the packed weights.
# Tolerances: the same stated bars; the measured values are printed (profiles/r02/parity_report.txt).
# measured (profiles/r02/parity_report.txt), worst over the four stresses: fp32 1.2e-6 / 8e-7, fp16 1.13e-3 / 5.1e-4,
# bf16 8.9e-3 / 4.1e-3 (unclamped single call / K-step mel); stated = about twice that
TOL_STRESS = {"fp32": dict(norm=1e-4, mel=1e-4), "fp16": dict(norm=2.5e-3, mel=1e-3), "bf16": dict(norm=2e-2, mel=8e-3)}


@pytest.mark.parametrize("precision", PRECS)
@pytest.mark.parametrize("stress", ["weights_x4", "biases_pm8", "cond_x4", "all"])
def test_dynamic_range_vs_oracle(stress, precision):
    c = Case("LJSpeech", "naive", False, 2, 300, wseed=5, iseed=61)
    rg = np.random.default_rng(3)
    W = dict(c.W)
    if stress in ("weights_x4", "all"):
        for k in W:
            if k.endswith("output_projection.conv.weight") and k.startswith("residual_layers"):
                W[k] = W[k] * np.float32(4.0)
            elif k.endswith("conditioner_projection.conv.weight") or k.endswith("conv_layer.conv.weight"):
                W[k] = W[k] * np.float32(2.0)
    if stress in ("biases_pm8", "all"):
        for k in W:
            if k.endswith(".bias") and k.startswith("residual_layers"):
                W[k] = W[k] + rg.uniform(-8.0, 8.0, W[k].shape).astype(np.float32)
    if stress in ("cond_x4", "all"):
        c.inp = dict(c.inp)
        c.inp["cond"] = c.inp["cond"] * np.float32(4.0)
    c.W = W
    from oracle.diffusion import DiffusionOracle
    from mixgan_tts_b200 import configs
    c.oracle = DiffusionOracle(W, model="naive", denoiser_cfg=c.mc["denoiser"], spec_min=[configs.SPEC_MIN] * 80,
                               spec_max=[configs.SPEC_MAX] * 80)
    gd = build(c, precision)
    x, cond = cu(c.t("x_T")), cu(c.t("cond"))
    t = torch.tensor([3, 1], device="cuda")
    with torch.no_grad():
        out = gd.denoise_fn(x, t, cond.transpose(1, 2), None)            # unclamped network output: no clip to hide behind
    from oracle.denoiser import denoiser_forward
    ref = denoiser_forward(c.oracle.W, c.t("x_T"), t.cpu(), c.t("cond").transpose(1, 2), None)
    assert torch.isfinite(out).all()
    e = rel_l2(out, ref)
    print(f"dynamic range [{stress}] {precision}: |ref| rms {float(ref.pow(2).mean().sqrt()):.3f}  rel L2 {e:.3e}")
    assert e < TOL_STRESS[precision]["norm"], (stress, e)
    mel = gd(None, cond, None, cu(c.t("pad_mask")), x_T=x, noises=cu(c.t("noises")))[0]
    final = c.oracle_forward()[0]
    e_mel = rel_l2(mel, final)
    print(f"dynamic range [{stress}] {precision}: K-step mel rel L2 {e_mel:.3e}")
    assert e_mel < TOL_STRESS[precision]["mel"], (stress, e_mel)


@pytest.mark.parametrize("precision", PRECS)
def test_long_multispeaker_utterances_config4(precision):
    """BASELINE configs[3] shape class: AISHELL3 shallow, multi-speaker, T = 1500 (not a multiple of any tile size).
    One utterance against the oracle at full length, then size-independent properties on a batch: determinism and
    bit-exact utterance sharding (a tile may span two utterances on the batch row axis)."""
    c1 = Case("AISHELL3", "shallow", True, 1, 1500, wseed=7, iseed=41)
    gd = build(c1, precision)
    kw = lambda c: dict(coarse_mel=cu(c.t("coarse_mel")), x_T=cu(c.t("x_T")), noises=cu(c.t("noises")),
                        start_noise=cu(c.t("start_noise")))
    m1 = gd(None, cu(c1.t("cond")), cu(c1.t("spk")), cu(c1.t("pad_mask")), **kw(c1))[0]
    assert rel_l2(m1, c1.oracle_forward()[0]) < TOL[precision]["mel"]
    c = Case("AISHELL3", "shallow", True, 6, 1500, wseed=7, iseed=42)
    cond, spk, pad = cu(c.t("cond")), cu(c.t("spk")), cu(c.t("pad_mask"))
    a = gd(None, cond, spk, pad, **kw(c))[0]
    assert torch.equal(a, gd(None, cond, spk, pad, **kw(c))[0])
    parts = []
    for sl in (slice(0, 1), slice(1, 4), slice(4, 6)):
        k = kw(c)
        parts.append(gd(None, cond[sl], spk[sl], pad[sl], coarse_mel=k["coarse_mel"][sl].contiguous(),
                        x_T=k["x_T"][sl].contiguous(), noises=k["noises"][:, sl].contiguous(),
                        start_noise=k["start_noise"][sl].contiguous())[0])
    assert torch.equal(torch.cat(parts), a)
    assert torch.isfinite(a).all() and float(a[pad].abs().max()) == 0.0


@pytest.mark.parametrize("precision", PRECS)
def test_sampling_loop_matches_step_by_step_p_sample(precision):
    """`forward`/`sampling` (all K steps prepared once, uniform timestep, constant k_l added inside the GEMM) must agree
    with K explicit `p_sample` calls (per-utterance timesteps, k_l added by the epilogue): two different kernels."""
    c = Case("LJSpeech", "naive", False, 3, 300, wseed=0, iseed=8)
    gd = build(c, precision)
    cond, pad = cu(c.t("cond")), cu(c.t("pad_mask"))
    x, noises = cu(c.t("x_T")), cu(c.t("noises"))
    mel = gd(None, cond, None, pad, x_T=x, noises=noises)[0]
    for i in reversed(range(gd.num_timesteps)):
        t = torch.full((c.B,), i, dtype=torch.long, device="cuda")
        x = gd.p_sample(x, t, cond.transpose(1, 2), None, noise=noises[i])
    ref = gd.denorm_spec(x[:, 0].transpose(1, 2)) * (~pad).unsqueeze(-1)
    assert rel_l2(mel, ref) < {"fp32": 1e-5, "fp16": 1e-4, "bf16": 2e-3}[precision]


@pytest.mark.parametrize("precision", PRECS)
@pytest.mark.parametrize("name", list(TRAIN_CASES))
def test_training_branch_forward_values_vs_golden(name, precision):
    """`forward(mel given)` under no_grad (evaluate.py's use): the reference's 5-tuple, every draw injected."""
    g = load_golden(name)
    c, ex = train_case(name)
    gd = build(c, precision)
    with torch.no_grad():
        out = gd(cu(ex["mel"]), cu(c.t("cond")), cu(c.t("spk")), cu(c.t("pad_mask")), coarse_mel=cu(c.t("coarse_mel")),
                 t=cu(ex["t"]), noise_t=cu(ex["noise_t"]), noise_prev=cu(ex["noise_prev"]), post_noise=cu(ex["post_noise"]))
    assert torch.equal(out[4].cpu(), torch.from_numpy(g["t"]))
    for k, v in zip(TRAIN_KEYS, out):
        tol = 1e-5 if k in ("x_t", "x_t_prev") else TOL[precision]["norm"]     # the two diffused states involve no network
        assert rel_l2(v, g[k]) < tol, k


@pytest.mark.parametrize("B,T", [(3, 48), (2, 131), (1, 1)])
def test_training_branch_fused_kernels_vs_torch_composition(B, T):
    """mgb_train_diffuse / mgb_train_posterior(+backward) against the torch composition of diffusion.py:206-220 they replace:
    the two diffused states bit for bit, the posterior outputs and d/d denoiser_out to fp32 rounding (the sigma table is
    exponentiated on the host), with t = 0 rows (x_t_prev = x_start, no posterior noise), padding and clamped values."""
    from mixgan_tts_b200.diffusion import _PosteriorFn
    c, _ = train_case("train_naive_lj_B3_T48")
    gd = build(c, "fp32")
    K, M = gd.num_timesteps, gd.mel_bins
    g = torch.Generator().manual_seed(B * 100 + T)
    mel = (torch.randn(B, T, M, generator=g) * 2 - 5).cuda()
    t = (torch.arange(B) % K).cuda()                       # includes t = 0
    pad = (torch.arange(T)[None, :] >= torch.tensor([max(1, T - 3 * i) for i in range(B)])[:, None]).cuda()
    nt, npv, pn = (torch.randn(B, 1, M, T, generator=g).cuda() for _ in range(3))
    valid = (~pad)[:, None, None, :]
    ref_xt = gd.diffuse_fn(mel, t.clone(), noise=nt) * valid
    ref_prev = gd.diffuse_fn(mel, t - 1, noise=npv) * valid
    lib = _lib.load()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    x_t, x_prev = torch.empty_like(nt), torch.empty_like(nt)
    sa, sn = gd._train_tables(mel.device)
    _lib.check(lib.mgb_train_diffuse(_lib.ptr(mel), _lib.ptr(nt), _lib.ptr(npv), _lib.ptr(gd.spec_min.float().reshape(-1).contiguous()),
                                     _lib.ptr(gd.spec_max.float().reshape(-1).contiguous()), _lib.ptr(sa), _lib.ptr(sn), _lib.ptr(t),
                                     _lib.ptr(pad.to(torch.uint8)), _lib.ptr(x_t), _lib.ptr(x_prev), B, T, M, K, st), "train_diffuse")
    assert torch.equal(x_t, ref_xt) and torch.equal(x_prev, ref_prev)
    # posterior + clamp, forward and backward
    den = (torch.randn(B, 1, M, T, generator=g) * 1.2).cuda()
    d1, d2 = den.clone().requires_grad_(True), den.clone().requires_grad_(True)
    x0_ref = (d1 * valid).clamp(-1., 1.)
    prev_ref = gd.q_posterior_sample(x_start=x0_ref, x_t=ref_xt, t=t, noise=pn) * valid
    x0, prev = _PosteriorFn.apply(d2, x_t, pn, gd._sched(mel.device), t, pad.to(torch.uint8), True, K)
    assert torch.equal(x0, x0_ref)
    assert torch.allclose(prev, prev_ref, rtol=1e-6, atol=1e-6)
    r0, r1 = (torch.randn(B, 1, M, T, generator=g).cuda() for _ in range(2))
    ((x0_ref * r0).sum() + (prev_ref * r1).sum()).backward()
    ((x0 * r0).sum() + (prev * r1).sum()).backward()
    assert torch.allclose(d2.grad, d1.grad, rtol=1e-6, atol=1e-6)
    assert float(d1.grad.abs().sum()) > 0


def test_shallow_start_and_denorm_elementwise_exact():
    c = golden_case("shallow_lj_B2_T130")
    lib = _lib.load()
    B, T, M = c.B, c.T, 80
    _, _, _, start = c.oracle_forward()
    coarse, sn, pad = cu(c.t("coarse_mel")), cu(c.t("start_noise")), cu(c.t("pad_mask")).to(torch.uint8)
    smin = torch.full((M,), -11.5129, device="cuda")
    smax = torch.full((M,), 2.0, device="cuda")
    xT = torch.empty((B, 1, M, T), device="cuda")
    sa = float(c.oracle.buf["sqrt_alphas_cumprod"][c.K - 1])
    s1 = float(c.oracle.buf["sqrt_one_minus_alphas_cumprod"][c.K - 1])
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.mgb_shallow_start(_lib.ptr(coarse), _lib.ptr(sn), _lib.ptr(smin), _lib.ptr(smax), sa, s1,
                                     _lib.ptr(pad), _lib.ptr(xT), B, T, M, st), "shallow_start")
    assert torch.allclose(xT.cpu(), start, rtol=0, atol=2e-7)
    mel = torch.empty((B, T, M), device="cuda")
    _lib.check(lib.mgb_denorm_mask(_lib.ptr(xT), _lib.ptr(smin), _lib.ptr(smax), _lib.ptr(pad), _lib.ptr(mel),
                                   B, T, M, st), "denorm_mask")
    ref = c.oracle.denorm_spec(start[:, 0].transpose(1, 2)) * (~c.t("pad_mask")).unsqueeze(-1)
    assert torch.allclose(mel.cpu(), ref, rtol=0, atol=2e-6)


@pytest.mark.parametrize("B,T", [(2, 800), (3, 132), (1, 16), (2, 20), (2, 130), (1, 7)])
def test_elementwise_kernels_bitwise_vs_torch_expressions(B, T):
    """mgb_shallow_start / mgb_denorm_mask against the torch expressions of diffusion.py:147-153, 177-185, 228-232 evaluated
    in fp32 on the GPU, bit for bit, on both code paths: the register-transpose kernels (T % 4 == 0) and the tiled ones
    (ragged T), with a padding mask."""
    lib = _lib.load()
    M = 80
    g = torch.Generator().manual_seed(B * 1000 + T)
    coarse = (torch.randn(B, T, M, generator=g) * 2 - 5).cuda()
    noise = torch.randn(B, 1, M, T, generator=g).cuda()
    lens = torch.tensor([max(1, T - 5 * i) for i in range(B)])
    pad = (torch.arange(T)[None, :] >= lens[:, None]).cuda()
    smin = (torch.rand(M, generator=g) * 2 - 12).cuda()
    smax = (torch.rand(M, generator=g) * 2 + 1).cuda()
    sa, sn = 0.8123456, 0.5831234
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    xT = torch.empty((B, 1, M, T), device="cuda")
    _lib.check(lib.mgb_shallow_start(_lib.ptr(coarse), _lib.ptr(noise), _lib.ptr(smin), _lib.ptr(smax), sa, sn,
                                     _lib.ptr(pad.to(torch.uint8)), _lib.ptr(xT), B, T, M, st), "shallow_start")
    saf, snf = torch.tensor(sa, dtype=torch.float32).cuda(), torch.tensor(sn, dtype=torch.float32).cuda()
    norm = ((coarse - smin) / (smax - smin) * 2 - 1).transpose(1, 2)[:, None]
    ref = (saf * norm + snf * noise) * (~pad)[:, None, None, :]
    assert torch.equal(xT, ref)
    mel = torch.empty((B, T, M), device="cuda")
    _lib.check(lib.mgb_denorm_mask(_lib.ptr(xT), _lib.ptr(smin), _lib.ptr(smax), _lib.ptr(pad.to(torch.uint8)), _lib.ptr(mel),
                                   B, T, M, st), "denorm_mask")
    ref2 = ((xT[:, 0].transpose(1, 2) + 1) / 2 * (smax - smin) + smin) * (~pad).unsqueeze(-1)
    assert torch.equal(mel, ref2)
    # no mask
    _lib.check(lib.mgb_denorm_mask(_lib.ptr(xT), _lib.ptr(smin), _lib.ptr(smax), None, _lib.ptr(mel), B, T, M, st), "denorm_mask")
    assert torch.equal(mel, (xT[:, 0].transpose(1, 2) + 1) / 2 * (smax - smin) + smin)


def test_abi_error_codes():
    lib = _lib.load()
    dims = _lib.ModelDims(80, 256, 256, 20, 0)
    assert lib.mgb_device_check(0) == 0
    x = torch.zeros(16, device="cuda")
    rc = lib.mgb_denoiser_forward(C.byref(dims), 0, None, _lib.ptr(x), _lib.ptr(x), _lib.ptr(x), None, _lib.ptr(x),
                                  1, 8, _lib.ptr(x), 64, None)
    assert rc == _lib.E_ARG and b"NULL" in lib.mgb_last_error()
    rc = lib.mgb_denoiser_forward(C.byref(dims), 0, _lib.ptr(x), _lib.ptr(x), _lib.ptr(x), _lib.ptr(x), None,
                                  _lib.ptr(x), 1, 8, _lib.ptr(x), 64, None)
    assert rc == _lib.E_WORKSPACE
    bad = _lib.ModelDims(80, 192, 256, 20, 0)
    assert lib.mgb_packed_bytes(C.byref(bad), 0) == 0


def test_length_regulator_bit_exact():
    """mgb_length_regulate / mgb_durations_from_log / mask against (i) the golden outputs of the REAL reference
    (tests/golden/length_regulator.npz) and (ii) the oracle on further random cases: bit for bit."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_golden_lr import LR_CASES, logd_case, lr_case
    from mixgan_tts_b200.length_regulator import durations_from_log, get_mask_from_lengths
    from oracle.length_regulator import length_regulate
    g = load_golden("length_regulator")
    lr = LengthRegulator()
    for name in LR_CASES:
        x, dur, max_len = lr_case(name)
        out, mel_len, mask = lr.regulate_with_mask(torch.from_numpy(x).cuda(), torch.from_numpy(dur).cuda(), max_len)
        assert np.array_equal(mel_len.cpu().numpy(), g[f"{name}/mel_len"]), name
        assert tuple(out.shape) == g[f"{name}/out"].shape, name
        assert np.array_equal(out.cpu().numpy().view(np.uint32), g[f"{name}/out"].view(np.uint32)), name
        if f"{name}/mask_w" in g.files:
            assert np.array_equal(mask.cpu().numpy(), g[f"{name}/mask_w"]), name
            assert np.array_equal(get_mask_from_lengths(mel_len).cpu().numpy(), g[f"{name}/mask"]), name
        out2, mel_len2 = lr(torch.from_numpy(x).cuda(), torch.from_numpy(dur).cuda(), max_len)     # the reference's signature
        assert torch.equal(out2, out) and torch.equal(mel_len2, mel_len)
    log_d, controls = logd_case()
    for c in controls:
        d = durations_from_log(torch.from_numpy(log_d).cuda(), c)
        assert d.dtype == torch.int64 and np.array_equal(d.cpu().numpy(), g[f"dur_from_log/{c}"]), c
    rg = np.random.default_rng(5)
    for B, S, D, max_len in [(3, 7, 5, None), (4, 33, 256, 400), (1, 1, 8, 10), (2, 50, 256, None), (5, 64, 130, 97)]:
        x = rg.standard_normal((B, S, D)).astype(np.float32)
        dur = rg.integers(-2, 9, (B, S)).astype(np.int64)
        ref, ref_len = length_regulate(x, dur, max_len)
        out, mel_len = lr(torch.from_numpy(x).cuda(), torch.from_numpy(dur).cuda(), max_len)
        assert np.array_equal(mel_len.cpu().numpy(), ref_len)
        assert out.shape == ref.shape
        assert np.array_equal(out.cpu().numpy().view(np.uint32), ref.view(np.uint32))


def test_length_regulator_is_differentiable_like_the_reference():
    """The reference's expand + cat + pad carries gradient back to the encoder; the drop-in must too (segment sum)."""
    from oracle.length_regulator import length_regulate_backward
    rg = np.random.default_rng(11)
    for B, S, D, max_len in [(2, 6, 4, None), (3, 20, 256, 60), (2, 9, 16, 11)]:
        x = torch.from_numpy(rg.standard_normal((B, S, D)).astype(np.float32)).cuda().requires_grad_(True)
        dur = torch.from_numpy(rg.integers(-1, 7, (B, S)).astype(np.int64)).cuda()
        out, mel_len = LengthRegulator()(x, dur, max_len)
        assert out.requires_grad and not mel_len.requires_grad
        w = torch.from_numpy(rg.standard_normal(tuple(out.shape)).astype(np.float32)).cuda()
        (out * w).sum().backward()
        ref = length_regulate_backward(w.cpu().numpy(), dur.cpu().numpy())
        assert np.allclose(x.grad.cpu().numpy(), ref, rtol=1e-5, atol=1e-5)
