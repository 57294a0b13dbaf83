"""Generate the golden vectors under ``tests/golden/`` by running the REAL reference.

Run in the build container only (needs ``/root/reference``):

    python tests/golden/make_golden.py

Each ``.npz`` holds the reference's outputs for inputs and weights that are
regenerated from seeds by ``mixgan_tts_b200.synth`` (the 55 MB of weights are not
committed; their sha256 is, so generator drift is detected).  The reference's
fp32 schedule buffers are stored verbatim as known-answer values.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from mixgan_tts_b200 import configs, synth  # noqa: E402
from oracle import ref_loader  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: dataset, model, multi_speaker, B, T, weight seed, input seed
    "naive_lj_B2_T64": ("LJSpeech", "naive", False, 2, 64, 0, 1234),
    "naive_lj_B3_T200": ("LJSpeech", "naive", False, 3, 200, 0, 4321),
    "shallow_aishell_spk_B2_T77": ("AISHELL3", "shallow", True, 2, 77, 7, 99),
    "shallow_lj_B2_T130": ("LJSpeech", "shallow", False, 2, 130, 0, 5),
}

SCHED_KEYS = ["betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod",
              "sqrt_one_minus_alphas_cumprod", "log_one_minus_alphas_cumprod",
              "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_variance",
              "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2"]


def run_case(name, dataset, model, multi, B, T, wseed, iseed):
    torch.set_num_threads(1)  # fixed summation order
    args, pc, mc, tc = configs.make_configs(dataset, model, multi)
    W = synth.make_denoiser_weights(wseed, layers=mc["denoiser"]["residual_layers"],
                                    multi_speaker=multi)
    gd = ref_loader.build_reference_diffusion(args, pc, mc, tc, W)
    K = gd.num_timesteps
    inp = synth.make_inputs(iseed, B, T, K, multi_speaker=multi, shallow=(model == "shallow"))
    tt = lambda a: None if a is None else torch.from_numpy(a)
    cond, spk, pad = tt(inp["cond"]), tt(inp["spk"]), tt(inp["pad_mask"])
    x_T, noises = tt(inp["x_T"]), tt(inp["noises"])

    out = {"weights_sha256": synth.weights_digest(W), "torch_version": torch.__version__,
           "numpy_version": np.__version__, "K": K}
    for k in SCHED_KEYS:
        out[f"sched_{k}"] = getattr(gd, k).numpy()

    # (1) one bare Denoiser.forward at mixed timesteps (model/modules.py:420)
    t_mixed = torch.tensor([(K - 1 - b) % K for b in range(B)], dtype=torch.long)
    with torch.no_grad():
        den = gd.denoise_fn(x_T, t_mixed, cond.transpose(1, 2), spk)
    out["denoiser_t"] = t_mixed.numpy()
    out["denoiser_out"] = den.numpy()

    # (2) one p_sample at t = K-1 (model/diffusion.py:122)
    tK = torch.full((B,), K - 1, dtype=torch.long)
    with ref_loader.injected_noise(noise_like_seq=[noises[K - 1]]):
        ps = gd.p_sample(x_T.clone(), tK, cond.transpose(1, 2), spk)
    out["p_sample_out"] = ps.numpy()

    # (3) the full inference forward (model/diffusion.py:187-200), noise injected in call order
    nl_seq = [noises[i] for i in reversed(range(K))]
    if model == "shallow":
        coarse = tt(inp["coarse_mel"])
        with ref_loader.injected_noise(noise_like_seq=nl_seq, randn_like_seq=[tt(inp["start_noise"])]):
            res = gd(None, cond, spk, pad, coarse_mel=coarse)
    else:
        # naive: sampling() draws x_T itself when noise is None, so drive sampling() directly
        # the way forward() does (diffusion.py:190-200) with x_T injected.
        gd.cond = cond.transpose(1, 2)
        gd.spk_emb = spk
        with ref_loader.injected_noise(noise_like_seq=nl_seq):
            states = gd.sampling(noise=x_T)
        res = (states[-1] * (~pad.unsqueeze(-1)),)
        out["state_after_first_step"] = states[1].numpy()
    out["final_mel"] = res[0].numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "K", K, "final", tuple(res[0].shape), "rms", float(res[0].pow(2).mean().sqrt()))


TRAIN_CASES = {
    # name: dataset, model, multi_speaker, B, T, weight seed, input seed   (training branch, forward values only)
    "train_naive_lj_B3_T48": ("LJSpeech", "naive", False, 3, 48, 0, 21),
    "train_shallow_aishell_spk_B2_T40": ("AISHELL3", "shallow", True, 2, 40, 7, 22),
}


def run_train_case(name, dataset, model, multi, B, T, wseed, iseed):
    """model/diffusion.py:201-225 with every random draw injected in call order: randint (:203), two randn_like
    (diffuse_fn :182 via :206-207) and one noise_like (q_posterior_sample :116)."""
    torch.set_num_threads(1)
    args, pc, mc, tc = configs.make_configs(dataset, model, multi)
    W = synth.make_denoiser_weights(wseed, layers=mc["denoiser"]["residual_layers"], multi_speaker=multi)
    gd = ref_loader.build_reference_diffusion(args, pc, mc, tc, W)
    K = gd.num_timesteps
    inp = synth.make_inputs(iseed, B, T, K, multi_speaker=multi, shallow=(model == "shallow"))
    ex = synth.make_train_extras(iseed + 1000, B, T, K)
    tt = lambda a: None if a is None else torch.from_numpy(a)
    with ref_loader.injected_noise(noise_like_seq=[tt(ex["post_noise"])],
                                   randn_like_seq=[tt(ex["noise_t"]), tt(ex["noise_prev"])],
                                   randint_seq=[tt(ex["t"]).clone()]):
        with torch.no_grad():
            res = gd(tt(ex["mel"]), tt(inp["cond"]), tt(inp["spk"]), tt(inp["pad_mask"]), coarse_mel=tt(inp["coarse_mel"]))
    out = {"weights_sha256": synth.weights_digest(W), "torch_version": torch.__version__, "K": K,
           "x_0_pred": res[0].numpy(), "x_t": res[1].numpy(), "x_t_prev": res[2].numpy(),
           "x_t_prev_pred": res[3].numpy(), "t": res[4].numpy()}
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "K", K, "t", res[4].tolist(), "x0 rms", float(res[0].pow(2).mean().sqrt()))


def run_grad_case(name, dataset, model, multi, B, T, wseed, iseed):
    """Gradients of the REAL reference (torch autograd through model/diffusion.py:201-225 and the Denoiser) for the
    linear probe loss of ``synth.grad_probe``: full d loss / d cond and d loss / d spk, and per parameter the gradient's
    L2 norm plus a strided sample (the full 55 MB of parameter gradients are not committed)."""
    torch.set_num_threads(1)
    args, pc, mc, tc = configs.make_configs(dataset, model, multi)
    W = synth.make_denoiser_weights(wseed, layers=mc["denoiser"]["residual_layers"], multi_speaker=multi)
    gd = ref_loader.build_reference_diffusion(args, pc, mc, tc, W)
    K = gd.num_timesteps
    inp = synth.make_inputs(iseed, B, T, K, multi_speaker=multi, shallow=(model == "shallow"))
    ex = synth.make_train_extras(iseed + 1000, B, T, K)
    pr = synth.grad_probe(iseed + 2000, B, T)
    tt = lambda a: None if a is None else torch.from_numpy(a)
    cond = tt(inp["cond"]).requires_grad_(True)
    spk = tt(inp["spk"])
    if spk is not None:
        spk.requires_grad_(True)
    for p in gd.denoise_fn.parameters():
        p.requires_grad_(True)
    with ref_loader.injected_noise(noise_like_seq=[tt(ex["post_noise"])],
                                   randn_like_seq=[tt(ex["noise_t"]), tt(ex["noise_prev"])],
                                   randint_seq=[tt(ex["t"]).clone()]):
        res = gd(tt(ex["mel"]), cond, spk, tt(inp["pad_mask"]), coarse_mel=tt(inp["coarse_mel"]))
    loss = (res[0] * tt(pr["r0"])).sum() + (res[3] * tt(pr["r1"])).sum()
    loss.backward()
    out = {"weights_sha256": synth.weights_digest(W), "torch_version": torch.__version__, "K": K,
           "loss": np.float64(loss.item()), "grad_cond": cond.grad.numpy()}
    if spk is not None:
        out["grad_spk"] = spk.grad.numpy()
    for k, p in gd.denoise_fn.named_parameters():
        g = p.grad.detach().numpy().reshape(-1)
        out["gnorm/" + k] = np.float64(np.sqrt((g.astype(np.float64) ** 2).sum()))
        out["gsample/" + k] = g[synth.grad_sample_index(g.size)]
    np.savez_compressed(os.path.join(HERE, name.replace("train_", "grad_") + ".npz"), **out)
    print(name, "loss", loss.item(), "|d cond|", float(cond.grad.norm()))


if __name__ == "__main__":
    if "--grad-only" in sys.argv:
        for name, spec in TRAIN_CASES.items():
            run_grad_case(name, *spec)
        sys.exit(0)
    only_train = "--train-only" in sys.argv
    if not only_train:
        for name, spec in CASES.items():
            run_case(name, *spec)
    for name, spec in TRAIN_CASES.items():
        run_train_case(name, *spec)
        run_grad_case(name, *spec)
