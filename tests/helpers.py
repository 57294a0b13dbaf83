"""Shared helpers for the parity tests (oracle side and case construction)."""
from __future__ import annotations

import os

import numpy as np
import torch

from mixgan_tts_b200 import configs, synth
from oracle.diffusion import DiffusionOracle

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# name: dataset, model, multi_speaker, B, T, weight seed, input seed  (tests/golden/make_golden.py)
GOLDEN_CASES = {
    "naive_lj_B2_T64": ("LJSpeech", "naive", False, 2, 64, 0, 1234),
    "naive_lj_B3_T200": ("LJSpeech", "naive", False, 3, 200, 0, 4321),
    "shallow_aishell_spk_B2_T77": ("AISHELL3", "shallow", True, 2, 77, 7, 99),
    "shallow_lj_B2_T130": ("LJSpeech", "shallow", False, 2, 130, 0, 5),
}


# training-branch goldens (forward values): same tuple layout
TRAIN_CASES = {
    "train_naive_lj_B3_T48": ("LJSpeech", "naive", False, 3, 48, 0, 21),
    "train_shallow_aishell_spk_B2_T40": ("AISHELL3", "shallow", True, 2, 40, 7, 22),
}
TRAIN_KEYS = ("x_0_pred", "x_t", "x_t_prev", "x_t_prev_pred")


def train_case(name):
    """(Case, extras dict of torch tensors) of one training-branch golden."""
    spec = TRAIN_CASES[name]
    c = Case(*spec)
    ex = synth.make_train_extras(spec[6] + 1000, c.B, c.T, c.K)
    return c, {k: torch.from_numpy(v) for k, v in ex.items()}


def rel_l2(a, b) -> float:
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


class Case:
    """One synthetic workload: configs, weights, inputs, and the oracle built on them."""

    def __init__(self, dataset, model, multi, B, T, wseed, iseed, layers=None, timesteps=None):
        self.dataset, self.model, self.multi, self.B, self.T = dataset, model, multi, B, T
        self.args, self.pc, self.mc, self.tc = configs.make_configs(
            dataset, model, multi, residual_layers=layers, timesteps=timesteps)
        den = self.mc["denoiser"]
        self.W = synth.make_denoiser_weights(wseed, layers=den["residual_layers"], multi_speaker=multi)
        self.K = den["timesteps" if model == "naive" else "shallow_timesteps"]
        self.inp = synth.make_inputs(iseed, B, T, self.K, multi_speaker=multi,
                                     shallow=(model == "shallow"))
        self.oracle = DiffusionOracle(self.W, model=model, denoiser_cfg=den,
                                      spec_min=[configs.SPEC_MIN] * 80, spec_max=[configs.SPEC_MAX] * 80)

    def t(self, key):
        v = self.inp[key]
        return None if v is None else torch.from_numpy(np.ascontiguousarray(v))

    def oracle_forward(self):
        return self.oracle.forward_inference(
            self.t("cond"), self.t("spk"), self.t("pad_mask"), x_T=self.t("x_T"),
            noises=self.t("noises"), coarse_mel=self.t("coarse_mel"), start_noise=self.t("start_noise"))


def golden_case(name) -> Case:
    return Case(*GOLDEN_CASES[name])


def load_golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


# ---- gradient parity (training path) -----------------------------------------------------------
def grad_golden_name(train_name: str) -> str:
    return train_name.replace("train_", "grad_")


def oracle_training_grads(c: Case, ex: dict, probe: dict, dtype=torch.float32):
    """torch autograd through the oracle's training branch for the linear probe loss of ``synth.grad_probe``.
    Returns (loss, outputs 5-tuple, {param key: grad}, grad_cond [B,T,H], grad_spk|None)."""
    cast = lambda v: None if v is None else (v.to(dtype) if v.is_floating_point() else v)
    W = {k: torch.from_numpy(v).to(dtype).requires_grad_(True) for k, v in c.W.items()}
    cond = cast(c.t("cond")).requires_grad_(True)
    spk = cast(c.t("spk"))
    if spk is not None:
        spk.requires_grad_(True)
    orc = c.oracle
    saved = {k: v for k, v in orc.buf.items()}, orc.spec_min, orc.spec_max
    if dtype != torch.float32:   # a higher-precision truth for error attribution
        orc.buf = {k: v.to(dtype) for k, v in orc.buf.items()}
        orc.spec_min, orc.spec_max = orc.spec_min.to(dtype), orc.spec_max.to(dtype)
    try:
        out = orc.forward_training_graph(
            cast(ex["mel"]), cond, spk, c.t("pad_mask"), t=ex["t"].clone(), noise_t=cast(ex["noise_t"]),
            noise_prev=cast(ex["noise_prev"]), post_noise=cast(ex["post_noise"]), coarse_mel=cast(c.t("coarse_mel")), W=W)
        loss = (out[0] * cast(probe["r0"])).sum() + (out[3] * cast(probe["r1"])).sum()
        loss.backward()
    finally:
        orc.buf, orc.spec_min, orc.spec_max = saved
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in W.items()}
    return loss.detach(), out, grads, cond.grad, (spk.grad if spk is not None else None)


def check_grads_against_golden(g, grads: dict, gcond, gspk, tol: float):
    """Compare a full set of gradients with the reference's committed summary (norm + strided sample per parameter,
    full d/d cond and d/d spk).  Returns the worst relative error seen."""
    worst = rel_l2(gcond, g["grad_cond"])
    assert worst < tol, ("grad_cond", worst)
    if "grad_spk" in g.files:
        e = rel_l2(gspk, g["grad_spk"])
        assert e < tol, ("grad_spk", e)
        worst = max(worst, e)
    total = float(np.sqrt(sum(float(g[k]) ** 2 for k in g.files if k.startswith("gnorm/"))))
    for k in g.files:
        if not k.startswith("gsample/"):
            continue
        key = k[len("gsample/"):]
        mine = torch.as_tensor(grads[key]).detach().cpu().double().reshape(-1)
        ref_norm = float(g["gnorm/" + key])
        idx = torch.from_numpy(synth.grad_sample_index(mine.numel()))
        ref_s = torch.from_numpy(g[k]).double()
        # a parameter whose gradient is (numerically) zero in the reference must be (numerically) zero here too
        scale = max(ref_norm, 1e-6 * total)
        e_norm = abs(float(mine.norm()) - ref_norm) / scale
        e_samp = float((mine[idx] - ref_s).norm()) / max(float(ref_s.norm()), 1e-6 * total * (len(idx) / mine.numel()) ** 0.5)
        assert e_norm < tol, (key, "norm", e_norm)
        assert e_samp < 10 * tol, (key, "sample", e_samp)
        worst = max(worst, e_norm)
    return worst
