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
