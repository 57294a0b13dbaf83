"""Noise schedules and the posterior constants of the reverse process (host side).

Same semantics as the reference's ``get_noise_schedule_list`` (``utils/tools.py:425-445``)
and the buffer block of ``GaussianDiffusion.__init__`` (``model/diffusion.py:53-83``): all
arithmetic in float64 numpy, each derived array cast to fp32 exactly once, so the buffers
(and therefore the per-step coefficients the kernels gather) match the reference bit for bit.
"""
from __future__ import annotations

import numpy as np


def noise_schedule_list(schedule_mode: str, timesteps: int, min_beta: float = 0.0,
                        max_beta: float = 0.01, s: float = 0.008) -> np.ndarray:
    if schedule_mode == "linear":
        return np.linspace(1e-4, max_beta, timesteps)
    if schedule_mode == "cosine":
        n = timesteps + 1
        grid = np.linspace(0, n, n)
        bar = np.cos(((grid / n) + s) / (1 + s) * np.pi * 0.5) ** 2
        bar = bar / bar[0]
        return np.clip(1 - (bar[1:] / bar[:-1]), a_min=0, a_max=0.999)
    if schedule_mode == "vpsde":
        T = timesteps
        t = np.arange(1, T + 1, dtype=np.int64)
        out = np.empty(T, dtype=np.float64)
        for i, ti in enumerate(t.tolist()):
            out[i] = 1.0 - np.exp(-min_beta / T - 0.5 * (max_beta - min_beta) * ((2 * ti - 1) / (T ** 2)))
        return out
    raise NotImplementedError(schedule_mode)


BUFFER_NAMES = (
    "betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod",
    "sqrt_one_minus_alphas_cumprod", "log_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod",
    "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
    "posterior_mean_coef1", "posterior_mean_coef2",
)


def posterior_buffers(betas) -> dict:
    b = np.asarray(betas, dtype=np.float64)
    a = 1.0 - b
    bar = np.cumprod(a, axis=0)
    bar_prev = np.append(1.0, bar[:-1])
    var = b * (1.0 - bar_prev) / (1.0 - bar)
    with np.errstate(divide="ignore"):
        vals = (
            b, bar, bar_prev, np.sqrt(bar), np.sqrt(1.0 - bar), np.log(1.0 - bar), np.sqrt(1.0 / bar),
            np.sqrt(1.0 / bar - 1), var, np.log(np.maximum(var, 1e-20)),
            b * np.sqrt(bar_prev) / (1.0 - bar), (1.0 - bar_prev) * np.sqrt(a) / (1.0 - bar),
        )
    return {k: np.asarray(v, dtype=np.float64).astype(np.float32) for k, v in zip(BUFFER_NAMES, vals)}
