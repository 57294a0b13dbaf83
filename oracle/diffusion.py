"""CPU restatement of ``GaussianDiffusion`` inference (torch fp32, host).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Noise is always INJECTED (device and host RNG streams cannot match):
``noises[t]`` is what the reference draws inside ``q_posterior_sample`` at
timestep ``t`` (``model/diffusion.py:116``; the draw at ``t == 0`` is multiplied
by zero, ``:118-119``), ``x_T`` the start state of ``sampling`` (``:160``) and
``start_noise`` the ``randn_like`` of the shallow start (``:182``).
"""
from __future__ import annotations

import numpy as np
import torch

from . import schedule
from .denoiser import as_torch, denoiser_forward, denoiser_forward_graph


class DiffusionOracle:
    def __init__(self, W: dict, *, model: str = "naive", denoiser_cfg: dict,
                 spec_min, spec_max):
        self.W = as_torch(W)
        self.model = model
        K = denoiser_cfg["timesteps" if model == "naive" else "shallow_timesteps"]   # diffusion.py:47
        betas = schedule.noise_schedule(denoiser_cfg["noise_schedule_naive"], K,
                                        denoiser_cfg["min_beta"], denoiser_cfg["max_beta"],
                                        denoiser_cfg["s"])
        self.K = int(K)
        self.buf = {k: torch.from_numpy(v) for k, v in schedule.diffusion_buffers(betas).items()}
        keep = denoiser_cfg["keep_bins"]
        self.spec_min = torch.tensor(np.asarray(spec_min, dtype=np.float32))[None, None, :keep]
        self.spec_max = torch.tensor(np.asarray(spec_max, dtype=np.float32))[None, None, :keep]

    # model/diffusion.py:228-232
    def norm_spec(self, x):
        return (x - self.spec_min) / (self.spec_max - self.spec_min) * 2 - 1

    def denorm_spec(self, x):
        return (x + 1) / 2 * (self.spec_max - self.spec_min) + self.spec_min

    def _extract(self, name, t, ndim=4):
        # model/diffusion.py:26-29
        return self.buf[name].gather(-1, t).reshape(t.shape[0], *((1,) * (ndim - 1)))

    def q_posterior(self, x_start, x_t, t):
        # model/diffusion.py:104-111
        mean = (self._extract("posterior_mean_coef1", t) * x_start
                + self._extract("posterior_mean_coef2", t) * x_t)
        return mean, self._extract("posterior_variance", t), \
            self._extract("posterior_log_variance_clipped", t)

    def q_posterior_sample(self, x_start, x_t, t, noise):
        # model/diffusion.py:113-119
        mean, _, logvar = self.q_posterior(x_start, x_t, t)
        nonzero = (1 - (t == 0).float()).reshape(-1, 1, 1, 1)
        return mean + nonzero * (0.5 * logvar).exp() * noise

    @torch.no_grad()
    def p_sample(self, x_t, t, cond, spk, noise, clip_denoised=True):
        # model/diffusion.py:121-129 — the network predicts x_0 directly
        x0 = denoiser_forward(self.W, x_t, t, cond, spk)
        if clip_denoised:
            x0 = x0.clamp(-1.0, 1.0)
        return self.q_posterior_sample(x0, x_t, t, noise), x0

    def q_sample(self, x_start, t, noise):
        # model/diffusion.py:147-153
        return (self._extract("sqrt_alphas_cumprod", t) * x_start
                + self._extract("sqrt_one_minus_alphas_cumprod", t) * noise)

    def diffuse_fn(self, x_start, t, noise):
        # model/diffusion.py:177-185 ([B,T,M] -> [B,1,M,T]; t < 0 returns the clean mel)
        x_start = self.norm_spec(x_start).transpose(1, 2)[:, None, :, :]
        neg = t < 0
        t = t.clamp(min=0)
        out = self.q_sample(x_start, t, noise)
        out[neg] = x_start[neg]
        return out

    @torch.no_grad()
    def sampling(self, cond_bht, spk, x_T, noises):
        """model/diffusion.py:155-165.  Returns ``(states, x0_preds)``: the K+1
        denormalised ``[B,T,M]`` states and, per step, the clamped normalised x0."""
        B = cond_bht.shape[0]
        xs, x0s = [x_T], []
        for i in reversed(range(self.K)):
            t = torch.full((B,), i, dtype=torch.long)
            x, x0 = self.p_sample(xs[-1], t, cond_bht, spk, noises[i])
            xs.append(x)
            x0s.append(x0)
        return [self.denorm_spec(x[:, 0].transpose(1, 2)) for x in xs], x0s

    @torch.no_grad()
    def forward_inference(self, cond, spk, pad_mask, *, x_T=None, noises=None,
                          coarse_mel=None, start_noise=None):
        """The ``mel is None`` branch of ``GaussianDiffusion.forward``
        (model/diffusion.py:187-200).  ``cond [B,T,H]``; ``pad_mask [B,T]`` True = padding."""
        B = cond.shape[0]
        valid = ~pad_mask.unsqueeze(-1)                       # :190  [B,T,1], True = valid
        cond_bht = cond.transpose(1, 2)                       # :191
        if self.model != "shallow":
            start = x_T
        else:
            t = torch.full((B,), self.K - 1, dtype=torch.long)
            start = self.diffuse_fn(coarse_mel, t, start_noise) * valid.unsqueeze(-1).transpose(1, -1)  # :198
        states, x0s = self.sampling(cond_bht, spk, start, noises)
        return states[-1] * valid, states, x0s, start

    @torch.no_grad()
    def forward_training(self, mel, cond, spk, pad_mask, **kw):
        """Forward values of ``forward_training_graph`` (what ``evaluate.py`` uses under ``no_grad``)."""
        return self.forward_training_graph(mel, cond, spk, pad_mask, **kw)

    def forward_training_graph(self, mel, cond, spk, pad_mask, *, t, noise_t, noise_prev, post_noise,
                               coarse_mel=None, clip_denoised=True, W=None):
        """The ``mel is not None`` branch of ``GaussianDiffusion.forward`` (model/diffusion.py:201-225),
        differentiable by torch autograd (pass ``W`` with ``requires_grad`` tensors to get parameter gradients).
        ``t`` is what the reference draws with ``torch.randint`` (:203),
        ``noise_t`` / ``noise_prev`` the two ``randn_like`` draws of ``diffuse_fn`` (:206-207, :182) and
        ``post_noise`` the draw inside ``q_posterior_sample`` (:116).  Returns the reference's 5-tuple."""
        valid = (~pad_mask)[:, None, None, :]                                   # :190, :202  [B,1,1,T]
        cond_bht = cond.transpose(1, 2)                                         # :191
        x_t = self.diffuse_fn(mel, t.clone(), noise_t) * valid                   # :206
        x_t_prev = self.diffuse_fn(mel, t - 1, noise_prev) * valid               # :207 (t-1 = -1 -> the clean mel)
        x0 = denoiser_forward_graph(W if W is not None else self.W, x_t, t, cond_bht, spk) * valid   # :210
        if clip_denoised:
            x0 = x0.clamp(-1.0, 1.0)                                            # :211-212
        if self.model != "shallow":
            x_start = x0                                                        # :215-216
        else:
            x_start = self.norm_spec(coarse_mel).transpose(1, 2)[:, None, :, :]  # :218-219
        x_prev_pred = self.q_posterior_sample(x_start, x_t, t, post_noise) * valid   # :220
        tr = lambda x: x[:, 0].transpose(1, 2)                                  # :222-225
        return tr(x0), tr(x_t), tr(x_t_prev), tr(x_prev_pred), t
