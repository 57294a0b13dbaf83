"""Drop-in ``GaussianDiffusion`` (reference: ``model/diffusion.py:38-235``) whose reverse process
runs in the sm_100a library.

Same constructor, attributes (``denoise_fn``, ``num_timesteps``, ``mel_bins``, ``cond``, ``spk_emb``),
buffers/``state_dict`` keys and method signatures as the reference.  Every method that draws
noise additionally accepts the noise as a keyword argument (a superset of the reference API) so
that parity tests and the bench can inject identical noise; when it is omitted, noise is drawn
with ``torch.randn`` on the device, as the reference does.

Hot methods (``p_sample``, ``sampling``, inference ``forward``) call the C ABI; the light
training-time helpers (``q_sample``, ``q_posterior``, ``norm_spec`` ...) are plain torch on the
registered buffers.  The training branch of ``forward`` (``mel`` given) returns the reference's 5-tuple;
with autograd enabled the Denoiser's backward runs in the library as well (``mgb_denoiser_backward``).
"""
from __future__ import annotations

import ctypes as C
import json
import os

import numpy as np
import torch
from torch import nn

from . import _lib
from .modules import PRECISIONS, Denoiser
from .schedule import noise_schedule_list, posterior_buffers


def _extract(a, t, x_shape):
    b = t.shape[0]
    return a.gather(-1, t).reshape(b, *((1,) * (len(x_shape) - 1)))


class _PosteriorFn(torch.autograd.Function):
    """``x_0_pred = clamp(denoiser_out * valid)``, ``x_t_prev_pred = q_posterior_sample(x_0_pred, x_t, t) * valid`` of the
    'naive' training branch (diffusion.py:210-212, :220) as one library kernel, and one more for d/d denoiser_out."""

    @staticmethod
    def forward(ctx, den_out, x_t, noise, sched, t, pad, clip, K):
        lib = _lib.load()
        den = den_out.detach().float().contiguous()
        B, _, M, T = den.shape
        x0, prev = torch.empty_like(den), torch.empty_like(den)
        st = C.c_void_p(torch.cuda.current_stream(den.device).cuda_stream)
        with torch.cuda.device(den.device):
            _lib.check(lib.mgb_train_posterior(_lib.ptr(den), _lib.ptr(x_t), _lib.ptr(noise), _lib.ptr(sched), _lib.ptr(t),
                                               _lib.ptr(pad), int(clip), _lib.ptr(x0), _lib.ptr(prev), B, T, M, K, st),
                       "mgb_train_posterior")
        ctx.save_for_backward(den, sched, t, pad)
        ctx.cfg = (int(clip), K)
        return x0, prev

    @staticmethod
    def backward(ctx, g_x0, g_prev):
        den, sched, t, pad = ctx.saved_tensors
        clip, K = ctx.cfg
        lib = _lib.load()
        B, _, M, T = den.shape
        g0 = None if g_x0 is None else g_x0.float().contiguous()
        g1 = None if g_prev is None else g_prev.float().contiguous()
        g = torch.empty_like(den)
        st = C.c_void_p(torch.cuda.current_stream(den.device).cuda_stream)
        with torch.cuda.device(den.device):
            _lib.check(lib.mgb_train_posterior_backward(_lib.ptr(g0), _lib.ptr(g1), _lib.ptr(den), _lib.ptr(sched), _lib.ptr(t),
                                                        _lib.ptr(pad), clip, _lib.ptr(g), B, T, M, K, st),
                       "mgb_train_posterior_backward")
        return g, None, None, None, None, None, None, None


class GaussianDiffusion(nn.Module):
    def __init__(self, args, preprocess_config, model_config, train_config, precision: str | None = None):
        super().__init__()
        self.model = args.model
        self.denoise_fn = Denoiser(preprocess_config, model_config, precision=precision)
        self.mel_bins = preprocess_config["preprocessing"]["mel"]["n_mel_channels"]
        den = model_config["denoiser"]
        betas = noise_schedule_list(
            schedule_mode=den["noise_schedule_naive"],
            timesteps=den["timesteps" if self.model == "naive" else "shallow_timesteps"],
            min_beta=den["min_beta"], max_beta=den["max_beta"], s=den["s"])
        self.num_timesteps = int(betas.shape[0])
        self.loss_type = train_config["loss"]["noise_loss"]
        for name, val in posterior_buffers(betas).items():
            self.register_buffer(name, torch.tensor(val, dtype=torch.float32))
        with open(os.path.join(preprocess_config["path"]["preprocessed_path"], "stats.json")) as f:
            stats = json.load(f)
        keep = den["keep_bins"]
        self.register_buffer("spec_min", torch.FloatTensor(stats["spec_min"])[None, None, :keep])
        self.register_buffer("spec_max", torch.FloatTensor(stats["spec_max"])[None, None, :keep])
        self.cond = None
        self.spk_emb = None
        self._sched_cache = None

    # ------------------------------------------------------------------ light helpers (torch)
    def q_mean_variance(self, x_start, t):
        mean = _extract(self.sqrt_alphas_cumprod, t, x_start.shape) * x_start
        variance = _extract(1. - self.alphas_cumprod, t, x_start.shape)
        log_variance = _extract(self.log_one_minus_alphas_cumprod, t, x_start.shape)
        return mean, variance, log_variance

    def predict_start_from_noise(self, x_t, t, noise):
        return (_extract(self.sqrt_recip_alphas_cumprod, t, x_t.shape) * x_t
                - _extract(self.sqrt_recipm1_alphas_cumprod, t, x_t.shape) * noise)

    def q_posterior(self, x_start, x_t, t):
        mean = (_extract(self.posterior_mean_coef1, t, x_t.shape) * x_start
                + _extract(self.posterior_mean_coef2, t, x_t.shape) * x_t)
        return (mean, _extract(self.posterior_variance, t, x_t.shape),
                _extract(self.posterior_log_variance_clipped, t, x_t.shape))

    def q_posterior_sample(self, x_start, x_t, t, repeat_noise=False, noise=None):
        b = x_start.shape[0]
        mean, _, logvar = self.q_posterior(x_start=x_start, x_t=x_t, t=t)
        if noise is None:
            noise = (torch.randn((1, *x_start.shape[1:]), device=x_start.device).repeat(b, 1, 1, 1)
                     if repeat_noise else torch.randn(x_start.shape, device=x_start.device))
        nonzero = (1 - (t == 0).float()).reshape(b, *((1,) * (len(x_start.shape) - 1)))
        return mean + nonzero * (0.5 * logvar).exp() * noise

    def q_sample(self, x_start, t, noise=None):
        if noise is None:
            noise = torch.randn_like(x_start)
        return (_extract(self.sqrt_alphas_cumprod, t, x_start.shape) * x_start
                + _extract(self.sqrt_one_minus_alphas_cumprod, t, x_start.shape) * noise)

    def diffuse_fn(self, x_start, t, noise=None):
        x_start = self.norm_spec(x_start).transpose(1, 2)[:, None, :, :]   # [B,1,M,T]
        # diffusion.py:177-185 without its two boolean-mask index operations: those size their result on the host
        # (a device synchronisation per call, and not capturable in a CUDA graph); masked_fill_ / where give the same values
        neg = t < 0
        t.masked_fill_(neg, 0)                                             # in place, as the reference
        out = self.q_sample(x_start=x_start, t=t, noise=noise)
        return torch.where(neg.reshape(-1, 1, 1, 1), x_start, out)

    def diffuse_trace(self, x_start, mask):
        b, device = x_start.shape[0], x_start.device
        trace = [self.norm_spec(x_start).clamp_(-1., 1.) * ~mask.unsqueeze(-1)]
        for t in range(self.num_timesteps):
            tt = torch.full((b,), t, device=device, dtype=torch.long)
            trace.append(self.diffuse_fn(x_start, tt)[:, 0].transpose(1, 2) * ~mask.unsqueeze(-1))
        return trace

    def norm_spec(self, x):
        return (x - self.spec_min) / (self.spec_max - self.spec_min) * 2 - 1

    def denorm_spec(self, x):
        return (x + 1) / 2 * (self.spec_max - self.spec_min) + self.spec_min

    def out2mel(self, x):
        return x

    # ------------------------------------------------------------------ C-ABI plumbing
    def _sched(self, device) -> torch.Tensor:
        """float [3][K]: coef1 | coef2 | sigma (sigma[t] = [t != 0] * exp(0.5 * logvar[t]))."""
        key = (device, self.posterior_mean_coef1._version, self.posterior_mean_coef1.data_ptr())
        if self._sched_cache is None or self._sched_cache[0] != key:
            c1 = self.posterior_mean_coef1.detach().float().cpu()
            c2 = self.posterior_mean_coef2.detach().float().cpu()
            sig = (0.5 * self.posterior_log_variance_clipped.detach().float().cpu()).exp()
            sig[0] = 0.0
            host = (self.sqrt_alphas_cumprod.detach().float().cpu().tolist(),
                    self.sqrt_one_minus_alphas_cumprod.detach().float().cpu().tolist())
            self._sched_cache = (key, torch.stack([c1, c2, sig]).contiguous().to(device), host)
        return self._sched_cache[1]

    def _train_tables(self, device):
        """(sqrt_alphas_cumprod, sqrt_one_minus_alphas_cumprod) as contiguous fp32 device tensors (the q_sample tables)."""
        key = (device, self.sqrt_alphas_cumprod._version, self.sqrt_alphas_cumprod.data_ptr())
        c = getattr(self, "_train_tab_cache", None)
        if c is None or c[0] != key:
            c = (key, self.sqrt_alphas_cumprod.detach().float().contiguous().to(device),
                 self.sqrt_one_minus_alphas_cumprod.detach().float().contiguous().to(device))
            self._train_tab_cache = c
        return c[1], c[2]

    def _call_ctx(self, ref):
        if ref.device.type != "cuda":
            raise RuntimeError("mixgan_tts_b200.GaussianDiffusion needs CUDA tensors (no CPU fallback)")
        den = self.denoise_fn
        return _lib.load(), den, PRECISIONS[den.precision], C.c_void_p(torch.cuda.current_stream(ref.device).cuda_stream)

    @staticmethod
    def _f32c(t):
        return None if t is None else t.detach().float().contiguous()

    def _spk(self, spk_emb):
        if self.denoise_fn.dims.multi_speaker:
            if spk_emb is None:
                raise TypeError("multi_speaker Denoiser needs spk_emb")
            return self._f32c(spk_emb)
        return None

    # ------------------------------------------------------------------ hot path
    @torch.no_grad()
    def p_sample(self, x_t, t, cond, spk_emb, clip_denoised=True, repeat_noise=False, noise=None,
                 return_x0=False):
        """One reverse step (diffusion.py:121-129).  ``cond`` is ``[B,H,T]`` as in the reference."""
        lib, den, prec, stream = self._call_ctx(x_t)
        B, _, M, T = x_t.shape
        with torch.cuda.device(x_t.device):
            if noise is None:
                noise = (torch.randn((1, 1, M, T), device=x_t.device).repeat(B, 1, 1, 1) if repeat_noise
                         else torch.randn(x_t.shape, device=x_t.device))
            x = self._f32c(x_t)
            cond_bth = self._f32c(cond.transpose(1, 2))
            spk = self._spk(spk_emb)
            tt = t.detach().to(torch.int64).contiguous()
            # the reference's `extract` (a.gather(-1, t)) raises on a timestep outside [0, K); so does this (one host read)
            if tt.numel() and not bool(((tt >= 0) & (tt < self.num_timesteps)).all()):
                raise IndexError(f"p_sample: timestep outside [0, {self.num_timesteps})")
            nz = self._f32c(noise)
            out = torch.empty_like(x)
            x0 = torch.empty_like(x) if return_x0 else None
            ws = den.workspace(B, T, self.num_timesteps, x.device)
            _lib.check(lib.mgb_reverse_step(
                C.byref(den.dims), prec, _lib.ptr(den.packed_weights()), _lib.ptr(x), _lib.ptr(tt),
                _lib.ptr(cond_bth), _lib.ptr(spk), _lib.ptr(nz), _lib.ptr(self._sched(x.device)),
                self.num_timesteps, int(bool(clip_denoised)), _lib.ptr(out), _lib.ptr(x0), B, T,
                _lib.ptr(ws), ws.numel(), stream), "mgb_reverse_step")
        return (out, x0) if return_x0 else out

    def _sample_core(self, cond_bth, spk, x_T, noises, pad_mask, want_states, clip=True):
        """Runs K reverse steps + denorm(+mask) in the library.  All tensors fp32 contiguous on one
        CUDA device: ``cond_bth [B,T,H]``, ``x_T [B,1,M,T]``, ``noises [K,B,1,M,T]``,
        ``pad_mask`` uint8 ``[B,T]`` (1 = padding) or None."""
        lib, den, prec, stream = self._call_ctx(x_T)
        B, _, M, T = x_T.shape
        K = self.num_timesteps
        dev = x_T.device
        mel = torch.empty((B, T, M), dtype=torch.float32, device=dev)
        states = torch.empty((K + 1, B, T, M), dtype=torch.float32, device=dev) if want_states else None
        ws = den.workspace(B, T, K, dev)
        smin = self.spec_min.detach().float().reshape(-1).contiguous()
        smax = self.spec_max.detach().float().reshape(-1).contiguous()
        _lib.check(lib.mgb_sample(
            C.byref(den.dims), prec, _lib.ptr(den.packed_weights()), _lib.ptr(x_T), _lib.ptr(cond_bth),
            _lib.ptr(spk), _lib.ptr(noises), _lib.ptr(self._sched(dev)), K, int(bool(clip)), _lib.ptr(smin),
            _lib.ptr(smax), _lib.ptr(pad_mask), _lib.ptr(states), _lib.ptr(mel), None, B, T,
            _lib.ptr(ws), ws.numel(), stream), "mgb_sample")
        return mel, states

    def _draw(self, B, T, dev, x_T, noises):
        M, K = self.mel_bins, self.num_timesteps
        if x_T is None:
            x_T = torch.randn((B, 1, M, T), device=dev)
        if noises is None:
            noises = torch.randn((K, B, 1, M, T), device=dev)   # the reference draws one per step too
        if tuple(noises.shape) != (K, B, 1, M, T):
            raise ValueError(f"noises must be [K,B,1,M,T] = {(K, B, 1, M, T)}, got {tuple(noises.shape)}")
        return self._f32c(x_T), self._f32c(noises)

    @torch.no_grad()
    def sampling(self, noise=None, noises=None):
        """diffusion.py:155-165: returns the K+1 denormalised ``[B,T,M]`` states (start first).
        ``noise`` is the start state x_T (reference name); ``noises[t]`` the per-step draws."""
        cond = self.cond                                   # [B,H,T], stashed by forward()
        B, _, T = cond.shape
        dev = cond.device
        with torch.cuda.device(dev):
            x_T, noises = self._draw(B, T, dev, noise, noises)
            _, states = self._sample_core(self._f32c(cond.transpose(1, 2)), self._spk(self.spk_emb), x_T, noises,
                                          None, want_states=True)
        return list(states.unbind(0))

    @torch.no_grad()
    def interpolate(self, x1, x2, t, cond, spk_emb, lam=0.5):
        b, device = x1.shape[0], x1.device
        t = self.num_timesteps - 1 if t is None else t
        assert x1.shape == x2.shape
        tb = torch.full((b,), t, device=device, dtype=torch.long)
        x = (1 - lam) * self.q_sample(x1, t=tb) + lam * self.q_sample(x2, t=tb)
        for i in reversed(range(0, t)):
            x = self.p_sample(x, torch.full((b,), i, device=device, dtype=torch.long), cond, spk_emb)
        return self.denorm_spec(x[:, 0].transpose(1, 2))

    def forward(self, mel, cond, spk_emb, mel_mask, coarse_mel=None, clip_denoised=True, *,
                x_T=None, noises=None, start_noise=None, t=None, noise_t=None, noise_prev=None, post_noise=None):
        """diffusion.py:187-226.  ``cond [B,T,H]``; ``mel_mask [B,T]`` True = padding.
        Inference (``mel is None``) returns ``(x_0_pred [B,T,M], None, None, None, t)``; with ``mel`` given the
        reference's training-branch 5-tuple ``(x_0_pred, x_t, x_t_prev, x_t_prev_pred, t)``."""
        b, device = cond.shape[0], cond.device
        self.cond = cond.transpose(1, 2).detach()
        self.spk_emb = spk_emb.detach() if spk_emb is not None else None
        if device.type != "cuda":
            raise RuntimeError("mixgan_tts_b200.GaussianDiffusion needs CUDA tensors (no CPU fallback)")
        if mel is not None:
            return self._forward_training(mel, cond, spk_emb, mel_mask, coarse_mel, clip_denoised, t, noise_t,
                                          noise_prev, post_noise)
        t = None
        T = cond.shape[1]
        with torch.no_grad(), torch.cuda.device(device):
            lib, den, prec, stream = self._call_ctx(cond)
            pad = mel_mask.detach().to(torch.uint8).contiguous()
            cond_bth = self._f32c(cond)
            spk = self._spk(spk_emb)
            if self.model == "shallow":   # x_T is produced by the shallow start, never drawn
                x_T = torch.empty((b, 1, self.mel_bins, T), dtype=torch.float32, device=device)
            x_T, noises = self._draw(b, T, device, x_T, noises)
            if self.model == "shallow":
                t = torch.full((b,), self.num_timesteps - 1, device=device, dtype=torch.long)
                if start_noise is None:
                    start_noise = torch.randn((b, 1, self.mel_bins, T), device=device)
                k = self.num_timesteps - 1
                self._sched(device)
                sa, sn = self._sched_cache[2][0][k], self._sched_cache[2][1][k]
                _lib.check(lib.mgb_shallow_start(
                    _lib.ptr(self._f32c(coarse_mel)), _lib.ptr(self._f32c(start_noise)),
                    _lib.ptr(self.spec_min.detach().float().reshape(-1).contiguous()),
                    _lib.ptr(self.spec_max.detach().float().reshape(-1).contiguous()),
                    sa, sn,
                    _lib.ptr(pad), _lib.ptr(x_T), b, T, self.mel_bins, stream), "mgb_shallow_start")
            x_0_pred, _ = self._sample_core(cond_bth, spk, x_T, noises, pad, want_states=False,
                                            clip=clip_denoised)
        return x_0_pred, None, None, None, t

    def _forward_training(self, mel, cond, spk_emb, mel_mask, coarse_mel, clip_denoised, t, noise_t, noise_prev,
                          post_noise):
        """diffusion.py:201-225.  The Denoiser call runs in the library (per-utterance timesteps; with autograd
        enabled its backward runs there too, see ``modules._DenoiserGradFn``).  The elementwise steps around it are two
        library kernels (``mgb_train_diffuse`` before, ``mgb_train_posterior`` + its backward after: ``_PosteriorFn``)
        when ``mel`` and the noises carry no gradient; otherwise — and for the 'shallow' model's posterior, which starts
        from the coarse mel — the torch composition on the registered buffers, through which torch autograd chains."""
        b, device = cond.shape[0], cond.device
        with torch.cuda.device(device):
            t_given = t is not None
            if t is None:
                t = torch.randint(0, self.num_timesteps, (b,), device=device)       # :203
            t = t.long()
            tr = lambda x: x[:, 0].transpose(1, 2)
            fused = (mel.dtype == torch.float32 and not mel.requires_grad and mel_mask is not None
                     and all(n is None or not n.requires_grad for n in (noise_t, noise_prev, post_noise)))
            if fused:
                # :206-207 in one library kernel (norm_spec, transpose, both q_samples, the t - 1 < 0 rule, the mask); the
                # draws happen in the reference's order and shapes, so the random stream is the reference's
                lib = _lib.load()
                K, M, T = self.num_timesteps, self.mel_bins, mel.shape[1]
                # a caller-supplied timestep outside [0, K) raises as the reference's gather does (one host read; skipped
                # for the internal draw and under graph capture, where the kernels clamp the table index instead)
                if t_given and not torch.cuda.is_current_stream_capturing() and not bool(((t >= 0) & (t < K)).all()):
                    raise IndexError(f"forward: timestep outside [0, {K})")
                shape = (b, 1, M, T)
                noise_t = torch.randn(shape, device=device) if noise_t is None else self._f32c(noise_t)
                noise_prev = torch.randn(shape, device=device) if noise_prev is None else self._f32c(noise_prev)
                pad = mel_mask.detach().to(torch.uint8).contiguous()
                x_t, x_t_prev = torch.empty(shape, device=device), torch.empty(shape, device=device)
                sched = self._sched(device)
                sa, sn = self._train_tables(device)
                st = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
                _lib.check(lib.mgb_train_diffuse(
                    _lib.ptr(self._f32c(mel)), _lib.ptr(noise_t), _lib.ptr(noise_prev),
                    _lib.ptr(self.spec_min.detach().float().reshape(-1).contiguous()),
                    _lib.ptr(self.spec_max.detach().float().reshape(-1).contiguous()), _lib.ptr(sa), _lib.ptr(sn), _lib.ptr(t),
                    _lib.ptr(pad), _lib.ptr(x_t), _lib.ptr(x_t_prev), b, T, M, K, st), "mgb_train_diffuse")
                den_out = self.denoise_fn(x_t, t, cond.transpose(1, 2), spk_emb)    # :210
                if self.model != "shallow":
                    post_noise = torch.randn(shape, device=device) if post_noise is None else self._f32c(post_noise)
                    x_0_pred, x_t_prev_pred = _PosteriorFn.apply(den_out, x_t, post_noise, sched, t, pad, bool(clip_denoised), K)
                    return tr(x_0_pred), tr(x_t), tr(x_t_prev), tr(x_t_prev_pred), t
                valid = (~mel_mask)[:, None, None, :]
                x_0_pred = den_out * valid
            else:
                valid = (~mel_mask)[:, None, None, :]                                   # :190, :202
                x_t = self.diffuse_fn(mel, t.clone(), noise=noise_t) * valid             # :206
                x_t_prev = self.diffuse_fn(mel, t - 1, noise=noise_prev) * valid         # :207
                x_0_pred = self.denoise_fn(x_t, t, cond.transpose(1, 2), spk_emb) * valid   # :210
            if clip_denoised:
                x_0_pred = x_0_pred.clamp(-1., 1.)                                  # :211-212
            if self.model != "shallow":
                x_start = x_0_pred
            else:
                x_start = self.norm_spec(coarse_mel).transpose(1, 2)[:, None, :, :]  # :218-219
            x_t_prev_pred = self.q_posterior_sample(x_start=x_start, x_t=x_t, t=t, noise=post_noise) * valid   # :220
        return tr(x_0_pred), tr(x_t), tr(x_t_prev), tr(x_t_prev_pred), t
