"""Drop-in ``LengthRegulator`` (reference: ``model/linguistic_encoder.py:383-416``) and the duration / mask helpers either
side of it (``linguistic_encoder.py:310-316``, ``utils/tools.py:144-153``) on the GPU.

Integer indexing, bit-exact: the expansion is an exclusive scan of the clamped durations followed by a row gather
(``mgb_length_regulate``), instead of the reference's Python double loop with one ``.item()`` device sync per phoneme.
Like the reference's ``expand`` + ``cat`` + ``pad`` the expansion is differentiable with respect to ``x`` (the path that
trains the encoder through the decoder): the backward is a segment sum (``mgb_length_regulate_backward``).
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from . import _lib


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _regulate(x, duration, max_len, want_mask=False):
    """Returns (out [B,L,D] fp32, mel_len [B] int64, mask uint8 [B,L] | None, scan scratch)."""
    if x.device.type != "cuda":
        raise RuntimeError("mixgan_tts_b200.LengthRegulator needs CUDA tensors (no CPU fallback)")
    lib = _lib.load()
    B, S, D = x.shape
    # expand() applies int() to every duration: a float duration is truncated towards zero (linguistic_encoder.py:408-409)
    dur = duration.detach().to(torch.int64).contiguous()
    xs = x.detach().float().contiguous()
    with torch.cuda.device(x.device):
        if not max_len:       # pad(): a falsy max_len means the batch maximum - the one host read the reference needs too
            max_len = int(dur.clamp(min=0).sum(dim=1).max().item())
        max_len = max(int(max_len), 1)
        out = torch.empty((B, max_len, D), dtype=torch.float32, device=x.device)
        mel_len = torch.empty((B,), dtype=torch.int64, device=x.device)
        mask = torch.empty((B, max_len), dtype=torch.uint8, device=x.device) if want_mask else None
        ws = torch.empty((B * (S + 1),), dtype=torch.int64, device=x.device)
        _lib.check(lib.mgb_length_regulate(_lib.ptr(xs), _lib.ptr(dur), _lib.ptr(out), _lib.ptr(mel_len), _lib.ptr(mask),
                                           B, S, D, max_len, _lib.ptr(ws), ws.numel() * 8, _stream(x.device)),
                   "mgb_length_regulate")
    return out, mel_len, mask, ws


class _LRFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, duration, max_len):
        out, mel_len, _, ws = _regulate(x, duration, max_len)
        ctx.save_for_backward(ws)
        ctx.shape = tuple(x.shape)
        ctx.max_len = out.shape[1]
        ctx.in_dtype = x.dtype
        ctx.mark_non_differentiable(mel_len)
        return out.to(x.dtype) if x.dtype != torch.float32 else out, mel_len

    @staticmethod
    def backward(ctx, gout, _gl):
        (ws,) = ctx.saved_tensors
        B, S, D = ctx.shape
        lib = _lib.load()
        g = gout.detach().float().contiguous()
        gx = torch.empty((B, S, D), dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            _lib.check(lib.mgb_length_regulate_backward(_lib.ptr(g), _lib.ptr(ws), _lib.ptr(gx), B, S, D, ctx.max_len,
                                                        _stream(g.device)), "mgb_length_regulate_backward")
        return gx.to(ctx.in_dtype), None, None


class LengthRegulator(nn.Module):
    def LR(self, x, duration, max_len):
        if torch.is_grad_enabled() and x.requires_grad:
            return _LRFn.apply(x, duration, max_len)
        out, mel_len, _, _ = _regulate(x, duration, max_len)
        return out, mel_len

    def forward(self, x, duration, max_len):
        return self.LR(x, duration, max_len)

    def regulate_with_mask(self, x, duration, max_len=None):
        """``(out, mel_len, mel_mask)`` in one pass: ``mel_mask`` is ``get_mask_from_lengths(mel_len, out.shape[1])``
        (True = valid frame), as ``LinguisticEncoder.forward`` builds it right after the regulator (:315-316)."""
        out, mel_len, mask, _ = _regulate(x, duration, max_len, want_mask=True)
        return out, mel_len, mask.bool()


def durations_from_log(log_duration: torch.Tensor, d_control: float = 1.0) -> torch.Tensor:
    """``clamp(round(exp(log_d) - 1) * d_control, min=0).long()`` (linguistic_encoder.py:310-314) in the library."""
    if log_duration.device.type != "cuda":
        raise RuntimeError("mixgan_tts_b200.durations_from_log needs CUDA tensors (no CPU fallback)")
    lib = _lib.load()
    ld = log_duration.detach().float().contiguous()
    dur = torch.empty(ld.shape, dtype=torch.int64, device=ld.device)
    if ld.numel():
        with torch.cuda.device(ld.device):
            _lib.check(lib.mgb_durations_from_log(_lib.ptr(ld), float(d_control), _lib.ptr(dur), ld.numel(), _stream(ld.device)),
                       "mgb_durations_from_log")
    return dur


def get_mask_from_lengths(lengths, max_len=None):
    """utils/tools.py:144-153: True = valid frame."""
    if lengths.device.type != "cuda":
        raise RuntimeError("mixgan_tts_b200.get_mask_from_lengths needs CUDA tensors (no CPU fallback)")
    if max_len is None:
        max_len = int(torch.max(lengths).item())
    lib = _lib.load()
    ln = lengths.detach().to(torch.int64).contiguous()
    B = ln.shape[0]
    mask = torch.empty((B, max(int(max_len), 0)), dtype=torch.uint8, device=ln.device)
    if mask.numel():
        with torch.cuda.device(ln.device):
            _lib.check(lib.mgb_mask_from_lengths(_lib.ptr(ln), _lib.ptr(mask), B, int(max_len), _stream(ln.device)),
                       "mgb_mask_from_lengths")
    return mask.bool()
