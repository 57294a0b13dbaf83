"""Drop-in ``LengthRegulator`` (reference: ``model/linguistic_encoder.py:383-416``) on the GPU.

Integer indexing, bit-exact: the expansion is an exclusive scan of the clamped durations followed
by a row gather (``mgb_length_regulate``), instead of the reference's Python double loop with one
``.item()`` device sync per phoneme.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from . import _lib


class LengthRegulator(nn.Module):
    def LR(self, x, duration, max_len):
        if x.device.type != "cuda":
            raise RuntimeError("mixgan_tts_b200.LengthRegulator needs CUDA tensors (no CPU fallback)")
        lib = _lib.load()
        B, S, D = x.shape
        dur = duration.detach().to(torch.int64).contiguous()
        xs = x.detach().float().contiguous()
        with torch.cuda.device(x.device):
            if max_len is None:   # batch maximum: the one host read the reference also needs (pad())
                max_len = int(dur.clamp(min=0).sum(dim=1).max().item())
            max_len = max(int(max_len), 1)
            out = torch.empty((B, max_len, D), dtype=torch.float32, device=x.device)
            mel_len = torch.empty((B,), dtype=torch.int64, device=x.device)
            ws = torch.empty((B * (S + 1),), dtype=torch.int64, device=x.device)
            stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
            _lib.check(lib.mgb_length_regulate(_lib.ptr(xs), _lib.ptr(dur), _lib.ptr(out), _lib.ptr(mel_len),
                                               B, S, D, max_len, _lib.ptr(ws), ws.numel() * 8, stream),
                       "mgb_length_regulate")
        return out, mel_len

    def forward(self, x, duration, max_len):
        return self.LR(x, duration, max_len)


def get_mask_from_lengths(lengths, max_len=None):
    """utils/tools.py:144-153: True = valid frame."""
    if max_len is None:
        max_len = int(torch.max(lengths).item())
    ids = torch.arange(0, max_len, device=lengths.device).unsqueeze(0)
    return ids < lengths.unsqueeze(1)
