"""CPU restatement of the reference LengthRegulator and mask helper (numpy, integer exact).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Follows ``model/linguistic_encoder.py:383-416`` (``LR`` / ``expand``: every source row is repeated
``max(int(duration), 0)`` times, utterances are zero-padded to ``max_len`` or the batch maximum),
``utils/tools.py:374-392`` (``pad``) and ``utils/tools.py:144-153`` (``get_mask_from_lengths``:
returns True = valid).  The duration rounding used at inference,
``clamp(round(exp(log_d) - 1) * d_control, min=0).long()`` is ``linguistic_encoder.py:310-316``.
"""
from __future__ import annotations

import numpy as np


def durations_from_log(log_d: np.ndarray, d_control: float = 1.0) -> np.ndarray:
    d = np.round(np.exp(log_d.astype(np.float32)) - np.float32(1.0)) * np.float32(d_control)
    return np.maximum(d, 0).astype(np.int64)


def length_regulate(x: np.ndarray, dur: np.ndarray, max_len: int | None = None):
    """x [B,S,D] float32, dur [B,S] int64 -> (out [B,L,D], mel_len [B] int64)."""
    B, S, D = x.shape
    outs, lens = [], []
    for b in range(B):
        rows = []
        for s in range(S):
            n = max(int(dur[b, s]), 0)
            rows.append(np.broadcast_to(x[b, s], (n, D)))
        e = np.concatenate(rows, 0) if rows else np.zeros((0, D), x.dtype)
        outs.append(e)
        lens.append(e.shape[0])
    L = max_len if max_len else max(lens)
    out = np.zeros((B, L, D), dtype=x.dtype)
    for b, e in enumerate(outs):
        n = min(e.shape[0], L)   # the GPU entry point truncates; the reference would raise on overflow
        out[b, :n] = e[:n]
    return out, np.asarray(lens, dtype=np.int64)


def mask_from_lengths(lengths: np.ndarray, max_len: int | None = None) -> np.ndarray:
    L = int(lengths.max()) if max_len is None else max_len
    return np.arange(L)[None, :] < lengths[:, None]
