"""CPU restatement of the reference LengthRegulator, duration rounding and mask helper (numpy, integer exact).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Follows ``model/linguistic_encoder.py:383-416`` (``LR`` / ``expand``: every source row is repeated
``max(int(duration), 0)`` times; ``int()`` truncates a float duration towards zero), ``utils/tools.py:374-392``
(``pad``: utterances are zero-padded to ``max_len`` when it is truthy, else to the batch maximum; an utterance LONGER
than ``max_len`` is silently truncated, because ``F.pad`` with a negative size crops) and ``utils/tools.py:144-153``
(``get_mask_from_lengths``: returns True = valid).  ``mel_len`` is the TRUE expanded length even when the rows were
cropped (``linguistic_encoder.py:396``).  The duration rounding used at inference,
``clamp(round(exp(log_d) - 1) * d_control, min=0).long()``, is ``linguistic_encoder.py:310-314`` (``torch.round`` rounds
half to even, ``.long()`` truncates towards zero).

Pinned against the real reference by ``tests/golden/length_regulator.npz`` (made by ``tests/golden/make_golden_lr.py``) and,
where ``/root/reference`` exists, live by ``tests/test_oracle_vs_reference.py``.
"""
from __future__ import annotations

import numpy as np


def durations_from_log(log_d: np.ndarray, d_control: float = 1.0) -> np.ndarray:
    d = np.round(np.exp(log_d.astype(np.float32)) - np.float32(1.0)) * np.float32(d_control)
    return np.trunc(np.maximum(d, np.float32(0))).astype(np.int64)


def length_regulate(x: np.ndarray, dur: np.ndarray, max_len: int | None = None):
    """x [B,S,D] float32, dur [B,S] (int64, or float: truncated towards zero) -> (out [B,L,D], mel_len [B] int64)."""
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
        n = min(e.shape[0], L)   # F.pad with a negative size crops (utils/tools.py:385-387); mel_len keeps the true length
        out[b, :n] = e[:n]
    return out, np.asarray(lens, dtype=np.int64)


def length_regulate_backward(grad_out: np.ndarray, dur: np.ndarray) -> np.ndarray:
    """d/dx of ``length_regulate``: the gradient rows of a source row's copies, summed (float64 accumulation)."""
    B, L, D = grad_out.shape
    S = dur.shape[1]
    gx = np.zeros((B, S, D), dtype=np.float64)
    for b in range(B):
        f = 0
        for s in range(S):
            n = max(int(dur[b, s]), 0)
            lo, hi = min(f, L), min(f + n, L)
            if hi > lo:
                gx[b, s] = grad_out[b, lo:hi].astype(np.float64).sum(0)
            f += n
    return gx


def mask_from_lengths(lengths: np.ndarray, max_len: int | None = None) -> np.ndarray:
    L = int(lengths.max()) if max_len is None else max_len
    return np.arange(L)[None, :] < lengths[:, None]
