"""CPU restatement of the reference aux decoder (torch fp32): ``model/mixgantts.py:139-143`` —
``Decoder`` ``transformer/Models.py:103-171`` (eval branch), ``FFTBlock`` ``transformer/Layers.py:11-31``,
``MultiHeadAttention`` / ``PositionwiseFeedForward`` ``transformer/SubLayers.py:8-97``, ``ScaledDotProductAttention``
``transformer/Modules.py:6-24``, ``PostNet`` ``transformer/Layers.py:67-137`` (BatchNorm1d with running statistics).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Functional form over a ``{MixGANTTS state_dict key: tensor}`` dict.
Pinned against the real modules by ``tests/golden/auxdec_*.npz`` (``tests/golden/make_golden_aux.py``) and live by
``tests/test_oracle_vs_reference.py``.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


def sinusoid_table(n_position: int, d_hid: int) -> torch.Tensor:           # transformer/Models.py:11-31
    pos = np.arange(n_position)[:, None].astype(np.float64)
    hid = np.arange(d_hid)[None, :]
    tab = pos / np.power(10000, 2 * (hid // 2) / d_hid)
    tab[:, 0::2] = np.sin(tab[:, 0::2])
    tab[:, 1::2] = np.cos(tab[:, 1::2])
    return torch.from_numpy(tab.astype(np.float32))


def multi_head_attention(W, p, x, key_mask, n_head):                        # SubLayers.py:31-59
    B, T, D = x.shape
    dk = D // n_head
    split = lambda y: y.view(B, T, n_head, dk).permute(2, 0, 1, 3).reshape(n_head * B, T, dk)
    q = split(F.linear(x, W[f"{p}.w_qs.weight"], W[f"{p}.w_qs.bias"]))
    k = split(F.linear(x, W[f"{p}.w_ks.weight"], W[f"{p}.w_ks.bias"]))
    v = split(F.linear(x, W[f"{p}.w_vs.weight"], W[f"{p}.w_vs.bias"]))
    attn = torch.bmm(q, k.transpose(1, 2)) / math.sqrt(dk)                  # Modules.py:16-17 (temperature = d_k ** 0.5)
    attn = attn.masked_fill(key_mask.unsqueeze(1).expand(-1, T, -1).repeat(n_head, 1, 1), -np.inf)   # :19-20
    out = torch.bmm(torch.softmax(attn, dim=2), v)
    out = out.view(n_head, B, T, dk).permute(1, 2, 0, 3).reshape(B, T, D)
    out = F.linear(out, W[f"{p}.fc.weight"], W[f"{p}.fc.bias"])             # dropout = identity (eval)
    return F.layer_norm(out + x, (D,), W[f"{p}.layer_norm.weight"], W[f"{p}.layer_norm.bias"])


def pos_ffn(W, p, x):                                                       # SubLayers.py:88-97
    k = W[f"{p}.w_1.weight"].shape[-1]
    h = F.relu(F.conv1d(x.transpose(1, 2), W[f"{p}.w_1.weight"], W[f"{p}.w_1.bias"], padding=(k - 1) // 2))
    out = F.conv1d(h, W[f"{p}.w_2.weight"], W[f"{p}.w_2.bias"]).transpose(1, 2)
    return F.layer_norm(out + x, (x.shape[-1],), W[f"{p}.layer_norm.weight"], W[f"{p}.layer_norm.bias"])


def decoder_forward(W, x, pad_mask, *, n_head=2, max_seq_len=1000):         # Models.py:137-171, eval
    B, T, D = x.shape
    if T > max_seq_len:
        x = x + sinusoid_table(T, D)[:T].unsqueeze(0)
    else:
        x = x + W["decoder.position_enc"][:, :T, :]
    layers = 1 + max(int(k.split(".")[2]) for k in W if k.startswith("decoder.layer_stack."))
    for i in range(layers):
        p = f"decoder.layer_stack.{i}"
        x = multi_head_attention(W, f"{p}.slf_attn", x, pad_mask, n_head).masked_fill(pad_mask.unsqueeze(-1), 0)   # Layers.py:24-27
        x = pos_ffn(W, f"{p}.pos_ffn", x).masked_fill(pad_mask.unsqueeze(-1), 0)                                    # :29-30
    return x


def postnet_forward(W, mel):                                                # Layers.py:128-137, eval
    x = mel.transpose(1, 2)
    n = 1 + max(int(k.split(".")[2]) for k in W if k.startswith("postnet.convolutions."))
    for i in range(n):
        c, b = f"postnet.convolutions.{i}.0.conv", f"postnet.convolutions.{i}.1"
        k = W[f"{c}.weight"].shape[-1]
        x = F.conv1d(x, W[f"{c}.weight"], W[f"{c}.bias"], padding=(k - 1) // 2)
        x = F.batch_norm(x, W[f"{b}.running_mean"], W[f"{b}.running_var"], W[f"{b}.weight"], W[f"{b}.bias"], False, 0.1, 1e-5)
        if i < n - 1:
            x = torch.tanh(x)
    return x.transpose(1, 2)


def aux_decoder_forward(W, x, pad_mask, *, n_head=2, max_seq_len=1000):
    """``(coarse_mels, decoder_output, mel_before_postnet)`` — model/mixgantts.py:139-143."""
    dec = decoder_forward(W, x, pad_mask, n_head=n_head, max_seq_len=max_seq_len)
    mel = F.linear(dec, W["mel_linear.weight"], W["mel_linear.bias"])
    return postnet_forward(W, mel) + mel, dec, mel
