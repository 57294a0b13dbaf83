"""CPU restatement of the reference ``Denoiser`` forward (torch fp32, host).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

``W`` is a ``{state_dict key: tensor}`` mapping with the reference's key names
(no ``denoise_fn.`` prefix), e.g. ``residual_layers.3.conv_layer.conv.weight``.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def as_torch(W: dict) -> dict:
    return {k: (v if isinstance(v, torch.Tensor) else torch.from_numpy(v)).float() for k, v in W.items()}


def num_layers(W: dict) -> int:
    n = 0
    while f"residual_layers.{n}.conv_layer.conv.weight" in W:
        n += 1
    return n


def diffusion_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    # model/blocks.py:906-913 — int64 [B] -> fp32 [B, dim], sin half then cos half
    half = dim // 2
    scale = math.log(10000) / (half - 1)
    freq = torch.exp(torch.arange(half, device=t.device) * -scale)
    arg = t[:, None] * freq[None, :]
    return torch.cat((arg.sin(), arg.cos()), dim=-1)


def mish(x: torch.Tensor) -> torch.Tensor:
    # model/blocks.py:894-896
    return x * torch.tanh(F.softplus(x))


def step_mlp(W: dict, t: torch.Tensor) -> torch.Tensor:
    # model/modules.py:433-434 — both LinearNorm layers are bias-free (blocks.py:281)
    C = W["mlp.2.linear.weight"].shape[0]
    e = diffusion_embedding(t, C)
    e = F.linear(e, W["mlp.0.linear.weight"])
    return F.linear(mish(e), W["mlp.2.linear.weight"])


def residual_block(W: dict, l: int, x, cond, d, spk):
    # model/blocks.py:1157-1176
    p = f"residual_layers.{l}"
    dl = F.linear(d, W[f"{p}.diffusion_projection.linear.weight"]).unsqueeze(-1)        # :1159
    c = F.conv1d(cond, W[f"{p}.conditioner_projection.conv.weight"],
                 W[f"{p}.conditioner_projection.conv.bias"])                            # :1160
    res = x + dl                                                                         # :1166
    y = res + c
    if f"{p}.speaker_projection.linear.weight" in W:                                     # :1161-1164
        s = F.linear(spk, W[f"{p}.speaker_projection.linear.weight"]).unsqueeze(-1)
        y = y + s
    y = F.conv1d(y, W[f"{p}.conv_layer.conv.weight"], W[f"{p}.conv_layer.conv.bias"],
                 padding=1)                                                              # :1167-1169
    gate, filt = torch.chunk(y, 2, dim=1)                                                # :1170
    y = torch.sigmoid(gate) * torch.tanh(filt)                                           # :1171
    y = F.conv1d(y, W[f"{p}.output_projection.conv.weight"],
                 W[f"{p}.output_projection.conv.bias"])                                  # :1173
    xo, skip = torch.chunk(y, 2, dim=1)                                                  # :1174
    return (xo + res) / math.sqrt(2.0), skip                                             # :1176


def denoiser_forward_graph(W: dict, mel, t, cond, spk=None):
    """``mel [B,1,M,T]``, ``t`` int64 ``[B]``, ``cond [B,H,T]``, ``spk [B,H]|None`` -> ``[B,1,M,T]``.

    model/modules.py:420-446.  Plain torch ops, so torch autograd differentiates it exactly as it differentiates
    the reference module: this is the gradient oracle of the training path.
    """
    L = num_layers(W)
    x = mel[:, 0]
    x = F.relu(F.conv1d(x, W["input_projection.0.conv.weight"], W["input_projection.0.conv.bias"]))
    d = step_mlp(W, t)
    skip_sum = None
    for l in range(L):
        x, skip = residual_block(W, l, x, cond, d, spk)
        skip_sum = skip if skip_sum is None else skip_sum + skip
    x = skip_sum / math.sqrt(L)                                                          # :441
    x = F.relu(F.conv1d(x, W["skip_projection.conv.weight"], W["skip_projection.conv.bias"]))
    x = F.conv1d(x, W["output_projection.conv.weight"], W["output_projection.conv.bias"])
    return x[:, None, :, :]


@torch.no_grad()
def denoiser_forward(W: dict, mel, t, cond, spk=None):
    """Forward values only (inference)."""
    return denoiser_forward_graph(W, mel, t, cond, spk)
