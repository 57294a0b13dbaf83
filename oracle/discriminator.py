"""CPU restatement of the reference JCU discriminator (torch fp32, differentiable): ``model/mixgantts.py:186-288``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Functional form over a ``{state_dict key: tensor}`` dict so that
torch autograd through it gives the reference's parameter gradients.  Pinned against the real ``JCUDiscriminator`` by
``tests/golden/jcu_*.npz`` (``tests/golden/make_golden_jcu.py``) and live by ``tests/test_oracle_vs_reference.py``.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def diffusion_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:          # model/blocks.py:906-913
    half = dim // 2
    e = math.log(10000) / (half - 1)
    e = torch.exp(torch.arange(half) * -e)
    e = t[:, None] * e[None, :]
    return torch.cat((e.sin(), e.cos()), dim=-1)


def mish(x):                                                                # model/blocks.py:894-896
    return x * torch.tanh(F.softplus(x))


def jcu_forward(W: dict, cfg: dict, x_ts, x_t_prevs, s, t, *, residual_channels=256, multi_speaker=False):
    """``cfg`` = ``model_config["discriminator"]``.  Returns ``(cond_feats, uncond_feats)`` as the reference."""
    n_layer, n_unc, n_cond = cfg["n_layer"], cfg["n_uncond_layer"], cfg["n_cond_layer"]
    ks, st = cfg["kernel_sizes"], cfg["strides"]

    def conv(prefix, i, x):
        k = ks[i]
        return F.conv1d(x, W[f"{prefix}.conv.weight"], W[f"{prefix}.conv.bias"], stride=st[i], padding=(k - 1) // 2)

    x = F.linear(torch.cat([x_t_prevs, x_ts], dim=-1), W["input_projection.linear.weight"]).transpose(1, 2)   # :263-265
    emb = diffusion_embedding(t, residual_channels)
    step = F.linear(mish(F.linear(emb, W["mlp.0.linear.weight"])), W["mlp.2.linear.weight"]).unsqueeze(-1)      # :266
    spk = F.linear(s, W["spk_mlp.0.linear.weight"]).unsqueeze(-1) if multi_speaker else None                   # :267-268
    cond_feats, uncond_feats = [], []
    for i in range(n_layer):                                                                                   # :272-275
        x = F.leaky_relu(conv(f"conv_block.{i}", i, x), 0.2)
        cond_feats.append(x)
        uncond_feats.append(x)
    x_cond = (x + step + spk) if multi_speaker else (x + step)                                                 # :277-278
    x_uncond = x
    for i in range(n_cond):                                                                                    # :281-283
        x_cond = F.leaky_relu(conv(f"cond_conv_block.{i}", n_layer + i, x_cond), 0.2)
        cond_feats.append(x_cond)
    for i in range(n_unc):                                                                                     # :285-287
        x_uncond = F.leaky_relu(conv(f"uncond_conv_block.{i}", n_layer + i, x_uncond), 0.2)
        uncond_feats.append(x_uncond)
    return cond_feats, uncond_feats


def lsgan_jcu_loss(logit_cond, logit_uncond, label: float):                 # model/loss.py:14-19 (mask=None)
    tgt_c, tgt_u = torch.full_like(logit_cond, label), torch.full_like(logit_uncond, label)
    return 0.5 * (F.mse_loss(logit_cond, tgt_c) + F.mse_loss(logit_uncond, tgt_u))


def fm_loss(D_real_cond, D_real_uncond, D_fake_cond, D_fake_uncond, n_layers):   # model/loss.py:221-227
    loss, w = 0, 4.0 / (n_layers + 1)
    for j in range(len(D_fake_cond) - 1):
        loss = loss + w * 0.5 * (F.l1_loss(D_real_cond[j].detach(), D_fake_cond[j]) + F.l1_loss(D_real_uncond[j].detach(), D_fake_uncond[j]))
    return loss
