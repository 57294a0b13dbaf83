"""CPU restatement of the reference HiFi-GAN generator (torch fp32), weight norm removed: ``hifigan/models.py:112-173``
(``Generator.forward`` :151-166, ``ResBlock.forward`` :96-103) as ``utils/model.py:103-121`` calls it.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Functional form over a ``{state_dict key: tensor}`` dict with plain
``weight`` / ``bias`` keys.  Pinned against the real ``hifigan.Generator`` by ``tests/golden/hifigan_*.npz``
(``tests/golden/make_golden_aux.py``) and live by ``tests/test_oracle_vs_reference.py``.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.1                                                           # hifigan/models.py:7


def generator_forward(W, x, cfg):
    """``x [B, n_mel, T]`` -> ``[B, 1, T * hop]``; ``cfg`` = the hifigan/config.json mapping."""
    rates, kernels = cfg["upsample_rates"], cfg["upsample_kernel_sizes"]
    rks, rds = cfg["resblock_kernel_sizes"], cfg["resblock_dilation_sizes"]
    nk = len(rks)
    x = F.conv1d(x, W["conv_pre.weight"], W["conv_pre.bias"], padding=3)
    for i, (u, k) in enumerate(zip(rates, kernels)):
        x = F.leaky_relu(x, LRELU_SLOPE)
        x = F.conv_transpose1d(x, W[f"ups.{i}.weight"], W[f"ups.{i}.bias"], stride=u, padding=(k - u) // 2)
        xs = None
        for j in range(nk):
            r = f"resblocks.{i * nk + j}"
            y = x
            for m, d in enumerate(rds[j]):                                  # ResBlock.forward :96-103
                kk = rks[j]
                xt = F.leaky_relu(y, LRELU_SLOPE)
                xt = F.conv1d(xt, W[f"{r}.convs1.{m}.weight"], W[f"{r}.convs1.{m}.bias"], dilation=d, padding=(kk * d - d) // 2)
                xt = F.leaky_relu(xt, LRELU_SLOPE)
                xt = F.conv1d(xt, W[f"{r}.convs2.{m}.weight"], W[f"{r}.convs2.{m}.bias"], padding=(kk - 1) // 2)
                y = xt + y
            xs = y if xs is None else xs + y
        x = xs / nk
    x = F.leaky_relu(x)                                                     # default slope 0.01 (:162)
    x = F.conv1d(x, W["conv_post.weight"], W["conv_post.bias"], padding=3)
    return torch.tanh(x)
