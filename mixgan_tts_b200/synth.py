"""Deterministic synthetic weights and inputs for the parity tests and the bench.

There is no network for checkpoints or datasets, so the Denoiser is random-init
and the inputs are random tensors of the LJSpeech / AISHELL3 shapes
(SURVEY.md §8d).  Everything is drawn from a ``numpy`` PCG64 stream so that the
same arrays come out in this container (where the goldens are made with the
real reference) and on the GPU box (where the reference does not exist).

Weight names and shapes are the reference's ``state_dict`` keys for
``Denoiser`` (``model/modules.py:385-418``, ``model/blocks.py:1133-1155``):
conv weights ``[out, in, k]``, linear weights ``[out, in]``.  Init scales follow
the reference constructors (PyTorch default conv init, xavier-uniform for the
bias-free ``LinearNorm``), except ``output_projection.conv.weight`` which the
reference zero-initialises (``model/modules.py:418``) and which is drawn
N(0, 0.05^2) here so that parity is not vacuous.
"""
from __future__ import annotations

import hashlib
import numpy as np


def _rng(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(seed))


def _uniform(g, shape, bound):
    return ((g.random(shape, dtype=np.float32) * 2.0 - 1.0) * np.float32(bound)).astype(np.float32)


def make_denoiser_weights(seed: int = 0, *, n_mel: int = 80, channels: int = 256,
                          d_encoder: int = 256, layers: int = 20,
                          multi_speaker: bool = False) -> dict:
    """Return ``{state_dict key: float32 ndarray}`` for one Denoiser."""
    g = _rng(seed)
    C, H, M = channels, d_encoder, n_mel
    w = {}

    def conv(name, cout, cin, k):
        bound = 1.0 / np.sqrt(cin * k)
        w[f"{name}.conv.weight"] = _uniform(g, (cout, cin, k), bound)
        w[f"{name}.conv.bias"] = _uniform(g, (cout,), bound)

    def linear(name, cout, cin):
        bound = np.sqrt(6.0 / (cin + cout))
        w[f"{name}.linear.weight"] = _uniform(g, (cout, cin), bound)

    conv("input_projection.0", C, M, 1)
    linear("mlp.0", 4 * C, C)
    linear("mlp.2", C, 4 * C)
    for l in range(layers):
        p = f"residual_layers.{l}"
        conv(f"{p}.conv_layer", 2 * C, C, 3)
        linear(f"{p}.diffusion_projection", C, C)
        if multi_speaker:
            linear(f"{p}.speaker_projection", C, H)
        conv(f"{p}.conditioner_projection", C, H, 1)
        conv(f"{p}.output_projection", 2 * C, C, 1)
    conv("skip_projection", C, C, 1)
    conv("output_projection", M, C, 1)
    w["output_projection.conv.weight"] = (
        g.standard_normal((M, C, 1), dtype=np.float32) * np.float32(0.05))
    return w


def make_discriminator_weights(seed: int = 0, *, n_mel: int = 80, residual_channels: int = 256, multi_speaker: bool = False,
                               cfg: dict | None = None, weight_std: float = 0.02) -> dict:
    """``{state_dict key: float32 ndarray}`` for one ``JCUDiscriminator`` (model/mixgantts.py:189-250): xavier-uniform
    bias-free linears, N(0, weight_std^2) conv weights (``weights_init``), PyTorch-default conv biases.  Tests use a larger
    ``weight_std`` than the reference's 0.02 so that the logits are not dominated by the biases."""
    cfg = cfg or {"n_layer": 3, "n_uncond_layer": 2, "n_cond_layer": 2, "n_channels": [64, 128, 512, 128, 1],
                  "kernel_sizes": [3, 5, 5, 5, 3], "strides": [1, 2, 2, 1, 1]}
    g = _rng(seed)
    w = {}

    def linear(name, cout, cin):
        w[f"{name}.linear.weight"] = _uniform(g, (cout, cin), np.sqrt(6.0 / (cin + cout)))

    def conv(name, cout, cin, k):
        w[f"{name}.conv.weight"] = (g.standard_normal((cout, cin, k), dtype=np.float32) * np.float32(weight_std))
        w[f"{name}.conv.bias"] = _uniform(g, (cout,), 1.0 / np.sqrt(cin * k))

    ch, ks, nl = cfg["n_channels"], cfg["kernel_sizes"], cfg["n_layer"]
    linear("input_projection", 2 * n_mel, 2 * n_mel)
    linear("mlp.0", 4 * residual_channels, residual_channels)
    linear("mlp.2", ch[nl - 1], 4 * residual_channels)
    if multi_speaker:
        linear("spk_mlp.0", ch[nl - 1], residual_channels)
    for i in range(nl):
        conv(f"conv_block.{i}", ch[i], ch[i - 1] if i else 2 * n_mel, ks[i])
    for i in range(nl, nl + cfg["n_uncond_layer"]):
        conv(f"uncond_conv_block.{i - nl}", ch[i], ch[i - 1], ks[i])
    for i in range(nl, nl + cfg["n_cond_layer"]):
        conv(f"cond_conv_block.{i - nl}", ch[i], ch[i - 1], ks[i])
    return w


def make_discriminator_inputs(seed: int, B: int, T: int, K: int, *, n_mel: int = 80, multi_speaker: bool = False) -> dict:
    """``x_ts``, ``x_t_prevs``, ``x_t_prev_preds`` ``[B,T,M]`` (normalised-mel scale), ``spk [B,256]`` or None, ``t [B]``."""
    g = _rng(seed)
    return {"x_ts": g.standard_normal((B, T, n_mel), dtype=np.float32),
            "x_t_prevs": g.standard_normal((B, T, n_mel), dtype=np.float32) * np.float32(0.7),
            "x_t_prev_preds": g.standard_normal((B, T, n_mel), dtype=np.float32) * np.float32(0.7),
            "spk": g.standard_normal((B, 256), dtype=np.float32) if multi_speaker else None,
            "t": np.array([(K - 1 - b) % K for b in range(B)], dtype=np.int64)}


AUXDEC_CFG = {"n_mel": 80, "d_model": 256, "n_head": 2, "d_inner": 1024, "ffn_kernel": 9, "layers": 6,
              "postnet_dim": 512, "postnet_kernel": 5, "postnet_layers": 5, "max_seq_len": 1000}


def sinusoid_table(n_position: int, d_hid: int) -> np.ndarray:
    """``get_sinusoid_encoding_table`` (transformer/Models.py:11-31), float64 numpy then one cast, as the reference does
    through ``torch.FloatTensor``."""
    pos = np.arange(n_position)[:, None].astype(np.float64)
    hid = np.arange(d_hid)[None, :]
    tab = pos / np.power(10000, 2 * (hid // 2) / d_hid)
    tab[:, 0::2] = np.sin(tab[:, 0::2])
    tab[:, 1::2] = np.cos(tab[:, 1::2])
    return tab.astype(np.float32)


def make_auxdec_weights(seed: int = 0, cfg: dict | None = None) -> dict:
    """``{key: float32 ndarray}`` for the aux decoder with the reference's ``MixGANTTS`` state_dict keys
    (``decoder.*`` transformer/Models.py:103-135, ``mel_linear.*`` model/mixgantts.py:59-62, ``postnet.*``
    transformer/Layers.py:67-126).  PyTorch-default uniform init for linears / convolutions; LayerNorm and BatchNorm affine
    parameters and the BatchNorm running statistics are perturbed so that folding them is actually tested."""
    c = dict(AUXDEC_CFG, **(cfg or {}))
    g = _rng(seed)
    D, H, M, P = c["d_model"], c["d_inner"], c["n_mel"], c["postnet_dim"]
    w = {"decoder.position_enc": sinusoid_table(c["max_seq_len"] + 1, D)[None]}

    def lin(name, cout, cin, k=None):
        fan = cin * (k or 1)
        shape = (cout, cin) if k is None else (cout, cin, k)
        w[f"{name}.weight"] = _uniform(g, shape, 1.0 / np.sqrt(fan))
        w[f"{name}.bias"] = _uniform(g, (cout,), 1.0 / np.sqrt(fan))

    def norm(name, n):
        w[f"{name}.weight"] = (1.0 + 0.1 * g.standard_normal((n,), dtype=np.float32)).astype(np.float32)
        w[f"{name}.bias"] = (0.1 * g.standard_normal((n,), dtype=np.float32)).astype(np.float32)

    for i in range(c["layers"]):
        p = f"decoder.layer_stack.{i}"
        lin(f"{p}.slf_attn.w_qs", D, D); lin(f"{p}.slf_attn.w_ks", D, D); lin(f"{p}.slf_attn.w_vs", D, D)
        norm(f"{p}.slf_attn.layer_norm", D)
        lin(f"{p}.slf_attn.fc", D, D)
        lin(f"{p}.pos_ffn.w_1", H, D, c["ffn_kernel"]); lin(f"{p}.pos_ffn.w_2", D, H, 1)
        norm(f"{p}.pos_ffn.layer_norm", D)
    lin("mel_linear", M, D)
    for i in range(c["postnet_layers"]):
        cin = M if i == 0 else P
        cout = M if i == c["postnet_layers"] - 1 else P
        lin(f"postnet.convolutions.{i}.0.conv", cout, cin, c["postnet_kernel"])
        b = f"postnet.convolutions.{i}.1"
        w[f"{b}.weight"] = (0.5 + g.random((cout,), dtype=np.float32)).astype(np.float32)
        w[f"{b}.bias"] = (0.1 * g.standard_normal((cout,), dtype=np.float32)).astype(np.float32)
        w[f"{b}.running_mean"] = (0.2 * g.standard_normal((cout,), dtype=np.float32)).astype(np.float32)
        w[f"{b}.running_var"] = (0.5 + g.random((cout,), dtype=np.float32)).astype(np.float32)
        w[f"{b}.num_batches_tracked"] = np.array(100, dtype=np.int64)
    return w


def make_auxdec_inputs(seed: int, B: int, T: int, *, d_model: int = 256, min_len_frac: float = 0.4) -> dict:
    """Decoder input ``x [B,T,256]`` (zero at padded frames, as the variance adaptor's length regulator leaves it),
    ``lens [B]`` (``lens[0] = T``) and ``pad_mask [B,T]`` (True = padding, the convention at ``self.decoder(output, mel_masks)``)."""
    g = _rng(seed)
    x = g.standard_normal((B, T, d_model), dtype=np.float32)
    lens = g.integers(max(1, int(T * min_len_frac)), T + 1, size=(B,), dtype=np.int64)
    lens[0] = T
    pad = np.arange(T)[None, :] >= lens[:, None]
    x[pad] = 0.0
    return {"x": x, "lens": lens, "pad_mask": pad}


HIFIGAN_CFG = {"resblock": "1", "upsample_rates": [8, 8, 2, 2], "upsample_kernel_sizes": [16, 16, 4, 4],
               "upsample_initial_channel": 512, "resblock_kernel_sizes": [3, 7, 11],
               "resblock_dilation_sizes": [[1, 3, 5], [1, 3, 5], [1, 3, 5]], "num_mels": 80}   # hifigan/config.json


def make_hifigan_weights(seed: int = 0, cfg: dict | None = None, gain: float = 1.0) -> dict:
    """``{key: float32 ndarray}`` for ``hifigan.Generator`` AFTER ``remove_weight_norm`` (plain ``weight`` / ``bias`` keys,
    hifigan/models.py:112-173).  The reference's N(0, 0.01) init makes every activation vanish after two layers, so weights
    are drawn variance-preserving (std = gain / sqrt(fan_in)) to keep the parity test meaningful at every depth."""
    c = dict(HIFIGAN_CFG, **(cfg or {}))
    g = _rng(seed)
    w = {}

    def conv(name, cout, cin, k, transposed=False, fan=None):
        fan = fan or cin * k
        shape = (cin, cout, k) if transposed else (cout, cin, k)
        w[f"{name}.weight"] = (g.standard_normal(shape, dtype=np.float32) * np.float32(gain / np.sqrt(fan)))
        w[f"{name}.bias"] = (0.1 * g.standard_normal((cout,), dtype=np.float32)).astype(np.float32)

    C0 = c["upsample_initial_channel"]
    conv("conv_pre", C0, c["num_mels"], 7)
    for i, (u, k) in enumerate(zip(c["upsample_rates"], c["upsample_kernel_sizes"])):
        conv(f"ups.{i}", C0 >> (i + 1), C0 >> i, k, transposed=True, fan=(C0 >> i) * k // u)
    nk = len(c["resblock_kernel_sizes"])
    for i in range(len(c["upsample_rates"])):
        ch = C0 >> (i + 1)
        for j, k in enumerate(c["resblock_kernel_sizes"]):
            for m in range(3):
                conv(f"resblocks.{i * nk + j}.convs1.{m}", ch, ch, k)
            for m in range(3):
                conv(f"resblocks.{i * nk + j}.convs2.{m}", ch, ch, k)
    conv("conv_post", 1, C0 >> len(c["upsample_rates"]), 7)
    w["conv_post.weight"] *= np.float32(0.3)       # keep the final tanh out of saturation (it would hide errors)
    return w


def make_mel(seed: int, B: int, T: int, n_mel: int = 80) -> np.ndarray:
    """A log-mel-like batch ``[B, T, n_mel]`` in the range the vocoder sees."""
    g = _rng(seed)
    m = g.standard_normal((B, T, n_mel), dtype=np.float32) * np.float32(2.0) - np.float32(5.0)
    return np.clip(m, -11.5129, 2.0).astype(np.float32)


def weights_digest(w: dict) -> str:
    h = hashlib.sha256()
    for k in sorted(w):
        h.update(k.encode())
        h.update(np.ascontiguousarray(w[k]).tobytes())
    return h.hexdigest()


def make_inputs(seed: int, B: int, T: int, K: int, *, n_mel: int = 80, d_encoder: int = 256,
                multi_speaker: bool = False, shallow: bool = False,
                min_len_frac: float = 0.5, spec_min: float = -11.5129,
                spec_max: float = 2.0) -> dict:
    """Synthetic batch (SURVEY.md §8d draw order).

    Returns float32 arrays: ``cond [B,T,H]``, ``lens [B]`` (int64, ``lens[0]=T``),
    ``pad_mask [B,T]`` (True = padding, the convention arriving at
    ``GaussianDiffusion.forward``, model/mixgantts.py:122,137), ``spk [B,H]`` or
    None, ``coarse_mel [B,T,M]`` and ``start_noise [B,1,M,T]`` (shallow only),
    ``x_T [B,1,M,T]`` and ``noises [K,B,1,M,T]`` (``noises[t]`` is the draw the
    reference makes inside ``q_posterior_sample`` at timestep ``t``).
    """
    g = _rng(seed)
    out = {}
    out["cond"] = g.standard_normal((B, T, d_encoder), dtype=np.float32)
    lo = max(1, int(T * min_len_frac))
    lens = g.integers(lo, T + 1, size=(B,), dtype=np.int64)
    lens[0] = T
    out["lens"] = lens
    out["pad_mask"] = np.arange(T)[None, :] >= lens[:, None]
    out["spk"] = g.standard_normal((B, d_encoder), dtype=np.float32) if multi_speaker else None
    if shallow:
        cm = g.standard_normal((B, T, n_mel), dtype=np.float32) * np.float32(2.0) - np.float32(5.0)
        out["coarse_mel"] = np.clip(cm, spec_min, spec_max).astype(np.float32)
        out["start_noise"] = g.standard_normal((B, 1, n_mel, T), dtype=np.float32)
    else:
        out["coarse_mel"] = None
        out["start_noise"] = None
    out["x_T"] = g.standard_normal((B, 1, n_mel, T), dtype=np.float32)
    out["noises"] = g.standard_normal((K, B, 1, n_mel, T), dtype=np.float32)
    return out


def make_train_extras(seed: int, B: int, T: int, K: int, *, n_mel: int = 80, spec_min: float = -11.5129,
                      spec_max: float = 2.0) -> dict:
    """Extra draws of the training branch of ``GaussianDiffusion.forward`` (model/diffusion.py:201-225): the target
    ``mel [B,T,M]``, the timesteps ``t [B]`` (mixed, always containing 0 when B > 1 so that the ``t - 1 = -1`` case is hit),
    and the three noise tensors ``noise_t``, ``noise_prev``, ``post_noise`` ``[B,1,M,T]``."""
    g = _rng(seed)
    mel = g.standard_normal((B, T, n_mel), dtype=np.float32) * np.float32(2.0) - np.float32(5.0)
    t = np.array([(K - 1 - b) % K for b in range(B)], dtype=np.int64)   # mixed timesteps, deterministic
    if B > 1:
        t[-1] = 0
    return {"mel": np.clip(mel, spec_min, spec_max).astype(np.float32), "t": t,
            "noise_t": g.standard_normal((B, 1, n_mel, T), dtype=np.float32),
            "noise_prev": g.standard_normal((B, 1, n_mel, T), dtype=np.float32),
            "post_noise": g.standard_normal((B, 1, n_mel, T), dtype=np.float32)}


def grad_probe(seed: int, B: int, T: int, *, n_mel: int = 80) -> dict:
    """Fixed random weights ``r0, r1 [B,T,M]`` of the linear probe loss used by the gradient parity tests:
    ``loss = <x_0_pred, r0> + <x_t_prev_pred, r1>`` over the training branch's outputs."""
    g = _rng(seed)
    return {"r0": g.standard_normal((B, T, n_mel), dtype=np.float32),
            "r1": g.standard_normal((B, T, n_mel), dtype=np.float32)}


def grad_sample_index(n: int, k: int = 257) -> np.ndarray:
    """Indices of the strided sample of a flattened gradient that the gradient goldens store."""
    return np.unique(np.linspace(0, n - 1, min(n, k)).astype(np.int64))
