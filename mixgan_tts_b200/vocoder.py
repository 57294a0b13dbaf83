"""Drop-in HiFi-GAN ``Generator`` (reference hifigan/models.py:112-173, loaded and called by
utils/model.py:76-121).

Same constructor argument (the ``AttrDict`` / mapping of hifigan/config.json), the same module tree and — for a checkpoint
saved with weight norm applied — the same ``state_dict`` keys: ``conv_pre``, ``ups.{i}``, ``resblocks.{r}.convs1.{m}``,
``resblocks.{r}.convs2.{m}``, ``conv_post``, each with ``weight_g`` / ``weight_v`` / ``bias`` (folded to a plain ``weight``
while loading, which is what the reference's ``remove_weight_norm()`` does before inference) or with plain ``weight`` /
``bias``.  ``forward(x)`` takes the reference's ``[B, n_mel, T]`` mel and returns ``[B, 1, T * hop]``;
``forward_frames(mel)`` takes the ``[B, T, n_mel]`` tensor ``GaussianDiffusion.forward`` returns, without the transpose.
The torch sub-modules only hold parameters; the computation is ``mgb_hifigan_forward`` in the sm_100a library (tcgen05
implicit-GEMM convolutions).  Inference only, no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from . import _lib


class VocDims(C.Structure):
    _fields_ = [("n_mel", C.c_int32), ("initial_channel", C.c_int32), ("n_up", C.c_int32), ("up_rates", C.c_int32 * 8),
                ("up_kernels", C.c_int32 * 8), ("n_res", C.c_int32), ("res_kernels", C.c_int32 * 4),
                ("res_dilations", (C.c_int32 * 3) * 4), ("split_mode", C.c_int32)]


PRECISIONS = {"fp16": 0, "mixed": 1, "fp16x3": 2}


def _get(h, key, default=None):
    return h[key] if isinstance(h, dict) and key in h else getattr(h, key, default)


def get_padding(kernel_size, dilation=1):
    return int((kernel_size * dilation - dilation) / 2)


class ResBlock(nn.Module):
    def __init__(self, h, channels, kernel_size=3, dilation=(1, 3, 5)):
        super().__init__()
        self.convs1 = nn.ModuleList([nn.Conv1d(channels, channels, kernel_size, 1, dilation=d, padding=get_padding(kernel_size, d))
                                     for d in dilation])
        self.convs2 = nn.ModuleList([nn.Conv1d(channels, channels, kernel_size, 1, dilation=1, padding=get_padding(kernel_size, 1))
                                     for _ in dilation])


class Generator(nn.Module):
    def __init__(self, h, precision: str = "mixed"):
        """``precision``: ``"fp16"`` = fp16 operands everywhere; ``"mixed"`` (default) = fp16 hi + lo operand pairs in
        ``conv_pre``, the transposed convolutions and ``conv_post`` (a few percent of the time, ~40 % less error);
        ``"fp16x3"`` = hi + lo pairs in every layer, the parity mode (~1e-5 against the fp32 reference, ~2.5x the time)."""
        super().__init__()
        self.h = h
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(PRECISIONS)}")
        self.precision = precision
        if str(_get(h, "resblock", "1")) != "1":
            raise ValueError('only the "resblock": "1" generator (hifigan/config.json) is built')
        rates, kernels = list(_get(h, "upsample_rates")), list(_get(h, "upsample_kernel_sizes"))
        rks, rds = list(_get(h, "resblock_kernel_sizes")), [list(d) for d in _get(h, "resblock_dilation_sizes")]
        C0, n_mel = int(_get(h, "upsample_initial_channel")), int(_get(h, "num_mels", 80))
        self.num_kernels, self.num_upsamples = len(rks), len(rates)
        self.conv_pre = nn.Conv1d(n_mel, C0, 7, 1, padding=3)
        self.ups = nn.ModuleList([nn.ConvTranspose1d(C0 // (2 ** i), C0 // (2 ** (i + 1)), k, u, padding=(k - u) // 2)
                                  for i, (u, k) in enumerate(zip(rates, kernels))])
        self.resblocks = nn.ModuleList()
        ch = C0
        for i in range(len(self.ups)):
            ch = C0 // (2 ** (i + 1))
            for k, d in zip(rks, rds):
                if len(d) != 3:
                    raise ValueError("resblock dilation lists must have three entries")
                self.resblocks.append(ResBlock(h, ch, k, d))
        self.conv_post = nn.Conv1d(ch, 1, 7, 1, padding=3)
        d = VocDims()
        d.n_mel, d.initial_channel, d.n_up, d.n_res = n_mel, C0, len(rates), len(rks)
        d.split_mode = PRECISIONS[precision]
        for i, (u, k) in enumerate(zip(rates, kernels)):
            d.up_rates[i], d.up_kernels[i] = u, k
        for j, (k, dl) in enumerate(zip(rks, rds)):
            d.res_kernels[j] = k
            for m in range(3):
                d.res_dilations[j][m] = dl[m]
        self.dims = d
        self.hop = 1
        for u in rates:
            self.hop *= u
        self._packed, self._ws = {}, {}

    def remove_weight_norm(self):
        """The reference strips its weight-norm hooks here (hifigan/models.py:168-175); this module never carries them
        (``weight_g`` / ``weight_v`` pairs are folded while loading), so there is nothing to do."""
        return self

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        self._packed.clear()
        return super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)

    def load_state_dict(self, state_dict, strict=True, **kw):
        sd = dict(state_dict)
        for key in [k for k in sd if k.endswith(".weight_g")]:       # w = g * v / ||v||, norm over all dims but 0
            base = key[: -len(".weight_g")]
            g, v = sd.pop(key), sd.pop(base + ".weight_v")
            norm = v.reshape(v.shape[0], -1).norm(dim=1).reshape([-1] + [1] * (v.dim() - 1))
            sd[base + ".weight"] = v * (g / norm)
        self._packed.clear()
        return super().load_state_dict(sd, strict=strict, **kw)

    def _flat_list(self):
        out = [self.conv_pre.weight, self.conv_pre.bias]
        for u in self.ups:
            out += [u.weight, u.bias]
        for rb in self.resblocks:
            for c in list(rb.convs1) + list(rb.convs2):
                out += [c.weight, c.bias]
        return out + [self.conv_post.weight, self.conv_post.bias]

    def invalidate_packed(self):
        self._packed.clear()

    def packed_weights(self, dev):
        lib = _lib.load()
        plist = self._flat_list()
        fp = tuple((p.data_ptr(), p._version) for p in plist)
        hit = self._packed.get(dev)
        if hit is not None and hit[0] == fp:
            return hit[1]
        flat = torch.cat([p.detach().reshape(-1).float() for p in plist]).to(dev).contiguous()
        n = lib.mgb_hifigan_flat_count(C.byref(self.dims))
        if n == 0:
            raise ValueError("this HiFi-GAN configuration is not supported by the library (mgb_hifigan_pack names the reason)")
        if flat.numel() != n:
            raise RuntimeError(f"generator parameter count {flat.numel()} != library's {n}")
        packed = torch.empty(lib.mgb_hifigan_packed_bytes(C.byref(self.dims)), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.mgb_hifigan_pack(C.byref(self.dims), _lib.ptr(flat), _lib.ptr(packed), packed.numel(),
                                            C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "mgb_hifigan_pack")
        self._packed[dev] = (fp, packed)
        return packed

    def forward_frames(self, mel):
        """``mel [B, T, n_mel]`` (frames-major) -> ``wav [B, T * hop]``."""
        if mel.device.type != "cuda":
            raise RuntimeError("mixgan_tts_b200.Generator needs CUDA tensors (no CPU fallback)")
        lib = _lib.load()
        dev = mel.device
        x = mel.detach().float().contiguous()
        B, T, M = x.shape
        if M != self.dims.n_mel:
            raise ValueError(f"expected {self.dims.n_mel} mel bins, got {M}")
        with torch.cuda.device(dev):
            packed = self.packed_weights(dev)
            key = (dev, B, T)
            ws = self._ws.get(key)
            if ws is None:
                self._ws.clear()                       # the workspace is large (about 50 KB per mel frame): keep one
                nbytes = lib.mgb_hifigan_workspace_bytes(C.byref(self.dims), B, T)
                if nbytes == 0:
                    raise ValueError("unsupported shape or HiFi-GAN configuration")
                ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                self._ws[key] = ws
            wav = torch.empty((B, T * self.hop), dtype=torch.float32, device=dev)
            _lib.check(lib.mgb_hifigan_forward(C.byref(self.dims), _lib.ptr(packed), _lib.ptr(x), _lib.ptr(wav), B, T,
                                               _lib.ptr(ws), ws.numel(), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
                       "mgb_hifigan_forward")
        return wav

    def forward(self, x):
        """The reference's signature: ``x [B, n_mel, T]`` -> ``[B, 1, T * hop]`` (hifigan/models.py:151-166)."""
        return self.forward_frames(x.transpose(1, 2)).unsqueeze(1)
