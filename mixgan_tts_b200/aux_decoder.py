"""Drop-in aux decoder of the ``aux`` / ``shallow`` models: the three sub-modules
``MixGANTTS`` owns for it — ``decoder`` (transformer/Models.py:103-171), ``mel_linear`` (model/mixgantts.py:59-62) and
``postnet`` (transformer/Layers.py:67-137) — behind one module whose ``forward`` is model/mixgantts.py:139-143:

    coarse_mels = self.decoder(output, mel_masks)
    coarse_mels = self.mel_linear(coarse_mels)
    coarse_mels = self.postnet(coarse_mels) + coarse_mels

Parameter names, shapes and ``state_dict`` keys are the reference's (``decoder.position_enc``,
``decoder.layer_stack.{i}.slf_attn.{w_qs,w_ks,w_vs,fc,layer_norm}.*``, ``decoder.layer_stack.{i}.pos_ffn.{w_1,w_2,layer_norm}.*``,
``mel_linear.*``, ``postnet.convolutions.{i}.0.conv.*``, ``postnet.convolutions.{i}.1.*``), so the matching slice of a
``MixGANTTS`` checkpoint loads with ``strict=True``.  The torch sub-modules only HOLD the parameters: the computation is
``mgb_auxdec_forward`` in the sm_100a library (tcgen05 GEMMs, attention, fused LayerNorm / BatchNorm / masks).  Inference only
(the library has no backward for this stage; dropout is the identity and BatchNorm uses its running statistics, i.e. the
reference in ``eval()`` mode).  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
from torch import nn

from . import _lib
from .synth import sinusoid_table


class AuxDims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_mel", "d_model", "n_head", "d_inner", "ffn_kernel", "layers", "postnet_dim",
                                         "postnet_kernel", "postnet_layers")]


class _Attn(nn.Module):
    def __init__(self, d, n_head):
        super().__init__()
        self.w_qs, self.w_ks, self.w_vs = nn.Linear(d, d), nn.Linear(d, d), nn.Linear(d, d)
        self.layer_norm = nn.LayerNorm(d)
        self.fc = nn.Linear(d, d)


class _Ffn(nn.Module):
    def __init__(self, d, h, k):
        super().__init__()
        self.w_1 = nn.Conv1d(d, h, kernel_size=k[0], padding=(k[0] - 1) // 2)
        self.w_2 = nn.Conv1d(h, d, kernel_size=k[1], padding=(k[1] - 1) // 2)
        self.layer_norm = nn.LayerNorm(d)


class _FFTBlock(nn.Module):
    def __init__(self, d, n_head, h, k):
        super().__init__()
        self.slf_attn = _Attn(d, n_head)
        self.pos_ffn = _Ffn(d, h, k)


class _Decoder(nn.Module):
    def __init__(self, d, n_head, h, k, layers, max_seq_len):
        super().__init__()
        self.max_seq_len, self.d_model = max_seq_len, d
        self.position_enc = nn.Parameter(torch.from_numpy(sinusoid_table(max_seq_len + 1, d)).unsqueeze(0), requires_grad=False)
        self.layer_stack = nn.ModuleList([_FFTBlock(d, n_head, h, k) for _ in range(layers)])


class _ConvHolder(nn.Module):
    def __init__(self, cin, cout, k):
        super().__init__()
        self.conv = nn.Conv1d(cin, cout, kernel_size=k, padding=(k - 1) // 2)


class _PostNet(nn.Module):
    def __init__(self, n_mel=80, dim=512, k=5, n=5):
        super().__init__()
        chans = [n_mel] + [dim] * (n - 1) + [n_mel]
        self.convolutions = nn.ModuleList(
            [nn.Sequential(_ConvHolder(chans[i], chans[i + 1], k), nn.BatchNorm1d(chans[i + 1])) for i in range(n)])


class AuxDecoder(nn.Module):
    """``AuxDecoder(preprocess_config, model_config)``; ``forward(output, mel_masks) -> coarse_mels [B, T, n_mel]`` with
    ``mel_masks`` True = padding (the convention at model/mixgantts.py:137-139; a prefix mask ``arange(T) >= len``)."""

    def __init__(self, preprocess_config, model_config):
        super().__init__()
        tr = model_config["transformer"]
        d = tr["decoder_hidden"]
        k = tr["conv_kernel_size"]
        k = list(k) if isinstance(k, (list, tuple)) else [k, 1]
        n_mel = preprocess_config["preprocessing"]["mel"]["n_mel_channels"]
        if k[1] != 1:
            raise ValueError("the second FFN convolution must be k = 1 (config/*/model.yaml conv_kernel_size)")
        self.decoder = _Decoder(d, tr["decoder_head"], tr["conv_filter_size"], k, tr["decoder_layer"], model_config["max_seq_len"])
        self.mel_linear = nn.Linear(d, n_mel)
        self.postnet = _PostNet(n_mel)
        self.dims = AuxDims(n_mel, d, tr["decoder_head"], tr["conv_filter_size"], k[0], tr["decoder_layer"], 512, 5, 5)
        self._packed = {}       # device -> (fingerprint, packed tensor)
        self._ws = {}           # (device, B, T) -> workspace tensor
        self._pos = {}

    # ---- parameters in the library's flat order (include/mixgan_b200.h) ----
    def _flat_list(self):
        out = []
        for blk in self.decoder.layer_stack:
            a, f = blk.slf_attn, blk.pos_ffn
            out += [a.w_qs.weight, a.w_qs.bias, a.w_ks.weight, a.w_ks.bias, a.w_vs.weight, a.w_vs.bias, a.layer_norm.weight,
                    a.layer_norm.bias, a.fc.weight, a.fc.bias, f.w_1.weight, f.w_1.bias, f.w_2.weight, f.w_2.bias,
                    f.layer_norm.weight, f.layer_norm.bias]
        out += [self.mel_linear.weight, self.mel_linear.bias]
        for seq in self.postnet.convolutions:
            cv, bn = seq[0].conv, seq[1]
            out += [cv.weight, cv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var]
        return out

    def invalidate_packed(self):
        """Call after changing parameters through ``.data`` (in-place updates that do not bump ``_version``)."""
        self._packed.clear()

    def _load_from_state_dict(self, *a, **k):
        self._packed.clear()
        return super()._load_from_state_dict(*a, **k)

    def packed_weights(self, dev):
        lib = _lib.load()
        plist = self._flat_list()
        fp = tuple((p.data_ptr(), p._version) for p in plist)
        hit = self._packed.get(dev)
        if hit is not None and hit[0] == fp:
            return hit[1]
        flat = torch.cat([p.detach().reshape(-1).float() for p in plist]).to(dev).contiguous()
        n = lib.mgb_auxdec_flat_count(C.byref(self.dims))
        if flat.numel() != n:
            raise RuntimeError(f"aux decoder parameter count {flat.numel()} != library's {n}")
        packed = torch.empty(lib.mgb_auxdec_packed_bytes(C.byref(self.dims)), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.mgb_auxdec_pack(C.byref(self.dims), _lib.ptr(flat), _lib.ptr(packed), packed.numel(),
                                           C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "mgb_auxdec_pack")
        self._packed[dev] = (fp, packed)
        return packed

    def _position_rows(self, T, dev):
        key = (dev, T)
        if key not in self._pos:
            if T <= self.decoder.max_seq_len:      # Models.py:155-160
                pos = self.decoder.position_enc[0, :T].detach().float().to(dev).contiguous()
            else:                                  # Models.py:146-153 (eval, longer than max_seq_len)
                pos = torch.from_numpy(sinusoid_table(T, self.decoder.d_model)).to(dev)
            if len(self._pos) > 16:
                self._pos.clear()
            self._pos[key] = pos
        return self._pos[key]

    def forward(self, output, mel_masks=None, lens=None, return_intermediate=False):
        if self.training:
            raise RuntimeError("mixgan_tts_b200.AuxDecoder is inference-only: call .eval() (the library has no backward for it)")
        if output.device.type != "cuda":
            raise RuntimeError("mixgan_tts_b200.AuxDecoder needs CUDA tensors (no CPU fallback)")
        lib = _lib.load()
        dev = output.device
        x = output.detach().float().contiguous()
        B, T, D = x.shape
        if D != self.dims.d_model:
            raise ValueError(f"expected [B, T, {self.dims.d_model}] decoder input, got {tuple(x.shape)}")
        if lens is None and mel_masks is not None:
            lens = (~mel_masks.bool()).sum(dim=1)
        lens32 = None if lens is None else lens.to(device=dev, dtype=torch.int32).contiguous()
        with torch.cuda.device(dev):
            packed = self.packed_weights(dev)
            pos = self._position_rows(T, dev)
            key = (dev, B, T)
            ws = self._ws.get(key)
            if ws is None:
                if len(self._ws) > 4:
                    self._ws.clear()
                ws = torch.empty(lib.mgb_auxdec_workspace_bytes(C.byref(self.dims), B, T), dtype=torch.uint8, device=dev)
                self._ws[key] = ws
            coarse = torch.empty((B, T, self.dims.n_mel), dtype=torch.float32, device=dev)
            dec = torch.empty((B, T, D), dtype=torch.float32, device=dev) if return_intermediate else None
            mel0 = torch.empty_like(coarse) if return_intermediate else None
            _lib.check(lib.mgb_auxdec_forward(C.byref(self.dims), _lib.ptr(packed), _lib.ptr(x), _lib.ptr(pos), _lib.ptr(lens32),
                                              _lib.ptr(coarse), _lib.ptr(dec), _lib.ptr(mel0), B, T, _lib.ptr(ws), ws.numel(),
                                              C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "mgb_auxdec_forward")
        return (coarse, dec, mel0) if return_intermediate else coarse
