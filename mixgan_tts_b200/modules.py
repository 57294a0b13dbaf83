"""Drop-in ``Denoiser`` (reference: ``model/modules.py:382-446``) backed by the sm_100a library.

Same constructor arguments, ``forward`` signature, parameter names/shapes and ``state_dict``
keys as the reference, so ``load_state_dict`` of a reference checkpoint works unchanged:

    input_projection.0.conv.{weight,bias}, mlp.{0,2}.linear.weight,
    residual_layers.N.{conv_layer.conv, conditioner_projection.conv, output_projection.conv}.{weight,bias},
    residual_layers.N.{diffusion_projection, speaker_projection}.linear.weight,
    skip_projection.conv.{weight,bias}, output_projection.conv.{weight,bias}

The torch sub-modules below only HOLD parameters (and give reference-identical random init);
they are never called.  ``forward`` hands raw device pointers to ``mgb_denoiser_forward``.
"""
from __future__ import annotations

import ctypes as C
import os

import torch
from torch import nn

from . import _lib

PRECISIONS = {"fp32": _lib.PREC_FP32, "bf16": _lib.PREC_BF16}


def default_precision() -> str:
    return os.environ.get("MIXGAN_B200_PRECISION", "bf16")


class _Conv(nn.Module):
    """Parameter holder with the reference's ``ConvNorm`` key layout (``.conv.weight``)."""

    def __init__(self, cin, cout, k=1):
        super().__init__()
        self.conv = nn.Conv1d(cin, cout, kernel_size=k, padding=(k - 1) // 2)


class _Linear(nn.Module):
    """Parameter holder with the reference's bias-free ``LinearNorm`` key layout."""

    def __init__(self, cin, cout):
        super().__init__()
        self.linear = nn.Linear(cin, cout, bias=False)
        nn.init.xavier_uniform_(self.linear.weight)


class _Block(nn.Module):
    def __init__(self, d_encoder, channels, multi_speaker):
        super().__init__()
        self.conv_layer = _Conv(channels, 2 * channels, 3)
        self.diffusion_projection = _Linear(channels, channels)
        if multi_speaker:
            self.speaker_projection = _Linear(d_encoder, channels)
        self.conditioner_projection = _Conv(d_encoder, channels, 1)
        self.output_projection = _Conv(channels, 2 * channels, 1)


class _Workspace:
    """Caller-owned scratch for the library, grown on demand and reused across calls."""

    def __init__(self):
        self.buf = None

    def get(self, nbytes: int, device) -> torch.Tensor:
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        return self.buf


class Denoiser(nn.Module):
    """Conditional diffusion denoiser; computes on the GPU through the C ABI only."""

    def __init__(self, preprocess_config, model_config, precision: str | None = None):
        super().__init__()
        n_mel = preprocess_config["preprocessing"]["mel"]["n_mel_channels"]
        d_encoder = model_config["transformer"]["encoder_hidden"]
        channels = model_config["denoiser"]["residual_channels"]
        layers = model_config["denoiser"]["residual_layers"]
        multi_speaker = bool(model_config["multi_speaker"])
        self.dims = _lib.ModelDims(n_mel, channels, d_encoder, layers, int(multi_speaker))
        self.precision = precision or default_precision()

        self.input_projection = nn.Sequential(_Conv(n_mel, channels, 1), nn.ReLU())
        self.mlp = nn.Sequential(_Linear(channels, channels * 4), nn.Identity(), _Linear(channels * 4, channels))
        self.residual_layers = nn.ModuleList(_Block(d_encoder, channels, multi_speaker) for _ in range(layers))
        self.skip_projection = _Conv(channels, channels, 1)
        self.output_projection = _Conv(channels, n_mel, 1)
        nn.init.zeros_(self.output_projection.conv.weight)   # as the reference (modules.py:418)

        self._packed = {}          # precision -> (fingerprint, packed tensor)
        self._ws = _Workspace()

    # ---------------------------------------------------------------- weights
    def _ordered_params(self):
        """Canonical flat order of include/mixgan_b200.h."""
        ps = [self.input_projection[0].conv.weight, self.input_projection[0].conv.bias,
              self.mlp[0].linear.weight, self.mlp[2].linear.weight]
        for blk in self.residual_layers:
            ps += [blk.conv_layer.conv.weight, blk.conv_layer.conv.bias, blk.diffusion_projection.linear.weight]
            if self.dims.multi_speaker:
                ps.append(blk.speaker_projection.linear.weight)
            ps += [blk.conditioner_projection.conv.weight, blk.conditioner_projection.conv.bias,
                   blk.output_projection.conv.weight, blk.output_projection.conv.bias]
        ps += [self.skip_projection.conv.weight, self.skip_projection.conv.bias,
               self.output_projection.conv.weight, self.output_projection.conv.bias]
        return ps

    def packed_weights(self, precision: str | None = None) -> torch.Tensor:
        """Kernel-layout copy of the parameters; rebuilt when any parameter changes."""
        precision = precision or self.precision
        prec = PRECISIONS[precision]
        params = self._ordered_params()
        fp = tuple((p.data_ptr(), p._version) for p in params)
        hit = self._packed.get(precision)
        if hit is not None and hit[0] == fp:
            return hit[1]
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("mixgan_tts_b200.Denoiser runs on a CUDA device only (no CPU fallback); "
                               "move the module with .cuda() first")
        lib = _lib.load()
        with torch.cuda.device(dev):
            flat = torch.cat([p.detach().reshape(-1).float() for p in params]).contiguous()
            assert flat.numel() == lib.mgb_flat_weight_count(C.byref(self.dims))
            nbytes = lib.mgb_packed_bytes(C.byref(self.dims), prec)
            if nbytes == 0:
                raise RuntimeError(f"precision {precision!r} is not available in this build")
            packed = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.mgb_pack_weights(C.byref(self.dims), prec, _lib.ptr(flat), _lib.ptr(packed), nbytes,
                                            C.c_void_p(stream)), "mgb_pack_weights")
        self._packed[precision] = (fp, packed)
        return packed

    def workspace(self, B: int, T: int, K: int, device, precision: str | None = None) -> torch.Tensor:
        lib = _lib.load()
        n = lib.mgb_workspace_bytes(C.byref(self.dims), PRECISIONS[precision or self.precision], B, T, K)
        if n == 0:
            raise ValueError(f"unsupported shape B={B} T={T}")
        return self._ws.get(n, device)

    # ---------------------------------------------------------------- forward
    @staticmethod
    def _f32c(t):
        return None if t is None else t.detach().float().contiguous()

    def forward(self, mel, diffusion_step, conditioner, speaker_emb, mask=None):
        """``mel [B,1,M,T]``, ``diffusion_step [B]``, ``conditioner [B,H,T]``, ``speaker_emb [B,H]``
        -> ``[B,1,M,T]``.  ``mask`` is accepted and ignored, as in the reference."""
        if mel.device.type != "cuda":
            raise RuntimeError("mixgan_tts_b200.Denoiser needs CUDA tensors (no CPU fallback)")
        if torch.is_grad_enabled() and self.training and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError("the B200 Denoiser has no backward in this build: call .eval() or "
                                      "run under torch.no_grad() (outputs never carry a graph)")
        if self.dims.multi_speaker and speaker_emb is None:
            raise TypeError("multi_speaker Denoiser needs speaker_emb")   # reference: F.linear(None) TypeError
        B, _, M, T = mel.shape
        if M != self.dims.n_mel or conditioner.shape != (B, self.dims.d_encoder, T):
            raise ValueError(f"shape mismatch: mel {tuple(mel.shape)}, conditioner {tuple(conditioner.shape)}")
        lib = _lib.load()
        dev = mel.device
        with torch.cuda.device(dev):
            x = self._f32c(mel)
            cond = self._f32c(conditioner.transpose(1, 2))          # [B,T,H]; free if it was a transposed view
            spk = self._f32c(speaker_emb) if self.dims.multi_speaker else None
            t = diffusion_step.detach().to(torch.int64).contiguous()
            out = torch.empty_like(x)
            packed = self.packed_weights()
            ws = self.workspace(B, T, 1, dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.mgb_denoiser_forward(
                C.byref(self.dims), PRECISIONS[self.precision], _lib.ptr(packed), _lib.ptr(x), _lib.ptr(t),
                _lib.ptr(cond), _lib.ptr(spk), _lib.ptr(out), B, T, _lib.ptr(ws), ws.numel(),
                C.c_void_p(stream)), "mgb_denoiser_forward")
        return out
