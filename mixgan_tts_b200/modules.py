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
import weakref

import torch
from torch import nn

from . import _lib

# "fp16": the reference-precision (fp32, 1e-3) mode on the tensor cores - fp16 operands (TF32's 11-bit significand),
# fp32 accumulation and streams; inference entry points only (training uses "bf16" or "fp32").
PRECISIONS = {"fp32": _lib.PREC_FP32, "bf16": _lib.PREC_BF16, "fp16": _lib.PREC_FP16}
_PACK_KINDS = dict(PRECISIONS, fp32_tables=_lib.PACK_FP32_TABLES)

# Fused optimizers (torch.optim.Adam(fused=True) ...) update parameters in place WITHOUT bumping Tensor._version, so the
# (data_ptr, _version) fingerprint alone would keep serving stale kernel-layout weights after an optimizer step.  Every
# optimizer step therefore advances this epoch, which is part of the fingerprint.
_PARAM_EPOCH = [0]
_DENOISER_PARAM_IDS = set()          # id() of every parameter a live Denoiser owns (entries are removed by a finalizer)
_OPT_TOUCHES_DENOISER = weakref.WeakKeyDictionary()   # optimizer -> does it own a Denoiser parameter? (decided once)


def _on_optimizer_step(optimizer, *_args, **_kwargs):
    try:
        hit = _OPT_TOUCHES_DENOISER.get(optimizer)
        if hit is None:
            hit = any(id(p) in _DENOISER_PARAM_IDS for g in optimizer.param_groups for p in g["params"])
            _OPT_TOUCHES_DENOISER[optimizer] = hit
    except TypeError:                               # an optimizer that cannot be weakly referenced: be conservative
        hit = True
    if hit:                                         # e.g. the discriminator's optimizer does not force a repack
        _PARAM_EPOCH[0] += 1


try:
    from torch.optim.optimizer import register_optimizer_step_post_hook as _register_post_hook
    _register_post_hook(_on_optimizer_step)
except Exception:                                   # pragma: no cover - very old torch: fall back to never caching
    _PARAM_EPOCH = None


def _param_fingerprint(params):
    if _PARAM_EPOCH is None:
        return object()                             # never equal: rebuild every time
    return (_PARAM_EPOCH[0],) + tuple((p.data_ptr(), p._version) for p in params)


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
    """Caller-owned scratch for the library, grown on demand and reused across calls; one buffer PER DEVICE (replicas of
    one module on several GPUs - nn.DataParallel - share this object through ``__dict__``)."""

    def __init__(self):
        self.bufs = {}

    def get(self, nbytes: int, device) -> torch.Tensor:
        buf = self.bufs.get(device)
        if buf is None or buf.numel() < nbytes:
            buf = self.bufs[device] = torch.empty(nbytes, dtype=torch.uint8, device=device)
        return buf

    @property
    def buf(self):
        """The buffer of the current CUDA device (diagnostics)."""
        return self.bufs.get(torch.device("cuda", torch.cuda.current_device()))


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

        # Kernel-layout caches.  Every cache is keyed by DEVICE: nn.DataParallel replicas (the reference wraps the model
        # in one unconditionally, train.py:43-44) share these dicts through __dict__ and run in parallel threads.
        self._packed = {}          # (precision, device) -> (fingerprint, packed tensor)
        self._flat = {}            # device -> (fingerprint, flat fp32 parameter vector) of the last pack
        self._flat_store = {}      # device -> (persistent flat buffer, per-parameter views into it)
        self._pack_epoch = 0       # bumped by invalidate_packed()
        self._ws = _Workspace()
        self._train_ws = _Workspace()
        self.grad_sync = None      # optional mixgan_tts_b200.grad_sync.GradSync (data-parallel training)
        # Opt-in (MIXGAN_B200_TRAIN_GRAPHS=1 or `den.use_cuda_graphs = True`), meant for fixed-shape training: from the
        # third call with one (B, T) signature on, the library's ~220 launches per training step are replayed as CUDA
        # graphs over static buffers (at most _TrainGraph.MAX_SIGNATURES signatures; other shapes run eagerly).  Measured
        # on B200 at B=8 x T=800: 3.76 -> 3.73 ms per step (the step is device-bound); bit-identical to the eager path
        # (tests/test_gpu_train.py).
        self.use_cuda_graphs = os.environ.get("MIXGAN_B200_TRAIN_GRAPHS", "0") == "1"
        self._train_graphs = {}
        ids = [id(p) for p in self.parameters()]
        _DENOISER_PARAM_IDS.update(ids)
        weakref.finalize(self, _DENOISER_PARAM_IDS.difference_update, ids)

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

    def invalidate_packed(self):
        """Forget every kernel-layout copy of the weights.  The caches notice optimizer steps, ``load_state_dict``,
        ``.to()`` / ``.cuda()`` and ordinary in-place ops by themselves; call this after writing parameters through
        ``.data`` (``p.data.copy_()``, EMA swaps ...), which bumps no version counter."""
        self.__dict__["_pack_epoch"] = self.__dict__.get("_pack_epoch", 0) + 1
        self._packed.clear()
        self._flat.clear()

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self.invalidate_packed()

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self.invalidate_packed()
        return out

    def packed_weights(self, precision: str | None = None, params=None) -> torch.Tensor:
        """Kernel-layout copy of the parameters; rebuilt when any parameter changes."""
        return self.packed_and_flat(precision, params)[0]

    def packed_and_flat(self, precision: str | None = None, params=None):
        """``(packed, flat)``: the kernel-layout copy and the flat fp32 parameter vector it was built from (``flat`` is None
        when the packed copy came out of the cache and no flat vector of the same parameters is held)."""
        precision = precision or self.precision
        prec = _PACK_KINDS[precision]
        params = params if params is not None else self._ordered_params()
        fp = (self._pack_epoch, _param_fingerprint(params))
        dev = params[0].device
        hit = self._packed.get((precision, dev))
        if hit is not None and hit[0] == fp:
            fl = self._flat.get(dev)
            return hit[1], (fl[1] if fl is not None and fl[0] == fp else None)
        if dev.type != "cuda":
            raise RuntimeError("mixgan_tts_b200.Denoiser runs on a CUDA device only (no CPU fallback); "
                               "move the module with .cuda() first")
        lib = _lib.load()
        with torch.cuda.device(dev):
            fl = self._flat.get(dev)
            if fl is not None and fl[0] == fp:
                flat = fl[1]
            else:
                # persistent flat buffer + per-parameter views, refreshed with ONE multi-tensor copy (a torch.cat over
                # 162 reshaped views costs ~0.4 ms of host time per training step)
                store = self._flat_store.get(dev)
                if store is None or store[0].numel() != sum(p.numel() for p in params):
                    buf = torch.empty(sum(p.numel() for p in params), dtype=torch.float32, device=dev)
                    views, off = [], 0
                    for p in params:
                        views.append(buf[off:off + p.numel()].view(p.shape))
                        off += p.numel()
                    store = self._flat_store[dev] = (buf, views)
                with torch.no_grad():
                    torch._foreach_copy_(store[1], [p.detach() for p in params])
                flat = store[0]
                self._flat[dev] = (fp, flat)
            assert flat.numel() == lib.mgb_flat_weight_count(C.byref(self.dims))
            nbytes = lib.mgb_packed_bytes(C.byref(self.dims), prec)
            if nbytes == 0:
                raise RuntimeError(f"precision {precision!r} is not available in this build")
            packed = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.mgb_pack_weights(C.byref(self.dims), prec, _lib.ptr(flat), _lib.ptr(packed), nbytes,
                                            C.c_void_p(stream)), "mgb_pack_weights")
        self._packed[(precision, dev)] = (fp, packed)
        return packed, flat

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
        if self.dims.multi_speaker and speaker_emb is None:
            raise TypeError("multi_speaker Denoiser needs speaker_emb")   # reference: F.linear(None) TypeError
        B, _, M, T = mel.shape
        if M != self.dims.n_mel or conditioner.shape != (B, self.dims.d_encoder, T):
            raise ValueError(f"shape mismatch: mel {tuple(mel.shape)}, conditioner {tuple(conditioner.shape)}")
        if torch.is_grad_enabled() and (
                mel.requires_grad or conditioner.requires_grad
                or (speaker_emb is not None and speaker_emb.requires_grad)
                or any(p.requires_grad for p in self.parameters())):
            return self._forward_with_grad(mel, diffusion_step, conditioner, speaker_emb)
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

    # ---------------------------------------------------------------- training (autograd)
    @property
    def train_precision(self) -> str:
        """Arithmetic of the autograd path: "bf16" trains on tcgen05; the two reference-precision modes ("fp32", and
        "fp16" whose tensor-core kernel is inference-only) train in exact fp32."""
        return "bf16" if self.precision == "bf16" else "fp32"

    def flat_weights(self, params=None) -> torch.Tensor:
        """The parameters as one fp32 vector in the canonical order (built together with a weight pack)."""
        kind = "fp32" if self.train_precision == "fp32" else "fp32_tables"
        packed, flat = self.packed_and_flat(kind, params)
        if flat is None:                       # packed copy cached but the flat vector was dropped: rebuild both
            self._packed.pop((kind, (params or self._ordered_params())[0].device), None)
            packed, flat = self.packed_and_flat(kind, params)
        return flat

    def train_workspace(self, B: int, T: int, device) -> torch.Tensor:
        lib = _lib.load()
        n = lib.mgb_train_workspace_bytes(C.byref(self.dims), PRECISIONS[self.train_precision], B, T)
        if n == 0:
            raise ValueError(f"unsupported shape B={B} T={T}")
        return self._train_ws.get(n, device)

    def _forward_with_grad(self, mel, diffusion_step, conditioner, speaker_emb):
        """Forward that records a graph node whose backward runs in the library, in ``self.precision`` arithmetic
        (``fp32``: CUDA-core GEMMs, the parity mode; ``bf16``: every GEMM on tcgen05 with fp32 accumulation and fp32
        residual/skip/gradient streams).  The transposes/casts around it are ordinary torch ops."""
        cond_bth = conditioner.transpose(1, 2).float().contiguous()
        spk = speaker_emb.float().contiguous() if self.dims.multi_speaker else None
        params = self._ordered_params()
        return _DenoiserGradFn.apply(self, mel.float().contiguous(), diffusion_step, cond_bth, spk, *params)


class _TrainGraph:
    """Static buffers + captured CUDA graphs of the library's training forward / backward for one call signature.

    A graph replays fixed addresses, so inputs are copied into the static buffers and results are cloned out of them;
    the flat parameter vector is refreshed with one multi-tensor copy and the fp32 weight pack is part of the forward
    graph.  One forward may be in flight per signature (the activation stash is static); a second forward before the
    first one's backward simply takes the eager path."""

    WARMUP_CALLS = 2      # eager calls before capturing (first-use initialisation inside the library must not be captured)
    MAX_SIGNATURES = 4    # every (B, T) signature owns ~0.1 MB per frame of static buffers: further shapes run eagerly

    def __init__(self, den, B, T, prec, device):
        lib, dims = _lib.load(), den.dims
        M, H = dims.n_mel, dims.d_encoder
        f32 = dict(dtype=torch.float32, device=device)
        u8 = dict(dtype=torch.uint8, device=device)
        self.calls = 0
        self.x = torch.empty((B, 1, M, T), **f32)
        self.t = torch.empty((B,), dtype=torch.int64, device=device)
        self.cond = torch.empty((B, T, H), **f32)
        self.spk = torch.empty((B, H), **f32) if dims.multi_speaker else None
        self.out = torch.empty((B, 1, M, T), **f32)
        self.saved = torch.empty(lib.mgb_train_saved_bytes(C.byref(dims), prec, B, T), **u8)
        # its own workspace: the module's shared one is reallocated when a larger shape comes along, and a captured graph
        # must keep replaying the addresses it was captured with
        self.ws = torch.empty(lib.mgb_train_workspace_bytes(C.byref(dims), prec, B, T), **u8)
        nflat = lib.mgb_flat_weight_count(C.byref(dims))
        self.flat = torch.empty(nflat, **f32)
        self.flat_views, off = [], 0
        for p in den._ordered_params():
            self.flat_views.append(self.flat[off:off + p.numel()].view(p.shape))
            off += p.numel()
        self.flat_fp = None
        self.packed = torch.empty(lib.mgb_packed_bytes(C.byref(dims), _lib.PREC_FP32), **u8)
        self.gout = torch.empty((B, 1, M, T), **f32)
        self.gflat = torch.empty(nflat, **f32)
        self.gx = torch.empty((B, 1, M, T), **f32)
        self.gcond = torch.empty((B, T, H), **f32)
        self.gspk = torch.empty((B, H), **f32) if dims.multi_speaker else None
        self.fwd = None                 # (graph, launches)
        self.bwd = {}                   # (bucket plan, need flags) -> [(graph, launches, flat_begin, flat_end)]
        self.token_ref = None

    def busy(self) -> bool:
        return self.token_ref is not None and self.token_ref() is not None


class _Token:
    """Held by the autograd context of the forward that owns a _TrainGraph's activation stash."""
    __slots__ = ("__weakref__",)


def _capture(lib, fn):
    """Capture the launches `fn(stream)` enqueues into a CUDA graph; returns (graph, number of library launches)."""
    g = torch.cuda.CUDAGraph()
    n0 = lib.mgb_launch_count()
    with torch.cuda.graph(g):
        fn(C.c_void_p(torch.cuda.current_stream().cuda_stream))
    return g, lib.mgb_launch_count() - n0


def _grad_views(den, gflat, need):
    """Per-parameter gradient views of the flat gradient (one split call, then a reshape per parameter)."""
    meta = den.__dict__.get("_param_meta")
    if meta is None:
        ps = den._ordered_params()
        meta = den.__dict__["_param_meta"] = ([p.numel() for p in ps], [p.shape for p in ps])
    parts = gflat.split(meta[0])
    return [g.view(shp) if nd else None for g, shp, nd in zip(parts, meta[1], need)]


class _DenoiserGradFn(torch.autograd.Function):
    """autograd node for ``Denoiser.forward``: ``mgb_denoiser_train_forward`` / ``mgb_denoiser_backward``.

    The backward runs as gradient buckets (groups of backward segments); when the module has a ``grad_sync``
    the all-reduce of a finished bucket is started while the next bucket is computed."""

    @staticmethod
    def forward(ctx, den, x, t, cond_bth, spk, *params):
        lib = _lib.load()
        dev = x.device
        B, _, M, T = x.shape
        prec = PRECISIONS[den.train_precision]
        tg = None
        if den.use_cuda_graphs:
            key = (B, T, prec, dev)
            tg = den._train_graphs.get(key)
            if tg is None and len(den._train_graphs) < _TrainGraph.MAX_SIGNATURES:
                tg = den._train_graphs[key] = _TrainGraph(den, B, T, prec, dev)
            if tg is not None:
                tg.calls += 1
                if tg.calls <= _TrainGraph.WARMUP_CALLS or tg.busy():
                    tg = None
        if tg is not None:
            with torch.cuda.device(dev):
                fp = _param_fingerprint(params)
                if tg.flat_fp != fp:
                    torch._foreach_copy_(tg.flat_views, [p.detach() for p in params])
                    tg.flat_fp = fp
                tg.x.copy_(x)
                tg.t.copy_(t.detach())
                tg.cond.copy_(cond_bth)
                if tg.spk is not None:
                    tg.spk.copy_(spk)
                if tg.fwd is None:
                    ws = tg.ws

                    def enqueue(stream):
                        kind = _lib.PREC_FP32 if prec == _lib.PREC_FP32 else _lib.PACK_FP32_TABLES
                        _lib.check(lib.mgb_pack_weights(C.byref(den.dims), kind, _lib.ptr(tg.flat), _lib.ptr(tg.packed),
                                                        tg.packed.numel(), stream), "mgb_pack_weights")
                        _lib.check(lib.mgb_denoiser_train_forward(
                            C.byref(den.dims), prec, _lib.ptr(tg.packed), _lib.ptr(tg.flat), _lib.ptr(tg.x), _lib.ptr(tg.t),
                            _lib.ptr(tg.cond), _lib.ptr(tg.spk), _lib.ptr(tg.out), _lib.ptr(tg.saved), tg.saved.numel(), B, T,
                            _lib.ptr(ws), ws.numel(), stream), "mgb_denoiser_train_forward")
                    tg.fwd = _capture(lib, enqueue)
                tg.fwd[0].replay()
                lib.mgb_note_launches(tg.fwd[1])
                out = tg.out.clone()
            token = _Token()
            tg.token_ref = weakref.ref(token)
            ctx.den, ctx.tg, ctx.token, ctx.prec = den, tg, token, prec
            ctx.shape = (B, M, T)
            return out
        ctx.tg = None
        with torch.cuda.device(dev):
            # both precisions read fp32 per-utterance tables; the bf16 mode needs nothing else from the fp32 pack
            # The eager path hands the backward its OWN copy of the flat parameter vector: the module's persistent flat
            # buffer is rewritten in place by the next pack, and two forwards with different parameters before the first
            # backward (a D step and a G step sharing the Denoiser) would otherwise differentiate against the wrong weights.
            kind = "fp32" if den.train_precision == "fp32" else "fp32_tables"
            packed, flat = den.packed_and_flat(kind, params)
            if flat is None:
                den._packed.pop((kind, dev), None)
                packed, flat = den.packed_and_flat(kind, params)
            flat = flat.clone()
            tt = t.detach().to(torch.int64).contiguous()
            out = torch.empty_like(x)
            saved = torch.empty(lib.mgb_train_saved_bytes(C.byref(den.dims), prec, B, T), dtype=torch.uint8, device=dev)
            ws = den.train_workspace(B, T, dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.mgb_denoiser_train_forward(
                C.byref(den.dims), prec, _lib.ptr(packed), _lib.ptr(flat), _lib.ptr(x), _lib.ptr(tt), _lib.ptr(cond_bth),
                _lib.ptr(spk), _lib.ptr(out), _lib.ptr(saved), saved.numel(), B, T, _lib.ptr(ws), ws.numel(),
                C.c_void_p(stream)), "mgb_denoiser_train_forward")
        ctx.den, ctx.saved, ctx.tt, ctx.cond, ctx.spk = den, saved, tt, cond_bth, spk
        ctx.flat, ctx.prec = flat, prec
        ctx.shape = (B, M, T)
        return out

    @staticmethod
    def _segment_ranges(den):
        lib = _lib.load()
        ranges = []
        for s in range(lib.mgb_train_segments(C.byref(den.dims))):
            b, e = C.c_size_t(0), C.c_size_t(0)
            _lib.check(lib.mgb_train_segment_range(C.byref(den.dims), s, C.byref(b), C.byref(e)), "segment range")
            ranges.append((b.value, e.value))
        return ranges

    @staticmethod
    def _backward_graphed(ctx, gout):
        from .grad_sync import plan_buckets
        den, lib, tg = ctx.den, _lib.load(), ctx.tg
        if ctx.token is None:
            raise RuntimeError("mixgan_tts_b200.Denoiser: the activation stash is released by the first backward "
                               "(a second backward through the same forward / retain_graph=True is not supported)")
        B, M, T = ctx.shape
        dev = gout.device
        need = ctx.needs_input_grad
        with torch.cuda.device(dev):
            sync = den.grad_sync
            buckets = tuple(plan_buckets(_DenoiserGradFn._segment_ranges(den), sync.bucket_bytes if sync is not None else None))
            want = (bool(need[1]), bool(need[3]), bool(tg.spk is not None and need[4]))
            tg.gout.copy_(gout)
            graphs = tg.bwd.get((buckets, want))
            if graphs is None:
                ws = tg.ws
                graphs = []
                for sb, se, fb, fe in buckets:
                    def enqueue(stream, sb=sb, se=se):
                        _lib.check(lib.mgb_denoiser_backward(
                            C.byref(den.dims), ctx.prec, _lib.ptr(tg.flat), _lib.ptr(tg.saved), tg.saved.numel(), _lib.ptr(tg.t),
                            _lib.ptr(tg.cond), _lib.ptr(tg.spk), _lib.ptr(tg.gout), _lib.ptr(tg.gflat),
                            _lib.ptr(tg.gcond if want[1] else None), _lib.ptr(tg.gspk if want[2] else None),
                            _lib.ptr(tg.gx if want[0] else None), B, T, sb, se, _lib.ptr(ws), ws.numel(), stream),
                            "mgb_denoiser_backward")
                    g, n = _capture(lib, enqueue)
                    graphs.append((g, n, fb, fe))
                tg.bwd[(buckets, want)] = graphs
            for g, n, fb, fe in graphs:
                g.replay()
                lib.mgb_note_launches(n)
                if sync is not None:
                    sync.reduce_async(tg.gflat[fb:fe])
            if sync is not None:
                sync.finish()
            gflat = tg.gflat.clone()          # the static buffers are rewritten by the next step
            gx = tg.gx.clone() if want[0] else None
            gcond = tg.gcond.clone() if want[1] else None
            gspk = tg.gspk.clone() if want[2] else None
        grads = _grad_views(den, gflat, need[5:])
        ctx.token = None                      # releases the signature's activation stash
        return (None, gx, None, gcond, gspk, *grads)

    @staticmethod
    def backward(ctx, gout):
        from .grad_sync import plan_buckets
        if ctx.tg is not None:
            return _DenoiserGradFn._backward_graphed(ctx, gout.float().contiguous())
        den, lib = ctx.den, _lib.load()
        if ctx.saved is None:
            raise RuntimeError("mixgan_tts_b200.Denoiser: the activation stash is released by the first backward "
                               "(a second backward through the same forward / retain_graph=True is not supported)")
        B, M, T = ctx.shape
        dev = gout.device
        need = ctx.needs_input_grad          # (den, x, t, cond, spk, *params)
        with torch.cuda.device(dev):
            gout = gout.float().contiguous()
            gflat = torch.empty_like(ctx.flat)
            gx = torch.empty((B, 1, M, T), dtype=torch.float32, device=dev) if need[1] else None
            gcond = torch.empty_like(ctx.cond) if need[3] else None
            gspk = torch.empty_like(ctx.spk) if (ctx.spk is not None and need[4]) else None
            ws = den.train_workspace(B, T, dev)
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            sync = den.grad_sync
            nseg = lib.mgb_train_segments(C.byref(den.dims))
            ranges = []
            for s in range(nseg):
                b, e = C.c_size_t(0), C.c_size_t(0)
                _lib.check(lib.mgb_train_segment_range(C.byref(den.dims), s, C.byref(b), C.byref(e)), "segment range")
                ranges.append((b.value, e.value))
            buckets = plan_buckets(ranges, sync.bucket_bytes if sync is not None else None)
            for sb, se, fb, fe in buckets:
                _lib.check(lib.mgb_denoiser_backward(
                    C.byref(den.dims), ctx.prec, _lib.ptr(ctx.flat), _lib.ptr(ctx.saved), ctx.saved.numel(),
                    _lib.ptr(ctx.tt), _lib.ptr(ctx.cond), _lib.ptr(ctx.spk), _lib.ptr(gout), _lib.ptr(gflat),
                    _lib.ptr(gcond), _lib.ptr(gspk), _lib.ptr(gx), B, T, sb, se, _lib.ptr(ws), ws.numel(), stream),
                    "mgb_denoiser_backward")
                if sync is not None:
                    sync.reduce_async(gflat[fb:fe])
            if sync is not None:
                sync.finish()
        grads = _grad_views(den, gflat, need[5:])
        ctx.saved = None
        return (None, gx, None, gcond, gspk, *grads)
