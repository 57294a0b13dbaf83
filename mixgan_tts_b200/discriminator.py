"""Drop-in ``JCUDiscriminator`` (reference: ``model/mixgantts.py:186-288``) whose layers run in the sm_100a library, forward
AND backward (``train.py:126-184`` drives it 4x forward + 2x backward per training step).

Same constructor arguments, parameter names / shapes / ``state_dict`` keys and ``forward`` signature as the reference:

    input_projection.linear.weight, mlp.{0,2}.linear.weight, [spk_mlp.0.linear.weight,]
    conv_block.{0,1,2}.conv.{weight,bias}, uncond_conv_block.{0,1}.conv.{weight,bias}, cond_conv_block.{0,1}.conv.{weight,bias}

Every layer (Linear or strided Conv1d + bias + leaky_relu / Mish, and the "x + diffusion_step (+ speaker)" add fused into
the first conditional convolution's operand gather) is one ``torch.autograd.Function`` over ``mgb_conv1d_forward`` /
``mgb_conv1d_backward`` on frames-major ``[B, T, C]`` fp32 tensors; torch autograd only chains the nodes.  The returned
feature lists hold ``[B, C, T']`` views, as the reference's do.  fp32 accuracy (3 x TF32 tensor-core or exact-fp32
CUDA-core kernels, csrc/conv1d_f32.cu): the reference's gradients are the parity target.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from . import _lib

ACT_NONE, ACT_LEAKY, ACT_MISH, ACT_RELU = 0, 1, 2, 3


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


class _ConvFn(torch.autograd.Function):
    """y = act(conv1d(x (+ rowbias), w, stride, padding=(k-1)/2) + bias) on frames-major x [B,Tin,Cin] -> y [B,Tout,Cout]."""

    @staticmethod
    def forward(ctx, x, w, bias, rowbias, stride, act):
        if x.device.type != "cuda":
            raise RuntimeError("mixgan_tts_b200.JCUDiscriminator needs CUDA tensors (no CPU fallback)")
        lib = _lib.load()
        x = x.float().contiguous()
        wc = w.detach().float().contiguous()
        B, Tin, Cin = x.shape
        Cout, Cin_w, k = wc.shape
        if Cin_w != Cin:
            raise ValueError(f"conv weight expects {Cin_w} input channels, got {Cin}")
        Tout = lib.mgb_conv1d_out_len(Tin, k, stride)
        dev = x.device
        with torch.cuda.device(dev):
            y = torch.empty((B, Tout, Cout), dtype=torch.float32, device=dev)
            pre = torch.empty_like(y) if act == ACT_MISH else None
            ws = torch.empty(lib.mgb_conv1d_workspace_bytes(B, Tin, Cin, Cout, k, stride), dtype=torch.uint8, device=dev)
            b = None if bias is None else bias.detach().float().contiguous()
            rb = None if rowbias is None else rowbias.detach().float().contiguous()
            _lib.check(lib.mgb_conv1d_forward(_lib.ptr(x), _lib.ptr(wc), _lib.ptr(b), _lib.ptr(rb), _lib.ptr(y), _lib.ptr(pre),
                                              B, Tin, Cin, Cout, k, stride, act, _lib.ptr(ws), ws.numel(), _stream(dev)),
                       "mgb_conv1d_forward")
        ctx.save_for_backward(x.detach(), wc, rb, y, pre)
        ctx.cfg = (stride, act, bias is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, wc, rb, y, pre = ctx.saved_tensors
        stride, act, has_bias = ctx.cfg
        lib = _lib.load()
        B, Tin, Cin = x.shape
        Cout, _, k = wc.shape
        dev = x.device
        need_x, need_w, need_b, need_rb = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2], ctx.needs_input_grad[3]
        with torch.cuda.device(dev):
            gy = gy.float().contiguous()
            gx = torch.empty_like(x) if (need_x or need_rb) else None
            gw = torch.empty_like(wc) if need_w else None
            gb = torch.empty((Cout,), dtype=torch.float32, device=dev) if (has_bias and need_b) else None
            grb = torch.empty_like(rb) if (rb is not None and need_rb) else None
            ws = torch.empty(lib.mgb_conv1d_workspace_bytes(B, Tin, Cin, Cout, k, stride), dtype=torch.uint8, device=dev)
            _lib.check(lib.mgb_conv1d_backward(_lib.ptr(x), _lib.ptr(wc), _lib.ptr(rb), _lib.ptr(y), _lib.ptr(pre), _lib.ptr(gy),
                                               _lib.ptr(gx), _lib.ptr(gw), _lib.ptr(gb), _lib.ptr(grb), B, Tin, Cin, Cout, k,
                                               stride, act, _lib.ptr(ws), ws.numel(), _stream(dev)), "mgb_conv1d_backward")
        return (gx if need_x else None), gw, gb, grb, None, None


def conv1d_frames(x, weight, bias=None, rowbias=None, stride=1, act=ACT_NONE):
    """Library Conv1d on a frames-major ``[B, T, Cin]`` tensor (``weight`` in the torch layout ``[Cout, Cin, k]``)."""
    return _ConvFn.apply(x, weight, bias, rowbias, stride, act)


def linear_rows(x, weight, act=ACT_NONE):
    """Bias-free ``LinearNorm`` on ``[..., Cin]`` rows through the same kernels (a k = 1 convolution)."""
    lead = x.shape[:-1]
    y = _ConvFn.apply(x.reshape(1, -1, x.shape[-1]), weight.unsqueeze(-1), None, None, 1, act)
    return y.reshape(*lead, weight.shape[0])


def step_embedding(t, dim):
    """``DiffusionEmbedding`` (model/blocks.py:899-913): ``[B, dim]`` = [sin(t f) | cos(t f)]."""
    lib = _lib.load()
    tt = t.detach().to(torch.int64).contiguous()
    emb = torch.empty((tt.shape[0], dim), dtype=torch.float32, device=tt.device)
    with torch.cuda.device(tt.device):
        _lib.check(lib.mgb_step_embedding(_lib.ptr(tt), _lib.ptr(emb), tt.shape[0], dim, _stream(tt.device)), "mgb_step_embedding")
    return emb


class _Lin(nn.Module):
    """Parameter holder with the reference's bias-free ``LinearNorm`` key layout (``.linear.weight``)."""

    def __init__(self, cin, cout):
        super().__init__()
        self.linear = nn.Linear(cin, cout, bias=False)
        nn.init.xavier_uniform_(self.linear.weight)


class _Cv(nn.Module):
    """Parameter holder with the reference's ``ConvNorm`` key layout (``.conv.{weight,bias}``), N(0, 0.02) weights as
    ``JCUDiscriminator.weights_init`` sets them."""

    def __init__(self, cin, cout, k, stride):
        super().__init__()
        self.conv = nn.Conv1d(cin, cout, kernel_size=k, stride=stride, padding=(k - 1) // 2)
        self.conv.weight.data.normal_(0.0, 0.02)
        self.stride = stride


class JCUDiscriminator(nn.Module):
    """Joint conditional / unconditional discriminator; computes on the GPU through the C ABI only."""

    def __init__(self, preprocess_config, model_config, train_config=None):
        super().__init__()
        n_mel = preprocess_config["preprocessing"]["mel"]["n_mel_channels"]
        rc = model_config["denoiser"]["residual_channels"]
        d = model_config["discriminator"]
        n_layer, n_unc, n_cond = d["n_layer"], d["n_uncond_layer"], d["n_cond_layer"]
        ch, ks, st = d["n_channels"], d["kernel_sizes"], d["strides"]
        self.multi_speaker = bool(model_config["multi_speaker"])
        self.residual_channels = rc
        self.input_projection = _Lin(2 * n_mel, 2 * n_mel)
        self.mlp = nn.Sequential(_Lin(rc, rc * 4), nn.Identity(), _Lin(rc * 4, ch[n_layer - 1]))
        if self.multi_speaker:
            self.spk_mlp = nn.Sequential(_Lin(rc, ch[n_layer - 1]))
        self.conv_block = nn.ModuleList(
            _Cv(ch[i - 1] if i != 0 else 2 * n_mel, ch[i], ks[i], st[i]) for i in range(n_layer))
        self.uncond_conv_block = nn.ModuleList(
            _Cv(ch[i - 1], ch[i], ks[i], st[i]) for i in range(n_layer, n_layer + n_unc))
        self.cond_conv_block = nn.ModuleList(
            _Cv(ch[i - 1], ch[i], ks[i], st[i]) for i in range(n_layer, n_layer + n_cond))

    def forward(self, x_ts, x_t_prevs, s, t):
        """``x_ts, x_t_prevs [B,T,M]``, ``s [B,H]`` (multi-speaker only), ``t [B]`` -> ``(cond_feats, uncond_feats)``:
        two lists of ``[B, C_i, T_i]`` tensors (the first ``n_layer`` entries are shared), the last of each the logits."""
        if x_ts.device.type != "cuda":
            raise RuntimeError("mixgan_tts_b200.JCUDiscriminator needs CUDA tensors (no CPU fallback)")
        if self.multi_speaker and s is None:
            raise TypeError("multi_speaker JCUDiscriminator needs the speaker embedding")
        x = torch.cat([x_t_prevs, x_ts], dim=-1)                                         # mixgantts.py:263-265
        x = conv1d_frames(x, self.input_projection.linear.weight.unsqueeze(-1))          # LinearNorm, frames-major already
        emb = step_embedding(t, self.residual_channels)
        h = linear_rows(emb, self.mlp[0].linear.weight, act=ACT_MISH)                    # :266
        step = linear_rows(h, self.mlp[2].linear.weight)
        if self.multi_speaker:
            step = step + linear_rows(s.float(), self.spk_mlp[0].linear.weight)          # :267-268, :275
        cond_feats, uncond_feats = [], []
        for layer in self.conv_block:                                                    # :272-275
            x = conv1d_frames(x, layer.conv.weight, layer.conv.bias, None, layer.stride, ACT_LEAKY)
            cond_feats.append(x.transpose(1, 2))
            uncond_feats.append(x.transpose(1, 2))
        x_cond, x_uncond = x, x
        for i, layer in enumerate(self.cond_conv_block):                                 # :281-283 (the add of :277 fused)
            x_cond = conv1d_frames(x_cond, layer.conv.weight, layer.conv.bias, step if i == 0 else None, layer.stride, ACT_LEAKY)
            cond_feats.append(x_cond.transpose(1, 2))
        for layer in self.uncond_conv_block:                                             # :285-287
            x_uncond = conv1d_frames(x_uncond, layer.conv.weight, layer.conv.bias, None, layer.stride, ACT_LEAKY)
            uncond_feats.append(x_uncond.transpose(1, 2))
        return cond_feats, uncond_feats


def get_lsgan_losses_fn():
    """``model/loss.py:12-30`` verbatim in behaviour (plain torch on the logits; tiny)."""
    import torch.nn.functional as F

    def jcu_loss_fn(logit_cond, logit_uncond, label_fn, mask=None):
        cond = F.mse_loss(logit_cond, label_fn(logit_cond), reduction="none" if mask is not None else "mean")
        cond = (cond * mask).sum() / mask.sum() if mask is not None else cond
        unc = F.mse_loss(logit_uncond, label_fn(logit_uncond), reduction="none" if mask is not None else "mean")
        unc = (unc * mask).sum() / mask.sum() if mask is not None else unc
        return 0.5 * (cond + unc)

    def d_loss_fn(r_cond, r_unc, f_cond, f_unc, mask=None):
        return jcu_loss_fn(r_cond, r_unc, torch.ones_like, mask), jcu_loss_fn(f_cond, f_unc, torch.zeros_like, mask)

    def g_loss_fn(f_cond, f_unc, mask=None):
        return jcu_loss_fn(f_cond, f_unc, torch.ones_like, mask)

    return d_loss_fn, g_loss_fn


def feature_matching_loss(D_real_cond, D_real_uncond, D_fake_cond, D_fake_uncond, n_layers):
    """``MixGANTTSLoss.get_fm_loss`` (model/loss.py:221-227)."""
    import torch.nn.functional as F
    loss, w = 0, 4.0 / (n_layers + 1)
    for j in range(len(D_fake_cond) - 1):
        loss = loss + w * 0.5 * (F.l1_loss(D_real_cond[j].detach(), D_fake_cond[j]) + F.l1_loss(D_real_uncond[j].detach(), D_fake_uncond[j]))
    return loss
