"""GPU parity of the JCU discriminator: forward features and the D-step / G-step gradients of
train.py:126-184 against the goldens the REAL reference produced (tests/golden/jcu_*.npz) and against torch autograd through
the CPU oracle; the generic Conv1d kernels behind it against F.conv1d on ragged shapes.  fp32 on both sides: 1e-4."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from mixgan_tts_b200 import JCUDiscriminator, configs, synth
from mixgan_tts_b200.discriminator import conv1d_frames, feature_matching_loss, get_lsgan_losses_fn
from helpers import load_golden, rel_l2

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_golden_jcu import JCU_CASES, WEIGHT_STD  # noqa: E402

pytestmark = pytest.mark.gpu
TOL = 1e-4


def build(multi, wseed):
    args, pc, mc, tc = configs.make_configs("LJSpeech", "naive", multi)
    D = JCUDiscriminator(pc, mc, tc)
    W = synth.make_discriminator_weights(wseed, multi_speaker=multi, weight_std=WEIGHT_STD)
    D.load_state_dict({k: torch.from_numpy(v) for k, v in W.items()}, strict=True)
    return D.cuda().train(), mc


@pytest.mark.parametrize("name", list(JCU_CASES))
def test_features_and_gradients_vs_reference_golden(name):
    multi, B, T, wseed, iseed = JCU_CASES[name]
    g = load_golden(name)
    D, mc = build(multi, wseed)
    inp = synth.make_discriminator_inputs(iseed, B, T, 4, multi_speaker=multi)
    cu = lambda k: None if inp[k] is None else torch.from_numpy(inp[k]).cuda()
    x_ts, prevs, spk, t = cu("x_ts"), cu("x_t_prevs"), cu("spk"), cu("t")
    preds = cu("x_t_prev_preds").requires_grad_(True)
    fc, fu = D(x_ts, preds, spk, t)
    rc, ru = D(x_ts, prevs, spk, t)
    for lst, key in zip((fc, fu, rc, ru), ("fake_cond", "fake_uncond", "real_cond", "real_uncond")):
        assert len(lst) == 5
        for i, f in enumerate(lst):
            assert tuple(f.shape) == g[f"{key}/{i}"].shape, (key, i)
            assert rel_l2(f.detach(), g[f"{key}/{i}"]) < TOL, (key, i)
    d_loss_fn, g_loss_fn = get_lsgan_losses_fn()
    # D step (train.py:139-146)
    fcd, fud = D(x_ts, preds.detach(), spk, t)
    r, f = d_loss_fn(rc[-1], ru[-1], fcd[-1], fud[-1])
    d_loss = r + f
    assert abs(float(d_loss) - float(g["d_loss"])) < TOL * abs(float(g["d_loss"]))
    D.zero_grad(set_to_none=True)
    d_loss.backward(retain_graph=True)
    total = float(np.sqrt(sum(float(g[k]) ** 2 for k in g.files if k.startswith("gnorm/"))))
    for k, p in D.named_parameters():
        flat = p.grad.detach().reshape(-1).double().cpu()
        ref_norm = float(g[f"gnorm/{k}"])
        assert abs(float(flat.norm()) - ref_norm) < TOL * max(ref_norm, 1e-6 * total), (k, float(flat.norm()), ref_norm)
        idx = torch.from_numpy(synth.grad_sample_index(flat.numel()))
        ref_s = torch.from_numpy(g[f"gsample/{k}"])
        assert float((flat[idx] - ref_s).norm()) < 10 * TOL * max(float(ref_s.norm()), 1e-6 * total), k
    # G step (train.py:159-182): adversarial + feature matching, gradient into the Denoiser's x_t_prev_pred
    n_layers = mc["discriminator"]["n_layer"] + mc["discriminator"]["n_cond_layer"]
    g_loss = g_loss_fn(fc[-1], fu[-1]) + feature_matching_loss(rc, ru, fc, fu, n_layers)
    assert abs(float(g_loss) - float(g["g_loss"])) < TOL * abs(float(g["g_loss"]))
    D.zero_grad(set_to_none=True)
    g_loss.backward()
    assert rel_l2(preds.grad, g["g_grad_preds"]) < TOL


@pytest.mark.parametrize("multi", [False, True])
def test_one_call_on_the_concatenated_batch_equals_two_calls(multi):
    """bench.py's GAN step runs the fake and the real pair of a phase as ONE call on the 2B-utterance batch (train.py calls
    the discriminator twice): same features, same losses, same parameter and input gradients (utterances are independent
    in every layer; only fp32 summation orders differ)."""
    B, T = 3, 96
    D, mc = build(multi, 2)
    inp = synth.make_discriminator_inputs(31, B, T, 4, multi_speaker=multi)
    cu = lambda k: None if inp[k] is None else torch.from_numpy(inp[k]).cuda()
    x_ts, prevs, spk, t = cu("x_ts"), cu("x_t_prevs"), cu("spk"), cu("t")
    n_layers = mc["discriminator"]["n_layer"] + mc["discriminator"]["n_cond_layer"]
    d_loss_fn, g_loss_fn = get_lsgan_losses_fn()

    def losses(fc, fu, rc, ru):
        r, f = d_loss_fn(rc[-1], ru[-1], fc[-1], fu[-1])
        return r + f + g_loss_fn(fc[-1], fu[-1]) + feature_matching_loss(rc, ru, fc, fu, n_layers)

    p1 = cu("x_t_prev_preds").requires_grad_(True)
    fc, fu = D(x_ts, p1, spk, t)
    rc, ru = D(x_ts, prevs, spk, t)
    D.zero_grad(set_to_none=True)
    l1 = losses(fc, fu, rc, ru)
    l1.backward()
    g1 = {k: p.grad.clone() for k, p in D.named_parameters()}
    p2 = cu("x_t_prev_preds").requires_grad_(True)
    D.zero_grad(set_to_none=True)
    c, u = D(torch.cat([x_ts, x_ts]), torch.cat([p2, prevs]), None if spk is None else torch.cat([spk, spk]), torch.cat([t, t]))
    fc2, fu2, rc2, ru2 = [f[:B] for f in c], [f[:B] for f in u], [f[B:] for f in c], [f[B:] for f in u]
    for a, b in zip(fc + fu + rc + ru, fc2 + fu2 + rc2 + ru2):
        assert rel_l2(b.detach(), a.detach()) < 1e-6
    l2 = losses(fc2, fu2, rc2, ru2)
    l2.backward()
    assert abs(float(l1) - float(l2)) < 1e-6 * abs(float(l1))
    assert rel_l2(p2.grad, p1.grad) < 1e-5
    for k, p in D.named_parameters():
        assert rel_l2(p.grad, g1[k]) < 1e-5, k


def test_state_dict_keys_match_the_reference_layout():
    for multi in (False, True):
        D, _ = build(multi, 1)
        W = synth.make_discriminator_weights(1, multi_speaker=multi)
        sd = D.state_dict()
        assert set(sd) == set(W)
        assert all(tuple(sd[k].shape) == W[k].shape for k in W)


@pytest.mark.parametrize("B,T", [(1, 1), (1, 2), (2, 3), (1, 5), (2, 129), (3, 64)])
def test_edge_lengths_vs_oracle(B, T):
    from oracle.discriminator import jcu_forward
    D, mc = build(True, 5)
    W = {k: v.detach().cpu() for k, v in D.state_dict().items()}
    inp = synth.make_discriminator_inputs(100 + T, B, T, 4, multi_speaker=True)
    cu = lambda k: torch.from_numpy(inp[k]).cuda()
    with torch.no_grad():
        fc, fu = D(cu("x_ts"), cu("x_t_prevs"), cu("spk"), cu("t"))
    tt = lambda k: torch.from_numpy(inp[k])
    oc, ou = jcu_forward(W, mc["discriminator"], tt("x_ts"), tt("x_t_prevs"), tt("spk"), tt("t"), multi_speaker=True)
    for a, b in zip(fc + fu, oc + ou):
        assert tuple(a.shape) == tuple(b.shape)
        assert rel_l2(a, b) < TOL


@pytest.mark.parametrize("B,T,Cin,Cout,k,stride,act", [
    (2, 17, 8, 12, 3, 1, 0), (1, 40, 160, 64, 3, 1, 1), (3, 33, 64, 128, 5, 2, 1), (2, 9, 6, 1, 3, 1, 1),
    (1, 64, 128, 512, 5, 2, 1), (2, 31, 12, 20, 7, 3, 3), (4, 1, 256, 1024, 1, 1, 2), (2, 50, 512, 128, 5, 1, 1),
    # split-K forward / data-gradient plans (few output tiles, deep K) and a many-tile plan, at the training step's shapes
    (16, 200, 512, 128, 5, 1, 1), (4, 401, 128, 512, 5, 2, 1), (16, 1, 1024, 512, 1, 1, 0), (8, 200, 128, 1, 3, 1, 0),
    (6, 333, 160, 64, 3, 1, 1),
])
def test_generic_conv1d_forward_and_backward_vs_torch(B, T, Cin, Cout, k, stride, act):
    """The library Conv1d against F.conv1d + autograd (fp32, channels-first), incl. the fused row bias."""
    g = torch.Generator().manual_seed(B * 1000 + T)
    x = torch.randn(B, T, Cin, generator=g).cuda().requires_grad_(True)
    w = (torch.randn(Cout, Cin, k, generator=g) / (Cin * k) ** 0.5).cuda().requires_grad_(True)
    b = torch.randn(Cout, generator=g).cuda().requires_grad_(True)
    rb = torch.randn(B, Cin, generator=g).cuda().requires_grad_(True)
    y = conv1d_frames(x, w, b, rb, stride, act)
    # the torch side runs on the CPU: cuDNN convolutions use TF32 by default, which is not an fp32 reference
    xr, wr, br, rbr = (v.detach().cpu().clone().requires_grad_(True) for v in (x, w, b, rb))
    z = F.conv1d((xr + rbr[:, None, :]).transpose(1, 2), wr, br, stride=stride, padding=(k - 1) // 2)
    z = {0: lambda v: v, 1: lambda v: F.leaky_relu(v, 0.2), 2: lambda v: v * torch.tanh(F.softplus(v)), 3: F.relu}[act](z)
    ref = z.transpose(1, 2)
    assert tuple(y.shape) == tuple(ref.shape)
    assert rel_l2(y.detach(), ref.detach()) < 1e-5
    r = torch.randn(ref.shape, generator=g).cuda()
    (y * r).sum().backward()
    (ref * r.cpu()).sum().backward()
    for name, a, c in (("x", x, xr), ("w", w, wr), ("bias", b, br), ("rowbias", rb, rbr)):
        assert rel_l2(a.grad, c.grad) < 1e-5, name
    # deterministic: a second backward gives the same bits
    x.grad = w.grad = b.grad = rb.grad = None
    y2 = conv1d_frames(x, w, b, rb, stride, act)
    (y2 * r).sum().backward()
    g1 = w.grad.clone()
    w.grad = None
    y3 = conv1d_frames(x, w, b, rb, stride, act)
    (y3 * r).sum().backward()
    assert torch.equal(g1, w.grad)


def test_no_cpu_fallback():
    args, pc, mc, tc = configs.make_configs("LJSpeech", "naive")
    D = JCUDiscriminator(pc, mc, tc)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        D(torch.zeros(1, 8, 80), torch.zeros(1, 8, 80), None, torch.zeros(1, dtype=torch.long))
