"""GPU parity of the training path (BASELINE configs[4]): Denoiser forward + backward in the library against torch
autograd through the CPU oracle and against the gradients the REAL reference produced (tests/golden/grad_*.npz).

Tolerance (relative L2, fp32 arithmetic on both sides): 1e-4 per gradient tensor; measured ~1e-6."""
import ctypes as C

import numpy as np
import pytest
import torch

from mixgan_tts_b200 import GaussianDiffusion, _lib, synth
from helpers import (TRAIN_CASES, Case, check_grads_against_golden, grad_golden_name, load_golden,
                     oracle_training_grads, rel_l2, train_case)

pytestmark = pytest.mark.gpu
TOL = 1e-4
# Random edge shapes: a ReLU / clamp mask can legitimately flip for a pre-activation within rounding distance of the
# threshold (different fp32 summation order on the two sides); ONE flipped element moves a gradient by ~3e-4 relative
# at these small sizes, so the sweep uses north_star's fp32 bar instead of the 1e-4 the pinned golden cases meet.
TOL_EDGE = 1e-3


def build(case: Case, precision: str = "fp32") -> GaussianDiffusion:
    gd = GaussianDiffusion(case.args, case.pc, case.mc, case.tc, precision=precision)
    gd.denoise_fn.load_state_dict({k: torch.from_numpy(v) for k, v in case.W.items()})
    return gd.cuda().train()


def cu(t, grad=False):
    if t is None:
        return None
    t = t.cuda()
    return t.requires_grad_(True) if grad else t


def run_library(c, ex, probe, gd=None):
    gd = gd or build(c)
    gd.zero_grad(set_to_none=True)
    cond, spk = cu(c.t("cond"), True), cu(c.t("spk"), c.t("spk") is not None)
    out = gd(cu(ex["mel"]), cond, spk, cu(c.t("pad_mask")), coarse_mel=cu(c.t("coarse_mel")), t=cu(ex["t"]),
             noise_t=cu(ex["noise_t"]), noise_prev=cu(ex["noise_prev"]), post_noise=cu(ex["post_noise"]))
    loss = (out[0] * cu(probe["r0"])).sum() + (out[3] * cu(probe["r1"])).sum()
    loss.backward()
    torch.cuda.synchronize()
    grads = {k: p.grad.detach().cpu() for k, p in gd.denoise_fn.named_parameters()}
    return loss.detach().cpu(), out, grads, cond.grad.cpu(), (spk.grad.cpu() if spk is not None else None), gd


def probe_for(name, c):
    return {k: torch.from_numpy(v) for k, v in synth.grad_probe(TRAIN_CASES[name][6] + 2000, c.B, c.T).items()}


@pytest.mark.parametrize("name", list(TRAIN_CASES))
def test_gradients_vs_reference_golden_and_oracle(name):
    g = load_golden(grad_golden_name(name))
    c, ex = train_case(name)
    probe = probe_for(name, c)
    loss, out, grads, gcond, gspk, _ = run_library(c, ex, probe)
    assert abs(float(loss) - float(g["loss"])) < 1e-4 * max(1.0, abs(float(g["loss"])))
    check_grads_against_golden(g, grads, gcond, gspk, tol=TOL)
    # every element of every gradient against torch autograd through the oracle
    _, oout, ograds, ogcond, ogspk = oracle_training_grads(c, ex, probe)
    for a, b in zip(out[:4], oout[:4]):
        assert rel_l2(a.detach(), b.detach()) < TOL
    assert rel_l2(gcond, ogcond) < TOL
    if ogspk is not None:
        assert rel_l2(gspk, ogspk) < TOL
    total = float(torch.sqrt(sum(v.double().pow(2).sum() for v in ograds.values())))
    for k, ref in ograds.items():
        err = float((grads[k].double() - ref.double()).norm()) / max(float(ref.double().norm()), 1e-6 * total)
        assert err < TOL, (k, err)


@pytest.mark.parametrize("multi,B,T,L", [(False, 1, 1, 2), (False, 2, 3, 1), (True, 2, 129, 3), (False, 3, 257, 2)])
def test_gradients_edge_shapes_vs_oracle(multi, B, T, L):
    """Short / ragged / non-multiple-of-tile utterances and shallow stacks, incl. d loss / d mel of a bare Denoiser call."""
    c = Case("AISHELL3" if multi else "LJSpeech", "naive", multi, B, T, wseed=3, iseed=40 + T, layers=L)
    from oracle.denoiser import denoiser_forward_graph
    gd = build(c)
    x = cu(c.t("x_T"), True)
    cond = cu(c.t("cond").transpose(1, 2).contiguous(), True)
    spk = cu(c.t("spk"), multi)
    t = torch.arange(B, dtype=torch.long).cuda() % c.K
    r = torch.randn(B, 1, 80, T, generator=torch.Generator().manual_seed(5))
    out = gd.denoise_fn(x, t, cond, spk)
    (out * r.cuda()).sum().backward()
    W = {k: torch.from_numpy(v).requires_grad_(True) for k, v in c.W.items()}
    ox = c.t("x_T").requires_grad_(True)
    ocond = c.t("cond").transpose(1, 2).contiguous().requires_grad_(True)
    ospk = c.t("spk").requires_grad_(True) if multi else None
    oout = denoiser_forward_graph(W, ox, t.cpu(), ocond, ospk)
    (oout * r).sum().backward()
    assert rel_l2(out.detach(), oout.detach()) < TOL_EDGE
    assert rel_l2(x.grad, ox.grad) < TOL_EDGE
    assert rel_l2(cond.grad, ocond.grad) < TOL_EDGE
    if multi:
        assert rel_l2(spk.grad, ospk.grad) < TOL_EDGE
    total = float(torch.sqrt(sum(v.grad.double().pow(2).sum() for v in W.values() if v.grad is not None)))
    for k, p in gd.denoise_fn.named_parameters():
        ref = W[k].grad if W[k].grad is not None else torch.zeros_like(W[k])
        err = float((p.grad.cpu().double() - ref.double()).norm()) / max(float(ref.double().norm()), 1e-6 * total)
        assert err < TOL_EDGE, (k, err)


def test_segmented_backward_is_bitwise_the_one_shot_backward_and_deterministic():
    """Gradient buckets (what the all-reduce overlaps with) change the launch grouping only, not one bit of the result."""
    from mixgan_tts_b200.grad_sync import plan_buckets
    name = "train_naive_lj_B3_T48"
    c, ex = train_case(name)
    probe = probe_for(name, c)
    _, _, g1, gc1, _, gd = run_library(c, ex, probe)
    _, _, g2, gc2, _, _ = run_library(c, ex, probe, gd)
    assert all(torch.equal(g1[k], g2[k]) for k in g1) and torch.equal(gc1, gc2)

    class OneRankSync:            # buckets without communication
        bucket_bytes = 4 << 20
        n = 0
        def reduce_async(self, b): self.n += 1
        def finish(self): pass
    gd.denoise_fn.grad_sync = OneRankSync()
    _, _, g3, gc3, _, _ = run_library(c, ex, probe, gd)
    assert gd.denoise_fn.grad_sync.n > 3
    assert all(torch.equal(g1[k], g3[k]) for k in g1) and torch.equal(gc1, gc3)


def test_optimizer_step_repacks_weights_and_inference_sees_them():
    """After an optimizer step the cached kernel-layout weights are rebuilt (parameter versions change)."""
    c, ex = train_case("train_naive_lj_B3_T48")
    probe = probe_for("train_naive_lj_B3_T48", c)
    gd = build(c)
    opt = torch.optim.SGD(gd.parameters(), lr=1e-3)
    l0 = run_library(c, ex, probe, gd)[0]
    opt.step()
    l1 = run_library(c, ex, probe, gd)[0]
    assert float(l1) < float(l0)          # a descent step on a linear probe loss lowers it
    with torch.no_grad():
        gd.eval()
        y = gd.denoise_fn(cu(c.t("x_T")), torch.zeros(c.B, dtype=torch.long).cuda(), cu(c.t("cond")).transpose(1, 2), None)
    W2 = {k: v.detach().cpu().numpy() for k, v in gd.denoise_fn.state_dict().items()}
    from oracle.denoiser import denoiser_forward
    ref = denoiser_forward({k: torch.from_numpy(v) for k, v in W2.items()}, c.t("x_T"), torch.zeros(c.B, dtype=torch.long),
                           c.t("cond").transpose(1, 2), None)
    assert rel_l2(y, ref) < TOL


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_fused_adam_updates_reach_the_kernels(precision):
    """Regression: torch.optim.Adam(fused=True) does not bump Tensor._version; the library must still see the new weights
    in the next training forward AND in inference afterwards."""
    c, ex = train_case("train_naive_lj_B3_T48")
    probe = probe_for("train_naive_lj_B3_T48", c)
    gd = build(c, precision)
    opt = torch.optim.Adam(gd.parameters(), lr=1e-5, fused=True)
    losses = []
    for _ in range(4):
        losses.append(float(run_library(c, ex, probe, gd)[0]))
        opt.step()
    assert losses[3] < losses[2] < losses[1] < losses[0], losses     # every forward saw the previous step's update
    with torch.no_grad():
        gd.eval()
        y = gd.denoise_fn(cu(c.t("x_T")), torch.zeros(c.B, dtype=torch.long).cuda(), cu(c.t("cond")).transpose(1, 2), None)
    from oracle.denoiser import denoiser_forward
    W2 = {k: v.detach().cpu() for k, v in gd.denoise_fn.state_dict().items()}
    ref = denoiser_forward(W2, c.t("x_T"), torch.zeros(c.B, dtype=torch.long), c.t("cond").transpose(1, 2), None)
    assert rel_l2(y, ref) < (TOL if precision == "fp32" else TOL_BF16_OUT)


def test_train_abi_errors():
    lib = _lib.load()
    d = _lib.ModelDims(80, 256, 256, 20, 0)
    assert lib.mgb_train_segments(C.byref(d)) == 22
    assert lib.mgb_denoiser_backward(C.byref(d), _lib.PREC_FP32, None, None, 0, None, None, None, None, None, None, None,
                                     None, 1, 8, 0, 22, None, 0, None) == _lib.E_ARG
    x = torch.zeros(16, device="cuda")
    rc = lib.mgb_denoiser_train_forward(C.byref(d), 7, _lib.ptr(x), _lib.ptr(x), _lib.ptr(x), _lib.ptr(x), _lib.ptr(x), None,
                                        _lib.ptr(x), _lib.ptr(x), 64, 1, 8, _lib.ptr(x), 64, None)
    assert rc == _lib.E_ARG
    for prec in (_lib.PREC_FP32, _lib.PREC_BF16):
        rc = lib.mgb_denoiser_train_forward(C.byref(d), prec, _lib.ptr(x), _lib.ptr(x), _lib.ptr(x), _lib.ptr(x), _lib.ptr(x), None,
                                            _lib.ptr(x), _lib.ptr(x), 64, 1, 8, _lib.ptr(x), 64, None)
        assert rc == _lib.E_WORKSPACE


# ---------------------------------------------------------------------------------------------------------------
# bf16 tensor-core training path.  Stated tolerances (relative L2 against fp32 torch autograd through the oracle; the
# measured values are written to gpurun_out/train_bf16_parity.txt):
#   outputs 8e-3 (the inference path's bf16 bar; measured 2.4-2.7e-3);
#   gradients 2e-2 per tensor when the two ReLU masks are stable (biases shifted so that no pre-activation sits near 0):
#     this isolates the arithmetic of the GEMMs/epilogues — bf16 operand rounding (2^-9 per element) through 20 blocks;
#   gradients 8e-2 per tensor with the random-init weights (measured 4.7-6.5e-2): there ~0.25 % of the skip-projection / input-projection
#     pre-activations lie within the bf16 forward error of zero, their ReLU masks flip against the fp32 run, and a
#     fraction p of flipped units moves a gradient by ~sqrt(p) = 5 % in L2 (measured 5-6 %, uniform over layers).  Any
#     bf16 training of this network has this property; it is not an error of the kernels.
TOL_BF16_OUT, TOL_BF16_GRAD, TOL_BF16_GRAD_RELU = 8e-3, 2e-2, 8e-2


def _status(gd, B, T, ws=None):
    lib = _lib.load()
    st = C.c_int(-1)
    den = gd.denoise_fn
    _lib.check(lib.mgb_train_debug_status(C.byref(den.dims), B, T, _lib.ptr(ws if ws is not None else den._train_ws.buf),
                                          C.byref(st)), "status")
    return st.value


def _report(line):
    import os
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/train_bf16_parity.txt", "a") as f:
        f.write(line + "\n")
    print(line)


@pytest.mark.parametrize("name", list(TRAIN_CASES))
def test_bf16_training_path_vs_oracle(name):
    c, ex = train_case(name)
    probe = probe_for(name, c)
    gd = build(c, "bf16")
    loss, out, grads, gcond, gspk, _ = run_library(c, ex, probe, gd)
    assert _status(gd, c.B, c.T) == 0, "a bf16 training kernel hit its watchdog"
    oloss, oout, ograds, ogcond, ogspk = oracle_training_grads(c, ex, probe)
    e_out = max(rel_l2(a.detach(), b.detach()) for a, b in zip(out[:4], oout[:4]))
    e_cond = rel_l2(gcond, ogcond)
    _report(f"{name}: loss {float(loss):.5f} (oracle {float(oloss):.5f})  out {e_out:.3e}  grad_cond {e_cond:.3e}"
            + (f"  grad_spk {rel_l2(gspk, ogspk):.3e}" if ogspk is not None else ""))
    total = float(torch.sqrt(sum(v.double().pow(2).sum() for v in ograds.values())))
    worst = ("", 0.0)
    for k, ref in ograds.items():
        err = float((grads[k].double() - ref.double()).norm()) / max(float(ref.double().norm()), 1e-3 * total)
        if err > worst[1]:
            worst = (k, err)
        if not k.startswith("residual_layers.") or k.startswith("residual_layers.0.") or k.startswith("residual_layers.19."):
            _report(f"    {k:55s} {err:.3e}  |ref| {float(ref.norm()):.3e}")
    _report(f"    worst: {worst[0]} {worst[1]:.3e}")
    assert e_out < TOL_BF16_OUT and e_cond < TOL_BF16_GRAD_RELU
    if ogspk is not None:
        assert rel_l2(gspk, ogspk) < TOL_BF16_GRAD_RELU
    assert worst[1] < TOL_BF16_GRAD_RELU, worst


@pytest.mark.parametrize("multi,B,T,L", [(False, 1, 1, 2), (False, 2, 3, 1), (True, 2, 129, 3), (False, 3, 257, 2),
                                         (True, 4, 200, 20), (False, 8, 800, 20), (True, 12, 1500, 20)])
def test_bf16_training_edge_shapes_vs_fp32_path(multi, B, T, L):
    """bf16 path against the library's own fp32 path (which the tests above pin to the oracle) on ragged shapes, with the
    two ReLU masks made stable (+8 on the input- and skip-projection biases) so that the comparison measures arithmetic."""
    c = Case("AISHELL3" if multi else "LJSpeech", "naive", multi, B, T, wseed=3, iseed=40 + T, layers=L)
    c.W = dict(c.W)
    for k in ("input_projection.0.conv.bias", "skip_projection.conv.bias"):
        c.W[k] = c.W[k] + np.float32(8.0)
    res = {}
    for prec in ("fp32", "bf16"):
        gd = build(c, prec)
        x = cu(c.t("x_T"), True)
        cond = cu(c.t("cond").transpose(1, 2).contiguous(), True)
        spk = cu(c.t("spk"), multi)
        t = torch.arange(B, dtype=torch.long).cuda() % c.K
        r = torch.randn(B, 1, 80, T, generator=torch.Generator().manual_seed(5)).cuda()
        out = gd.denoise_fn(x, t, cond, spk)
        (out * r).sum().backward()
        torch.cuda.synchronize()
        if prec == "bf16":
            assert _status(gd, B, T) == 0
        res[prec] = (out.detach(), x.grad, cond.grad, spk.grad if multi else None,
                     {k: p.grad.detach() for k, p in gd.denoise_fn.named_parameters()})
    a, b = res["bf16"], res["fp32"]
    errs = {"out": rel_l2(a[0], b[0]), "dx": rel_l2(a[1], b[1]), "dcond": rel_l2(a[2], b[2])}
    if multi:
        errs["dspk"] = rel_l2(a[3], b[3])
    total = float(torch.sqrt(sum(v.double().pow(2).sum() for v in b[4].values())))
    worst = ("", 0.0)
    for k, ref in b[4].items():
        err = float((a[4][k].double() - ref.double()).norm()) / max(float(ref.double().norm()), 1e-3 * total)
        if err > worst[1]:
            worst = (k, err)
    _report(f"stable-mask multi={multi} B={B} T={T} L={L}: " + " ".join(f"{k} {v:.3e}" for k, v in errs.items())
            + f"  worst param grad {worst[0]} {worst[1]:.3e}")
    assert errs["out"] < TOL_BF16_OUT
    assert all(v < TOL_BF16_GRAD for k, v in errs.items() if k != "out"), errs
    assert worst[1] < TOL_BF16_GRAD, worst


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_cuda_graph_replay_matches_eager_bit_for_bit(precision):
    """From the third call on, the library's launch sequences are replayed as CUDA graphs (static buffers); the
    gradients and the parameter trajectory under an optimizer must be bit-identical to the eager path."""
    name = "train_shallow_aishell_spk_B2_T40"
    c, ex = train_case(name)
    probe = probe_for(name, c)
    hist = {}
    for graphs in (False, True):
        gd = build(c, precision)
        gd.denoise_fn.use_cuda_graphs = graphs
        opt = torch.optim.SGD(gd.parameters(), lr=1e-3)
        losses = []
        for step in range(6):
            loss, _, grads, gcond, gspk, _ = run_library(c, ex, probe, gd)
            opt.step()
            losses.append((float(loss), gcond.clone(), gspk.clone(), {k: v.clone() for k, v in grads.items()}))
        hist[graphs] = losses
        if graphs:
            tgs = list(gd.denoise_fn._train_graphs.values())
            assert len(tgs) == 1 and tgs[0].fwd is not None and len(tgs[0].bwd) == 1, "the graph path was not taken"
            if precision == "bf16":
                assert _status(gd, c.B, c.T) == 0 and _status(gd, c.B, c.T, tgs[0].ws) == 0
    for (la, ca, sa, ga), (lb, cb, sb, gb) in zip(hist[False], hist[True]):
        assert la == lb
        assert torch.equal(ca, cb) and torch.equal(sa, sb)
        assert all(torch.equal(ga[k], gb[k]) for k in ga)
    assert hist[True][-1][0] < hist[True][0][0]      # and the optimizer actually moved the parameters


def test_second_forward_before_backward_falls_back_to_eager():
    """The graph path owns ONE activation stash per signature: a second forward in flight takes the eager path."""
    c = Case("LJSpeech", "naive", False, 2, 33, wseed=3, iseed=9, layers=2)
    gd = build(c, "fp32")
    den = gd.denoise_fn
    x, cond = cu(c.t("x_T")), cu(c.t("cond").transpose(1, 2).contiguous())
    t = torch.zeros(2, dtype=torch.long).cuda()
    outs = []
    for _ in range(3):                                   # warm-up eager calls, then one graphed forward
        cr = cond.clone().requires_grad_(True)
        o = den(x, t, cr, None)
        o.sum().backward()
        outs.append(cr.grad.clone())
    c1, c2 = cond.clone().requires_grad_(True), cond.clone().requires_grad_(True)
    o1 = den(x, t, c1, None)                             # graphed, holds the stash
    o2 = den(x, t, c2, None)                             # must not clobber it
    o2.sum().backward()
    o1.sum().backward()
    assert torch.equal(c1.grad, outs[0]) and torch.equal(c2.grad, outs[0])
