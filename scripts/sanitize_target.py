"""Small workload for compute-sanitizer (scripts/gpu_sanitize.sh): the smoke shapes through every hand-written kernel family -
fused_pair_kernel (bf16 + fp16 operands, per-utterance and uniform-timestep variants, two layer groups), the elementwise
kernels, the length regulator, and the bf16 / fp32 training forward + backward (fgemm_kernel, wgemm_kernel, small ops)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import torch
from helpers import Case
from mixgan_tts_b200 import GaussianDiffusion
from mixgan_tts_b200.length_regulator import LengthRegulator, durations_from_log

which = sys.argv[1] if len(sys.argv) > 1 else "all"
cu = lambda t: None if t is None else t.cuda()

def build(c, prec, train=False):
    gd = GaussianDiffusion(c.args, c.pc, c.mc, c.tc, precision=prec)
    gd.denoise_fn.load_state_dict({k: torch.from_numpy(v) for k, v in c.W.items()})
    gd = gd.cuda()
    return gd.train() if train else gd.eval()

if which in ("all", "sample"):
    for prec in ("bf16", "fp16"):
        c = Case("LJSpeech", "naive", False, 2, 200, wseed=0, iseed=7)                       # KUNI kernels, 2 groups
        gd = build(c, prec)
        with torch.no_grad():
            mel = gd(None, cu(c.t("cond")), None, cu(c.t("pad_mask")), x_T=cu(c.t("x_T")), noises=cu(c.t("noises")))[0]
            t = torch.tensor([3, 1], device="cuda")
            x0 = gd.denoise_fn(cu(c.t("x_T")), t, cu(c.t("cond")).transpose(1, 2), None)      # per-utterance k_l kernels
        c2 = Case("AISHELL3", "shallow", True, 2, 77, wseed=7, iseed=99)                      # shallow start, multi-speaker
        gd2 = build(c2, prec)
        with torch.no_grad():
            m2 = gd2(None, cu(c2.t("cond")), cu(c2.t("spk")), cu(c2.t("pad_mask")), coarse_mel=cu(c2.t("coarse_mel")),
                     x_T=cu(c2.t("x_T")), noises=cu(c2.t("noises")), start_noise=cu(c2.t("start_noise")))[0]
        torch.cuda.synchronize()
        print(prec, "sample ok", float(mel.abs().mean()), float(x0.abs().mean()), float(m2.abs().mean()))
if which in ("all", "lr"):
    x = torch.randn(3, 9, 256, device="cuda", requires_grad=True)
    dur = durations_from_log(torch.randn(3, 9, device="cuda") + 1.0, 1.2)
    out, ml = LengthRegulator()(x, dur, None)
    out.sum().backward()
    torch.cuda.synchronize()
    print("lr ok", tuple(out.shape), ml.tolist())
if which in ("all", "train"):
    for prec in ("bf16", "fp32"):
        c = Case("LJSpeech", "naive", False, 2, 96, wseed=3, iseed=50, layers=3)
        gd = build(c, prec, train=True)
        cond = cu(c.t("cond")).transpose(1, 2).contiguous().requires_grad_(True)
        out = gd.denoise_fn(cu(c.t("x_T")), torch.tensor([3, 0], device="cuda"), cond, None)
        out.sum().backward()
        torch.cuda.synchronize()
        print(prec, "train ok", float(out.abs().mean()), float(cond.grad.abs().mean()))
