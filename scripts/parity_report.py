"""Measured parity of every precision mode against the CPU oracle (normalised x0 and denormalised mel, relative L2).
Run on the GPU box: python scripts/parity_report.py > gpurun_out/parity_report.txt"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import torch
from helpers import Case, rel_l2
from mixgan_tts_b200 import GaussianDiffusion

def build(case, precision):
    gd = GaussianDiffusion(case.args, case.pc, case.mc, case.tc, precision=precision)
    gd.denoise_fn.load_state_dict({k: torch.from_numpy(v) for k, v in case.W.items()}, strict=True)
    return gd.cuda().eval()

cu = lambda t: None if t is None else t.cuda()
cases = [("naive LJ B=3 T=200 K=4", Case("LJSpeech", "naive", False, 3, 200, 0, 4321)),
         ("naive LJ B=2 T=800 K=4", Case("LJSpeech", "naive", False, 2, 800, 0, 77)),
         ("shallow AISHELL3 multi-spk B=2 T=300 K=1", Case("AISHELL3", "shallow", True, 2, 300, 7, 99))]
print(f"torch {torch.__version__}  device {torch.cuda.get_device_name(0)}")
for name, c in cases:
    final, states, x0s, start = c.oracle_forward()
    valid = ~c.t("pad_mask")
    x0_ref = x0s[-1][:, 0].transpose(1, 2)[valid]
    for prec in ("fp32", "fp16", "bf16"):
        gd = build(c, prec)
        mel = gd(None, cu(c.t("cond")), cu(c.t("spk")), cu(c.t("pad_mask")), coarse_mel=cu(c.t("coarse_mel")),
                 x_T=cu(c.t("x_T")), noises=cu(c.t("noises")), start_noise=cu(c.t("start_noise")))[0]
        x0 = gd.norm_spec(mel).cpu()[valid]
        print(f"{name:45s} {prec:5s} normalised x0 {rel_l2(x0, x0_ref):.3e}   mel {rel_l2(mel, final):.3e}")
