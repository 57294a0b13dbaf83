"""bf16 path diagnostics: error of one Denoiser call / full sampling vs the CPU oracle, watchdog word."""
import ctypes as C
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import Case, rel_l2  # noqa: E402
from mixgan_tts_b200 import GaussianDiffusion, _lib  # noqa: E402
from oracle.denoiser import denoiser_forward  # noqa: E402

lib = _lib.load()


def status(gd, B, T):
    den = gd.denoise_fn
    ws = den._ws.buf
    st = C.c_int(0)
    lib.mgb_debug_status(C.byref(den.dims), 1, B, T, _lib.ptr(ws), C.byref(st))
    return st.value


def run(model, multi, B, T, layers=None, dataset="LJSpeech"):
    c = Case(dataset, model, multi, B, T, wseed=0, iseed=T + B, layers=layers)
    gd = GaussianDiffusion(c.args, c.pc, c.mc, c.tc, precision="bf16")
    gd.denoise_fn.load_state_dict({k: torch.from_numpy(v) for k, v in c.W.items()})
    gd = gd.cuda().eval()
    cu = lambda k: None if c.t(k) is None else c.t(k).cuda()
    K = gd.num_timesteps
    t = torch.tensor([(K - 1 - b) % K for b in range(B)], dtype=torch.long)
    ref = denoiser_forward(c.oracle.W, c.t("x_T"), t, c.t("cond").transpose(1, 2), c.t("spk"))
    t0 = time.time()
    out = gd.denoise_fn(cu("x_T"), t.cuda(), cu("cond").transpose(1, 2), cu("spk"))
    torch.cuda.synchronize()
    st = status(gd, B, T)
    e_den = rel_l2(out, ref)
    per_b = [rel_l2(out[i], ref[i]) for i in range(B)]
    mel = gd(None, cu("cond"), cu("spk"), cu("pad_mask"), coarse_mel=cu("coarse_mel"), x_T=cu("x_T"),
             noises=cu("noises"), start_noise=cu("start_noise"))[0]
    torch.cuda.synchronize()
    st2 = status(gd, B, T)
    final, states, x0s, _ = c.oracle_forward()
    valid = ~c.t("pad_mask")
    e_mel = rel_l2(mel, final)
    e_norm = rel_l2(gd.norm_spec(mel).cpu()[valid], x0s[-1][:, 0].transpose(1, 2)[valid])
    print(f"{model:8s} multi={int(multi)} B={B} T={T:4d} L={layers or 20}: denoiser rel_l2={e_den:.3e} "
          f"(per-utt max {max(per_b):.3e}) status={st} | sampling mel={e_mel:.3e} norm-x0={e_norm:.3e} status={st2}"
          f" nan={bool(torch.isnan(out).any())}", flush=True)
    if e_den > 5e-2:
        d = (out.cpu() - ref)[0, 0]          # [M, T]
        per_t = d.pow(2).mean(0).sqrt()
        print("   per-frame rms err (utt 0):", " ".join(f"{v:.2f}" for v in per_t[:min(T, 160)].tolist()))


if __name__ == "__main__":
    print("MGB_GROUP_LAYERS =", os.environ.get("MGB_GROUP_LAYERS", "(default 10)"))
    run("naive", False, 1, 64, layers=1)
    run("naive", False, 1, 64, layers=2)
    run("naive", False, 2, 64)
    run("naive", False, 2, 200)
    run("naive", False, 1, 800)
    run("shallow", True, 2, 77, dataset="AISHELL3")
