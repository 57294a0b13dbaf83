"""Step-by-step run of the bf16 training path against the library's fp32 training path (debug aid)."""
import ctypes as C, faulthandler, sys, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
faulthandler.enable()
faulthandler.dump_traceback_later(40, exit=True)
import torch
from mixgan_tts_b200 import GaussianDiffusion, _lib, configs, synth

def log(*a):
    print(*a, flush=True)

B, T, L = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
multi = len(sys.argv) > 4 and sys.argv[4] == "multi"
cfg = configs.make_configs("AISHELL3" if multi else "LJSpeech", "naive", multi, residual_layers=L)
W = synth.make_denoiser_weights(3, layers=L, multi_speaker=multi)
inp = synth.make_inputs(41, B, T, 4, multi_speaker=multi)
rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
res = {}
for prec in ("fp32", "bf16"):
    gd = GaussianDiffusion(*cfg, precision=prec)
    gd.denoise_fn.load_state_dict({k: torch.from_numpy(v) for k, v in W.items()})
    gd = gd.cuda().train()
    den = gd.denoise_fn
    x = torch.from_numpy(inp["x_T"]).cuda().requires_grad_(True)
    cond = torch.from_numpy(inp["cond"]).cuda().transpose(1, 2).contiguous().requires_grad_(True)
    spk = torch.from_numpy(inp["spk"]).cuda().requires_grad_(True) if multi else None
    t = torch.arange(B, dtype=torch.long).cuda() % 4
    r = torch.randn(B, 1, 80, T, generator=torch.Generator().manual_seed(5)).cuda()
    log(prec, "forward ...")
    out = den(x, t, cond, spk)
    torch.cuda.synchronize()
    log(prec, "forward done, |out|", float(out.norm()))
    (out * r).sum().backward()
    torch.cuda.synchronize()
    log(prec, "backward done")
    if prec == "bf16":
        st = C.c_int(-1)
        _lib.load().mgb_train_debug_status(C.byref(den.dims), B, T, _lib.ptr(den._train_ws.buf), C.byref(st))
        log("watchdog status", st.value)
    res[prec] = (out.detach(), x.grad, cond.grad, {k: p.grad.detach().clone() for k, p in den.named_parameters()})
a, b = res["bf16"], res["fp32"]
log("out", rel(a[0], b[0]), "dx", rel(a[1], b[1]), "dcond", rel(a[2], b[2]))
tot = float(torch.sqrt(sum(v.double().pow(2).sum() for v in b[3].values())))
for k in b[3]:
    e = float((a[3][k].double() - b[3][k].double()).norm()) / max(float(b[3][k].double().norm()), 1e-3 * tot)
    if e > 2e-2 or not k.startswith("residual_layers") or ".0." in k:
        log(f"  {k:55s} {e:.3e} |ref| {float(b[3][k].norm()):.3e}")
