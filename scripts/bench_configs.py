"""Device-resident throughput of the other BASELINE.json configurations (not bench lines; context for DESIGN.md):
config 2 in fp32 parity mode, config 3 (LJSpeech shallow, K=1) and config 4 (AISHELL3 shallow, multi-speaker,
T=1500).  CUDA-event timing of GaussianDiffusion.forward with injected noise, inputs resident in HBM."""
import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from helpers import Case
from mixgan_tts_b200 import GaussianDiffusion

def run(name, ds, model, multi, B, T, prec, steps):
    c = Case(ds, model, multi, B, T, wseed=0, iseed=5)
    gd = GaussianDiffusion(c.args, c.pc, c.mc, c.tc, precision=prec)
    gd.denoise_fn.load_state_dict({k: torch.from_numpy(v) for k, v in c.W.items()})
    gd = gd.cuda().eval()
    cu = lambda k: None if c.inp[k] is None else c.t(k).cuda()
    args = (None, cu("cond"), cu("spk"), cu("pad_mask"))
    kw = dict(coarse_mel=cu("coarse_mel"), x_T=cu("x_T"), noises=cu("noises"), start_noise=cu("start_noise"))
    for _ in range(3):
        gd(*args, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        gd(*args, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    K = gd.num_timesteps
    print(f"{name:58s} {prec}: {ms:8.3f} ms/call  {B * T / ms * 1e3 / 1e6:8.2f} M frames/s  {B * T * K / ms * 1e3 / 1e6:8.2f} M frame-steps/s")

run("config 2: LJSpeech naive K=4, B=64 x T=800", "LJSpeech", "naive", False, 64, 800, "bf16", 20)
run("config 2: LJSpeech naive K=4, B=64 x T=800", "LJSpeech", "naive", False, 64, 800, "fp32", 2)
run("config 3: LJSpeech shallow K=1, B=64 x T=800", "LJSpeech", "shallow", False, 64, 800, "bf16", 40)
run("config 4: AISHELL3 shallow K=1 multi-speaker, B=32 x T=1500", "AISHELL3", "shallow", True, 32, 1500, "bf16", 40)
run("small batch: LJSpeech naive K=4, B=1 x T=800", "LJSpeech", "naive", False, 1, 800, "bf16", 20)
run("small batch: LJSpeech naive K=4, B=8 x T=800", "LJSpeech", "naive", False, 8, 800, "bf16", 20)
