"""Warm per-kernel durations of one training step (torch.profiler / CUPTI), as a cross-check of ncu's cold-cache list."""
import sys, os, collections
sys.path.insert(0, ".")
import torch
from torch.profiler import profile, ProfilerActivity
from mixgan_tts_b200 import GaussianDiffusion, configs, synth

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
B, T = 8, 800
cfg = configs.make_configs("LJSpeech", "naive")
gd = GaussianDiffusion(*cfg, precision=prec)
gd.denoise_fn.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_denoiser_weights(0).items()})
gd = gd.cuda().train()
opt = torch.optim.Adam(gd.denoise_fn.parameters(), lr=1e-5, fused=True)
inp, ex, pr = synth.make_inputs(77, B, T, 4), synth.make_train_extras(78, B, T, 4), synth.grad_probe(79, B, T)
to = lambda a: torch.from_numpy(a).cuda()
cond0, pad, mel, r0, r1 = to(inp["cond"]), to(inp["pad_mask"]), to(ex["mel"]), to(pr["r0"]), to(pr["r1"])

def step():
    opt.zero_grad(set_to_none=True)
    cond = cond0.detach().requires_grad_(True)
    out = gd(mel, cond, None, pad)
    ((out[0] * r0).sum() + (out[3] * r1).sum()).backward()
    opt.step()

for _ in range(3):
    step()
torch.cuda.synchronize()
N = 5
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as p:
    for _ in range(N):
        step()
    torch.cuda.synchronize()
tot = collections.defaultdict(lambda: [0.0, 0])
for e in p.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        import re
        m = re.search(r"(\w+_kernel(<[^>]*>)?)", e.name)
        n = m.group(1) if m else e.name.split("(")[0].split("::")[-1][:60]
        tot[n][0] += e.device_time if hasattr(e, "device_time") else e.cuda_time
        tot[n][1] += 1
T_all = sum(v[0] for v in tot.values())
for n, v in sorted(tot.items(), key=lambda kv: -kv[1][0])[:24]:
    print(f"{v[0] / N:9.1f} us/step {v[1] // N:4d}x {v[0] / max(v[1], 1):7.2f} us each {100 * v[0] / T_all:5.1f}%  {n}")
print(f"{T_all / N:9.1f} us of kernels per step")
