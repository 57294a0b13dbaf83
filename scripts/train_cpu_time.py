"""Host time needed to ENQUEUE one training step vs. the device time of the step (is the step launch-bound on the CPU?)."""
import sys, time
sys.path.insert(0, ".")
import torch
from mixgan_tts_b200 import GaussianDiffusion, configs, synth
B, T = 8, 800
cfg = configs.make_configs("LJSpeech", "naive")
gd = GaussianDiffusion(*cfg, precision="bf16")
gd.denoise_fn.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_denoiser_weights(0).items()})
gd = gd.cuda().train()
opt = torch.optim.Adam(gd.denoise_fn.parameters(), lr=1e-5, fused=True)
inp, ex, pr = synth.make_inputs(77, B, T, 4), synth.make_train_extras(78, B, T, 4), synth.grad_probe(79, B, T)
to = lambda a: torch.from_numpy(a).cuda()
cond0, pad, mel, r0, r1 = to(inp["cond"]), to(inp["pad_mask"]), to(ex["mel"]), to(pr["r0"]), to(pr["r1"])
parts = {"fwd": 0.0, "bwd": 0.0, "opt": 0.0}
def step(acc=False):
    t0 = time.perf_counter()
    opt.zero_grad(set_to_none=True)
    cond = cond0.detach().requires_grad_(True)
    out = gd(mel, cond, None, pad)
    loss = (out[0] * r0).sum() + (out[3] * r1).sum()
    t1 = time.perf_counter()
    loss.backward()
    t2 = time.perf_counter()
    opt.step()
    t3 = time.perf_counter()
    if acc:
        parts["fwd"] += t1 - t0; parts["bwd"] += t2 - t1; parts["opt"] += t3 - t2
for _ in range(5):
    step()
torch.cuda.synchronize()
N = 50
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
# (a) CPU enqueue time with an empty queue in front (sync before every step)
for _ in range(N):
    torch.cuda.synchronize()
    step(acc=True)
torch.cuda.synchronize()
print("host enqueue time per step (ms):", {k: round(1e3 * v / N, 3) for k, v in parts.items()}, "total", round(1e3 * sum(parts.values()) / N, 3))
e0.record()
for _ in range(N):
    step()
e1.record(); torch.cuda.synchronize()
print("device time per step, back to back (ms):", round(e0.elapsed_time(e1) / N, 3))
