import sys, cProfile, pstats, io
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
def step():
    opt.zero_grad(set_to_none=True)
    cond = cond0.detach().requires_grad_(True)
    out = gd(mel, cond, None, pad)
    ((out[0] * r0).sum() + (out[3] * r1).sum()).backward()
    opt.step()
for _ in range(5): step()
torch.cuda.synchronize()
pr_ = cProfile.Profile(); pr_.enable()
for _ in range(50): step()
torch.cuda.synchronize(); pr_.disable()
s = io.StringIO(); pstats.Stats(pr_, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
