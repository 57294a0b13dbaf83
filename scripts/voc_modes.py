import sys, os, time
import numpy as np, torch
sys.path.insert(0, '/root/repo')
from mixgan_tts_b200 import Generator, synth
from oracle import hifigan as oh
rel = lambda a, b: float((a.double().cpu() - b.double()).norm() / b.double().norm())
for seed, B, T in ((5, 1, 24), (31, 2, 40), (9, 2, 64)):
    W = synth.make_hifigan_weights(seed)
    mel = torch.from_numpy(synth.make_mel(seed + 1, B, T))
    want = oh.generator_forward({k: torch.from_numpy(v) for k, v in W.items()}, mel.transpose(1, 2), synth.HIFIGAN_CFG).squeeze(1)
    for prec in ("fp16", "mixed", "fp16x3"):
        gen = Generator(synth.HIFIGAN_CFG, precision=prec)
        gen.load_state_dict({k: torch.from_numpy(v) for k, v in W.items()})
        gen = gen.cuda().eval()
        got = gen.forward_frames(mel.cuda())
        torch.cuda.synchronize()
        print(f"seed {seed} B={B} T={T} {prec:7s} rel {rel(got, want):.3e}", flush=True)
W = synth.make_hifigan_weights(0)
mel = torch.from_numpy(synth.make_mel(1, 16, 800)).cuda()
for prec in ("fp16", "mixed", "fp16x3"):
    gen = Generator(synth.HIFIGAN_CFG, precision=prec)
    gen.load_state_dict({k: torch.from_numpy(v) for k, v in W.items()})
    gen = gen.cuda().eval()
    for _ in range(2): gen.forward_frames(mel)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(4): gen.forward_frames(mel)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 4
    print(f"{prec:7s} B=16 T=800: {dt*1e3:.2f} ms  {16*800/dt/1e6:.3f} M frames/s", flush=True)
    del gen; torch.cuda.empty_cache()
