"""Timing of the stages either side of the diffusion decoder (aux decoder, HiFi-GAN) on one GPU: CUDA events, warm-up,
rotating inputs.  Prints one JSON line per stage."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mixgan_tts_b200 import AuxDecoder, Generator, _lib, configs, synth  # noqa: E402

AUX_FLOP_PER_FRAME_FIXED = 6 * (4 * 2 * 256 * 256 + 2 * 256 * 1024 * 9 + 2 * 1024 * 256) + 2 * 256 * 80 \
    + 2 * 5 * (80 * 512 + 3 * 512 * 512 + 512 * 80)          # linears + FFN convs + mel_linear + PostNet (per frame)


def aux_flops(T, lens):
    att = sum(6 * 2 * 2 * 2 * 128 * float(l) * float(l) for l in lens)      # 6 layers x 2 heads x (QK^T + PV) over valid keys/queries
    return AUX_FLOP_PER_FRAME_FIXED * len(lens) * T + att


def voc_flops_per_frame(cfg=synth.HIFIGAN_CFG):
    C0, f, rate = cfg["upsample_initial_channel"], 0.0, 1
    f += 2 * cfg["num_mels"] * C0 * 7
    for i, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        cin, ch = C0 >> i, C0 >> (i + 1)
        f += 2 * cin * ch * k * rate            # each input sample meets every tap once
        rate *= u
        f += rate * sum(2 * ch * ch * kk * 6 for kk in cfg["resblock_kernel_sizes"])
    f += rate * 2 * (C0 >> len(cfg["upsample_rates"])) * 7
    return f


def timed(fn, steps, warmup):
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def bench_aux(B=64, T=800, steps=10, warmup=3):
    _, pc, mc, _ = configs.make_configs("LJSpeech", "shallow")
    m = AuxDecoder(pc, mc)
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synth.make_auxdec_weights(0).items()})
    m = m.cuda().eval()
    sets = []
    for i in range(3):
        inp = synth.make_auxdec_inputs(100 + i, B, T, min_len_frac=0.5)
        sets.append((torch.from_numpy(inp["x"]).cuda(), torch.from_numpy(inp["lens"]).cuda(), inp["lens"]))
    lib = _lib.load()
    n0 = lib.mgb_launch_count()
    ms = timed(lambda i: m(sets[i % 3][0], lens=sets[i % 3][1]), steps, warmup)
    launches = (lib.mgb_launch_count() - n0) / (steps + warmup)
    fl = np.mean([aux_flops(T, s[2]) for s in sets])
    return {"stage": "aux_decoder", "B": B, "T": T, "ms": ms, "frames_per_s": B * T / ms * 1e3, "tflops": fl / ms / 1e9,
            "launches_per_call": launches, "flop_per_frame_fixed": AUX_FLOP_PER_FRAME_FIXED}


def bench_voc(B=16, T=800, steps=5, warmup=2):
    gen = Generator(synth.HIFIGAN_CFG)
    gen.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_hifigan_weights(0).items()})
    gen = gen.cuda().eval()
    mels = [torch.from_numpy(synth.make_mel(200 + i, B, T)).cuda() for i in range(2)]
    lib = _lib.load()
    n0 = lib.mgb_launch_count()
    ms = timed(lambda i: gen.forward_frames(mels[i % 2]), steps, warmup)
    launches = (lib.mgb_launch_count() - n0) / (steps + warmup)
    fpf = voc_flops_per_frame()
    return {"stage": "hifigan", "B": B, "T": T, "ms": ms, "frames_per_s": B * T / ms * 1e3, "tflops": fpf * B * T / ms / 1e9,
            "audio_seconds_per_s": B * T * 256 / 22050 / ms * 1e3, "launches_per_call": launches, "flop_per_frame": fpf}


if __name__ == "__main__":
    # usage: bench_stages.py [aux|voc|both] [steps] [warmup]
    torch.cuda.set_device(0)
    what = sys.argv[1] if len(sys.argv) > 1 else "both"
    kw = {}
    if len(sys.argv) > 2:
        kw["steps"] = int(sys.argv[2])
    if len(sys.argv) > 3:
        kw["warmup"] = int(sys.argv[3])
    if what in ("aux", "both"):
        print(json.dumps(bench_aux(**kw)), flush=True)
    if what in ("voc", "both"):
        print(json.dumps(bench_voc(**kw)), flush=True)
