"""Diagnostic (not a test): the tensor-core engine stage by stage against the CPU oracle, printing relative errors."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mixgan_tts_b200 import AuxDecoder, Generator, _lib, configs, synth  # noqa: E402
from oracle import aux_decoder as oa, hifigan as oh  # noqa: E402

rel = lambda a, b: float((a.double().cpu() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def voc(B, T, cfg=None, seed=1):
    cfg = dict(synth.HIFIGAN_CFG, **(cfg or {}))
    gen = Generator(cfg)
    W = synth.make_hifigan_weights(seed, cfg)
    gen.load_state_dict({k: torch.from_numpy(v) for k, v in W.items()})
    gen = gen.cuda().eval()
    mel = torch.from_numpy(synth.make_mel(seed + 1, B, T))
    want = oh.generator_forward({k: torch.from_numpy(v) for k, v in W.items()}, mel.transpose(1, 2), cfg).squeeze(1)
    got = gen.forward_frames(mel.cuda())
    torch.cuda.synchronize()
    st = C.c_int(-1)
    _lib.load().mgb_hifigan_debug_status(C.byref(gen.dims), B, T, _lib.ptr(next(iter(gen._ws.values()))), C.byref(st))
    print(f"hifigan B={B} T={T} ups={cfg['upsample_rates']}: rel {rel(got, want):.3e}  status {st.value}  "
          f"nan {int(torch.isnan(got).sum())}", flush=True)


def aux(B, T, layers=6, seed=1):
    _, pc, mc, _ = configs.make_configs("LJSpeech", "shallow")
    mc["transformer"]["decoder_layer"] = layers
    m = AuxDecoder(pc, mc)
    W = synth.make_auxdec_weights(seed, {"layers": layers})
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in W.items()})
    m = m.cuda().eval()
    inp = synth.make_auxdec_inputs(seed + 1, B, T)
    x, pad = torch.from_numpy(inp["x"]), torch.from_numpy(inp["pad_mask"])
    want = oa.aux_decoder_forward({k: torch.from_numpy(np.asarray(v)) for k, v in W.items()}, x, pad)
    got = m(x.cuda(), pad.cuda(), return_intermediate=True)
    torch.cuda.synchronize()
    st = C.c_int(-1)
    _lib.load().mgb_auxdec_debug_status(C.byref(m.dims), B, T, _lib.ptr(next(iter(m._ws.values()))), C.byref(st))
    print(f"auxdec B={B} T={T} layers={layers}: coarse {rel(got[0], want[0]):.3e}  dec {rel(got[1], want[1]):.3e}  "
          f"mel {rel(got[2], want[2]):.3e}  status {st.value}", flush=True)


if __name__ == "__main__":
    what = sys.argv[1:] or ["voc", "aux"]
    if "voc" in what:
        voc(1, 8, {"upsample_rates": [8], "upsample_kernel_sizes": [16], "upsample_initial_channel": 512})
        voc(1, 8, {"upsample_rates": [8, 8], "upsample_kernel_sizes": [16, 16]})
        voc(2, 40)
    if "aux" in what:
        aux(2, 100, layers=1)
        aux(2, 300, layers=1)
        aux(3, 150)
