"""Debug: per-role SM-cycle timeline of the tensor-core conv kernel (CTA 0, first 64 work items of the LAST launch of a call).
Run with MIXGAN_B200_USE_DEBUG_LIB=1 MGB_TC_TRACE=1."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mixgan_tts_b200 import Generator, synth  # noqa: E402

import numpy as np
from mixgan_tts_b200 import AuxDecoder, configs
if len(sys.argv) > 1 and sys.argv[1] == "aux":
    _, pc, mc, _ = configs.make_configs("LJSpeech", "shallow")
    m = AuxDecoder(pc, mc)
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synth.make_auxdec_weights(0).items()})
    m = m.cuda().eval()
    inp = synth.make_auxdec_inputs(1, 64, 800, min_len_frac=0.5)
    m(torch.from_numpy(inp["x"]).cuda(), torch.from_numpy(inp["pad_mask"]).cuda())     # ONE call: run_conv calls 0.. = QKV, fc+LN, w_1, w_2+LN, ...
    torch.cuda.synchronize()
    ws = next(iter(m._ws.values()))
else:
    gen = Generator(synth.HIFIGAN_CFG)
    gen.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_hifigan_weights(0).items()})
    gen = gen.cuda().eval()
    mel = torch.from_numpy(synth.make_mel(1, 16, 800)).cuda()
    gen.forward_frames(mel)
    torch.cuda.synchronize()
    ws = next(iter(gen._ws.values()))
tr = ws[1024:1024 + 32 * 16 * 8].view(torch.int64).cpu().reshape(32, 16)
t0 = int(tr[0, 0])
names = ["prod_first_load", "mma_before_acc_empty", "mma_start", "mma_committed", "epi_before_wait", "epi_start", "epi_done", "-", "cg0_ld", "cg0_math", "cg0_stored", "cg1_ld", "cg1_math", "cg1_stored"]
print("cols: prod mma_pre mma_start mma_commit epi_pre epi_start cg0_ld cg0_math cg0_stored cg1_ld cg1_math cg1_stored epi_done"); print("work item: " + "  ".join(names) + "   (cycles since the producer's first load; last launch = conv_post, 7 taps, N=32)")
for j in range(12):
    print(j, "  ".join(f"{int(tr[j, k]) - t0:8d}" for k in (0,1,2,3,4,5,8,9,10,11,12,13,6)))
nz = int((tr[:, 3] != 0).sum())
d = tr[2:max(4, nz - 1)]
print(f"{nz} items recorded; mean per-item deltas: mma wait acc_empty", float((d[:, 2] - d[:, 1]).float().mean()),
      "| mma issue", float((d[:, 3] - d[:, 2]).float().mean()), "| commit->epi start", float((d[:, 5] - d[:, 3]).float().mean()),
      "| epi body", float((d[:, 6] - d[:, 5]).float().mean()), "| item period", float((d[1:, 3] - d[:-1, 3]).float().mean()))
