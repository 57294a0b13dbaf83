#!/bin/bash
set -u
MGB_PROFILE=1 timeout 300 python - <<'PY' 2>&1 | grep -E "mgb profile|mgb timeline|Error|error" | head -12
import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from helpers import Case
from mixgan_tts_b200 import GaussianDiffusion
c = Case("LJSpeech", "naive", False, 64, 800, wseed=0, iseed=5)
gd = GaussianDiffusion(c.args, c.pc, c.mc, c.tc, precision="bf16")
gd.denoise_fn.load_state_dict({k: torch.from_numpy(v) for k, v in c.W.items()})
gd = gd.cuda().eval()
cu = lambda k: c.t(k).cuda()
t = torch.full((64,), 3, dtype=torch.long, device="cuda")
for _ in range(2):
    out = gd.denoise_fn(cu("x_T"), t, cu("cond").transpose(1, 2), None)
torch.cuda.synchronize()
PY
