#!/bin/bash
# In-kernel role profile (debug library, MGB_PROFILE=1) of one full sampling call (KUNI kernels) at B=64, T=800.
# usage: scripts/gpu_prof.sh [bf16|fp16]
set -u
PREC=${1:-bf16}
MIXGAN_B200_USE_DEBUG_LIB=1 MGB_PROFILE=1 PREC=$PREC timeout 300 python - <<'PY' 2>&1 | grep -E "mgb profile|mgb timeline|mgb boundary|mgb ring|mgb tile|Error|error" | tail -${LINES_OUT:-6}
import os, sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from helpers import Case
from mixgan_tts_b200 import GaussianDiffusion
B = 64
c = Case("LJSpeech", "naive", False, B, 800, wseed=0, iseed=5)
gd = GaussianDiffusion(c.args, c.pc, c.mc, c.tc, precision=os.environ["PREC"])
gd.denoise_fn.load_state_dict({k: torch.from_numpy(v) for k, v in c.W.items()})
gd = gd.cuda().eval()
cu = lambda k: c.t(k).cuda()
out = gd(None, cu("cond"), None, cu("pad_mask"), x_T=cu("x_T"), noises=cu("noises"))[0]
torch.cuda.synchronize()
PY
