#!/bin/bash
# Full picture of the current build on one B200: parity, both bench arms, in-kernel role profile,
# ncu launch list and one full ncu capture of the fused kernel.  Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader > gpurun_out/gpu.txt
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "== bench (ours)"; timeout 600 python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; tail -c 2500 gpurun_out/bench_ours.json
echo "== bench (reference arm)"; timeout 600 python bench.py --impl reference --steps 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
echo "== role profile"
MGB_PROFILE=1 timeout 300 python - <<'PY' 2>&1 | grep -E "mgb profile|mgb timeline|mgb boundary|mgb ring|mgb tile|Error|error" | tail -8 | tee gpurun_out/role_profile.txt
import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from helpers import Case
from mixgan_tts_b200 import GaussianDiffusion
B = 64
c = Case("LJSpeech", "naive", False, B, 800, wseed=0, iseed=5)
gd = GaussianDiffusion(c.args, c.pc, c.mc, c.tc, precision="bf16")
gd.denoise_fn.load_state_dict({k: torch.from_numpy(v) for k, v in c.W.items()})
gd = gd.cuda().eval()
cu = lambda k: c.t(k).cuda()
t = torch.full((B,), 3, dtype=torch.long, device="cuda")
for _ in range(2):
    out = gd.denoise_fn(cu("x_T"), t, cu("cond").transpose(1, 2), None)
torch.cuda.synchronize()
PY
echo "== ncu launch list"
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_l.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_bf16.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
tail -3 gpurun_out/launches_bf16.csv | cut -c1-200
echo "== ncu full (fused kernel)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_pair_kernel -s 8 -c 2 \
    -o gpurun_out/prof_pair -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fused.log 2>&1
tail -2 gpurun_out/ncu_fused.log | cut -c1-300
echo "== done"
