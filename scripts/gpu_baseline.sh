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
LINES_OUT=8 bash scripts/gpu_prof.sh 2>&1 | tee gpurun_out/role_profile.txt | cut -c1-300
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
