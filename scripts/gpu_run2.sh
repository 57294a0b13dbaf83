#!/bin/bash
set -u
mkdir -p gpurun_out
for g in 20 10 5; do
  echo "== diag MGB_GROUP_LAYERS=$g"
  MGB_GROUP_LAYERS=$g timeout 300 python scripts/bf16_diag.py 2>&1 | tail -25
done | tee gpurun_out/bf16_diag.log
echo "== parity (both precisions)"
timeout 900 python -m pytest tests/test_gpu_parity.py -q 2>&1 | tail -30 | tee gpurun_out/parity2.log
echo "== bench bf16"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; tail -c 2500 gpurun_out/bench_bf16.json; tail -3 gpurun_out/bench_bf16.err
echo "== done"
