#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_ncu.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_pair_kernel -s 8 -c 1 \
    -o gpurun_out/prof_pair -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fused.log 2>&1
tail -2 gpurun_out/ncu_fused.log | cut -c1-300
