#!/bin/bash
# ncu launch list of one bench run + one full capture of the fused kernel (second group launch of a call)
set -u
mkdir -p gpurun_out
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_l.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 58 -c 40 --csv \
    --log-file gpurun_out/launches_bf16.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
tail -2 gpurun_out/launches_bf16.csv | cut -c1-160
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_pair_kernel -s 8 -c 2 \
    -o gpurun_out/prof_pair -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fused.log 2>&1
tail -2 gpurun_out/ncu_fused.log | cut -c1-200
