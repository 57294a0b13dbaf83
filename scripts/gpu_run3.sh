#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== bench bf16 plain"
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain3.log 2>&1 || { tail -5 gpurun_out/plain3.log; exit 1; }
tail -c 600 gpurun_out/plain3.log
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv \
    --log-file gpurun_out/launches_bf16.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu3a.log 2>&1
tail -2 gpurun_out/ncu3a.log
echo "== ncu full fused kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_group_kernel -s 8 -c 2 \
    -o gpurun_out/prof_bf16 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu3b.log 2>&1
tail -2 gpurun_out/ncu3b.log
echo "== done"
