#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_l.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 102 -c 68 --csv \
    --log-file gpurun_out/launches_bf16.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
tail -1 gpurun_out/ncu_l.log | cut -c1-200
