#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python bench.py --workload train --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/train_gan.json 2> gpurun_out/train_gan.err || { tail -5 gpurun_out/train_gan.err; exit 1; }
cat gpurun_out/train_gan.json | cut -c1-600
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/launches_train_gan.csv \
  python bench.py --workload train --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gan.log 2>&1
tail -2 gpurun_out/ncu_gan.log | cut -c1-300
wc -l gpurun_out/launches_train_gan.csv
