#!/bin/bash
# launch list (gpu__time_duration) of the aux decoder and HiFi-GAN stages; run only after bench_stages.py exits 0 without ncu
set -u
mkdir -p gpurun_out
timeout 300 python scripts/bench_stages.py > gpurun_out/stages.json 2> gpurun_out/stages.err || { tail -5 gpurun_out/stages.err; exit 1; }
cat gpurun_out/stages.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_stages.csv \
  python scripts/bench_stages.py > gpurun_out/ncu_stages.log 2>&1
tail -2 gpurun_out/ncu_stages.log
