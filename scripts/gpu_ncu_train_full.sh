#!/bin/bash
# ncu --set full of the training path's GEMM kernels (one launch each of the conv forward, conv^T backward, weight grad)
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fgemm_kernel|wgemm_kernel" -s ${SKIP:-400} -c ${COUNT:-12} \
    -o gpurun_out/prof_train -f python bench.py --workload train --precision bf16 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_train_full.log 2>&1
tail -2 gpurun_out/ncu_train_full.log | cut -c1-200
