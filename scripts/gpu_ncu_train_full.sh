#!/bin/bash
# ncu --set full of the training path's GEMM kernels: eight consecutive frames-GEMM launches and two weight-gradient launches
set -u
mkdir -p gpurun_out
if [ "${ONLY_WGEMM:-0}" != "1" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fgemm_kernel" -s ${SKIP:-560} -c ${COUNT:-8} \
    -o gpurun_out/prof_train -f python bench.py --workload train --precision bf16 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_train_full.log 2>&1
tail -1 gpurun_out/ncu_train_full.log | cut -c1-200
fi
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^wgemm_kernel$" -s ${WSKIP:-27} -c 2 \
    -o gpurun_out/prof_train_wgemm -f python bench.py --workload train --precision bf16 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_train_wgemm.log 2>&1
tail -1 gpurun_out/ncu_train_wgemm.log | cut -c1-200
