#!/bin/bash
set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:tcconv_kernel --launch-skip 0 --launch-count 4 -o gpurun_out/ncu_aux_convs -f \
  python scripts/bench_stages.py aux 1 0 > gpurun_out/ncu_aux.log 2>&1
ls -la gpurun_out/ncu_aux_convs.ncu-rep
