#!/bin/bash
# ncu --set full of selected tcconv / attention launches (one vocoder call, one aux call, no warm-up)
set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
# vocoder: tcconv launches 59 / 60 of the call = first resblock pair of the 32-channel stage (c1 k=3, c2 k=3 + residual)
timeout 600 $NCU -k regex:tcconv_kernel --launch-skip 59 --launch-count 2 -o gpurun_out/ncu_voc_ch32 -f \
  python scripts/bench_stages.py voc 1 0 > gpurun_out/ncu_voc.log 2>&1
tail -1 gpurun_out/ncu_voc.log
# aux decoder (pack launches come first: 6 layers x 6 + 1 + 5 = 42 tcconv-free pack kernels, so count tcconv only):
# launch 0 = QKV, 1 = fc+LN, 2 = w_1, 3 = w_2+LN
timeout 600 $NCU -k regex:tcconv_kernel --launch-skip 0 --launch-count 4 -o gpurun_out/ncu_aux_convs -f \
  python scripts/bench_stages.py aux 1 0 > gpurun_out/ncu_aux.log 2>&1
timeout 600 $NCU -k regex:attn_kernel --launch-skip 0 --launch-count 1 -o gpurun_out/ncu_aux_attn -f \
  python scripts/bench_stages.py aux 1 0 > gpurun_out/ncu_attn.log 2>&1
ls -la gpurun_out/*.ncu-rep
