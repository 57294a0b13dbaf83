#!/bin/bash
# First GPU contact: fp32 parity, smoke, bench, ncu launch list + one full capture, tcgen05 probes last.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== parity" ; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -25 | tee gpurun_out/parity.log
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench" ; timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== ncu launches"
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 273 -c 300 --csv \
    --log-file gpurun_out/launches_fp32.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
tail -3 gpurun_out/ncu1.log
echo "== ncu full (k=3 conv+gate kernel)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s 400 -c 3 \
    -o gpurun_out/prof_fp32 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
echo "== umma probe" ; timeout 300 python -m pytest tests/test_umma_probe.py -q 2>&1 | tail -30 | tee gpurun_out/probe.log
cat gpurun_out/umma_probe.txt 2>/dev/null
echo "== done"
