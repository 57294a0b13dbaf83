#!/bin/bash
# One GPU call that regenerates the raw material of profiles/r02 (run AFTER the plain commands have passed).
set -u
mkdir -p gpurun_out
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/gputest.log
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo rc=$?
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_arm.json 2>/dev/null; echo rc=$?
echo "== role profiles"
LINES_OUT=14 bash scripts/gpu_prof.sh bf16 > gpurun_out/role_profile_bf16.txt 2>&1
LINES_OUT=14 bash scripts/gpu_prof.sh fp16 > gpurun_out/role_profile_fp16.txt 2>&1
echo "== sampling launch list + fused kernel capture"; bash scripts/gpu_ncu_both.sh
echo "== stages launch list"; bash scripts/gpu_ncu_stages.sh | tail -3
echo "== tcconv / attention captures"; bash scripts/gpu_ncu_tcconv.sh | tail -4
echo "== GAN step launch list"; MIXGAN_B200_BENCH_NO_GRAPH=1 bash scripts/gpu_ncu_gan.sh | tail -2
echo "== tcconv traces"
for n in 0 1 2 3; do MIXGAN_B200_USE_DEBUG_LIB=1 MGB_TC_TRACE=$n timeout 200 python scripts/tc_trace.py aux 2>&1 | tail -1; done > gpurun_out/tc_trace_aux.txt
MIXGAN_B200_USE_DEBUG_LIB=1 MGB_TC_TRACE=all timeout 200 python scripts/tc_trace.py 2>&1 | tail -1 > gpurun_out/tc_trace_voc.txt
cat gpurun_out/tc_trace_aux.txt gpurun_out/tc_trace_voc.txt | cut -c1-260
ls gpurun_out | head -60
echo "== discriminator Conv1d layers (warm-L2 per-kernel durations)"
timeout 300 ncu --cache-control none --metrics gpu__time_duration.sum --clock-control none -k regex:"conv_gemm|conv_wgrad|splitk|wgrad_reduce|colsum|act_bwd|pack_w" --csv --log-file gpurun_out/conv1d_launches.csv python scripts/bench_conv1d.py 16 > gpurun_out/conv1d_ncu.log 2>&1; tail -1 gpurun_out/conv1d_ncu.log
echo "== elementwise kernels"; timeout 200 python scripts/bench_elementwise.py > gpurun_out/elementwise.txt 2>&1; tail -3 gpurun_out/elementwise.txt | cut -c1-200
