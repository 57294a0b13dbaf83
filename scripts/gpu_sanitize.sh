#!/bin/bash
# compute-sanitizer over the smoke shapes: memcheck, racecheck, synccheck, initcheck.
# Run on the GPU box:  bash scripts/gpu_sanitize.sh   -> gpurun_out/sanitizer_<tool>.log (+ a summary line per tool)
set -u
mkdir -p gpurun_out
CS=${CS:-/usr/local/cuda/bin/compute-sanitizer}
for tool in ${TOOLS:-memcheck synccheck racecheck initcheck}; do
  for part in ${PARTS:-sample lr train}; do
    log=gpurun_out/sanitizer_${tool}_${part}.log
    timeout ${SAN_TIMEOUT:-900} $CS --tool $tool --print-limit 20 --error-exitcode 1 python scripts/sanitize_target.py $part > $log 2>&1
    rc=$?
    echo "== $tool / $part: exit $rc  $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1)"
    grep -E " ok " $log | sed 's/^/     /'
  done
done
