#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== parity quick"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -3
for b in 4 8 18 64; do
  echo "== bench batch=$b"
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --batch $b 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('ms/step',round(d['ms_per_step'],3),'kernel_ms',round(r['kernel_ms'],4),'launches',r['kernel_launches_timed'],'share',round(r['kernel_share_of_step'],3),'frac',round(r['frac'],3),'value',int(d['value']),'e2e',int(d['e2e']['value']))
    else: print(l.rstrip()[:300])
"
done
echo "== done"
