#!/bin/bash
# ncu launch list of one training step (bench.py --workload train): per-kernel durations, cold-cache and serialised.
set -u
mkdir -p gpurun_out
PREC=${PREC:-bf16}
timeout 200 python bench.py --workload train --precision $PREC --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/train_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-1000} -c ${COUNT:-900} --csv \
    --log-file gpurun_out/launches_train_$PREC.csv python bench.py --workload train --precision $PREC --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_train.log 2>&1
python - <<'PY'
import csv, collections, os
prec = os.environ.get("PREC", "bf16")
rows = list(csv.reader(open(f"gpurun_out/launches_train_{prec}.csv")))
for i, r in enumerate(rows):
    if r and r[0] == "ID":
        hdr, start = r, i + 1
        break
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
body = [r for r in rows[start:] if len(r) > iv]
# one complete training step = from the first weight-pack launch of a forward to the next forward's first weight-pack launch
marks = [i for i, r in enumerate(body) if "pack_tiles_kernel" in r[ik] and (i == 0 or "pack_tiles_kernel" not in body[i - 1][ik])]
if len(marks) >= 2:
    body = body[marks[0]:marks[1]]
    print(f"one training step: launches {marks[0]}..{marks[1]} of the captured window")
tot = collections.OrderedDict()
for r in body:
    n = r[ik].split("(")[0].split("::")[-1][:60]
    tot.setdefault(n, [0.0, 0]); tot[n][0] += float(r[iv].replace(",", "")); tot[n][1] += 1
T = sum(v[0] for v in tot.values())
for n, v in sorted(tot.items(), key=lambda kv: -kv[1][0])[:28]:
    print(f"{v[0]/1e3:9.1f} us {v[1]:4d}x {100*v[0]/T:5.1f}%  {n}")
print(f"{T/1e3:9.1f} us total over {sum(v[1] for v in tot.values())} launches")
PY
