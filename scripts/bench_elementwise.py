"""bench.py's elementwise record alone (HBM GB/s of shallow_start / denorm_mask / cond_pack), for ncu or quick loops."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
ctx = bench.Ctx()
r = bench.bench_elementwise(ctx, reps=int(os.environ.get("REPS", "20")))
for k, v in r.items():
    print(k, json.dumps(v) if isinstance(v, dict) else v)
