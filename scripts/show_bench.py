"""Pretty-print the JSON line(s) of bench.py:  python scripts/show_bench.py gpurun_out/bench_full.json"""
import json, sys
for path in sys.argv[1:]:
    for l in open(path):
        if not l.startswith("{"):
            continue
        d = json.loads(l)
        print(f"== {path}: {d.get('impl', 'ours')} n_gpus={d.get('n_gpus')} metric={d.get('metric')} dtype={d.get('dtype')}")
        print("value", int(d["value"]), "ms/step", round(d["ms_per_step"], 3), "e2e", int(d["e2e"]["value"]) if d.get("e2e") else None,
              "launches", d.get("gpu_launches"))
        r = d.get("roofline")
        if r:
            print("  roofline: kernel_ms", round(r["kernel_ms"], 4), "[", round(r.get("kernel_ms_min", 0), 4), round(r.get("kernel_ms_max", 0), 4), "] n",
                  r["kernel_launches_timed"], "share", r["kernel_share_of_step"], "achieved", round(r["achieved"], 1), "peak", r["peak"],
                  "frac", round(r["frac"], 3), "sm_mhz", r.get("sm_mhz_in_kernel"), "cycle-frac", r.get("frac_of_cycle_peak_at_kernel_clock"))
        if d.get("clocks"):
            print("  clocks", d["clocks"])
        for k, v in (d.get("sub") or {}).items():
            if v is None:
                continue
            if k == "elementwise":
                for kk, vv in v.items():
                    if isinstance(vv, dict):
                        print(f"  elementwise {kk}: {vv['ms'] * 1e3:.1f} us, {vv['gbs']:.0f} GB/s, frac {vv['frac']:.2f}")
            elif "value" not in v:           # a record of records (stages)
                for kk, vv in v.items():
                    if isinstance(vv, dict) and "value" in vv:
                        print(f"  sub {k}.{kk}: value {int(vv['value'])} {vv.get('unit', '')} frac {(vv.get('roofline') or {}).get('frac')}")
            else:
                rr = v.get("roofline", {})
                print(f"  sub {k}: value {int(v['value'])} ms {v['ms_per_step']:.3f} e2e {int(v['e2e']['value']) if 'e2e' in v else None} "
                      f"frac {rr.get('frac')} mhz {rr.get('sm_mhz_in_kernel')} dp_check {v.get('dp_check')} exposed {v.get('allreduce_exposed_ms_per_step')}")
        if d.get("cpu_baseline"):
            print("  cpu_baseline", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], "cores")
