"""Fold the ncu launch list of scripts/bench_conv1d.py: the LAST forward+backward of each layer, one line per kernel."""
import csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = [(r["Kernel Name"], r["Grid Size"], float(r["Metric Value"]) / 1000) for r in csv.DictReader(lines)]
pat = re.compile(r"(conv_gemm_f32_kernel<.>|conv_gemm_tc_kernel<[^>]*>|conv_wgrad_f32_kernel|wgrad_reduce_f32_kernel|splitk_epilogue_kernel|colsum_\w+|act_bwd_kernel|pack_w_tc_kernel|pack_w_kernel)")
lib = [(pat.search(n).group(1), g, t) for n, g, t in rows if pat.search(n)]
# split into layers x iterations at every forward conv_gemm<0>
iters, cur = [], []
for k in lib:
    if k[0] in ("pack_w_kernel", "pack_w_tc_kernel") and cur and any((c[0] == "conv_gemm_f32_kernel<0>" or c[0].startswith("conv_gemm_tc_kernel<0")) for c in cur) and any(c[0].startswith("conv_wgrad") for c in cur):
        iters.append(cur); cur = []
    cur.append(k)
iters.append(cur)
names = ["input_proj", "conv0", "conv1", "conv2", "cond0", "cond1", "mlp0", "mlp2"]
tot = 0.0
for li, it in enumerate(iters[2::3]):
    s = sum(t for _, _, t in it)
    tot += s
    print(f"{names[li] if li < len(names) else li:11s} {s:7.1f} us : " + "  ".join(f"{n.replace('_f32_kernel','').replace('_kernel','')}{g.replace(' ','')} {t:.1f}" for n, g, t in it if t > 8))
print(f"total fwd+bwd kernels {tot:.1f} us")
