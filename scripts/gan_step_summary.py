"""Fold the ncu launch list of `bench.py --workload train` (scripts/gpu_ncu_gan.sh) into one eager GAN step: per-kernel totals,
library / torch split, discriminator / Denoiser split.   python scripts/gan_step_summary.py gpurun_out/launches_train_gan.csv"""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = [(r["Kernel Name"], r["Grid Size"], float(r["Metric Value"]) / 1000) for r in csv.DictReader(lines)]
adam = [i for i, r in enumerate(rows) if "FusedOptimizerTensorListMetadata<4>" in r[0]]
# an eager GAN step ends with the generator's Adam launches; the previous step's end is the Adam group before the
# discriminator's one
groups = []
for i in adam:
    if groups and i - groups[-1][-1] == 1:
        groups[-1].append(i)
    else:
        groups.append([i])
end = groups[-1][-1] + 1
start = groups[-3][-1] + 1
seg = rows[start:end]


def short(n):
    n = re.sub(r"mgb::|<unnamed>::|void |at::native::|at::", "", n)
    return re.sub(r"\(.*", "", n)[:78]


DISC = ("conv_gemm_f32", "conv_gemm_tc", "pack_w_tc", "conv_wgrad_f32", "wgrad_reduce_f32", "splitk_epilogue", "colsum_part", "colsum_final", "act_bwd", "pack_w_kernel", "step_embedding")
agg = collections.OrderedDict()
for n, g, t in seg:
    k = short(n)
    agg.setdefault(k, [0, 0.0])
    agg[k][0] += 1
    agg[k][1] += t
tot = sum(t for _, _, t in seg)
print(f"one eager GAN step: {len(seg)} launches, {tot:.1f} us of kernels (serialised)")
lib = [(n, t) for n, _, t in seg if "mgb::" in n]
disc = [(n, t) for n, t in lib if any(d in n for d in DISC)]
print(f"  library {len(lib)} launches {sum(t for _, t in lib):.1f} us  [discriminator {len(disc)} / {sum(t for _, t in disc):.1f} us, "
      f"Denoiser {len(lib) - len(disc)} / {sum(t for _, t in lib) - sum(t for _, t in disc):.1f} us];  torch {len(seg) - len(lib)} launches {tot - sum(t for _, t in lib):.1f} us")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{v[1]:9.1f} us x{v[0]:4d} {100 * v[1] / tot:5.1f}%  {k}")
