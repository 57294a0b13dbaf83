"""Turn the scratch artefacts of scripts/gpu_profiles_r02.sh (gpurun_out/) into the tracked summaries under profiles/r02/.
Run here (no GPU needed): python scripts/summarize_r02.py"""
import collections, csv, json, os, re, shutil, subprocess, sys

G, out = "gpurun_out", os.path.join("profiles", "r02")
os.makedirs(out, exist_ok=True)


def short(n):
    n = re.sub(r"\(.*", "", n).replace("void ", "").replace("mgb::", "").replace("<unnamed>::", "").replace("tcnet::", "")
    return n[:72]


def read_launches(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    return [(r["Kernel Name"], r["Grid Size"], float(r["Metric Value"].replace(",", ""))) for r in csv.DictReader(lines)]


def write_agg(f, seg, by_grid=False):
    agg = collections.OrderedDict()
    for n, g, t in seg:
        k = (short(n), g) if by_grid else (short(n), "")
        agg.setdefault(k, [0, 0.0])
        agg[k][0] += 1
        agg[k][1] += t
    tot = sum(v[1] for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{v[1] / 1e3:10.1f} us x{v[0]:4d} {100 * v[1] / tot:5.1f}%  {k[0]} {k[1]}\n")
    f.write(f"{tot / 1e3:10.1f} us total over {len(seg)} launches\n")


# ---- bench lines
for src, dst in (("bench_1gpu.json", "bench_1gpu.json"), ("bench_2gpu.json", "bench_2gpu.json"), ("bench_reference_arm.json", "bench_reference_arm.json"),
                 ("stages.json", "bench_stages_standalone.json"), ("role_profile_bf16.txt", "role_profile_bf16.txt"),
                 ("role_profile_fp16.txt", "role_profile_fp16.txt"), ("tc_trace_aux.txt", "tcconv_role_trace_aux.txt"),
                 ("tc_trace_voc.txt", "tcconv_role_trace_vocoder.txt"), ("gputest.log", "gpu_test_suite.log")):
    p = os.path.join(G, src)
    if os.path.exists(p):
        shutil.copy(p, os.path.join(out, dst))

# ---- launch lists
p = os.path.join(G, "launches_bf16.csv")
if os.path.exists(p):
    L = read_launches(p)
    with open(os.path.join(out, "launches_sampling_summary.txt"), "w") as f:
        f.write("ncu --metrics gpu__time_duration.sum --clock-control none -s 58 -c 40 python bench.py --steps 2 --warmup 3 --no-cpu-baseline\n"
                "(40 consecutive launches inside the sampling loop, B=64 x T=800, bf16; cold-cache, serialised: compare SHARES)\n\n")
        write_agg(f, L)
p = os.path.join(G, "launches_stages.csv")
if os.path.exists(p):
    L = read_launches(p)
    idx = [i for i, r in enumerate(L) if "pack_rows" in r[0]]
    with open(os.path.join(out, "launches_stages_summary.txt"), "w") as f:
        f.write("ncu --metrics gpu__time_duration.sum --clock-control none python scripts/bench_stages.py\n"
                "(one aux-decoder call at B=64 x T=800 and one HiFi-GAN call at B=16 x T=800; cold-cache, serialised)\n\n")
        # aux calls have 38 launches between pack_rows, vocoder calls 79
        gaps = [(idx[k + 1] - idx[k], idx[k]) for k in range(len(idx) - 1)]
        aux = [s for n, s in gaps if n == 38]
        voc = [s for n, s in gaps if n == 79]
        if aux:
            f.write("== aux decoder, one call (38 launches)\n")
            write_agg(f, L[aux[-1]:aux[-1] + 38], by_grid=True)
        if voc:
            f.write("\n== HiFi-GAN, one call (79 launches)\n")
            write_agg(f, L[voc[-1]:voc[-1] + 79], by_grid=True)
p = os.path.join(G, "launches_train_gan.csv")
if os.path.exists(p):
    txt = subprocess.run([sys.executable, "scripts/gan_step_summary.py", p, "60"], capture_output=True, text=True).stdout
    if txt.strip():
        with open(os.path.join(out, "launches_train_gan_summary.txt"), "w") as f:
            f.write("ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --workload train --steps 1 --warmup 3 (eager)\n"
                    "(the last eager GAN training step of train.py:126-184, B=8 x T=800; cold-cache, serialised: compare SHARES)\n\n" + txt)
p = os.path.join(G, "conv1d_launches.csv")
if os.path.exists(p):
    txt = subprocess.run([sys.executable, "scripts/conv1d_ncu_summary.py", p], capture_output=True, text=True).stdout
    if txt.strip():
        with open(os.path.join(out, "conv1d_f32_layers_summary.txt"), "w") as f:
            f.write("ncu --cache-control none --metrics gpu__time_duration.sum --clock-control none python scripts/bench_conv1d.py 16\n"
                    "(forward + backward of every JCU-discriminator layer at the batched training shape, 16 utterances x T=800; warm L2;\n"
                    " per layer: total kernel time, then every kernel above 8 us with its grid)\n\n" + txt)
p = os.path.join(G, "elementwise.txt")
if os.path.exists(p):
    shutil.copy(p, os.path.join(out, "elementwise_hbm.txt"))

# ---- ncu --set full captures
def ncu_summary(rep, dst, header):
    rep = os.path.join(G, rep)
    if not os.path.exists(rep):
        return
    txt = subprocess.run([sys.executable, "scripts/ncu_summary.py", rep], capture_output=True, text=True).stdout
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    sass = {k: sum(1 for l in src.splitlines() if k in l) for k in ("UTCHMMA", "UTMALDG", "UBLKCP", "LDTM", "UTCBAR", "SYNCS", "STL", "LDL")}
    with open(os.path.join(out, dst), "w") as f:
        f.write(header + "\n" + txt + f"\nSASS mnemonics in the captured kernels (static): {sass}\n")


ncu_summary("prof_pair.ncu-rep", "ncu_fused_pair_kernel_summary.txt",
            "ncu --set full --clock-control none --import-source on -k regex:fused_pair_kernel -s 8 -c 2 python bench.py --steps 2 --warmup 3\n"
            "(launch 1 = layers [0,10) of a Denoiser call, launch 2 = layers [10,20) + tail; cold caches between replays)")
ncu_summary("ncu_aux_convs.ncu-rep", "ncu_tcconv_aux_summary.txt",
            "ncu --set full --clock-control none --import-source on -k regex:tcconv_kernel --launch-skip 0 --launch-count 4 python scripts/bench_stages.py aux 1 0\n"
            "(first FFT block at B=64 x T=800: QKV GEMM (N=768), fc + residual + LayerNorm, conv k=9 256->1024 + ReLU, conv k=1 1024->256 + residual + LayerNorm)")
ncu_summary("ncu_aux_attn.ncu-rep", "ncu_attention_summary.txt",
            "ncu --set full --clock-control none --import-source on -k regex:attn_kernel --launch-count 1 python scripts/bench_stages.py aux 1 0\n"
            "(self-attention of the first FFT block, 64 utterances x 2 heads x 7 query tiles, ragged lengths)")
ncu_summary("ncu_voc_ch32.ncu-rep", "ncu_tcconv_vocoder_ch32_summary.txt",
            "ncu --set full --clock-control none --import-source on -k regex:tcconv_kernel --launch-skip 59 --launch-count 2 python scripts/bench_stages.py voc 1 0\n"
            "(HiFi-GAN 32-channel stage at B=16 x T=800, first resblock pair: c1 k=3 and c2 k=3 + residual; 3.3 M rows)")
rep = os.path.join(G, "prof_pair.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per = [float(r[ir]) * unit[rows[1][ir]] + float(r[iw]) * unit[rows[1][iw]] for r in rows[2:]]
    json.dump({"kernel": "fused_pair_kernel", "dram_bytes_per_launch": per, "mean_bytes_per_launch": sum(per) / len(per),
               "source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, cold caches (workload B=64 x T=800)"},
              open(os.path.join(out, "ncu_traffic.json"), "w"), indent=1)
print("wrote", sorted(os.listdir(out)))
