"""Print / save the key metrics of every launch in an .ncu-rep (raw page) + the top stall sites (source page).
usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep [out.txt]"""
import collections, csv, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"]
lines = []
for i, h in enumerate(hdr):
    if h in want:
        vals = [r[i] for r in rows[2:]]
        if h == "Kernel Name":
            vals = [v.split("(")[0][-40:] for v in vals]
        lines.append(f"{h} [{rows[1][i]}] = {' | '.join(vals)}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(src.splitlines()))
# per kernel blocks: header rows start with "Kernel Name"? find header line containing '# Samples'
blocks, cur, sh = [], None, None
for r in srows:
    if "# Samples" in r or "Sampling Data (All)" in r:
        sh = r
        cur = []
        blocks.append((sh, cur))
    elif cur is not None and len(r) == len(sh):
        cur.append(r)
for bi, (sh, body) in enumerate(blocks):
    col = "# Samples" if "# Samples" in sh else "Sampling Data (All)"
    isamp, isrc = sh.index(col), sh.index("Source")
    tot = sum(int(r[isamp] or 0) for r in body if (r[isamp] or "0").isdigit())
    top = sorted((r for r in body if (r[isamp] or "0").isdigit()), key=lambda r: -int(r[isamp] or 0))[:14]
    lines.append(f"--- launch {bi}: top stall-sample sites of {tot} samples")
    for r in top:
        lines.append(f"   {100 * int(r[isamp]) / max(tot, 1):5.1f}%  {r[isrc].strip()[:110]}")
txt = "\n".join(lines)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(f"{rep}\n{txt}\n")
