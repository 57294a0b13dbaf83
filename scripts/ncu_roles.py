"""Summarise an ncu source-page CSV of fused_group_kernel: stall samples per barrier wait and per role."""
import csv, re, collections, subprocess, sys
rep = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/prof_bf16.ncu-rep"
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
isrc, isamp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
first = []
for r in rows[2:]:
    if r and r[0] == 'Kernel Name': break
    first.append(r)
N = len(first)
samp = [int(r[isamp]) if r[isamp].isdigit() else 0 for r in first]
BASE = 0x38400
names = {}
for i in range(6): names[BASE + 8 * i] = f'FULL{i}'; names[BASE + 0x30 + 8 * i] = f'EMPTY{i}'
names.update({BASE + 0x60: 'TFULL0', BASE + 0x68: 'TFULL1', BASE + 0x70: 'TEMPTY0', BASE + 0x78: 'TEMPTY1', BASE + 0x80: 'AREADY',
              BASE + 0x88: 'GREADY0', BASE + 0x90: 'GREADY1', BASE + 0x98: 'GREADY2', BASE + 0xa0: 'GREADY3', BASE + 0xa8: 'SKIPDONE'})
role = ['setup'] * N
last = 'setup'
for i, r in enumerate(first):
    s = r[isrc]
    if 'UBLKCP' in s: last = 'producer'
    elif 'UTCHMMA' in s: last = 'mma'
    elif 'LDTM' in s or 'MUFU' in s: last = 'epilogue'
    role[i] = last
tot = sum(samp)
byrole = collections.Counter()
for i in range(N): byrole[role[i]] += samp[i]
print('samples by role region:', dict(byrole), 'total', tot)
loops = collections.Counter()
for i in range(N):
    m = re.search(r'TRYWAIT P\d, \[(.*?)\]', first[i][isrc])
    if m:
        off = re.search(r'0x([0-9a-f]+)', m.group(1))
        nm = names.get(int(off.group(1), 16), 'dyn') if off else 'dyn'
        loops[(role[i], nm)] += sum(samp[max(0, i - 1):i + 14])
for (ro, nm), w in loops.most_common(16):
    print(f"{w:7d} {100 * w / tot:5.1f}%  {ro:9s} wait {nm}")
print('top instructions:')
for s, i in sorted(((samp[i], i) for i in range(N)), reverse=True)[:14]:
    print(f"{s:7d} {100 * s / tot:5.1f}% #{i:5d} {role[i]:9s} {first[i][isrc][:90]}")
