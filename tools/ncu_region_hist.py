#!/usr/bin/env python
"""Instruction / sample share per source REGION (line ranges given as name:lo-hi ...) from an .ncu-rep."""
import csv, io, subprocess, sys
rep, kern, fname_want = sys.argv[1], sys.argv[2], sys.argv[3]
regions = []
for a in sys.argv[4:]:
    n, r = a.split(":"); lo, hi = r.split("-"); regions.append((n, int(lo), int(hi)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
fname = None; hdr = None; lines = []
for r in rows:
    if not r: continue
    if r[0] in ('File Path', 'File Name'): fname = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit(): lines.append((fname, r))
iI = hdr.index('Instructions Executed'); iSm = hdr.index('# Samples')
tot = sum(int(r[iI]) for f, r in lines); ts = sum(int(r[iSm]) for f, r in lines)
acc = {n: [0, 0] for n, _, _ in regions}; other = {}
for f, r in lines:
    n_i, n_s = int(r[iI]), int(r[iSm])
    if f == fname_want:
        hit = False
        for n, lo, hi in regions:
            if lo <= int(r[0]) <= hi:
                acc[n][0] += n_i; acc[n][1] += n_s; hit = True; break
        if hit: continue
    o = other.setdefault(f, [0, 0]); o[0] += n_i; o[1] += n_s
print('total warp-inst', tot, 'samples', ts)
for n, _, _ in regions:
    print('%-22s %5.1f%% inst %5.1f%% samples' % (n, 100 * acc[n][0] / tot, 100 * acc[n][1] / max(ts, 1)))
for f, (a, b) in sorted(other.items(), key=lambda kv: -kv[1][0]):
    print('%-22s %5.1f%% inst %5.1f%% samples' % ('[' + str(f) + ']', 100 * a / tot, 100 * b / max(ts, 1)))
