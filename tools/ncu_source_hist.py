#!/usr/bin/env python
"""Per-source-line instruction / stall-sample histogram from an .ncu-rep (needs -lineinfo)."""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
fname = None; hdr = None; lines = []; seen_fn = 0
for r in rows:
    if not r: continue
    if r[0] == 'File Path': fname = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name':
        continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        lines.append((fname, r))
iI = hdr.index('Instructions Executed'); iSm = hdr.index('# Samples')
tot = sum(int(r[iI]) for f, r in lines); ts = sum(int(r[iSm]) for f, r in lines)
print('total warp-inst', tot, 'samples', ts)
agg = sorted(((int(r[iI]), int(r[iSm]), f, r[0], r[1].strip()[:100]) for f, r in lines if int(r[iI]) > 0), reverse=True)
for n, s, f, l, src in agg[:top]:
    print('%5.1f%% inst %5.1f%% smp  %-18s %4s  %s' % (100 * n / tot, 100 * s / max(ts, 1), f, l, src))
