#!/usr/bin/env python
"""Opcode histogram (weighted by executed warp-instructions) of one kernel in an .ncu-rep."""
import csv, io, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]   # kern: substring of the demangled name, e.g. "64, (int)1"
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = None; cur = None; per = collections.OrderedDict()
for r in rows:
    if not r: continue
    if r[0] == 'Kernel Name': cur = r[1]; per.setdefault(cur, []); continue
    if r[0] == 'Address': hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].startswith('0x'): per[cur].append(r)
iI = hdr.index('Instructions Executed'); iS = hdr.index('Source'); iSm = hdr.index('# Samples')
for k, lines in per.items():
    if kern not in k: continue
    tot = sum(int(r[iI]) for r in lines); ts = sum(int(r[iSm]) for r in lines)
    print('==', k[:100], 'static', len(lines), 'executed warp-inst', tot, 'samples', ts)
    ops = collections.Counter(); smp = collections.Counter()
    for r in lines:
        s = r[iS].strip()
        if s.startswith('@'): s = s.split(None, 1)[1]
        op = s.split()[0].split('.')[0]
        ops[op] += int(r[iI]); smp[op] += int(r[iSm])
    for op, n in ops.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 25):
        print('   %-10s %6.2f%% inst  %6.2f%% samples' % (op, 100.0 * n / max(tot, 1), 100.0 * smp[op] / max(ts, 1)))
    if '--once' in sys.argv: break
