#!/usr/bin/env python3
"""Per-source-line executed warp instructions of the first kernel in an .ncu-rep (needs -lineinfo + --import-source on).
usage: ncu_lines.py REPORT [min_millions]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
thr = float(sys.argv[2]) * 1e6 if len(sys.argv) > 2 else 4e6
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; sec = 0; per = {}; files = {}
for i, r in enumerate(rows):
    if not r: continue
    if r[0] == 'Line No':
        sec += 1; cur = None
        files[sec] = rows[i - 1][1] if i and len(rows[i - 1]) > 1 else '?'
        continue
    if len(r) < 8: continue
    if r[0].strip().isdigit(): cur = (sec, int(r[0]), r[1])
    if r[2].startswith('0x') and cur:
        try: n = int(r[7])
        except ValueError: continue
        per[cur] = per.get(cur, 0) + n
tot = sum(per.values())
print('total (with inlining double counts) %.1fM' % (tot / 1e6))
for k in sorted(per):
    if per[k] > thr:
        print('%-28s %5d %7.1fM %5.1f%%  %s' % (files[k[0]].split('/')[-1], k[1], per[k] / 1e6, 100 * per[k] / tot, k[2].strip()[:100]))
