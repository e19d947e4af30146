#!/usr/bin/env python3
"""Executed warp instructions of the search kernel by source region of mb_search.cu (innermost line attribution).
usage: ncu_regions.py REPORT  — regions are found from marker comments / function names in the current source, so run it
on a report of the current build."""
import csv, io, subprocess, sys, re
rep = sys.argv[1]
src = open('molar_b200/csrc/mb_search.cu').read().split('\n')
def line_of(pat, start=0):
    for i in range(start, len(src)):
        if pat in src[i]: return i + 1
    raise SystemExit('marker not found: ' + pat)
marks = [
    ('helpers', 1),
    ('flush', line_of('__device__ __forceinline__ void flush_chunk_copy')),
    ('fp helpers (tests)', line_of('// ---- packed f32x2 helpers')),
    ('row helpers', line_of('__device__ __forceinline__ int div_k')),
    ('fill_run_table', line_of('__device__ __forceinline__ void fill_run_table')),
    ('scan_step', line_of('__device__ __forceinline__ void scan_step')),
    ('emit_group', line_of('__device__ __forceinline__ void emit_group')),
    ('kernel prologue / tile', line_of('search_cells_kernel(const __grid_constant__ SearchParams P)')),
    ('home batch load', line_of('for (unsigned hb = hs; hb < he; hb += 32)')),
    ('cursor / fetch', line_of('auto fetch = [&]')),
    ('fused: fetch + flags', line_of('// ---- fused double step')),
    ('fused: tests', line_of('const unsigned long long nrc22 = pk2_once(-rc2, -rc2);', line_of('// ---- fused double step'))),
    ('fused: masks / scan / setup', line_of('fused = !__any_sync(0xffffffffu, tmin <= P.band);')),
    ('fused: emission loop', line_of('uint4 hqa = lds128u(ha), hqb = lds128u(ha + 16u);', line_of('// ---- fused double step'))),
    ('single step: fetch', line_of('runs_before = rb_save;')),
    ('step flags / pack', line_of('unsigned m0 = 0, m1 = 0;', line_of('runs_before = rb_save;'))),
    ('direct tests', line_of('} else if (!any_wrapped) {')),
    ('mixed tests', line_of('// ---- mixed step: wrapped cell pairs')),
    ('count / scan (single step)', line_of('if (MODE == 3) {', line_of('// ---- mixed step'))),
    ('emission (single step)', line_of('const bool single = tot <= STAGE_CAP;', line_of('// inclusive scan of the hits per lane'))),
    ('after emission', line_of('stage_n += need;')),
    ('end', line_of('// general all-pairs kernel')),
]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
sec = 0; cur = None; per = {}
for r in rows:
    if not r: continue
    if r[0] == 'Line No': sec += 1; cur = None; continue
    if len(r) < 8: continue
    if r[0].strip().isdigit(): cur = (sec, int(r[0]))
    if r[2].startswith('0x') and cur:
        try: n = int(r[7])
        except ValueError: continue
        per.setdefault(cur, {})[r[2]] = n
# the section of mb_search.cu: the one with the most distinct lines above 600
cnt = {}
for (s, l) in per:
    if l > 600: cnt[s] = cnt.get(s, 0) + 1
ms = max(cnt, key=cnt.get)
best = {}; alladdr = {}
for (s, l), v in per.items():
    for a, n in v.items():
        alladdr[a] = n
        if s == ms and (a not in best or l > best[a][0]): best[a] = (l, n)
tot = sum(alladdr.values())
print('total %.1fM, attributed to mb_search.cu lines %.1fM' % (tot / 1e6, sum(n for l, n in best.values()) / 1e6))
for i in range(len(marks) - 1):
    lo, hi = marks[i][1], marks[i + 1][1]
    c = sum(n for l, n in best.values() if lo <= l < hi)
    print('%-28s lines %4d-%4d %8.1fM %5.1f%%' % (marks[i][0], lo, hi - 1, c / 1e6, 100 * c / tot))
