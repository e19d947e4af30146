#!/usr/bin/env python3
"""One line per kernel launch of an .ncu-rep: duration, DRAM bytes, achieved DRAM GB/s, issue utilisation, registers,
occupancy, top stall reasons.  usage: ncu_multi.py REPORT..."""
import csv, io, re, subprocess, sys
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, data = rows[0], rows[2:]
    def col(r, k, d=0.0):
        try: return float(r[h.index(k)].replace(',', ''))
        except Exception: return d
    u = rows[1]
    print('# ' + rep)
    for r in data:
        name = r[h.index('Kernel Name')][:60]
        dur_unit = u[h.index('gpu__time_duration.sum')]
        dur = col(r, 'gpu__time_duration.sum') * {'us': 1.0, 'ms': 1e3, 'ns': 1e-3, 'usecond': 1.0, 'msecond': 1e3, 'nsecond': 1e-3, 's': 1e6}.get(dur_unit, 1.0)
        def nbytes(k):
            un = u[h.index(k)]
            return col(r, k) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(un, 1)
        rd, wr = nbytes('dram__bytes_read.sum'), nbytes('dram__bytes_write.sum')
        st = []
        for i, k in enumerate(h):
            m = re.match(r'smsp__average_warps_issue_stalled_(.*)_per_issue_active.ratio', k)
            if m:
                try: st.append((float(r[i]), m.group(1)))
                except ValueError: pass
        st.sort(reverse=True)
        print('%-60s %9.1f us  dram r %8.2f MB w %8.2f MB = %6.0f GB/s  issue %4.1f%%  regs %3d  warps %4.1f%%  inst %8.2fM  stalls: %s' % (
            name, dur, rd / 1e6, wr / 1e6, (rd + wr) / dur / 1e3 if dur else 0, col(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
            int(col(r, 'launch__registers_per_thread')), col(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'),
            col(r, 'smsp__inst_executed.sum') / 1e6, ', '.join('%s %.1f' % (n, v) for v, n in st[:3])))
