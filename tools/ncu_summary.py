#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, without a GPU) into a small text file for profiles/."""
import collections
import csv
import io
import re
import subprocess
import sys


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct",
]


def main():
    rep = sys.argv[1]
    hdr, units, data = raw(rep)
    lines = [f"# ncu --set full summary of {rep}"]
    names = [r[hdr.index("Kernel Name")] for r in data]
    lines.append(f"kernels: {names}")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            lines.append(f"{k:75s} {units[i]:12s} {[r[i] for r in data]}")
    lines.append("-- warp stall reasons (avg warps stalled per issue-active cycle) --")
    st = []
    for i, h in enumerate(hdr):
        m = re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active.ratio", h)
        if m:
            st.append((float(data[0][i]), m.group(1)))
    for v, n in sorted(st, reverse=True):
        lines.append(f"  {n:24s} {v:8.3f}")
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    if secs:
        s = secs[0]
        e = secs[1] if len(secs) > 1 else len(rows)
        h = rows[s + 1]
        ix, isrc, isamp = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
        body = rows[s + 2:e]
        tot = sum(int(r[ix]) for r in body)
        tsamp = sum(int(r[isamp]) for r in body)
        ops, samp = collections.Counter(), collections.Counter()
        for r in body:
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc])
            op = m.group(2).split(".")[0] if m else "?"
            ops[op] += int(r[ix])
            samp[op] += int(r[isamp])
        lines.append(f"-- executed warp instructions by opcode (total {tot}, {len(body)} SASS instrs, samples {tsamp}) --")
        for op, c in ops.most_common(22):
            lines.append(f"  {op:10s} {c / 1e6:10.1f}M {100 * c / tot:5.1f}%   samples {100 * samp[op] / max(tsamp, 1):5.1f}%")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
