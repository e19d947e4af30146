#!/usr/bin/env python3
"""Measurement of the SURVEY §8(f) rows built after the headline path (not the driver's bench contract):
periodic reductions / inertia, DCD decode, XTC decode.  One JSON line per workload on stdout:
value in the row's own unit, the kernel's algorithmic-byte roofline (CUDA events inside the library,
mb_get_stat "traj_*", or wall clock around the synchronous C-ABI call for the reductions) and the oracle
timed beside it on the host (a reported baseline, bounded sample).

    python tools/bench_extra.py [--atoms 1000000] [--frames 16] [--reps 20]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)

TRIC = np.array([[21.5, -2.7, -2.7], [0.0, 21.5, -2.7], [0.0, 0.0, 21.5]], np.float32)


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def roof(alg_bytes, ms, kernel):
    a = alg_bytes / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": kernel, "achieved": a, "peak": peak(), "unit": "GB/s", "frac": a / peak(),
            "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": ms, "traffic": None}


def bench_pbc(mb, orc, n, reps):
    xyz = orc.synth_frame(20260, 0, n, TRIC, stray_permille=10)
    m = orc.synth_masses(20260, n)
    s = mb.System(xyz, masses=m, box=TRIC)
    sel = s()
    out = []
    for name, fn, passes, cpu in (
            ("center_of_mass_pbc", lambda: sel.com(dims=[True, True, True]), 1,
             lambda: orc.center_pbc(xyz, m, orc.Box(matrix=TRIC), 7, prec="f32")),
            ("gyration_pbc", lambda: sel.gyration(pbc=True), 2,
             lambda: orc.gyration_pbc(xyz, m, orc.Box(matrix=TRIC), prec="f32")),
            ("inertia_pbc", lambda: sel.inertia(pbc=True), 2,
             lambda: orc.inertia(xyz, m, orc.Box(matrix=TRIC), prec="f32"))):
        for _ in range(3):
            fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        ms = (time.perf_counter() - t0) * 1e3 / reps
        t0 = time.perf_counter()
        cpu()
        cms = (time.perf_counter() - t0) * 1e3
        out.append({"workload": f"{name}, {n} atoms, triclinic box, 1% stray atoms", "metric": "calls/sec",
                    "value": 1e3 / ms, "unit": "calls/s", "ms_per_call": ms,
                    "roofline": roof(passes * 16.0 * n, ms, "center_pbc_kernel" + ("+tensor_kernel" if passes == 2 else "")
                                     + " (wall clock of the synchronous call: launch + sync latency included)"),
                    "cpu_baseline": {"value": 1e3 / cms, "unit": "calls/s", "cores": 1, "kind": "port",
                                     "sample": "1 call of the oracle's f32 restatement (serial, as the reference)"}})
    s.close()
    return out


def bench_dcd(mb, T, n, frames):
    rng = np.random.default_rng(0)
    one = (rng.random((1, n, 3)) * 20.0).astype(np.float32)
    cell = np.array([215.0, 0.0, 215.0, 0.0, 0.0, 215.0])
    buf = T.write_dcd(np.repeat(one, frames, axis=0), boxes=[cell] * frames)
    traj = mb.Trajectory()
    traj.load(buf, "dcd")
    t0 = time.perf_counter()
    traj.load(buf, "dcd")
    wall = time.perf_counter() - t0
    dec, h2d = traj.stat("traj_decode_ms"), traj.stat("traj_h2d_ms")
    t0 = time.perf_counter()
    T.read_dcd(buf[: len(buf) // frames * min(frames, 4) + 400])
    cpu_s = (time.perf_counter() - t0) / min(frames, 4)
    traj.close()
    return [{"workload": f"DCD decode, {n} atoms x {frames} frames (x[],y[],z[] f32 -> [atom][3] nm)",
             "metric": "frames/sec (decode kernel)", "value": frames / (dec * 1e-3), "unit": "frames/s",
             "roofline": roof(24.0 * n * frames, dec, "dcd_unpack_kernel (one launch, all frames)"),
             "e2e": {"value": frames / wall, "unit": "frames/s", "h2d_bytes_per_step": len(buf), "d2h_bytes_per_step": 0,
                     "h2d_ms": h2d, "note": "host byte buffer (pageable) -> device frames, whole call"},
             "cpu_baseline": {"value": 1.0 / cpu_s, "unit": "frames/s", "cores": 1, "kind": "port",
                              "sample": "numpy restatement of DcdFileHandler::read_state on 4 frames"}}]


def bench_xtc(mb, T, n, frames):
    rng = np.random.default_rng(1)
    L = (n / 100.0) ** (1.0 / 3.0)
    nmol = n // 3
    o = rng.random((nmol, 3)) * L
    xyz = (np.repeat(o, 3, axis=0) + rng.normal(0.0, 0.03, (3 * nmol, 3))).astype(np.float32)
    box = np.diag([L, L, L]).astype(np.float32)
    t0 = time.perf_counter()
    fr = T.write_xtc_frame(xyz, box)
    enc_s = time.perf_counter() - t0
    buf = fr * frames
    traj = mb.Trajectory()
    traj.load(buf, "xtc")
    t0 = time.perf_counter()
    traj.load(buf, "xtc")
    wall = time.perf_counter() - t0
    dec, scan, h2d = traj.stat("traj_decode_ms"), traj.stat("traj_scan_ms"), traj.stat("traj_h2d_ms")
    ok = bool(np.array_equal(traj.frames(0, 1)[0][:3000], T.read_xtc_frame(fr)[0]["xyz"][:3000])) if n <= 200_000 else None
    cpu = None
    if n <= 200_000:
        t0 = time.perf_counter()
        T.read_xtc_frame(fr)
        cpu = time.perf_counter() - t0
    traj.close()
    na = len(xyz)
    return [{"workload": f"XTC decode, {na} atoms x {frames} frames, water-like synthetic frames "
                         f"({len(fr) / na:.2f} B/atom compressed)",
             "metric": "frames/sec (scan + decode kernels)", "value": frames / (dec * 1e-3), "unit": "frames/s",
             "scan_ms": scan, "decode_ms": dec - scan, "checked_against_oracle": ok,
             "roofline": roof((len(fr) + 12.0 * na) * frames, dec, "xtc_scan_kernel + xtc_decode_kernel"),
             "e2e": {"value": frames / wall, "unit": "frames/s", "h2d_bytes_per_step": len(buf), "d2h_bytes_per_step": 0,
                     "h2d_ms": h2d, "note": "host byte buffer (pageable) -> device frames, whole call"},
             "cpu_baseline": ({"value": 1.0 / cpu, "unit": "frames/s", "cores": 1, "kind": "port",
                               "sample": "pure-Python oracle decoder, 1 frame (a checker, not a fast CPU decoder)"}
                              if cpu else None),
             "input_synthesis_s": enc_s}]


def bench_connect(mb, orc, n):
    """CSR adjacency of the 1.2 nm pair list, and unwrap of a water-like system (3-atom molecules, 0.12 nm bonds)."""
    out = []
    box = TRIC
    xyz = orc.synth_frame(20260, 0, n, box)
    s = mb.System(xyz, box=box)
    s.set_option("with_dist", 0)
    lib, h = s._lib, s._h
    npairs = mb._capi.check(lib.mb_search_single(h, 1.2, None, n, 7))
    lib.mb_connectivity(h, n, None)  # warm-up: allocations
    t0 = time.perf_counter()
    nnz = mb._capi.check(lib.mb_connectivity(h, n, None))
    ms = (time.perf_counter() - t0) * 1e3
    out.append({"workload": f"SearchConnectivity (CSR) of the 1.2 nm pair list, {n} atoms, {npairs} pairs",
                "metric": "calls/sec", "value": 1e3 / ms, "unit": "calls/s", "ms_per_call": ms,
                "roofline": roof(8.0 * npairs * 2 + 8.0 * npairs, ms, "csr_count_kernel + scan + csr_fill_kernel "
                                 "(pairs read twice, 2P neighbour ids written; wall clock of the synchronous call)"),
                "cpu_baseline": None})
    import ctypes as C
    lib.mb_search_connectivity(h, C.c_float(1.2), None, n, 7, None)  # warm-up: plan, allocations
    t0 = time.perf_counter()
    nnz2 = mb._capi.check(lib.mb_search_connectivity(h, C.c_float(1.2), None, n, 7, None))
    ms2 = (time.perf_counter() - t0) * 1e3
    out.append({"workload": f"SearchConnectivity as neighbour rows written by the search kernel (full shell: count pass, "
                            f"scan, fill pass), {n} atoms, {nnz2} entries",
                "metric": "calls/sec", "value": 1e3 / ms2, "unit": "calls/s", "ms_per_call": ms2,
                "same_entries_as_pair_list_csr": bool(nnz2 == nnz),
                "roofline": roof(12.0 * n + 4.0 * nnz2, ms2, "search_cells_kernel<4> + scan + search_cells_kernel<5> "
                                 "(frame read, 4-byte neighbour ids written; wall clock of the synchronous call)"),
                "cpu_baseline": None})
    s.close()
    rng = np.random.default_rng(2)
    L = (n / 100.0) ** (1.0 / 3.0)
    nmol = n // 3
    o = rng.random((nmol, 3)) * L
    w = np.repeat(o, 3, axis=0) + rng.normal(0.0, 0.03, (3 * nmol, 3))
    wrapped = np.mod(w, L).astype(np.float32)
    bx = np.diag([L, L, L]).astype(np.float32)
    s = mb.System(wrapped, box=bx)
    sel = s()
    roots = np.zeros(len(wrapped), np.int64)
    s._lib.mb_unwrap_connectivity(s._h, 0.12, None, len(wrapped), 7, roots.ctypes.data_as(mb._capi.i64p))  # warm-up
    s.set_state(wrapped, box=bx)  # the call moves the atoms: time it on the wrapped frame again
    t0 = time.perf_counter()
    ncomp = mb._capi.check(s._lib.mb_unwrap_connectivity(s._h, 0.12, None, len(wrapped), 7,
                                                         roots.ctypes.data_as(mb._capi.i64p)))
    ms = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    sub = wrapped[: min(len(wrapped), 150_000)]
    orc.unwrap_connectivity(0.12, sub, orc.Box(matrix=bx))
    cpu_ms = (time.perf_counter() - t0) * 1e3 * len(wrapped) / len(sub)
    out.append({"workload": f"unwrap_connectivity, {len(wrapped)} atoms in 3-atom molecules, cutoff 0.12 nm "
                            f"({ncomp} components)", "metric": "calls/sec", "value": 1e3 / ms, "unit": "calls/s",
                "ms_per_call": ms,
                "roofline": roof(24.0 * len(wrapped), ms, "search + CSR + union-find + unwrap_bfs_kernel (whole call, "
                                 "wall clock; algorithmic bytes = frame read + written)"),
                "cpu_baseline": {"value": 1e3 / cpu_ms, "unit": "calls/s", "cores": 4, "kind": "port",
                                 "sample": f"oracle on the first {len(sub)} atoms, scaled linearly"}})
    s.close()
    return out


def bench_many(mb, orc, n, reps):
    """Thousands of small selections: one mb_reduce_many launch vs one mb_center_of_mass call per selection vs the
    oracle's serial loop (what a rayon worker does per selection in the reference)."""
    xyz = orc.synth_frame(20260, 0, n, TRIC)
    m = orc.synth_masses(20260, n)
    s = mb.System(xyz, masses=m, box=TRIC)
    per = 20
    nsel = n // per
    ids = np.arange(nsel * per, dtype=np.uint64)
    offsets = np.arange(0, nsel * per + 1, per, dtype=np.uint64)
    for _ in range(2):
        s.reduce_many((ids, offsets), "com_gyration")
    t0 = time.perf_counter()
    for _ in range(reps):
        s.reduce_many((ids, offsets), "com_gyration")
    ms = (time.perf_counter() - t0) * 1e3 / reps
    k = 200
    t0 = time.perf_counter()
    for j in range(k):
        s((j * per, j * per + per - 1)).com()
    single_us = (time.perf_counter() - t0) * 1e6 / k
    t0 = time.perf_counter()
    for j in range(2000):
        sel = ids[j * per:(j + 1) * per]
        orc.center_of_mass(xyz, m, sel, prec="f32")
        orc.gyration(xyz, m, sel, prec="f32")
    cpu_us = (time.perf_counter() - t0) * 1e6 / 2000
    s.close()
    return [{"workload": f"COM + gyration of {nsel} selections of {per} atoms ({n} atoms), ONE mb_reduce_many call "
                         f"(ids + offsets uploaded per call)", "metric": "selections/sec", "value": nsel / (ms * 1e-3),
             "unit": "selections/s", "ms_per_call": ms, "per_selection_call_us": single_us,
             "roofline": roof(16.0 * nsel * per + 8.0 * nsel * per, ms, "reduce_many_kernel (wall clock of the "
                              "synchronous call incl. the H2D of 8 B/atom of ids; 16 B/atom gathered on the device)"),
             "cpu_baseline": {"value": 1e6 / cpu_us, "unit": "selections/s", "cores": 1, "kind": "port",
                              "sample": "oracle center_of_mass + gyration on 2000 selections through ctypes "
                                        "(call overhead included)"}}]


def bench_vdw(mb, orc, n):
    """vdW contact search through the cell kernel vs the plain two-set search at the vdW grid cutoff."""
    M = (TRIC * np.float32((n / 1.0e6) ** (1.0 / 3.0))).astype(np.float32)
    xyz = orc.synth_frame(20260, 0, n, M)
    rng = np.random.default_rng(9)
    vdw = (0.10 + 0.11 * rng.random(n)).astype(np.float32)
    ids1 = np.arange(0, n, 2, dtype=np.uint64)
    ids2 = np.arange(1, n, 2, dtype=np.uint64)
    s = mb.System(xyz, box=M, vdw=vdw)
    s.set_option("with_dist", 0)
    C = mb._capi
    p1, p2 = ids1.ctypes.data_as(C.u64p), ids2.ctypes.data_as(C.u64p)
    v1 = np.ascontiguousarray(vdw[ids1.astype(int)])
    v2 = np.ascontiguousarray(vdw[ids2.astype(int)])
    cut = float(v1.max() + v2.max() + np.finfo(np.float32).eps)

    def timed(fn):
        fn()
        t0 = time.perf_counter()
        for _ in range(3):
            r = fn()
        return (time.perf_counter() - t0) * 1e3 / 3, r

    t_v, nv = timed(lambda: C.check(s._lib.mb_search_double_vdw(s._h, p1, len(ids1), v1.ctypes.data_as(C.f32p), p2,
                                                                len(ids2), v2.ctypes.data_as(C.f32p), 0, 7)))
    t_p, npl = timed(lambda: C.check(s._lib.mb_search_double(s._h, cut, p1, len(ids1), p2, len(ids2), 0, 7)))
    s.close()
    return [{"workload": f"vdW contact search, {len(ids1)} x {len(ids2)} atoms, radii 0.10-0.21 nm, periodic",
             "metric": "calls/sec", "value": 1e3 / t_v, "unit": "calls/s", "ms_per_call": t_v, "pairs": int(nv),
             "plain_two_set_search_same_grid_cutoff_ms": t_p, "plain_pairs": int(npl),
             "roofline": roof(12.0 * n + 8.0 * n + 8.0 * nv, t_v, "bin + scan + scatter (x2) + search_cells_kernel<VDW> "
                              "(whole synchronous call incl. id / radius upload)"), "cpu_baseline": None}]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--atoms", type=int, default=1_000_000)
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    import molar_b200 as mb
    from oracle import oracle_py as orc
    from oracle import traj_oracle as T
    lines = []
    if a.only in ("", "pbc"):
        lines += bench_pbc(mb, orc, a.atoms, a.reps)
    if a.only in ("", "many"):
        lines += bench_many(mb, orc, a.atoms, a.reps)
    if a.only in ("", "vdw"):
        lines += bench_vdw(mb, orc, min(a.atoms, 400_000))
    if a.only in ("", "connect"):
        lines += bench_connect(mb, orc, a.atoms)
    if a.only in ("", "dcd"):
        lines += bench_dcd(mb, T, a.atoms, a.frames)
    if a.only in ("", "xtc"):
        lines += bench_xtc(mb, T, min(a.atoms, 150_000), max(a.frames, 256))
        lines += bench_xtc(mb, T, a.atoms, max(a.frames, 96))
    for ln in lines:
        print(json.dumps(ln))


if __name__ == "__main__":
    main()
