#!/usr/bin/env python3
"""bench.py — frames/sec of the hot path on synthetic frames (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

One rank per GPU (torchrun launches the processes for N>1), frames sharded across ranks (weak scaling:
every rank owns `--frames` resident frames and one step = one pass of the hot path over all of them).
Prints ONE JSON line on rank 0.  Everything — kernels, CUDA-event timing, the NCCL communicator, the
gather of the per-frame scalars, barrier and max over ranks — goes through libmolar_b200.so's C ABI;
torch is not imported.

Workloads (BASELINE.json configs):
  search1m   configs[2]  1M-atom triclinic box, 1.2 nm neighbour-pair enumeration   (default, headline)
  search100k configs[1]  100k-atom orthorhombic box, 1.2 nm
  fit500k    configs[3]  500k-atom Kabsch fit + superposition + RMSD to frame 0
  pipeline1m configs[4]  1M-atom COM + gyration + 1.2 nm contact count, NCCL scalar gather
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20260
TRIC = np.array([[21.5, -2.7, -2.7], [0.0, 21.5, -2.7], [0.0, 0.0, 21.5]], np.float32)
ORTHO = np.diag([10.0, 10.0, 10.0]).astype(np.float32)
CUTOFF = 1.2

WORKLOADS = {
    "search1m": dict(n_atoms=1_000_000, box=TRIC, frames=16, kind="search",
                     desc="1M-atom triclinic box, 1.2 nm neighbour-pair enumeration (configs[2])"),
    "search100k": dict(n_atoms=100_000, box=ORTHO, frames=128, kind="search",
                       desc="100k-atom orthorhombic box, 1.2 nm neighbour search (configs[1])"),
    "fit500k": dict(n_atoms=500_000, box=TRIC, frames=256, kind="fit",
                    desc="500k-atom Kabsch fit + superposition + RMSD to frame 0 (configs[3])"),
    "pipeline1m": dict(n_atoms=1_000_000, box=TRIC, frames=16, kind="pipeline",
                       desc="1M-atom COM + gyration + 1.2 nm contact count (configs[4])"),
}


def measured_traffic(workload):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return float(json.load(f)[workload]["dram_bytes_per_launch"])
    except Exception:
        return None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


_REAL_STDOUT = None


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ---------------------------------------------------------------------------------------------
# reference arm: the reference's CPU algorithm (C++ restatement, all host threads)
# ---------------------------------------------------------------------------------------------
def cpu_frames_per_sec(wl, n_frames, nthreads, first_frame=0, budget_s=60.0):
    """Times the oracle (kind "port": MolAR itself is Rust and cannot be built here) on up to
    n_frames frames, stopping early once budget_s of CPU work has been spent (bounded sample)."""
    from oracle import oracle_py as orc
    n, box = wl["n_atoms"], wl["box"]
    b = orc.Box(matrix=box)
    n_gen = min(n_frames, 4)
    frames = [orc.synth_frame(SEED, first_frame + f, n, box) for f in range(n_gen)]
    masses = orc.synth_masses(SEED, n)
    ref = frames[0]
    done = 0
    t0 = time.perf_counter()
    for it in range(n_frames):
        xyz = frames[it % n_gen]
        if done and time.perf_counter() - t0 > budget_s:
            break
        done += 1
        if wl["kind"] == "search":
            h = orc.lib().orc_search_single_pbc(CUTOFF, xyz.ctypes.data_as(orc._f32p), None, n, b.h, 7, nthreads)
            orc.lib().orc_result_free(h)
        elif wl["kind"] == "fit":
            rc, R, t = orc.fit_transform(xyz, masses, None, ref, masses, None, prec="f32")
            moved = orc.apply_transform_f32(xyz, None, R, t)
            orc.rmsd(moved, None, ref, None, prec="f32")
        else:
            orc.center_of_mass(xyz, masses, prec="f32")
            orc.gyration(xyz, masses, prec="f32")
            h = orc.lib().orc_search_single_pbc(CUTOFF, xyz.ctypes.data_as(orc._f32p), None, n, b.h, 7, nthreads)
            orc.lib().orc_result_free(h)
    dt = time.perf_counter() - t0
    return done / dt, dt, done


def run_reference(args, wl):
    """The reference's own CPU path for this workload on the host cores (oracle port: MolAR is Rust and there is
    no cargo here).  Loads nothing of the product: no libmolar_b200.so, no CUDA."""
    rank, world, local = dist_env()
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    nthreads = cores if wl["kind"] != "fit" else 1  # measure.rs paths are serial in the reference
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_frames_per_sec(wl, 1, nthreads)
    fps, dt, done = cpu_frames_per_sec(wl, max(1, args.steps), nthreads, first_frame=1, budget_s=60.0)
    line = {
        "impl": "reference", "metric": "frames/sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": done, "steps_requested": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": 1000.0 * dt / done,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "n_atoms": wl["n_atoms"], "cutoff_nm": CUTOFF,
                   "frames_per_step": 1, "note": "one step = one frame on the host cores; steps = frames actually "
                                                 "timed (bounded sample of 60 s of CPU work)"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": nthreads, "kind": "port",
                         "sample": f"{done} frames of the workload ({dt:.1f} s); C++ restatement of MolAR's "
                                   f"CPU algorithm (MolAR is Rust; no cargo in this image)"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)
    return 0


# ---------------------------------------------------------------------------------------------
# our arm: everything through libmolar_b200.so (C ABI via ctypes); no torch anywhere on this path.
# Ranks are the launcher's processes (torchrun exports RANK / LOCAL_RANK / WORLD_SIZE); the rank
# plumbing — NCCL communicator, scalar gather, barrier, max over ranks — is the library's mb_comm_*.
# ---------------------------------------------------------------------------------------------
FP32_LANE_OPS_PER_TEST = 9  # 3 sub, 3 mul, 2 add (as fma by 1.0), 1 sub against cutoff^2 — one f32 lane op each


def fp32_peak_lane_ops(sm_count, sm_mhz):
    """Non-tensor FP32 issue peak: 128 FP32 lanes per SM, one op per lane per clock (an FMA counts as ONE op here,
    since the bit-exact test may not fuse).  Source: 4 SMSP x 32 lanes (B300_MICROARCH.md) x measured SM clock."""
    return sm_count * 128.0 * sm_mhz * 1e6


def run_ours(args, wl):
    import molar_b200 as mb
    from molar_b200 import comm as mbc

    rank, world, local = dist_env()
    n, box, F, kind = wl["n_atoms"], wl["box"], args.frames or wl["frames"], wl["kind"]
    traj = mb.Trajectory(device=local)  # raises without a CUDA device: there is no CPU fallback
    cm = mbc.Comm(traj, rank, world)    # NCCL communicator inside the library (no-op for one rank)
    f_first, f_last = mbc.frame_block(rank, F)  # rank r owns global frames [r*F, (r+1)*F)
    traj.synth(SEED, f_first, F, n, box, mass_seed=SEED)
    for kv in filter(None, args.opts.split(",")):
        key, val = kv.split("=")
        traj.set_option(key, float(val))
    traj.set_option("profile", 1)

    def step():
        if kind == "search":
            return traj.search(CUTOFF)
        if kind == "fit":
            return traj.fit(ref_frame=0, superpose=True)
        return traj.pipeline(CUTOFF)

    def barrier():
        cm.barrier()
        traj.synchronize()

    ncol = {"search": 1, "fit": 1, "pipeline": 5}[kind]

    def gather(res):
        # per-frame scalars -> every rank (NCCL all-gather over NVLink inside mb_gather_scalars; the only collective)
        return cm.gather(np.asarray(res, dtype=np.float64).reshape(F, ncol))

    for _ in range(max(args.warmup, 3)):
        res = step()
    allrows = gather(res)  # warm the communicator outside the timed region
    traj.set_option("profile", 1)  # reset the kernel-time accumulators
    l0 = traj.launch_count()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    barrier()
    traj.timer_record(0)
    for _ in range(args.steps):
        res = step()
    allrows = gather(res)
    traj.timer_record(1)
    barrier()
    ms = float(cm.max([traj.timer_ms(0, 1)])[0])  # device time, max over ranks
    clk = clocks.stop() if rank == 0 else None
    launches = traj.launch_count() - l0
    k_ms = traj.stat("search_kernel_ms")
    k_n = traj.stat("search_kernel_launches")
    tests_per_frame = traj.stat("search_tests_per_frame") if kind == "pipeline" else 0.0
    sm_count = traj.stat("sm_count")
    fps = world * F * args.steps / (ms / 1000.0)
    assert allrows.shape == (world * F, ncol)

    if args.no_e2e:
        if rank == 0:
            emit({"tuning": True, "opts": args.opts, "value": fps, "ms_per_frame": ms / (F * args.steps),
                  "search_kernel_ms": k_ms / max(k_n, 1), "workload": args.workload})
        traj.close()
        return 0

    # ---- end to end through the public API with HOST buffers (page-locked), rank-local --------------------------
    # host frames come from the library's own generator (mb_batch_synth + mb_batch_download): the oracle stays out
    # of the product arm until the cpu_baseline leg
    e2e_frames = min(F, 4)
    host = mbc.pinned_empty((e2e_frames, n, 3), np.float32)
    host[:] = traj.frames(0, e2e_frames)
    gen = mb.Trajectory(device=local)
    gen.synth(SEED, 0, 1, n, box, mass_seed=SEED)
    masses_h = gen.masses_host()
    gen.close()
    sysm = mb.System(host[0], masses=masses_h, box=box, device=local)
    refsys = mb.System(host[0], masses=masses_h, box=box, device=local) if kind == "fit" else None
    e2e_t = {"set": 0.0}

    def e2e_once(x):
        ta = time.perf_counter()
        sysm.set_state(x, box)  # H2D of the frame (12 B/atom) from pinned memory
        e2e_t["set"] += time.perf_counter() - ta
        if kind == "search":
            lib, h = sysm._lib, sysm._h
            return mb._capi.check(lib.mb_search_single(h, CUTOFF, None, n, 7))  # D2H: the pair count
        if kind == "fit":
            tr = mb.fit_transform(sysm(), refsys())
            sysm().apply_transform(tr)
            return mb.rmsd(sysm(), refsys())
        c = sysm().com()
        g = sysm().gyration()
        lib, h = sysm._lib, sysm._h
        return (c, g, mb._capi.check(lib.mb_count_single(h, CUTOFF, None, n, 7)))

    sysm.set_option("with_dist", 0)
    e2e_once(host[0])
    barrier()
    t0 = time.perf_counter()
    reps = max(1, min(args.steps, 4))
    for _ in range(reps):
        for f in range(e2e_frames):
            e2e_once(host[f])
    sysm.synchronize()
    e2e_s = float(cm.max([time.perf_counter() - t0])[0])
    per_call_fps = world * reps * e2e_frames / e2e_s
    d2h = {"search": 8, "fit": 8 + 96, "pipeline": 40}[kind]
    e2e_extra = {"per_call_value": per_call_fps}

    # what a drop-in distance_search_single_pbc call returns is the PAIR LIST (distance_search.rs:948-953): search +
    # fetch of the canonical (i<j) pairs into page-locked host memory as u32 x 2 (8 B/pair; the usize widening of
    # mb_fill_pairs doubles the PCIe bytes and is the binding's choice)
    if kind == "search":
        cnt = int(e2e_once(host[0]))
        pairs_h = mbc.pinned_empty((cnt + cnt // 8 + 4096, 2), np.uint32)
        sysm.fill_pairs_u32(pairs_h)
        barrier()
        t0 = time.perf_counter()
        nfr = min(e2e_frames, 3)
        tot_pairs = 0
        for f in range(nfr):
            c = int(e2e_once(host[f]))
            sysm.fill_pairs_u32(pairs_h)
            tot_pairs += c
        wp_s = float(cm.max([time.perf_counter() - t0])[0])
        e2e_extra["with_pairs_value"] = world * nfr / wp_s
        e2e_extra["with_pairs_d2h_bytes_per_step"] = 8.0 * tot_pairs / nfr
        e2e_extra["with_pairs_note"] = ("mb_set_frame + mb_search_single + mb_fill_pairs_u32 per frame: the pair "
                                        "list (u32 x 2, canonical i<j) lands in page-locked host memory; PCIe-bound")
        mbc.pinned_free(pairs_h)

    # the trajectory-loop entry points: host frames in, per-frame results out, uploads overlapped with the work on
    # the previous chunks (mb_stream_*; the reference overlaps IO with analysis the same way, io.rs:209-233)
    ns = {"search": 32, "fit": 128, "pipeline": 8}[kind] if n >= 500_000 else 64
    blockh = mbc.pinned_empty((ns, n, 3), np.float32)
    for f in range(ns):
        blockh[f] = host[f % e2e_frames]
    st = mb.Trajectory(device=local)
    for kv in filter(None, args.opts.split(",")):
        key, val = kv.split("=")
        st.set_option(key, float(val))

    # pure upload bandwidth of this rank while all ranks upload at once (what bounds e2e at N > 1)
    barrier()
    t0 = time.perf_counter()
    st.upload(blockh[:min(ns, 16)], box=box)
    up_s = float(cm.max([time.perf_counter() - t0])[0])
    e2e_extra["h2d_gbs_per_rank_all_ranks_uploading"] = min(ns, 16) * n * 12 / up_s / 1e9

    def stream_once():
        if kind == "search":
            return st.stream_search(blockh, CUTOFF, box)
        if kind == "fit":
            return st.stream_fit(blockh, masses_h)
        return st.stream_pipeline(blockh, CUTOFF, box, masses=masses_h)

    stream_once()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        stream_once()
    st.synchronize()
    s_s = float(cm.max([time.perf_counter() - t0])[0])
    e2e_fps = world * reps * ns / s_s
    e2e_extra["h2d_gbs_per_rank_streaming"] = reps * ns * n * 12 / s_s / 1e9
    e2e_note = ("mb_stream_%s: host frames (page-locked) in, per-frame results on the host; uploads overlap the work "
                "on the previous chunks; per_call_value = blocking per-frame calls (mb_set_frame + ...)" % kind)
    st.close()
    mbc.pinned_free(blockh)

    if rank == 0:
        peak, which = peaks()
        if kind in ("search", "pipeline"):
            pairs = float(np.mean(res)) if kind == "search" else float(np.mean(np.asarray(res)[:, 4]))
            # pipeline: the count-only search reads the frame (12 B/atom); masses are L2-resident
            alg_bytes = 12.0 * n + (8.0 * pairs if kind == "search" else 0.0)
            kern = "search_cells_kernel"
            k_avg_ms = k_ms / max(k_n, 1)
        else:
            pairs = 0.0
            alg_bytes = 24.0 * n
            kern = "fit (whole step: moments + superposition)"
            k_avg_ms = ms / (F * args.steps)
        # launches of consecutive frames overlap on alternating streams (small frames: up to three at once), which
        # stretches each launch's own duration; the time the kernel costs per frame is bounded by the step time
        k_event_ms = k_avg_ms
        k_avg_ms = min(k_avg_ms, ms / (F * args.steps))
        achieved = alg_bytes / (k_avg_ms * 1e-3) / 1e9 if k_avg_ms > 0 else 0.0
        roof = {"bound": "hbm", "kernel": kern, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": measured_traffic(args.workload), "peak_source": which,
                "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": k_avg_ms,
                "avg_launch_event_ms": k_event_ms,
                "kernel_share_of_step": min(1.0, k_ms / ms) if kind != "fit" else 1.0,
                "note": "avg_launch_event_ms = CUDA-event time of a launch on its own stream; consecutive "
                        "frames run on alternating streams and overlap, so avg_launch_ms = min(event time, "
                        "step time per frame)"}
        if kind == "pipeline":
            # the count-only search is bound by the FP32 pipes, not by HBM
            sm_mhz = (clk or {}).get("sm_mhz") or 1965.0
            fpeak = fp32_peak_lane_ops(sm_count, sm_mhz) / FP32_LANE_OPS_PER_TEST
            tps = tests_per_frame / (k_avg_ms * 1e-3) if k_avg_ms > 0 else 0.0
            roof = {"bound": "fp32", "kernel": "search_cells_kernel<count only>", "achieved": tps / 1e12,
                    "peak": fpeak / 1e12, "unit": "Ttests/s", "frac": tps / fpeak if fpeak else 0.0, "traffic": None,
                    "tests_per_launch": tests_per_frame, "avg_launch_ms": k_avg_ms,
                    "peak_source": f"non-tensor FP32 issue peak: {int(sm_count)} SMs x 128 lanes x {sm_mhz:.0f} MHz "
                                   f"(SM clock sampled during the run) / {FP32_LANE_OPS_PER_TEST} lane ops per "
                                   f"unfused distance test",
                    "hbm": {"achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                            "algorithmic_bytes_per_launch": alg_bytes, "peak_source": which},
                    "kernel_share_of_step": min(1.0, k_ms / ms)}
        line = {
            "metric": "frames/sec", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "n_atoms": n, "cutoff_nm": CUTOFF, "frames_per_step_per_gpu": F,
                       "pairs_per_frame": pairs, "l2": f"inputs {F * n * 12 / 1e6:.0f} MB/GPU > 126 MB L2"
                       if F * n * 12 > 126e6 else "pair output per frame exceeds L2; inputs re-streamed",
                       "parallelism": f"frames sharded over {world} GPU(s); scalar gather by mb_gather_scalars "
                                      f"(NCCL {cm.info()[2]})"},
            "roofline": roof,
            "e2e": dict({"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": 12 * n, "d2h_bytes_per_step": d2h,
                         "note": e2e_note}, **e2e_extra),
            "gpu_launches": int(launches),
            "clocks": clk,
        }
        if not args.no_cpu and world == 1:
            cores = os.cpu_count() or 1
            nthreads = cores if kind != "fit" else 1
            nfr = {"search": 2, "fit": 20, "pipeline": 2}[kind] if n >= 500_000 else 20
            cfps, cdt, nfr = cpu_frames_per_sec(wl, nfr, nthreads, budget_s=25.0)
            line["cpu_baseline"] = {"value": cfps, "unit": "frames/s", "cores": nthreads, "kind": "port",
                                    "sample": f"{nfr} frames of the same workload ({cdt:.1f} s); C++ restatement "
                                              f"of MolAR's CPU algorithm, not MolAR itself"}
        emit(line)
    mbc.pinned_free(host)
    traj.close()
    sysm.close()
    if refsys:
        refsys.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="search1m", choices=sorted(WORKLOADS))
    ap.add_argument("--frames", type=int, default=0, help="resident frames per GPU (one step = one pass over them)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end (host buffer) leg (tuning runs only)")
    ap.add_argument("--opts", default="", help="library tuning options, e.g. subdiv_x=3,slice_x=2 (tuning runs only)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    # stdout must carry exactly ONE JSON line: native libraries (NCCL prints its version banner there) write to
    # fd 1 directly, so fd 1 is pointed at stderr and the JSON line goes to a private copy of the real stdout
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        # the reference arm builds and loads ONLY the CPU oracle: the product library stays out of this process
        from oracle import oracle_py
        oracle_py.build()
        return run_reference(args, wl)
    import __graft_entry__
    __graft_entry__.build()
    return run_ours(args, wl)


if __name__ == "__main__":
    sys.exit(main())
