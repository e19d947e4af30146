#!/usr/bin/env python3
"""bench.py — frames/sec of the hot path on synthetic frames (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

One rank per GPU (torchrun for N>1), frames sharded across ranks (weak scaling: every rank owns
`--frames` resident frames and one step = one pass of the hot path over all of them).  Prints ONE
JSON line on rank 0.  torch is used for plumbing only (torch.distributed barrier / NCCL gather of
the per-frame scalars, CUDA events on the library's own stream); every number is produced by
libmolar_b200.so through its C ABI.

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
    rank, world, local = dist_env()
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    nthreads = cores if wl["kind"] != "fit" else 1  # measure.rs paths are serial in the reference
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_frames_per_sec(wl, 1, nthreads)
    fps, dt, done = cpu_frames_per_sec(wl, max(1, args.steps), nthreads, first_frame=1, budget_s=90.0)
    line = {
        "impl": "reference", "metric": "frames/sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * dt / done,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "n_atoms": wl["n_atoms"], "cutoff_nm": CUTOFF,
                   "frames_per_step": 1, "note": "one step = one frame on the host cores"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": nthreads, "kind": "port",
                         "sample": f"{done} frames of the workload ({dt:.1f} s); C++ restatement of MolAR's "
                                   f"CPU algorithm (MolAR is Rust; no cargo in this image)"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)
    return 0


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args, wl):
    import torch
    import torch.distributed as dist
    import molar_b200 as mb

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libmolar_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, box, F, kind = wl["n_atoms"], wl["box"], args.frames or wl["frames"], wl["kind"]

    import shard
    traj = mb.Trajectory(device=local)
    f_first, f_last = shard.frame_block(rank, F)  # rank r owns global frames [r*F, (r+1)*F)
    traj.synth(SEED, f_first, F, n, box, mass_seed=SEED)
    for kv in filter(None, args.opts.split(",")):
        key, val = kv.split("=")
        traj.set_option(key, float(val))
    traj.set_option("profile", 1)
    ext = torch.cuda.ExternalStream(traj.stream(), device=local)

    def step():
        if kind == "search":
            return traj.search(CUTOFF)
        if kind == "fit":
            return traj.fit(ref_frame=0, superpose=True)
        return traj.pipeline(CUTOFF)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    dev = torch.device("cuda", local)
    ncol = {"search": 1, "fit": 1, "pipeline": 5}[kind]
    scal_host = torch.zeros((F, ncol), dtype=torch.float64).pin_memory()
    gathered = torch.empty((world * F, ncol), dtype=torch.float64, device=dev)

    def gather(res):
        # per-frame scalars -> every rank (NCCL over NVLink; the only collective on the path)
        return shard.gather_rows(np.asarray(res, dtype=np.float64).reshape(F, ncol), world, device=dev,
                                 out=gathered, stage=scal_host)

    for _ in range(max(args.warmup, 3)):
        res = step()
    gather(res)  # warm the torch allocator / NCCL communicator outside the timed region
    traj.set_option("profile", 1)  # reset the kernel-time accumulators
    l0 = traj.launch_count()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for _ in range(args.steps):
        res = step()
    gather(res)
    torch.cuda.current_stream().synchronize()
    e1.record(ext)
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clk = clocks.stop() if rank == 0 else None
    launches = traj.launch_count() - l0
    k_ms = traj.stat("search_kernel_ms")
    k_n = traj.stat("search_kernel_launches")
    fps = world * F * args.steps / (ms / 1000.0)

    if args.no_e2e:
        if rank == 0:
            emit({"tuning": True, "opts": args.opts, "value": fps, "ms_per_frame": ms / (F * args.steps),
                              "search_kernel_ms": k_ms / max(k_n, 1), "workload": args.workload})
        traj.close()
        if world > 1:
            dist.destroy_process_group()
        return 0
    # ---- end-to-end through the public per-call API with HOST buffers (pinned), rank-local -------
    from oracle import oracle_py as orc  # only to synthesise host-side input frames
    e2e_frames = min(F, 4)
    host = [torch.from_numpy(orc.synth_frame(SEED, rank * F + f, n, box)).pin_memory() for f in range(e2e_frames)]
    sysm = mb.System(host[0].numpy(), masses=orc.synth_masses(SEED, n), box=box, device=local)
    refsys = (mb.System(host[0].numpy(), masses=orc.synth_masses(SEED, n), box=box, device=local)
              if kind == "fit" else None)

    e2e_t = {"set": 0.0, "call": 0.0}

    def e2e_once(x):
        ta = time.perf_counter()
        sysm.set_state(x.numpy(), box)  # H2D of the frame (12 B/atom) from pinned memory
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
        for x in host:
            e2e_once(x)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if os.environ.get("MB_DEBUG_TIMING"):
        sys.stderr.write(f"[bench] rank {rank}: e2e {e2e_s * 1e3:.1f} ms for {reps * e2e_frames} calls, "
                         f"set_state {e2e_t['set'] * 1e3:.1f} ms\n")
    if world > 1:
        t = torch.tensor([e2e_s], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_fps = world * reps * e2e_frames / e2e_s
    d2h = {"search": 8, "fit": 8 + 96, "pipeline": 40}[kind]
    e2e_extra = {}
    # the trajectory-loop entry points: host frames in, per-frame results out, uploads overlapped with the work on
    # the previous chunk (mb_stream_*; the reference overlaps IO with analysis the same way, io.rs:209-233)
    e2e_extra["per_call_value"] = e2e_fps
    ns = {"search": 32, "fit": 128, "pipeline": 8}[kind] if n >= 500_000 else 64
    blockh = torch.from_numpy(np.stack([host[f % e2e_frames].numpy() for f in range(ns)])).pin_memory()
    st = mb.Trajectory(device=local)
    for kv in filter(None, args.opts.split(",")):
        key, val = kv.split("=")
        st.set_option(key, float(val))
    masses_h = orc.synth_masses(SEED, n)

    def stream_once():
        if kind == "search":
            return st.stream_search(blockh.numpy(), CUTOFF, box)
        if kind == "fit":
            return st.stream_fit(blockh.numpy(), masses_h)
        return st.stream_pipeline(blockh.numpy(), CUTOFF, box, masses=masses_h)

    stream_once()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        stream_once()
    torch.cuda.synchronize()
    s_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([s_s], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        s_s = float(t.item())
    e2e_fps = world * reps * ns / s_s
    e2e_note = ("mb_stream_%s: host frames (pinned) in, per-frame results on the host; uploads overlap the work on the "
                "previous chunk; per_call_value = blocking per-frame calls (mb_set_frame + ...)" % kind)
    st.close()

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
            kern = "fit_moments_kernel+superpose_rmsd_kernel (whole step)"
            k_avg_ms = ms / (F * args.steps)
        # launches of consecutive frames overlap on alternating streams (small frames: up to three at once), which
        # stretches each launch's own duration; the time the kernel costs per frame is bounded by the step time
        k_event_ms = k_avg_ms
        k_avg_ms = min(k_avg_ms, ms / (F * args.steps))
        achieved = alg_bytes / (k_avg_ms * 1e-3) / 1e9 if k_avg_ms > 0 else 0.0
        line = {
            "metric": "frames/sec", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "n_atoms": n, "cutoff_nm": CUTOFF, "frames_per_step_per_gpu": F,
                       "pairs_per_frame": pairs, "l2": f"inputs {F * n * 12 / 1e6:.0f} MB/GPU > 126 MB L2"
                       if F * n * 12 > 126e6 else "pair output per frame exceeds L2; inputs re-streamed",
                       "parallelism": f"frames sharded over {world} GPU(s)"},
            "roofline": {"bound": "hbm", "kernel": kern, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": measured_traffic(args.workload), "peak_source": which,
                         "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": k_avg_ms,
                         "avg_launch_event_ms": k_event_ms,
                         "kernel_share_of_step": min(1.0, k_ms / ms) if kind != "fit" else 1.0,
                         "note": "avg_launch_event_ms = CUDA-event time of a launch on its own stream; consecutive "
                                 "frames run on alternating streams and overlap, so avg_launch_ms = min(event time, "
                                 "step time per frame)"},
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
    traj.close()
    sysm.close()
    if refsys:
        refsys.close()
    if world > 1:
        dist.destroy_process_group()
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
    import __graft_entry__
    __graft_entry__.build()
    if args.impl == "reference":
        return run_reference(args, wl)
    return run_ours(args, wl)


if __name__ == "__main__":
    sys.exit(main())
