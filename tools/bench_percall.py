#!/usr/bin/env python3
"""Per-call timings of the single-selection search at 1M atoms through the C ABI: pairs only (with_dist=0), pairs +
distances (the default of the pymolar-style call) and count only; `within` of a 50k-atom inner set."""
import os, sys, time, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import molar_b200 as mb
from molar_b200 import _capi
TRIC = np.array([[21.5, -2.7, -2.7], [0.0, 21.5, -2.7], [0.0, 0.0, 21.5]], np.float32)
n = 1_000_000
t = mb.Trajectory(); t.synth(20260, 0, 1, n, TRIC); xyz = t.frame(0); t.close()
s = mb.System(xyz, box=TRIC)
def timed(fn, reps=5):
    fn(); s.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    s.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
out = {}
for wd in (0, 1):
    s.set_option("with_dist", wd)
    out["search_single with_dist=%d ms" % wd] = timed(lambda: _capi.check(s._lib.mb_search_single(s._h, 1.2, None, n, 7)))
out["count_single ms"] = timed(lambda: _capi.check(s._lib.mb_count_single(s._h, 1.2, None, n, 7)))
from molar_b200.api import within
inner = np.arange(0, n, 20, dtype=np.uint64)
out["within 1M vs 50k ms"] = timed(lambda: within(0.5, s(), s(inner), dims=7), reps=3)
print(json.dumps(out))
