"""Per-call latency (wall clock, through the Python shim) of the single-selection reductions: the small results come back
through mapped page-locked host memory, not through a D2H copy."""
import sys, time, json, numpy as np
sys.path.insert(0, '/root/repo')
import molar_b200 as mb
from oracle import oracle_py as orc
TRIC = np.array([[21.5, -2.7, -2.7], [0.0, 21.5, -2.7], [0.0, 0.0, 21.5]], np.float32)
out = {}
for n in (1_000_000, 20_000):
    xyz = orc.synth_frame(20260, 0, n, TRIC); m = orc.synth_masses(20260, n)
    s = mb.System(xyz, masses=m, box=TRIC)
    sel = s()
    ref = mb.System(orc.synth_frame(20260, 1, n, TRIC), masses=m)
    for name, fn in (("com", lambda: sel.com()), ("gyration", lambda: sel.gyration()), ("com_pbc", lambda: sel.com(dims=[True]*3)),
                     ("rmsd", lambda: mb.rmsd(sel, ref())), ("fit_transform", lambda: mb.fit_transform(sel, ref()))):
        fn()
        t0 = time.perf_counter()
        for _ in range(200): fn()
        out[f"{name} n={n} us"] = round((time.perf_counter() - t0) / 200 * 1e6, 1)
    s.close(); ref.close()
print(json.dumps(out))
