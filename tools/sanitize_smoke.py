"""Small-config run of every kernel family for compute-sanitizer (memcheck / racecheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import molar_b200 as mb
from molar_b200.api import within
from oracle import oracle_py as orc

TRIC = np.array([[21.5, -2.7, -2.7], [0.0, 21.5, -2.7], [0.0, 0.0, 21.5]], np.float32)
M = (TRIC * np.float32(0.25)).astype(np.float32)
n = 9000
xyz = orc.synth_frame(1, 0, n, M, stray_permille=10)
m = orc.synth_masses(1, n)
s = mb.System(xyz, masses=m, box=M, vdw=(0.1 + 0.1 * m / 16).astype(np.float32))
p, d = mb.distance_search(1.2, s(), dims=[True] * 3)                      # cell kernel, pairs + dist
ids1 = np.arange(0, n, 2, dtype=np.uint64); ids2 = np.arange(1, n, 3, dtype=np.uint64)
s.set_option("two_set_cells_min", 0)
p2, d2 = mb.distance_search(1.0, s(ids1), s(ids2), dims=[True] * 3)       # two-set cell kernel
w = within(0.7, s(), s(np.arange(50, 90, dtype=np.uint64)), dims=7)       # within flags
s.set_option("two_set_cells_min", 1e30)
p3, d3 = mb.distance_search(1.0, s(ids1), s(ids2))                        # all-pairs kernel
p4, d4 = mb.distance_search("vdw", s(ids1), s(ids2), dims=[True] * 3)     # vdW
com, rg = s().com(), s().gyration()
ref = mb.System(orc.synth_frame(1, 1, n, M), masses=m)
tr = mb.fit_transform(s(), ref()); s().apply_transform(tr); r = mb.rmsd(s(), ref()); rw = mb.rmsd_mw(s(), ref())
t = mb.Trajectory(); t.synth(1, 0, 4, 8000, M, mass_seed=1)
c = t.search(1.2); c2 = t.search(1.2, count_only=True); rows = t.pipeline(1.2); rr = t.fit(0)
t.set_option("fused_fit", 1); rr2 = t.fit(0)
t.set_option("fused_fit", 3); rr3 = t.fit(0)  # persistent kernel: solver warps, teams, L2-served lag
t.set_option("fused_fit", 4); rr4 = t.fit(0); t.set_option("fused_fit", 0)  # warp-specialised kernel: TMA rings, mbarrier hand-overs
assert np.allclose(rr4, rr, rtol=1e-9)
print("ok", len(p), len(p2), len(w), len(p3), len(p4), c.tolist(), c2.tolist(), float(r))
# ---- round 2 additions: periodic reductions, inertia, trajectory ingest, pair-list consumers ----
from oracle import traj_oracle as T
s2 = mb.System(xyz, masses=m, box=M)
pl, dl = mb.distance_search(1.2, s2(), dims=[True] * 3)                   # pairs + dist
s2.set_option("with_dist", 0)
npl2 = mb._capi.check(s2._lib.mb_search_single(s2._h, 1.2, None, n, 7))    # pairs only
pl2 = np.empty((npl2, 2), np.uint64); mb._capi.check(s2._lib.mb_fill_pairs(s2._h, pl2.ctypes.data, None))
rp, cols = s2.connectivity()                                               # CSR adjacency
rp2, cols2 = s2().search_connectivity(1.2, dims=[True] * 3)                # neighbour rows written by the search kernel (modes 4, 5)
assert np.array_equal(rp, rp2) and len(cols2) == len(cols)
cp = s2().com(dims=[True] * 3); cg = s2().cog(dims=[True, False, True]); gp = s2().gyration(pbc=True)
mom, ax = s2().inertia(pbc=True); mom2, ax2 = s2().inertia(); ptr = s2().principal_transform()
sels = s2(np.arange(0, n, 2, dtype=np.uint64)).unwrap_connectivity(0.25)   # union-find + persistent BFS kernel
fr = np.stack([xyz, xyz[::-1].copy(), xyz])
tt = mb.Trajectory(); cs = tt.stream_search(fr, 1.2, M); rf = tt.stream_fit(fr, m); pp = tt.stream_pipeline(fr, 1.2, M, masses=m)
ang = T.write_dcd(fr, boxes=[np.array([50.0, 0, 50.0, 0, 0, 50.0])] * 3, fixed=[1, 5, 77])
td = mb.load_trajectory(ang, "dcd"); fd = td.frames()
xb = T.write_xtc_frame(np.abs(xyz), M) + T.write_xtc_frame(np.abs(xyz[:7]), M)[:0]
tx = mb.load_trajectory(xb, "xtc"); fx = tx.frames()
print("ok2", len(pl), len(pl2), len(cols), cp.tolist(), float(gp), mom.tolist(), len(sels), cs.tolist(), fd.shape, fx.shape)
assert len(pl) == len(p) and len(pl2) == len(p)
