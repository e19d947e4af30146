"""GPU parity: neighbour search through the C ABI vs the CPU oracle.  Bit-exact pair sets
(sorted (min,max) sets) and bit-exact f32 distances."""
import os

import numpy as np
import pytest

from oracle import oracle_py as orc
from tests.helpers import SEED, TRIC, assert_same_pairs, gpu_canonical, oracle_single

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mb():
    import molar_b200
    return molar_b200


def run_single(mb, xyz, cutoff, box=None, dims=None, ids=None, **opts):
    s = mb.System(xyz, box=box)
    for k, v in opts.items():
        s.set_option(k, v)
    sel = s(ids) if ids is not None else s()
    pairs, dist = mb.distance_search(cutoff, sel, dims=dims)
    s.close()
    if dist is None:  # with_dist=0: the pairs-only kernel
        p = orc.canonical_pairs(pairs)
        assert len(p) == len(pairs), "GPU pair list contains duplicates"
        return p, None
    return gpu_canonical(pairs, dist)


def run_single_both(mb, xyz, cutoff, op, od, **kw):
    """The same search through the pairs + distances kernel (MODE 1) and the pairs-only kernel (MODE 0: fused filter
    + exact re-evaluation, per-home emission, bulk flush): both must give the oracle's pair set."""
    gp, gd = run_single(mb, xyz, cutoff, **kw)
    assert_same_pairs(gp, gd, op, od)
    gp0, _ = run_single(mb, xyz, cutoff, with_dist=0, **kw)
    assert gp0.shape == op.shape and np.array_equal(gp0, op), "pairs-only kernel: pair sets differ"


ORTHO_SMALL = np.diag([4.0, 4.4, 5.1]).astype(np.float32)


def test_single_pbc_small_general_kernel(mb):
    xyz = orc.synth_frame(SEED, 0, 1200, ORTHO_SMALL)
    op, od, dims = oracle_single(1.2, xyz, box=ORTHO_SMALL, pbc=7)
    gp, gd = run_single(mb, xyz, 1.2, box=ORTHO_SMALL, dims=[True] * 3)
    assert_same_pairs(gp, gd, op, od)


@pytest.mark.parametrize("diag", [(2.0, 2.9, 1.3), (1.3, 5.0, 2.5), (3.7, 1.0, 6.0)])
def test_single_pbc_degenerate_grid(mb, diag):
    # dims of 1 and 2: the same cell pair is reached directly AND wrapped (SURVEY §7 hard part 3)
    M = np.diag(diag).astype(np.float32)
    xyz = orc.synth_frame(SEED + 1, 0, 700, M, stray_permille=30)
    op, od, dims = oracle_single(1.2, xyz, box=M, pbc=7)
    assert min(dims) <= 2
    gp, gd = run_single(mb, xyz, 1.2, box=M, dims=[True] * 3)
    assert gp.shape == op.shape and np.array_equal(gp, op)
    # a degenerate pair can be reported with several distances by the reference; we keep the minimum
    assert np.array_equal(gd, od)


@pytest.mark.parametrize("subdiv", [0, 1, 2, 3])
def test_single_pbc_cells_orthorhombic(mb, subdiv):
    M = np.diag([6.0, 7.0, 8.0]).astype(np.float32)
    xyz = orc.synth_frame(SEED + 2, 0, 30000, M, stray_permille=10)
    op, od, dims = oracle_single(1.2, xyz, box=M, pbc=7)
    assert list(dims) == [5, 5, 6]
    run_single_both(mb, xyz, 1.2, op, od, box=M, dims=[True] * 3, subdiv=subdiv)


@pytest.mark.parametrize("with_dist", [0, 1])
def test_single_pbc_lattice_ties_at_cutoff(mb, with_dist):
    """Atoms on a 0.3 nm lattice, cutoff 1.5 nm = 5 lattice steps: thousands of pairs sit within a few ulps of the
    cutoff ((5,0,0), (3,4,0), ... in lattice units, each coordinate an f32 rounding of k * 0.3).  The fused-multiply-add
    filter of the pair kernel cannot decide those; they must come out as the reference's unfused expression decides
    them (with_dist=0: filter + exact re-evaluation of the step; with_dist=1: the MODE 1 kernel)."""
    k = 30
    g = (np.arange(k, dtype=np.float64) * 0.3).astype(np.float32)
    xyz = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    M = np.diag([k * 0.3] * 3).astype(np.float32)
    op, od, dims = oracle_single(1.5, xyz, box=M, pbc=7)
    d2 = od.astype(np.float64) ** 2
    assert (np.abs(d2 - 2.25) < 1e-5).sum() > 1000  # the case really has ties at the cutoff
    if with_dist:
        gp, gd = run_single(mb, xyz, 1.5, box=M, dims=[True] * 3)
        assert_same_pairs(gp, gd, op, od)
    else:
        gp, _ = run_single(mb, xyz, 1.5, box=M, dims=[True] * 3, with_dist=0)
        assert gp.shape == op.shape and np.array_equal(gp, op)


@pytest.mark.parametrize("subdiv", [0, 1, 2])
def test_single_pbc_cells_triclinic_config3_shape(mb, subdiv):
    M = (TRIC * np.float32(0.3)).astype(np.float32)
    xyz = orc.synth_frame(SEED + 3, 0, 27000, M, stray_permille=10)
    op, od, dims = oracle_single(1.2, xyz, box=M, pbc=7)
    run_single_both(mb, xyz, 1.2, op, od, box=M, dims=[True] * 3, subdiv=subdiv)


def test_single_pbc_cells_triclinic_adversarial_positive_shear(mb):
    # SURVEY §7 hard part 2: the reference grid misses pairs here; parity = bug-compatible
    L = 6.0
    M = np.array([[L, 0, L / 2], [0, L, L / 2], [0, 0, L / np.sqrt(2)]], np.float32)
    xyz = orc.synth_frame(SEED + 4, 0, 22000, M, stray_permille=10)
    op, od, dims = oracle_single(1.2, xyz, box=M, pbc=7)
    gp, gd = run_single(mb, xyz, 1.2, box=M, dims=[True] * 3)
    assert_same_pairs(gp, gd, op, od)
    gp2, gd2 = run_single(mb, xyz, 1.2, box=M, dims=[True] * 3, force_brute=1)
    assert_same_pairs(gp2, gd2, op, od)


def test_single_pbc_reference_triclinic_fixture(mb, golden_dir):
    # molar/tests/triclinic.pdb (CRYST1 93.753^3, 60/60/90), every 9th atom
    z = np.load(os.path.join(golden_dir, "triclinic_sub.npz"))
    a, b, c, al, be, ga = [float(v) for v in z["cryst1"]]
    box = orc.Box(vectors_angles=(a, b, c, al, be, ga))
    M = box.matrix
    xyz = z["xyz"]
    ij, d, dims = orc.search_single(1.0, xyz, None, box, 7, 4)
    op, od = orc.canonical_pairs(ij, d)
    gp, gd = run_single(mb, xyz, 1.0, box=M, dims=[True] * 3)
    assert_same_pairs(gp, gd, op, od)


@pytest.mark.parametrize("dims", [[True, True, False], [False, True, False], [True, False, True]])
def test_single_partial_pbc(mb, dims):
    M = np.diag([6.0, 7.0, 8.0]).astype(np.float32)
    xyz = orc.synth_frame(SEED + 5, 0, 12000, M, stray_permille=40)
    pbc = sum(1 << i for i in range(3) if dims[i])
    op, od, gd_ = oracle_single(1.2, xyz, box=M, pbc=pbc)
    run_single_both(mb, xyz, 1.2, op, od, box=M, dims=dims)


def test_single_nonperiodic_config1_2lao(mb, golden_dir):
    # BASELINE config 1: 2k-atom protein, distance_search within 1.0 nm + rmsd to self
    xyz = np.load(os.path.join(golden_dir, "2lao.npz"))["xyz"]
    op, od, dims = oracle_single(1.0, xyz)
    gp, gd = run_single(mb, xyz, 1.0)
    assert_same_pairs(gp, gd, op, od)
    s = mb.System(xyz)
    assert mb.rmsd(s(), s()) == 0.0
    s.close()


def test_single_nonperiodic_cells(mb):
    M = np.diag([6.0, 7.0, 8.0]).astype(np.float32)
    xyz = orc.synth_frame(SEED + 6, 0, 25000, M) - np.float32(2.5)  # negative coordinates too
    op, od, dims = oracle_single(0.9, xyz)
    run_single_both(mb, xyz, 0.9, op, od)


def test_single_with_selection_ids(mb):
    M = np.diag([6.0, 7.0, 8.0]).astype(np.float32)
    xyz = orc.synth_frame(SEED + 7, 0, 30000, M)
    ids = np.arange(0, 30000, 2, dtype=np.uint64)
    op, od, dims = oracle_single(1.2, xyz, ids=ids, box=M, pbc=7)
    gp, gd = run_single(mb, xyz, 1.2, box=M, dims=[True] * 3, ids=ids)
    assert_same_pairs(gp, gd, op, od)
    assert set(np.unique(gp).tolist()) <= set(ids.tolist())


def test_double_search_vs_oracle(mb):
    M = np.diag([5.0, 5.5, 6.0]).astype(np.float32)
    xyz = orc.synth_frame(SEED + 8, 0, 6000, M, stray_permille=10)
    ids1 = np.arange(0, 6000, 3, dtype=np.uint64)
    ids2 = np.arange(1, 6000, 2, dtype=np.uint64)
    box = orc.Box(matrix=M)
    for pbc in (0, 7):
        ij, d, dims = orc.search_double(1.0, xyz, ids1, xyz, ids2, box if pbc else None, pbc, 4)
        op, od = orc.ordered_pairs(ij, d)
        s = mb.System(xyz, box=M)
        pairs, dist = mb.distance_search(1.0, s(ids1), s(ids2), dims=[bool(pbc)] * 3)
        s.close()
        gp, gd = orc.ordered_pairs(pairs, dist)
        assert len(gp) == len(pairs)
        assert_same_pairs(gp, gd, op, od)


@pytest.mark.parametrize("pbc", [0, 7, 3])
@pytest.mark.parametrize("tric", [False, True])
def test_double_and_within_cell_path_vs_oracle(mb, pbc, tric):
    """Two-set searches large enough for the cell kernel (home tiles from set 1, candidate stream from
    set 2, full shell): ordered pair set + distances, and the `within` id set, against the oracle."""
    from molar_b200.api import within
    M = (TRIC * np.float32(0.33)).astype(np.float32) if tric else np.diag([6.0, 7.0, 8.0]).astype(np.float32)
    n = 40000
    xyz = orc.synth_frame(SEED + 31, 0, n, M, stray_permille=10)
    ids1 = np.arange(0, n, 2, dtype=np.uint64)
    ids2 = np.arange(1, n, 3, dtype=np.uint64)
    box = orc.Box(matrix=M)
    ij, d, dims = orc.search_double(1.0, xyz, ids1, xyz, ids2, box if pbc else None, pbc, 8)
    op, od = orc.ordered_pairs(ij, d)
    s = mb.System(xyz, box=M)
    s.set_option("two_set_cells_min", 0)
    pairs, dist = mb.distance_search(1.0, s(ids1), s(ids2), dims=[bool(pbc & 1), bool(pbc & 2), bool(pbc & 4)])
    gp, gd = orc.ordered_pairs(pairs, dist)
    assert len(gp) == len(pairs)
    assert_same_pairs(gp, gd, op, od)
    # the same search through the pairs-only kernel (MODE 0)
    s.set_option("with_dist", 0)
    pairs0, none = mb.distance_search(1.0, s(ids1), s(ids2), dims=[bool(pbc & 1), bool(pbc & 2), bool(pbc & 4)])
    s.set_option("with_dist", 1)
    assert none is None and len(pairs0) == len(op)
    k0 = np.sort((pairs0[:, 0].astype(np.uint64) << np.uint64(32)) | pairs0[:, 1].astype(np.uint64))
    ko = np.sort((op[:, 0].astype(np.uint64) << np.uint64(32)) | op[:, 1].astype(np.uint64))
    assert np.array_equal(k0, ko)
    # within: small inner set, periodic and non-periodic
    inner = np.arange(100, 400, dtype=np.uint64)
    if pbc:
        ref_ids = orc.search_within(0.7, xyz, None, xyz, inner, box=box, pbc=pbc, nthreads=4)
    else:
        lo, up = orc.within_bounds(0.7, xyz, None)
        ref_ids = orc.search_within(0.7, xyz, None, xyz, inner, lower=lo, upper=up, nthreads=4)
    got = within(0.7, s(), s(inner), dims=pbc)
    assert np.array_equal(got, np.unique(ref_ids))
    s.close()


@pytest.mark.parametrize("case", ["within_0.5_resid555", "within_0.5_pbc_resid555"])
def test_within_golden_vectors_through_cell_path(mb, golden_dir, case):
    from molar_b200.api import within
    z = np.load(os.path.join(golden_dir, "albumin_within.npz"))
    cutoff, pbc = z[case + "_params"]
    s = mb.System(z["xyz"], box=z["box9"].reshape(3, 3).T)
    s.set_option("two_set_cells_min", 0)
    got = within(float(np.float32(cutoff)), s(), s(z[case + "_inner"].astype(np.uint64)), dims=int(pbc))
    s.close()
    assert np.array_equal(got.astype(np.int64), z[case + "_answer"])


@pytest.mark.parametrize("pbc", [0, 7])
def test_double_vdw_search_vs_oracle(mb, pbc):
    """distance_search_double_vdw[_pbc] (distance_search.rs:767-879) through the pymolar-style
    `distance_search("vdw", sel1, sel2)`: local indices converted to global like pymolar does."""
    M = np.diag([4.0, 4.2, 4.4]).astype(np.float32)
    n = 9000
    xyz = orc.synth_frame(SEED + 21, 0, n, M, stray_permille=10)
    rng = np.random.default_rng(3)
    vdw = (0.12 + 0.1 * rng.random(n)).astype(np.float32)
    ids1 = np.arange(0, n, 2, dtype=np.uint64)
    ids2 = np.arange(1, n, 3, dtype=np.uint64)
    box = orc.Box(matrix=M)
    ij, d, dims = orc.search_double_vdw(xyz, ids1, vdw[ids1.astype(int)], xyz, ids2, vdw[ids2.astype(int)],
                                        box if pbc else None, pbc, 4)
    op, od = orc.ordered_pairs(ij, d)
    op = np.stack([ids1[op[:, 0].astype(int)], ids2[op[:, 1].astype(int)]], 1)
    s = mb.System(xyz, box=M, vdw=vdw)
    pairs, dist = mb.distance_search("vdw", s(ids1), s(ids2), dims=[bool(pbc)] * 3)
    s.close()
    gp, gd = orc.ordered_pairs(pairs, dist)
    assert len(gp) == len(pairs)
    order = np.lexsort((op[:, 1], op[:, 0]))
    assert_same_pairs(gp, gd, op[order], od[order])


@pytest.mark.parametrize("pbc", [0, 7])
def test_double_vdw_search_cell_path_100k_vs_oracle(mb, pbc):
    """vdW contact search at 10^5 x 10^5 atoms: the cell kernel with a per-pair cutoff (grid cutoff = max + max + eps,
    distance_search.rs:781-783,845-847) against the oracle, and its time against the plain two-set search at the same
    grid cutoff (must be within 2x: it was an all-pairs scan before)."""
    import time
    M = (TRIC * np.float32(0.47)).astype(np.float32)  # ~103k atoms at water density
    n = 200_000
    xyz = orc.synth_frame(SEED + 41, 0, n, M, stray_permille=5)
    rng = np.random.default_rng(9)
    vdw = (0.10 + 0.11 * rng.random(n)).astype(np.float32)
    ids1 = np.arange(0, n, 2, dtype=np.uint64)
    ids2 = np.arange(1, n, 2, dtype=np.uint64)
    box = orc.Box(matrix=M)
    ij, d, dims = orc.search_double_vdw(xyz, ids1, vdw[ids1.astype(int)], xyz, ids2, vdw[ids2.astype(int)],
                                        box if pbc else None, pbc, 8)
    op, od = orc.ordered_pairs(ij, d)
    op = np.stack([ids1[op[:, 0].astype(int)], ids2[op[:, 1].astype(int)]], 1)
    order = np.lexsort((op[:, 1], op[:, 0]))
    s = mb.System(xyz, box=M, vdw=vdw)
    pairs, dist = mb.distance_search("vdw", s(ids1), s(ids2), dims=[bool(pbc)] * 3)
    gp, gd = orc.ordered_pairs(pairs, dist)
    assert len(gp) == len(pairs) and len(gp) > 10000
    assert_same_pairs(gp, gd, op[order], od[order])
    # timing: vdW vs plain two-set search with the grid cutoff of the vdW search
    grid_cut = float(vdw[ids1.astype(int)].max() + vdw[ids2.astype(int)].max() + np.finfo(np.float32).eps)
    s.set_option("with_dist", 0)

    def timed(fn):
        fn()
        t0 = time.perf_counter()
        for _ in range(3):
            fn()
        return (time.perf_counter() - t0) / 3

    from molar_b200 import _capi
    p1, p2 = ids1.ctypes.data_as(_capi.u64p), ids2.ctypes.data_as(_capi.u64p)
    v1 = np.ascontiguousarray(vdw[ids1.astype(int)])
    v2 = np.ascontiguousarray(vdw[ids2.astype(int)])
    f32p = _capi.f32p
    t_vdw = timed(lambda: _capi.check(s._lib.mb_search_double_vdw(s._h, p1, len(ids1), v1.ctypes.data_as(f32p), p2,
                                                                   len(ids2), v2.ctypes.data_as(f32p), 0, pbc)))
    t_plain = timed(lambda: _capi.check(s._lib.mb_search_double(s._h, grid_cut, p1, len(ids1), p2, len(ids2), 0, pbc)))
    s.close()
    assert t_vdw < 2.0 * t_plain + 2e-3, (t_vdw, t_plain)


CASES = ["within_0.5_resid10", "within_0.3_resid20", "within_0.5_resid555", "within_0.5_pbc_resid555"]


@pytest.mark.parametrize("case", CASES)
def test_within_reference_golden_vectors(mb, golden_dir, case):
    """The reference's own golden answers (molar/tests/generated_vmd_tests.in:27,35;
    generated_pteros_tests.in:21,27) reproduced by the CUDA path."""
    from molar_b200.api import within
    z = np.load(os.path.join(golden_dir, "albumin_within.npz"))
    cutoff, pbc = z[case + "_params"]
    M = z["box9"].reshape(3, 3).T
    s = mb.System(z["xyz"], box=M)
    inner = s(z[case + "_inner"].astype(np.uint64))
    got = within(float(np.float32(cutoff)), s(), inner, dims=int(pbc))
    s.close()
    assert np.array_equal(got.astype(np.int64), z[case + "_answer"])


def test_empty_and_error_paths(mb):
    M = np.diag([6.0, 7.0, 8.0]).astype(np.float32)
    xyz = orc.synth_frame(SEED, 0, 100, M)
    s = mb.System(xyz)  # no box
    with pytest.raises(mb.MolarB200Error) as e:
        mb.distance_search(1.0, s(), dims=[True, True, True])
    assert e.value.code == -4  # NoPbc
    pairs, dist = mb.distance_search(1e-4, s())  # nothing within cutoff
    assert pairs.shape == (0, 2) and dist.shape == (0,)
    with pytest.raises(IndexError):
        s([1000])
    # the C ABI is the boundary a Rust binding calls directly: unsorted / repeated / out-of-range ids are rejected
    # there (MB_ERR_ARG) instead of reaching a kernel
    from molar_b200 import _capi
    for bad in ([5, 3, 9], [1, 1, 2], [0, 50, 100], [2 ** 40, 2 ** 40 + 1]):
        ids = np.asarray(bad, np.uint64)
        rc = s._lib.mb_search_single(s._h, 1.0, ids.ctypes.data_as(_capi.u64p), len(ids), 0)
        assert rc == _capi.MB_ERR_ARG, (bad, rc)
        out = (np.zeros(3))
        rc = s._lib.mb_center_of_geometry(s._h, ids.ctypes.data_as(_capi.u64p), len(ids), out.ctypes.data_as(_capi.f64p))
        assert rc == _capi.MB_ERR_ARG, (bad, rc)
    s.close()


# ---- full-size configs (BASELINE.json configs[1], configs[2]) ---------------------------------

def test_config2_100k_orthorhombic_exact(mb):
    M = np.diag([10.0, 10.0, 10.0]).astype(np.float32)
    xyz = orc.synth_frame(SEED, 0, 100000, M)
    op, od, dims = oracle_single(1.2, xyz, box=M, pbc=7, nthreads=8)
    assert list(dims) == [8, 8, 8]
    gp, gd = run_single(mb, xyz, 1.2, box=M, dims=[True] * 3)
    assert_same_pairs(gp, gd, op, od)


def test_batch_search_streams_and_last_pairs(mb):
    """Frames alternate over several stream slots; counts, checksums and the pair list of the last
    frame must not depend on the number of slots, and must equal the oracle's."""
    M = (TRIC * np.float32(0.4)).astype(np.float32)
    n, nf = 60000, 5
    res = {}
    for ns in (1, 2, 3):
        t = mb.Trajectory()
        t.synth(SEED, 0, nf, n, M, stray_permille=5)
        t.set_option("batch_streams", ns)
        counts = t.search(1.2)
        pairs = orc.canonical_pairs(t.last_pairs())
        counts2, chk = t.search(1.2, checksums=True)
        assert np.array_equal(counts, counts2)
        res[ns] = (counts, chk, pairs)
        t.close()
    for ns in (2, 3):
        assert np.array_equal(res[ns][0], res[1][0]) and np.array_equal(res[ns][1], res[1][1])
        assert np.array_equal(res[ns][2], res[1][2])
    xyz = orc.synth_frame(SEED, nf - 1, n, M, stray_permille=5)
    op, od, dims = oracle_single(1.2, xyz, box=M, pbc=7, nthreads=8)
    assert np.array_equal(res[2][2], op) and res[2][0][-1] == len(op)


def test_config3_1m_triclinic_properties(mb):
    """1M atoms: too slow for the oracle inside the test budget, so size-independent properties:
    count-only == enumerated count; the order-independent checksum and count do not depend on the
    traversal subdivision (k=1 walks exactly the reference's own cell pairs); the device generator
    produces the oracle's bits."""
    n = 1_000_000
    t = mb.Trajectory()
    t.synth(SEED, 0, 2, n, TRIC)
    assert np.array_equal(t.frame(1)[:5000], orc.synth_frame(SEED, 1, 5000, TRIC))
    counts, chk = t.search(1.2, checksums=True)
    counts_only = t.search(1.2, count_only=True)
    assert np.array_equal(counts, counts_only)
    expect = n * (n / abs(np.linalg.det(TRIC.astype(np.float64)))) * (4 / 3) * np.pi * 1.2 ** 3 / 2
    assert np.all(np.abs(counts - expect) / expect < 0.01)
    t.set_option("subdiv", 1)
    counts1, chk1 = t.search(1.2, f0=0, f1=1, checksums=True)
    assert counts1[0] == counts[0] and np.array_equal(chk1[0], chk[0])
    # the shifted-image filter for wrapped cell pairs must decide exactly like the reference's
    # PeriodicBox::distance_squared evaluated for every pair
    t.set_option("subdiv", 0)
    t.set_option("exact_pbc", 1)
    counts2, chk2 = t.search(1.2, f0=0, f1=2, checksums=True)
    assert np.array_equal(counts2, counts) and np.array_equal(chk2, chk)
    t.close()


@pytest.mark.parametrize("stray", [0, 10])
def test_config3_1m_triclinic_vs_oracle_checksum(mb, stray):
    """The headline configuration at FULL size against the oracle: one 1M-atom frame of the config-3 triclinic box
    (13 x 15 x 17 reference grid), with and without 1 % stray atoms.  The oracle hashes every pair it emits inside its
    worker pool (count, sum and xor of mix64((min << 32) | max) — the hash of checksum_kernel), so the 3.6e8 pairs
    never have to be materialised on the host.  distance_search.rs:928-954."""
    n = 1_000_000
    xyz = orc.synth_frame(SEED, 3, n, TRIC, stray_permille=stray)
    cnt, ssum, sxor, dims = orc.search_single_pbc_checksum(1.2, xyz, orc.Box(matrix=TRIC), 7, nthreads=os.cpu_count() or 4)
    assert dims == [13, 15, 17]
    t = mb.Trajectory()
    t.synth(SEED, 3, 1, n, TRIC, stray_permille=stray)
    assert np.array_equal(t.frame(0), xyz)
    counts, chk = t.search(1.2, checksums=True)
    assert int(counts[0]) == cnt
    assert int(chk[0, 0]) == ssum and int(chk[0, 1]) == sxor
    # the count-only kernel (config 5's contact count) at full size
    assert int(t.search(1.2, count_only=True)[0]) == cnt
    # ... and through the per-call entry point with the pair list checksummed on the device
    t.close()
    import ctypes as C
    s = mb.System(xyz, box=TRIC)
    s.set_option("with_dist", 0)
    assert s._lib.mb_search_single(s._h, 1.2, None, n, 7) == cnt
    out2 = (C.c_uint64 * 2)()
    assert s._lib.mb_pairs_checksum(s._h, out2) == 0
    assert (int(out2[0]), int(out2[1])) == (ssum, sxor)
    # ... and as neighbour rows written by the search kernel itself (full shell, count pass + fill pass): 2 x cnt
    # entries, and each half of the adjacency (row < column, row > column) hashes to the oracle's pair list
    assert s._lib.mb_search_connectivity(s._h, C.c_float(1.2), None, n, 7, None) == 2 * cnt
    assert s.connectivity_checksum() == (ssum, sxor, ssum, sxor)
    s.close()


def test_single_pbc_dense_system_multipass_emission(mb):
    """450 atoms/nm^3: one 64-candidate step finds more pairs than the per-warp staging buffer holds,
    so the masks are expanded in several passes."""
    M = np.diag([3.7, 3.8, 3.9]).astype(np.float32)
    xyz = orc.synth_frame(SEED + 13, 0, 24000, M, stray_permille=10)
    op, od, dims = oracle_single(1.2, xyz, box=M, pbc=7, nthreads=8)
    assert list(dims) == [3, 3, 3]
    for opts in ({}, {"subdiv": 1}, {"with_dist": 0}):
        s = mb.System(xyz, box=M)
        for k, v in opts.items():
            s.set_option(k, v)
        if opts.get("with_dist", 1):
            pairs, dist = mb.distance_search(1.2, s(), dims=[True] * 3)
            gp, gd = gpu_canonical(pairs, dist)
            assert_same_pairs(gp, gd, op, od)
        else:
            cnt = mb._capi.check(s._lib.mb_search_single(s._h, 1.2, None, len(xyz), 7))
            pairs = np.empty((cnt, 2), np.uint64)
            mb._capi.check(s._lib.mb_fill_pairs(s._h, pairs.ctypes.data, None))
            gp = orc.canonical_pairs(pairs)
            assert len(gp) == cnt and np.array_equal(gp, op)
        s.close()


_BIG_BOX_ORACLE = {}


def _big_box_case(which):
    """oracle answer of the two big-box cases, computed once for both parametrisations"""
    if which not in _BIG_BOX_ORACLE:
        if which == "tric":
            M = (TRIC * np.float32(0.45)).astype(np.float32)
            xyz = orc.synth_frame(SEED + 11, 0, 90000, M, stray_permille=20)
        else:
            M = np.diag([9.7, 10.3, 11.1]).astype(np.float32)
            xyz = orc.synth_frame(SEED + 12, 0, 100000, M, stray_permille=20)
        op, od, dims = oracle_single(1.2, xyz, box=M, pbc=7, nthreads=8)
        _BIG_BOX_ORACLE[which] = (M, xyz, op, od, dims)
    return _BIG_BOX_ORACLE[which]


@pytest.mark.parametrize("exact", [0, 1])
def test_single_pbc_big_box_filter_vs_exact_path(mb, exact):
    """Boxes large enough for the wrapped-pair filter (>= 4 reference cells per dim), strays
    included, against the oracle."""
    for which in ("tric", "ortho"):
        M, xyz, op, od, dims = _big_box_case(which)
        assert min(dims) >= 4
        gp, gd = run_single(mb, xyz, 1.2, box=M, dims=[True] * 3, exact_pbc=exact)
        assert_same_pairs(gp, gd, op, od)


def test_stream_search_host_frames(mb):
    """mb_stream_search: host frames uploaded chunk by chunk while the previous chunk is searched; counts and the
    last frame's pair list equal the per-call path."""
    n, nf = 60_000, 11
    box = np.diag([8.0, 9.0, 8.5]).astype(np.float32)
    frames = np.stack([orc.synth_frame(SEED + 3, f, n, box, stray_permille=5) for f in range(nf)])
    traj = mb.Trajectory()
    counts = traj.stream_search(frames, 1.2, box)
    s = mb.System(frames[0], box=box)
    for f in (0, 4, nf - 1):
        s.set_state(frames[f], box)
        pairs, dist = mb.distance_search(1.2, s(), dims=[True, True, True])
        assert counts[f] == len(pairs)
    lp = traj.last_pairs()
    got = np.unique(np.sort(lp, axis=1), axis=0)
    want = np.unique(np.sort(np.asarray(pairs, dtype=np.uint64), axis=1), axis=0)
    assert np.array_equal(got, want)
    counts2 = traj.stream_search(frames, 1.2, box, count_only=True)
    assert np.array_equal(counts, counts2)
    traj.close()
    s.close()
