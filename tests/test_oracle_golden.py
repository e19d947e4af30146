"""Pin the CPU oracle against the reference's own golden vectors and known-answer tests.

Golden `within` vectors: molar/tests/generated_vmd_tests.in:27,35 and
molar/tests/generated_pteros_tests.in:21,27 (evaluated by the reference through
selection/ast.rs:589-631 -> distance_search_within[_pbc], distance_search.rs:519-598).
PeriodicBox KATs: molar/src/periodic_box.rs:456-620.
"""
import os

import numpy as np
import pytest

from oracle import oracle_py as orc


@pytest.fixture(scope="module")
def albumin(golden_dir):
    return np.load(os.path.join(golden_dir, "albumin_within.npz"))


CASES = ["within_0.5_resid10", "within_0.3_resid20", "within_0.5_resid555", "within_0.5_pbc_resid555"]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("nthreads", [1, 4])
def test_within_golden(albumin, case, nthreads):
    xyz = albumin["xyz"]
    cutoff, pbc = albumin[case + "_params"]
    cutoff = float(np.float32(cutoff))
    pbc = int(pbc)
    inner = albumin[case + "_inner"]
    answer = albumin[case + "_answer"]
    box = orc.Box(matrix=albumin["box9"].reshape(3, 3).T)
    if pbc:
        ids = orc.search_within(cutoff, xyz, None, xyz, inner, box=box, pbc=pbc, nthreads=nthreads)
    else:
        lo, up = orc.within_bounds(cutoff, xyz, None)
        ids = orc.search_within(cutoff, xyz, None, xyz, inner, lower=lo, upper=up, nthreads=nthreads)
    got = np.unique(ids.astype(np.int64))
    assert np.array_equal(got, answer)


def test_within_golden_counts(albumin):
    # the four id-vector lengths quoted in SURVEY.md §4
    assert [len(albumin[c + "_answer"]) for c in CASES] == [204, 54, 221, 224]


# ---- PeriodicBox KATs (periodic_box.rs:456-620) -------------------------------------------

def _diag_box():
    return orc.Box(matrix=np.diag([10.0, 10.0, 10.0]))


@pytest.mark.parametrize("dims,expect", [
    (0, (8.0, 8.0, 8.0)),        # :457-464
    (7, (-2.0, -2.0, -2.0)),     # :467-474
    (1, (-2.0, 8.0, 8.0)),       # :477-485
    (3, (-2.0, -2.0, 8.0)),      # :488-496
])
def test_shortest_vector_dims_orthogonal(dims, expect):
    r = _diag_box().shortest_vector([8.0, 8.0, 8.0], dims)
    assert np.linalg.norm(r - np.asarray(expect, np.float32)) < 1e-6


def test_shortest_vector_pymolar_known_answer():
    """The reference's Python test suite holds one more known answer (molar_python/tests/test_2.py:233-245): box
    PeriodicBox([1, 2, 3], [90, 90, 90]) — a diagonal matrix by from_vectors_angles (periodic_box.rs:229-232) — maps
    (0.9, 0.5, 0.6) onto (-0.1, 0.5, 0.6), abs 1e-6."""
    r = orc.Box(matrix=np.diag([1.0, 2.0, 3.0])).shortest_vector([0.9, 0.5, 0.6], 7)
    assert np.allclose(r, [-0.1, 0.5, 0.6], rtol=0, atol=1e-6)


def _axes_match(axes_cols, ref_rows_descending, tol):
    """columns of axes_cols (ascending moments) against the reference's three axes (listed by descending moment);
    an axis is defined up to its sign"""
    ref = np.asarray(ref_rows_descending)[::-1]
    return all(min(np.abs(axes_cols[:, k] - ref[k]).max(), np.abs(axes_cols[:, k] + ref[k]).max()) < tol for k in range(3))


def test_inertia_axes_reference_known_answer(golden_dir):
    """The one known answer the reference holds for the measure path: molar/src/selection.rs:198-213 (`test_inertia`,
    all 4295 atoms of tests/protein.pdb) keeps the three inertia axes it expects as a comment (:209-211).  The oracle's
    inertia restatement (masses from the element column, mass-weighted centre, tensor, symmetric eigenproblem,
    measure.rs:88-98,573-610) reproduces them to the precision those numbers carry (5e-6: they come from an f32
    build), in every precision mode.  Fixture: tools/make_golden_inertia.py."""
    g = np.load(os.path.join(golden_dir, "protein_inertia.npz"))
    for prec in ("f64", "f32", "mixed"):
        rc, tensor, mom, axes, centre = orc.inertia(g["xyz"], g["masses"], box=None, prec=prec)
        assert rc == 0 and mom[0] < mom[1] < mom[2]
        assert _axes_match(axes, g["ref_axes"], 1e-5), prec


def test_orthogonal_has_no_tric_corrections():  # :546-551
    assert len(orc.Box(matrix=np.diag([10.0, 20.0, 30.0])).corrections) == 0


MDTRAJ_BOX = np.array([[10.0, 4.0, -4.0], [0.0, 10.0, 0.0], [0.0, 0.0, 10.0]])


def test_triclinic_mdtraj_box_matches_brute_force():  # :559-575
    b = orc.Box(matrix=MDTRAJ_BOX)
    d = np.sqrt(b.distance_squared([38.9214, 40.0078, -34.0795], [-26.6187, 40.8926, 30.9709], 7))
    assert abs(d - 5.353627) < 1e-3


def test_triclinic_corner_matches_brute_force():  # :580-603
    M = np.array([[6.0, 0.0, 3.0], [0.0, 6.0, 3.0], [0.0, 0.0, 6.0]])
    b = orc.Box(matrix=M)
    dx = np.array([2.9, 2.9, 2.9], np.float32)
    best = np.inf
    a, bb, c = M[:, 0], M[:, 1], M[:, 2]
    for i in range(-2, 3):
        for j in range(-2, 3):
            for k in range(-2, 3):
                best = min(best, np.linalg.norm(dx + i * a + j * bb + k * c))
    got = np.linalg.norm(b.shortest_vector(dx, 7))
    assert abs(got - best) < 1e-5


def test_triclinic_far_apart_reduction():  # :607-620
    b = orc.Box(matrix=MDTRAJ_BOX)
    d = np.sqrt(b.distance_squared([0.1, 0.2, 0.3], [60.1, 0.2, 0.3], 7))
    assert d < 1e-4


def test_invalid_from_vec_ang():  # :448-454
    with pytest.raises(ValueError):
        orc.Box(vectors_angles=(10.0, 0.2, 15.0, 90.0, 9.0, 90.0))


def test_inverse_orthorhombic_formula():
    # inv00 = (m11*m22)/(m00*(m11*m22)), not 1/m00
    m = np.array([7.7963, 9.4184, 11.0416], np.float32)
    b = orc.Box(matrix=np.diag(m))
    inv = b.inv
    minor = np.float32(m[1] * m[2])
    det = np.float32(m[0] * minor)
    assert inv[0, 0] == np.float32(minor / det)


# ---- search: restated cell-list vs brute force on the reference's own predicate -------------

def _brute_pairs(xyz, cutoff, box=None):
    x = xyz.astype(np.float64)
    n = len(x)
    out = []
    M = None if box is None else box.matrix.astype(np.float64)
    Minv = None if box is None else np.linalg.inv(M)
    for i in range(n - 1):
        d = x[i + 1:] - x[i]
        if box is not None:
            f = d @ Minv.T
            f -= np.round(f)
            d = f @ M.T
        r2 = (d * d).sum(1)
        for j in np.nonzero(r2 <= cutoff * cutoff)[0]:
            out.append((i, i + 1 + j, np.sqrt(r2[j])))
    return out


def test_single_pbc_orthorhombic_vs_bruteforce():
    M = np.diag([4.0, 4.4, 5.1]).astype(np.float32)
    xyz = orc.synth_frame(20260, 0, 1200, M)
    box = orc.Box(matrix=M)
    ij, d, dims = orc.search_single(1.2, xyz, box=box, pbc=7)
    assert list(dims) == [3, 3, 4]
    got, gd = orc.canonical_pairs(ij, d)
    bf = _brute_pairs(xyz, 1.2, box)
    # brute force is f64: allow disagreement only within a sliver around the cutoff
    gs = set(map(tuple, got.tolist()))
    bs = {(i, j) for i, j, r in bf}
    for (i, j, r) in bf:
        if (i, j) not in gs:
            assert abs(r - 1.2) < 1e-5
    assert len(gs - bs) <= 3 and len(bs - gs) <= 3
    assert len(gs) > 10000


def test_single_nonpbc_vs_bruteforce(golden_dir):
    xyz = np.load(os.path.join(golden_dir, "2lao.npz"))["xyz"]
    ij, d, dims = orc.search_single(1.0, xyz)
    got, gd = orc.canonical_pairs(ij, d)
    bf = _brute_pairs(xyz, 1.0)
    gs = set(map(tuple, got.tolist()))
    bs = {(i, j) for i, j, r in bf}
    assert len(gs ^ bs) <= 2
    # no duplicates in non-periodic search
    assert len(got) == len(ij)
    # config 1: rmsd to self is exactly 0
    rc, r = orc.rmsd(xyz, None, xyz, None, prec="f32")
    assert rc == 0 and r == 0.0


def test_double_vdw_vs_bruteforce():
    # distance_search_double_vdw[_pbc] (distance_search.rs:767-879): per-pair cutoff, LOCAL indices
    M = np.diag([3.0, 3.2, 3.4]).astype(np.float32)
    xyz = orc.synth_frame(1, 0, 3000, M)
    ids1 = np.arange(0, 3000, 2, dtype=np.uint64)
    ids2 = np.arange(1, 3000, 3, dtype=np.uint64)
    rng = np.random.default_rng(0)
    v1 = (0.1 + 0.1 * rng.random(len(ids1))).astype(np.float32)
    v2 = (0.12 + 0.08 * rng.random(len(ids2))).astype(np.float32)
    a = xyz[ids1.astype(int)].astype(np.float64)
    b = xyz[ids2.astype(int)].astype(np.float64)
    for pbc, box in ((0, None), (7, orc.Box(matrix=M))):
        ij, d, dims = orc.search_double_vdw(xyz, ids1, v1, xyz, ids2, v2, box, pbc, 2)
        got, gd = orc.ordered_pairs(ij, d)
        dd = a[:, None, :] - b[None, :, :]
        if pbc:
            L = np.diag(M).astype(np.float64)
            dd -= np.round(dd / L) * L
        r = np.sqrt((dd ** 2).sum(-1))
        cut = v1[:, None].astype(np.float64) + v2[None, :] + np.finfo(np.float32).eps
        bi, bj = np.nonzero(r <= cut)
        assert len(got) == len(bi)
        assert np.array_equal(got, np.stack([bi, bj], 1).astype(np.uint64))
        assert got[:, 0].max() < len(ids1) and got[:, 1].max() < len(ids2)  # local indices


def test_checksum_variant_equals_hash_of_materialised_pairs():
    """orc_search_single_pbc_checksum (used for the full-size config-3 parity test on the GPU) must be the
    count / sum / xor of mix64 over exactly the pairs orc_search_single_pbc materialises."""
    import numpy as np
    from oracle import oracle_py as orc
    M = (np.array([[21.5, -2.7, -2.7], [0.0, 21.5, -2.7], [0.0, 0.0, 21.5]], np.float32) * np.float32(0.3)).astype(np.float32)
    box = orc.Box(matrix=M)
    for stray in (0, 20):
        xyz = orc.synth_frame(20260, 1, 27000, M, stray_permille=stray)
        ij, d, dims = orc.search_single(1.2, xyz, None, box, 7, 4)
        want = orc.pairs_checksum(ij)
        for nt in (1, 4):
            cnt, ssum, sxor, dims2 = orc.search_single_pbc_checksum(1.2, xyz, box, 7, nthreads=nt)
            assert (cnt, ssum, sxor) == want and list(dims2) == [int(v) for v in dims]
    # the vectorised hash equals the scalar definition
    x = 0x0000000500000007
    def mix(v):
        m = (1 << 64) - 1
        v ^= v >> 33; v = (v * 0xff51afd7ed558ccd) & m; v ^= v >> 33; v = (v * 0xc4ceb9fe1a85ec53) & m; v ^= v >> 33
        return v
    assert int(orc.mix64(np.array([x], np.uint64))[0]) == mix(x)
