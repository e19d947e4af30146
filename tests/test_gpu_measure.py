"""GPU parity: reductions and Kabsch through the C ABI vs the f64 oracle.
Tolerance from BASELINE.json north_star: 1e-6 relative to the reference f64 path."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation

from oracle import oracle_py as orc
from tests.helpers import SEED, TRIC

pytestmark = pytest.mark.gpu
RTOL = 1e-6  # the stated tolerance; observed agreement is ~1e-12

M = np.diag([6.0, 7.0, 8.0]).astype(np.float32)


@pytest.fixture(scope="module")
def mb():
    import molar_b200
    return molar_b200


def _data(n, seed=0):
    return orc.synth_frame(SEED + seed, 0, n, M), orc.synth_masses(SEED + seed, n)


@pytest.mark.parametrize("n,ids", [(5000, None), (100003, None), (100000, "stride3"), (7, None), (1, None)])
def test_com_gyration(mb, n, ids):
    xyz, m = _data(n, 1)
    sel_ids = np.arange(0, n, 3, dtype=np.uint64) if ids else None
    s = mb.System(xyz, masses=m)
    sel = s(sel_ids) if sel_ids is not None else s()
    rc, com = orc.center_of_mass(xyz, m, sel_ids)
    rc, rg = orc.gyration(xyz, m, sel_ids)
    assert np.allclose(sel.com(), com, rtol=RTOL, atol=0)
    assert abs(sel.gyration() - rg) <= RTOL * max(rg, 1e-30) + 1e-12
    s.close()


def test_com_far_from_origin_small_rg(mb):
    xyz, m = _data(20000, 2)
    xyz = (xyz * np.float32(0.01) + np.float32(500.0)).astype(np.float32)
    s = mb.System(xyz, masses=m)
    rc, rg = orc.gyration(xyz, m)
    assert abs(s().gyration() - rg) / rg < RTOL
    s.close()


def test_zero_mass_and_sizes_errors(mb):
    xyz, m = _data(100, 3)
    s = mb.System(xyz, masses=np.zeros(100, np.float32))
    with pytest.raises(mb.MolarB200Error) as e:
        s().com()
    assert e.value.code == -1  # ZeroMass (measure.rs:70-71)
    with pytest.raises(mb.MolarB200Error) as e:
        mb.rmsd(s([0, 1, 2]), s([3, 4]))
    assert e.value.code == -2  # Sizes (measure.rs:494-496)
    s.close()


def test_rmsd_and_rmsd_mw(mb):
    a, m = _data(50000, 4)
    b = orc.synth_frame(999, 1, 50000, M)
    s1 = mb.System(a, masses=m)
    s2 = mb.System(b, masses=m)
    rc, r = orc.rmsd(a, None, b, None)
    rc, rw = orc.rmsd(a, None, b, None, masses1=m)
    assert abs(mb.rmsd(s1(), s2()) - r) / r < RTOL
    assert abs(mb.rmsd_mw(s1(), s2()) - rw) / rw < RTOL
    ids1 = np.arange(0, 30000, 2, dtype=np.uint64)
    ids2 = np.arange(10000, 25000, dtype=np.uint64)
    rc, r = orc.rmsd(a, ids1, b, ids2)
    assert abs(mb.rmsd(s1(ids1), s2(ids2)) - r) / r < RTOL
    s1.close()
    s2.close()


@pytest.mark.parametrize("angle", [80.0, 179.0, 3.0])
def test_fit_transform_apply_rmsd(mb, angle):
    # the reference's eigen_test scenario (selection.rs:148-172): rotate by 80 deg, fit, rmsd ~ 0
    xyz, m = _data(40000, 5)
    rot = Rotation.from_euler("zyx", [angle, 20.0, -35.0], degrees=True).as_matrix()
    moved = ((xyz.astype(np.float64) @ rot.T) + np.array([1.0, -2.0, 0.5])).astype(np.float32)
    s1 = mb.System(moved, masses=m)
    s2 = mb.System(xyz, masses=m)
    tr = mb.fit_transform(s1(), s2())
    rc, R, t = orc.fit_transform(moved, m, None, xyz, m, None)
    assert np.allclose(tr.R, R, rtol=0, atol=RTOL) and np.allclose(tr.t, t, rtol=RTOL, atol=1e-6 * np.abs(t).max())
    assert np.allclose(tr.R.T @ tr.R, np.eye(3), atol=1e-12) and abs(np.linalg.det(tr.R) - 1) < 1e-12
    s1().apply_transform(tr)
    fitted = s1.coords()
    expect = orc.apply_transform_f64(moved, None, R, t)
    assert np.allclose(fitted, expect, rtol=RTOL, atol=1e-6)
    assert mb.rmsd(s1(), s2()) < 1e-5
    s1.close()
    s2.close()


def test_fit_transform_subsets_separate_masses_and_origin(mb):
    xyz, m = _data(30000, 6)
    other = orc.synth_frame(4242, 3, 30000, M)
    ids1 = np.arange(100, 20100, dtype=np.uint64)
    ids2 = np.arange(5000, 25000, dtype=np.uint64)
    s1 = mb.System(xyz, masses=m)
    s2 = mb.System(other, masses=m)
    for at_origin in (False, True):
        tr = mb.fit_transform(s1(ids1), s2(ids2), at_origin=at_origin)
        rc, R, t = orc.fit_transform(xyz, m, ids1, other, m, ids2, at_origin=at_origin)
        assert rc == 0
        assert np.allclose(tr.R, R, atol=1e-6) and np.allclose(tr.t, t, atol=1e-5)
    s1.close()
    s2.close()


def test_fit_reflection(mb):
    xyz, m = _data(5000, 7)
    mir = xyz.copy()
    mir[:, 0] *= -1
    s1 = mb.System(mir, masses=m)
    s2 = mb.System(xyz, masses=m)
    tr = mb.fit_transform(s1(), s2())
    rc, R, t = orc.fit_transform(mir, m, None, xyz, m, None)
    assert abs(np.linalg.det(tr.R) - 1) < 1e-12 and np.allclose(tr.R, R, atol=1e-6)
    s1.close()
    s2.close()


def test_batch_fit_config4_shape(mb):
    """config 4 at reduced frame count: fit every frame onto frame 0, superpose, RMSD."""
    n, nf = 500_000, 6
    t = mb.Trajectory()
    t.synth(SEED, 0, nf, n, TRIC, mass_seed=SEED)
    m = orc.synth_masses(SEED, n)
    ref = orc.synth_frame(SEED, 0, n, TRIC)
    before = [t.frame(f) for f in range(nf)]
    assert np.array_equal(before[0], ref)
    r = t.fit(ref_frame=0, superpose=True)
    for f in (0, 1, nf - 1):
        rc, R, tt = orc.fit_transform(before[f], m, None, ref, m, None)
        exp = orc.apply_transform_f64(before[f], None, R, tt)
        exp_r = np.sqrt(((exp - ref) ** 2).sum(1).mean())
        assert abs(r[f] - exp_r) <= RTOL * exp_r + 1e-9
        assert np.allclose(t.frame(f), exp, rtol=RTOL, atol=2e-6)
    assert r[0] < 1e-6
    t.close()


@pytest.mark.parametrize("n,nf", [(100_000, 9), (99_999, 4), (2048, 5), (500_000, 13)])
def test_batch_fit_fused_vs_two_kernel_path(mb, n, nf):
    """The single-pass kernels (fused_fit=1: TMA ring, finisher role rotates over the slice CTAs; 2: TMA ring, worker CTAs
    + dedicated finisher CTAs; 3: one persistent kernel whose second pass lags behind and is served by L2) and the
    generic two-kernel path must agree with the oracle and with each other (n % 4 != 0: modes 1 and 2 fall back to the
    generic path, mode 3 takes its scalar loop; 4: the warp-specialised persistent kernel — TMA producers, pass-1 and
    pass-2 warp groups, solver warp)."""
    m = orc.synth_masses(SEED, n)
    ref = orc.synth_frame(SEED, 0, n, TRIC)
    out = {}
    for mode in (0, 1, 2, 3, 4):
        t = mb.Trajectory()
        t.synth(SEED, 0, nf, n, TRIC, mass_seed=SEED)
        t.set_option("fused_fit", mode)
        before = t.frame(nf - 1)
        r = t.fit(ref_frame=0, superpose=True)
        rc, R, tt = orc.fit_transform(before, m, None, ref, m, None)
        exp = orc.apply_transform_f64(before, None, R, tt)
        exp_r = np.sqrt(((exp - ref) ** 2).sum(1).mean())
        assert abs(r[nf - 1] - exp_r) <= RTOL * exp_r + 1e-9
        assert np.allclose(t.frame(nf - 1), exp, rtol=RTOL, atol=2e-6)
        out[mode] = (r, t.frame(1))
        # fit only (no superposition) leaves the frames untouched
        t2 = mb.Trajectory()
        t2.synth(SEED, 0, 2, n, TRIC, mass_seed=SEED)
        t2.set_option("fused_fit", mode)
        r2 = t2.fit(ref_frame=0, superpose=False)
        assert np.array_equal(t2.frame(1), orc.synth_frame(SEED, 1, n, TRIC)) and abs(r2[1] - r[1]) <= 1e-9 * r[1]
        # a second pass over the (now superposed) batch: RMSD unchanged up to f32 rounding of the stored frames
        r3 = t.fit(ref_frame=0, superpose=True)
        assert np.allclose(r3[1:], r[1:], rtol=1e-4)
        t.close()
        t2.close()
    for mode in (1, 2, 3, 4):
        assert np.allclose(out[0][0], out[mode][0], rtol=1e-10) and np.allclose(out[0][1], out[mode][1], atol=1e-6)


def test_batch_pipeline_config5_shape(mb):
    n, nf = 200_000, 3
    box = (TRIC * np.float32(0.6)).astype(np.float32)
    t = mb.Trajectory()
    t.synth(SEED, 5, nf, n, box, mass_seed=SEED)
    rows = t.pipeline(1.2)
    m = orc.synth_masses(SEED, n)
    b = orc.Box(matrix=box)
    for f in range(nf):
        xyz = orc.synth_frame(SEED, 5 + f, n, box)
        rc, com = orc.center_of_mass(xyz, m)
        rc, rg = orc.gyration(xyz, m)
        assert np.allclose(rows[f, :3], com, rtol=RTOL) and abs(rows[f, 3] - rg) / rg < RTOL
        if f == 0:
            ij, d, dims = orc.search_single(1.2, xyz, None, b, 7, 8)
            assert int(rows[f, 4]) == len(orc.canonical_pairs(ij))
    t.close()


def test_stream_fit_and_pipeline_host_frames(mb):
    """mb_stream_fit / mb_stream_pipeline: host frames uploaded chunk by chunk; results equal the resident-batch path."""
    n, nf = 40_000, 21
    frames = np.stack([orc.synth_frame(SEED + 9, f, n, M) for f in range(nf)])
    m = orc.synth_masses(SEED + 9, n)
    a = mb.Trajectory()
    a.upload(frames, box=M, masses=m)
    want_rmsd = a.fit(ref_frame=0, superpose=True)
    b = mb.Trajectory()
    got_rmsd = b.stream_fit(frames, m)
    assert np.allclose(got_rmsd, want_rmsd, rtol=1e-12, atol=1e-12)
    rc, R, t = orc.fit_transform(frames[5], m, None, frames[0], m, None)
    rc, r5 = orc.rmsd(orc.apply_transform_f64(frames[5], None, R, t).astype(np.float32), None, frames[0], None)
    assert abs(got_rmsd[5] - r5) / r5 < 1e-5
    a.upload(frames, box=M, masses=m)
    want = a.pipeline(1.2)
    got = b.stream_pipeline(frames, 1.2, M, masses=m)
    assert np.allclose(got, want, rtol=1e-12, atol=0)
    a.close()
    b.close()


def test_stream_entry_points_multi_chunk_500k(mb):
    """500k-atom frames: the streaming calls work in several chunks (ring of two chunks, upload of chunk k+1
    overlapped with the work on chunk k); every result must equal the resident-batch path."""
    n, nf = 500_000, 10
    box = np.diag([17.0, 17.0, 17.3]).astype(np.float32)
    frames = np.stack([orc.synth_frame(SEED + 21, f, n, box) for f in range(nf)])
    m = orc.synth_masses(SEED + 21, n)
    a = mb.Trajectory()
    a.upload(frames, box=box, masses=m)
    want_counts = a.search(1.2, count_only=True)
    want_rows = a.pipeline(1.2)
    want_rmsd = a.fit(ref_frame=0, superpose=True)
    a.close()
    b = mb.Trajectory()
    assert np.array_equal(b.stream_search(frames, 1.2, box), want_counts)
    assert np.array_equal(b.stream_search(frames, 1.2, box, count_only=True), want_counts)
    assert np.allclose(b.stream_pipeline(frames, 1.2, box, masses=m), want_rows, rtol=1e-12, atol=0)
    assert np.allclose(b.stream_fit(frames, m), want_rmsd, rtol=1e-12, atol=1e-12)
    b.close()


def test_reduce_many_small_selections_vs_oracle(mb):
    """Thousands of per-residue selections in ONE launch (mb_reduce_many): every row equals the oracle's f64 result for
    that selection; zero-mass and empty selections are reported per selection (MeasureError::ZeroMass)."""
    rng = np.random.default_rng(11)
    n = 60_000
    xyz = orc.synth_frame(SEED + 3, 0, n, M)
    m = orc.synth_masses(SEED + 3, n)
    m[100:110] = 0.0  # a massless residue
    s = mb.System(xyz, masses=m, box=M)
    # residues of 3..40 consecutive atoms, plus a few scattered and a few large selections
    bounds = [0]
    while bounds[-1] < n - 50:
        bounds.append(bounds[-1] + int(rng.integers(3, 41)))
    sels = [np.arange(a, b, dtype=np.uint64) for a, b in zip(bounds[:-1], bounds[1:])]
    sels += [np.sort(rng.choice(n, size=k, replace=False)).astype(np.uint64) for k in (1, 2, 33, 500, 5000)]
    sels = [a for a in sels if not (a.min() >= 100 and a.max() < 110)]
    assert len(sels) > 2000
    com = s.reduce_many(sels, "com")
    cog = s.reduce_many(sels, "cog")
    rg = s.reduce_many(sels, "gyration")
    both = s.reduce_many(sels, "com_gyration")
    for k in list(range(0, len(sels), 97)) + list(range(len(sels) - 5, len(sels))):
        ids = sels[k]
        rc, c = orc.center_of_mass(xyz, m, ids)
        rc2, g = orc.gyration(xyz, m, ids)
        assert rc == 0 and np.allclose(com[k], c, rtol=RTOL, atol=1e-12)
        assert np.allclose(cog[k], xyz[ids.astype(np.int64)].astype(np.float64).mean(0), rtol=RTOL, atol=1e-9)
        assert abs(rg[k, 0] - g) <= RTOL * max(g, 1e-9) + 1e-7
        assert np.array_equal(both[k, :3], com[k]) and both[k, 3] == rg[k, 0]
    # the same through the CTA-per-selection kernel (large mean size)
    big = [np.sort(rng.choice(n, size=3000, replace=False)).astype(np.uint64) for _ in range(6)]
    cb = s.reduce_many(big, "com_gyration")
    for k, ids in enumerate(big):
        rc, c = orc.center_of_mass(xyz, m, ids)
        rc2, g = orc.gyration(xyz, m, ids)
        assert np.allclose(cb[k, :3], c, rtol=RTOL) and abs(cb[k, 3] - g) <= RTOL * g
    # per-selection errors: zero mass -> MB_ERR_ZERO_MASS with that row NaN and the others intact
    bad = [np.arange(0, 10, dtype=np.uint64), np.arange(100, 110, dtype=np.uint64), np.arange(20, 25, dtype=np.uint64)]
    out, status = s.reduce_many(bad, "com", return_status=True)
    assert list(status) == [0, 1, 0] and np.isnan(out[1]).all() and np.array_equal(out[0], s(bad[0]).com())
    with pytest.raises(mb.MolarB200Error) as e:
        s.reduce_many(bad, "com")
    assert e.value.code == -1
    # cog needs no masses
    assert np.isfinite(s.reduce_many(bad, "cog")).all()
    s.close()
