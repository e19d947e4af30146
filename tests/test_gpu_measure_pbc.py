"""GPU parity: periodic reductions, inertia tensor and principal axes through the C ABI vs the oracle.

The CUDA path chooses every atom's periodic image with the reference's f32 arithmetic and sums in f64, so
it is compared (tightly) with the oracle's "mixed" statement of the same thing, and — at the north-star
tolerance of 1e-6 relative — with the oracle's f64 restatement of the reference.
"""
import numpy as np
import pytest

from oracle import oracle_py as orc
from tests.helpers import SEED, TRIC

pytestmark = pytest.mark.gpu
RTOL = 1e-6

ORTHO = np.diag([6.0, 7.0, 8.0]).astype(np.float32)
TRIC_SMALL = np.array([[6.0, -1.1, -0.9], [0.0, 6.5, -1.3], [0.0, 0.0, 7.0]], np.float32)


@pytest.fixture(scope="module")
def mb():
    import molar_b200
    return molar_b200


def _cluster(box, n, seed=3, spread=0.9):
    rng = np.random.default_rng(seed)
    M = box.astype(np.float64)
    centre = M @ np.array([0.97, 0.03, 0.51])
    whole = centre + rng.normal(0.0, spread / 3.0, size=(n, 3))
    frac = np.linalg.solve(M, whole.T).T
    wrapped = (M @ (frac - np.floor(frac)).T).T
    return wrapped.astype(np.float32), (1.0 + 15.0 * rng.random(n)).astype(np.float32)


def _close(a, b, rtol, scale=None):
    a, b = np.asarray(a, float), np.asarray(b, float)
    s = np.abs(b).max() if scale is None else scale
    return np.abs(a - b).max() <= rtol * s


@pytest.mark.parametrize("box", [ORTHO, TRIC_SMALL], ids=["ortho", "tric"])
@pytest.mark.parametrize("n,stride", [(200003, 1), (50000, 3), (9, 1), (1, 1)])
def test_center_pbc(mb, box, n, stride):
    xyz, m = _cluster(box, n)
    ids = np.arange(0, n, stride, dtype=np.uint64) if stride > 1 else None
    b = orc.Box(matrix=box)
    s = mb.System(xyz, masses=m, box=box)
    sel = s(ids) if ids is not None else s()
    L = float(np.abs(box).max())
    for dims, bits in (([True, True, True], 7), ([True, False, True], 5)):
        rc, want = orc.center_pbc(xyz, m, b, bits, ids, prec="mixed")
        assert _close(sel.com(dims=dims), want, 1e-11, L)
        rc, want64 = orc.center_pbc(xyz, m, b, bits, ids, prec="f64")
        assert _close(sel.com(dims=dims), want64, RTOL, L)
        rc, wantg = orc.center_pbc(xyz, None, b, bits, ids, prec="mixed")
        assert _close(sel.cog(dims=dims), wantg, 1e-11, L)
    assert _close(sel.cog(), orc.center_of_geometry(xyz, ids), 1e-12, L)
    s.close()


@pytest.mark.parametrize("box", [ORTHO, TRIC_SMALL], ids=["ortho", "tric"])
def test_gyration_and_inertia_pbc(mb, box):
    n = 150000
    xyz, m = _cluster(box, n, seed=5)
    ids = np.arange(1, n, 2, dtype=np.uint64)
    b = orc.Box(matrix=box)
    s = mb.System(xyz, masses=m, box=box)
    for sel, sid in ((s(), None), (s(ids), ids)):
        rc, rg = orc.gyration_pbc(xyz, m, b, sid, prec="mixed")
        assert abs(sel.gyration(pbc=True) - rg) <= 1e-10 * rg
        rc, rg64 = orc.gyration_pbc(xyz, m, b, sid, prec="f64")
        assert abs(sel.gyration_pbc() - rg64) <= RTOL * rg64
        rc, tens, mom, axes, centre = orc.inertia(xyz, m, b, sid, prec="mixed")
        gmom, gaxes = sel.inertia(pbc=True)
        assert _close(gmom, mom, 1e-9)
        rc, tens64, mom64, axes64, _ = orc.inertia(xyz, m, b, sid, prec="f64")
        assert _close(gmom, mom64, RTOL)
        # axes: orthonormal, right-handed, diagonalise the oracle's tensor, equal to the oracle's up to sign
        assert np.allclose(gaxes.T @ gaxes, np.eye(3), atol=1e-12) and np.linalg.det(gaxes) > 0
        assert np.allclose(gaxes.T @ tens @ gaxes, np.diag(mom), atol=1e-8 * np.abs(tens).max())
        for k in range(3):
            assert min(np.abs(gaxes[:, k] - axes[:, k]).max(), np.abs(gaxes[:, k] + axes[:, k]).max()) < 1e-6
    s.close()


def test_inertia_and_principal_transform_nonperiodic(mb):
    n = 120000
    rng = np.random.default_rng(11)
    # an elongated, tilted cloud away from the origin
    xyz = (rng.normal(size=(n, 3)) * np.array([3.0, 1.0, 0.4])) @ np.array(
        [[0.8, -0.6, 0.0], [0.6, 0.8, 0.0], [0.0, 0.0, 1.0]]) + np.array([12.0, -7.0, 30.0])
    xyz = xyz.astype(np.float32)
    m = (1.0 + 15.0 * rng.random(n)).astype(np.float32)
    s = mb.System(xyz, masses=m)
    rc, tens, mom, axes, centre = orc.inertia(xyz, m, None, prec="f64")
    gmom, gaxes = s().inertia()
    assert _close(gmom, mom, RTOL)
    for k in range(3):
        assert min(np.abs(gaxes[:, k] - axes[:, k]).max(), np.abs(gaxes[:, k] + axes[:, k]).max()) < 1e-6
    tr = s().principal_transform()
    assert np.allclose(tr.R @ tr.R.T, np.eye(3), atol=1e-12) and np.linalg.det(tr.R) > 0
    # the transform keeps the centre of mass and makes the inertia tensor diagonal with the same moments
    assert np.allclose(tr.R @ centre + tr.t, centre, atol=1e-9)
    s().apply_transform(tr)
    moved = s.coords()
    rc, tens2, mom2, axes2, centre2 = orc.inertia(moved, m, None, prec="f64")
    assert np.allclose(centre2, centre, atol=1e-4)
    off = tens2 - np.diag(np.diag(tens2))
    assert np.abs(off).max() < 1e-5 * np.abs(tens2).max()
    assert np.allclose(np.diag(tens2), mom, rtol=1e-5)
    s.close()


def test_config5_scale_pbc_reductions_1m(mb):
    """1M atoms in the config-3 triclinic box with 1 % stray atoms (outside the box): every quantity
    against the oracle's mixed statement."""
    n = 1_000_000
    xyz = orc.synth_frame(SEED, 0, n, TRIC, stray_permille=10)
    m = orc.synth_masses(SEED, n)
    b = orc.Box(matrix=TRIC)
    s = mb.System(xyz, masses=m, box=TRIC)
    rc, want = orc.center_pbc(xyz, m, b, 7, prec="mixed")
    assert _close(s().com(dims=[True, True, True]), want, 1e-10, 21.5)
    rc, rg = orc.gyration_pbc(xyz, m, b, prec="mixed")
    assert abs(s().gyration(pbc=True) - rg) <= 1e-9 * rg
    s.close()


def test_inertia_axes_reference_known_answer(mb, golden_dir):
    """molar/src/selection.rs:198-213: the reference's `test_inertia` (all atoms of tests/protein.pdb) keeps the three
    inertia axes it expects as a comment; the CUDA path (moments1_kernel -> tensor_kernel -> host eigenproblem)
    reproduces them to the precision those numbers carry, and its moments agree with the oracle's f64 build to 1e-9."""
    import os
    g = np.load(os.path.join(golden_dir, "protein_inertia.npz"))
    s = mb.System(g["xyz"], masses=g["masses"], box=g["box"])
    mom, axes = s().inertia()
    ref = g["ref_axes"][::-1]  # the comment lists the axes by descending moment
    for k in range(3):
        assert min(np.abs(axes[:, k] - ref[k]).max(), np.abs(axes[:, k] + ref[k]).max()) < 1e-5
    rc, tensor, omom, oaxes, centre = orc.inertia(g["xyz"], g["masses"], box=None, prec="f64")
    assert np.allclose(mom, omom, rtol=1e-9)
    s.close()


def test_center_pbc_through_the_pymolar_shortest_vector_known_answer(mb):
    """molar_python/tests/test_2.py:233-245: in the 1 x 2 x 3 box the shortest image of (0.9, 0.5, 0.6) is
    (-0.1, 0.5, 0.6).  center_of_geometry_pbc of the two atoms {origin, (0.9, 0.5, 0.6)} is the first atom plus half of
    that vector (measure.rs:142-170), so the reference's own known answer goes through center_pbc_kernel."""
    xyz = np.array([[0.0, 0.0, 0.0], [0.9, 0.5, 0.6]], np.float32)
    s = mb.System(xyz, box=np.diag([1.0, 2.0, 3.0]).astype(np.float32))
    assert np.allclose(s().cog(dims=[True] * 3), [-0.05, 0.25, 0.3], rtol=0, atol=1e-6)
    s.close()


def test_errors(mb):
    xyz, m = _cluster(ORTHO, 100)
    s = mb.System(xyz, masses=m)  # no box
    with pytest.raises(mb.MolarB200Error) as e:
        s().com(dims=[True, True, True])
    assert e.value.code == -4  # MeasureError::Pbc(NoPbc) (require_box, measure.rs:176)
    with pytest.raises(mb.MolarB200Error) as e:
        s().gyration(pbc=True)
    assert e.value.code == -4
    s.close()
    s = mb.System(xyz, masses=np.zeros(100, np.float32), box=ORTHO)
    with pytest.raises(mb.MolarB200Error) as e:
        s().com(dims=[True, True, True])
    assert e.value.code == -1  # ZeroMass (measure.rs:191-193)
    s.close()
