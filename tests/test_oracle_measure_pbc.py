"""Pin the oracle's restatement of the periodic reductions, inertia tensor and principal axes
(measure.rs:37-45,88-108,142-252,573-610; periodic_box.rs:286-330).

No test of the reference exercises these functions: PARITY UNPINNED by the reference.  They are pinned
here against an independent numpy f64 implementation (minimum image by explicit search over the 27
neighbouring lattice translations) and against invariants (a cluster cut by the box faces has the
centre, radius of gyration and inertia tensor of the whole cluster).
"""
import itertools

import numpy as np
import pytest

from oracle import oracle_py as orc

ORTHO = np.diag([6.0, 7.0, 8.0]).astype(np.float32)
TRIC = np.array([[6.0, -1.1, -0.9], [0.0, 6.5, -1.3], [0.0, 0.0, 7.0]], np.float32)


def _cluster(box, n=4000, seed=3, spread=0.9):
    """A compact cluster (radius << half box) centred near a box corner, wrapped into the box."""
    rng = np.random.default_rng(seed)
    M = box.astype(np.float64)
    centre = M @ np.array([0.97, 0.03, 0.51])
    whole = centre + rng.normal(0.0, spread / 3.0, size=(n, 3))
    frac = np.linalg.solve(M, whole.T).T
    wrapped = (M @ (frac - np.floor(frac)).T).T
    masses = (1.0 + 15.0 * rng.random(n)).astype(np.float32)
    return whole, wrapped.astype(np.float32), masses


def _min_image(M, d):
    """brute-force minimum image of the rows of d (search over 27 translations)"""
    best = d.copy()
    best2 = (d ** 2).sum(1)
    for s in itertools.product((-1, 0, 1), repeat=3):
        cand = d + M @ np.array(s, float)
        c2 = (cand ** 2).sum(1)
        m = c2 < best2 - 1e-12
        best[m] = cand[m]
        best2[m] = c2[m]
    return best


def _np_center_pbc(xyz, masses, M):
    x = xyz.astype(np.float64)
    d = _min_image(M, x - x[0])
    im = x[0] + d
    if masses is None:
        return im.sum(0) / len(x)
    w = masses.astype(np.float64)
    # reference quirk (measure.rs:180,200): the first atom enters the numerator with weight one
    return (x[0] + (im[1:] * w[1:, None]).sum(0)) / w.sum()


@pytest.mark.parametrize("box", [ORTHO, TRIC], ids=["ortho", "tric"])
def test_center_pbc_vs_numpy(box):
    whole, xyz, m = _cluster(box)
    b = orc.Box(matrix=box)
    M = box.astype(np.float64)
    for masses in (m, None):
        want = _np_center_pbc(xyz, masses, M)
        for prec, tol in (("f64", 1e-9), ("mixed", 1e-6), ("f32", 3e-4)):
            rc, c = orc.center_pbc(xyz, masses, b, 7, prec=prec)
            assert rc == 0
            assert np.allclose(c, want, rtol=0, atol=tol * 10.0), (prec, c, want)
    # the periodic centre of the cut cluster is a lattice image of the centre of the whole cluster
    rc, c = orc.center_pbc(xyz, None, b, 7, prec="f64")
    delta = _min_image(M, (c - whole.mean(0))[None, :])[0]
    assert np.abs(delta).max() < 1e-5
    # ... while the plain centre of geometry of the cut cluster is far away from it
    assert np.abs(_min_image(M, (orc.center_of_geometry(xyz) - whole.mean(0))[None, :])[0]).max() > 0.5


def test_center_of_geometry_and_first_atom_quirk():
    whole, xyz, m = _cluster(ORTHO, n=500)
    assert np.allclose(orc.center_of_geometry(xyz), xyz.astype(np.float64).mean(0), rtol=1e-12)
    # with every atom inside one image the periodic COM differs from the plain one only by the quirk
    x = (xyz * np.float32(0.05) + np.float32(3.0)).astype(np.float32)
    b = orc.Box(matrix=ORTHO)
    rc, c = orc.center_pbc(x, m, b, 7, prec="f64")
    w = m.astype(np.float64)
    plain = (x.astype(np.float64) * w[:, None]).sum(0) / w.sum()
    quirk = plain + x[0].astype(np.float64) * (1.0 - w[0]) / w.sum()
    assert np.allclose(c, quirk, rtol=1e-9) and not np.allclose(c, plain, rtol=1e-7)


def test_center_pbc_dims_partial():
    whole, xyz, m = _cluster(ORTHO)
    b = orc.Box(matrix=ORTHO)
    x = xyz.astype(np.float64)
    d = x - x[0]
    L = np.diag(ORTHO).astype(np.float64)
    for dims in (1, 2, 5):
        dd = d.copy()
        for k in range(3):
            if (dims >> k) & 1:
                dd[:, k] -= L[k] * np.round(dd[:, k] / L[k])
        want = (x[0] + dd).sum(0) / len(x)
        rc, c = orc.center_pbc(xyz, None, b, dims, prec="f64")
        assert rc == 0 and np.allclose(c, want, atol=1e-9)


def test_zero_mass():
    whole, xyz, m = _cluster(ORTHO, n=50)
    rc, _ = orc.center_pbc(xyz, np.zeros(50, np.float32), orc.Box(matrix=ORTHO), 7)
    assert rc == 1  # MeasureError::ZeroMass (measure.rs:191-193)


@pytest.mark.parametrize("box", [ORTHO, TRIC], ids=["ortho", "tric"])
def test_gyration_and_inertia_pbc_equal_those_of_the_whole_cluster(box):
    whole, xyz, m = _cluster(box)
    b = orc.Box(matrix=box)
    w = m.astype(np.float64)
    M = box.astype(np.float64)
    # numpy statement of gyration_pbc / inertia_pbc on the wrapped coordinates
    c = _np_center_pbc(xyz, m, M)
    d = _min_image(M, xyz.astype(np.float64) - c.astype(np.float32).astype(np.float64))
    rg = np.sqrt((w * (d ** 2).sum(1)).sum() / w.sum())
    T = np.zeros((3, 3))
    for a in range(3):
        for bb in range(3):
            T[a, bb] = ((d ** 2).sum(1) * w).sum() * (a == bb) - (w * d[:, a] * d[:, bb]).sum()
    for prec, tol in (("f64", 1e-7), ("mixed", 1e-6)):
        rc, g = orc.gyration_pbc(xyz, m, b, prec=prec)
        assert rc == 0 and abs(g - rg) / rg < tol
        rc, tens, mom, axes, centre = orc.inertia(xyz, m, b, prec=prec)
        assert rc == 0 and np.allclose(tens, T, rtol=0, atol=tol * np.abs(T).max())
        ev = np.linalg.eigvalsh(T)
        assert np.allclose(mom, ev, rtol=10 * tol)
        assert np.allclose(axes.T @ axes, np.eye(3), atol=1e-9) and np.linalg.det(axes) > 0.999999
        assert np.allclose(axes.T @ T @ axes, np.diag(ev), atol=1e-5 * np.abs(T).max())
    # the cut cluster has (nearly: the centre differs by the first-atom quirk) the Rg of the whole one
    cw = (whole * w[:, None]).sum(0) / w.sum()
    rg_whole = np.sqrt((w * ((whole - cw) ** 2).sum(1)).sum() / w.sum())
    assert abs(rg - rg_whole) / rg_whole < 1e-3


def test_inertia_nonperiodic_vs_numpy():
    whole, xyz, m = _cluster(ORTHO, n=3000, seed=9)
    x = xyz.astype(np.float64)
    w = m.astype(np.float64)
    c = (x * w[:, None]).sum(0) / w.sum()
    d = x - c
    T = np.eye(3) * (w * (d ** 2).sum(1)).sum() - np.einsum("k,ka,kb->ab", w, d, d)
    rc, tens, mom, axes, centre = orc.inertia(xyz, m, None, prec="f64")
    assert rc == 0 and np.allclose(centre, c, rtol=1e-12)
    # the reference rounds the centre to Float before subtracting: f64 build => exact
    assert np.allclose(tens, T, rtol=1e-10, atol=1e-9 * np.abs(T).max())
    assert np.allclose(mom, np.linalg.eigvalsh(T), rtol=1e-9)
    rc, tens32, mom32, axes32, _ = orc.inertia(xyz, m, None, prec="f32")
    assert np.allclose(mom32, mom, rtol=2e-3)
