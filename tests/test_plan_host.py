"""Host-side planning logic of the cell kernels, checked on the CPU through mb_plan_describe (no GPU):
the neighbour-row table must reach EVERY fine cell that can hold an atom within the cutoff of an atom of the home
tile — from exactly one side for a single-set search (half shell), from the home side for a two-set search (full
shell) — on orthorhombic and triclinic boxes, for both tile-size policies."""
import ctypes as C
import itertools

import numpy as np
import pytest

from molar_b200 import _capi

TRIC = np.array([[21.5, -2.7, -2.7], [0.0, 21.5, -2.7], [0.0, 0.0, 21.5]], np.float32)


def describe(M, cutoff, n, full_shell=0):
    L = _capi.load()
    b9 = np.ascontiguousarray(np.asarray(M, np.float32).T.reshape(9))
    oi = (C.c_int * 16)()
    rows = np.zeros(4 * 1024, np.int8)
    band = np.zeros(2, np.float32)
    _capi.check(L.mb_plan_describe(b9.ctypes.data_as(_capi.f32p), cutoff, 7, n, full_shell, oi, rows.ctypes.data,
                                   band.ctypes.data_as(_capi.f32p)))
    o = list(oi)
    return dict(use_cells=o[0], dims=o[1:4], k=o[4:7], hx=o[7], fd=np.array(o[8:11]), nrows=o[11], fast_pbc=o[12],
                rows=rows[: 4 * o[11]].reshape(-1, 4).astype(int), band=band)


def close_pairs(M, pts_frac, cutoff):
    """all (i, j), i != j, whose minimum-image distance is <= cutoff (brute force over 27 images)"""
    x = pts_frac @ M.T
    out = []
    for s in itertools.product((-1, 0, 1), repeat=3):
        sh = M @ np.array(s, float)
        d2 = ((x[:, None, :] - x[None, :, :] - sh) ** 2).sum(-1)
        i, j = np.nonzero(d2 <= cutoff * cutoff)
        keep = i != j
        out.append(np.stack([i[keep], j[keep]], 1))
    return np.unique(np.concatenate(out), axis=0)


CASES = [
    ("ortho100k", np.diag([10.0, 10.0, 10.0]), 1.2, 100_000),
    ("tric1m", TRIC.astype(np.float64), 1.2, 1_000_000),
    ("tric_small", TRIC.astype(np.float64) * 0.3, 1.2, 27_000),
    ("ortho_flat", np.diag([9.0, 14.0, 5.5]), 1.0, 60_000),
    ("short_cutoff", np.diag([8.0, 8.0, 8.0]), 0.35, 50_000),
]


@pytest.mark.parametrize("name,M,cutoff,n", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("full_shell", [0, 1])
def test_rows_cover_every_pair_within_cutoff(name, M, cutoff, n, full_shell):
    pl = describe(M, cutoff, n, full_shell)
    assert pl["use_cells"] == 1 and pl["nrows"] > 0
    fd, hx = pl["fd"], pl["hx"]
    assert fd[0] % hx == 0 and all(fd[d] == pl["dims"][d] * pl["k"][d] for d in range(3))
    rows = {(r[0], r[1]): (r[2], r[3]) for r in pl["rows"]}
    assert len(rows) == pl["nrows"]  # one row per (dy, dz)
    rng = np.random.default_rng(7)
    # points concentrated in a slab of a few cutoffs (plus its periodic neighbourhood) keep the brute force small
    pts = rng.random((1500, 3))
    width = min(1.0, 3.0 * cutoff / np.linalg.norm(M, axis=0).min())
    pts[:, 0] = (pts[:, 0] * width + 0.97) % 1.0
    pts[:, 1] = (pts[:, 1] * width + 0.95) % 1.0
    pts[:, 2] = (pts[:, 2] * width + 0.96) % 1.0
    pairs = close_pairs(np.asarray(M, float), pts, cutoff)
    assert len(pairs) > 2000
    cell = np.minimum(np.floor(pts * fd).astype(int), fd - 1)
    tile0 = cell.copy()
    tile0[:, 0] = (cell[:, 0] // hx) * hx

    def covered(i, j):
        d = cell[j] - tile0[i]
        d = (d + fd // 2) % fd - fd // 2  # nearest image in cell space (offsets are unique modulo the grid)
        if d[1] == 0 and d[2] == 0 and 0 <= d[0] < hx and not full_shell:
            return True  # home tile against itself
        r = rows.get((d[1], d[2]))
        return r is not None and r[0] <= d[0] <= r[1]

    for i, j in pairs:
        a, b = covered(i, j), covered(j, i)
        if full_shell:
            assert a, f"{name}: cell of {j} not reachable from the tile of {i}"
        else:
            same = np.array_equal(tile0[i], tile0[j])
            assert a or b or same, f"{name}: pair ({i},{j}) reachable from neither side"
            if not same:
                assert not (a and b), f"{name}: pair ({i},{j}) reachable from both sides (would be reported twice)"


def test_filter_band_brackets_the_cutoff_and_degenerate_grids_fall_back():
    pl = describe(TRIC, 1.2, 1_000_000)
    assert pl["fast_pbc"] == 1 and pl["band"][0] < np.float32(1.2) ** 2 < pl["band"][1]
    assert (pl["band"][1] - pl["band"][0]) / 1.44 < 0.01
    # 3 reference cells along z: wrapped pairs are always evaluated exactly
    L = 6.0
    M = np.array([[L, 0, L / 2], [0, L, L / 2], [0, 0, L / np.sqrt(2)]])
    pl = describe(M, 1.2, 22_000)
    assert pl["use_cells"] == 1 and pl["dims"][2] == 3 and pl["fast_pbc"] == 0
    # one or two reference cells in a periodic dimension: general all-pairs kernel
    assert describe(np.diag([2.0, 9.0, 9.0]), 1.2, 50_000)["use_cells"] == 0
    # tiny selections do not use the cell path
    assert describe(np.diag([9.0, 9.0, 9.0]), 1.2, 1000)["use_cells"] == 0


def test_reference_grid_dims_match_the_oracle_on_random_boxes():
    """Grid::dims = max(floor(lab_extent / cutoff), 1) with lab extents = row sums of the box matrix
    (distance_search.rs:103-110, periodic_box.rs:369-375): the planner and the oracle must agree, including where
    extent / cutoff lands on an integer in f32."""
    from oracle import oracle_py as orc
    rng = np.random.default_rng(11)
    pts = rng.random((64, 3)).astype(np.float32)
    boxes = [np.diag([21.6, 14.4, 9.6]), np.diag([12.0, 12.0, 12.0]), TRIC.astype(np.float64)]
    for _ in range(40):
        a, b, c = rng.uniform(4.0, 25.0, 3)
        sh = rng.uniform(-0.3, 0.3, 3)
        boxes.append(np.array([[a, sh[0] * b, sh[1] * c], [0.0, b, sh[2] * c], [0.0, 0.0, c]]))
    for M in boxes:
        for cutoff in (1.2, 0.8, 2.0):
            M32 = np.asarray(M, np.float32)
            ij, d, dims = orc.search_single(cutoff, (pts @ M32.T).astype(np.float32), None, orc.Box(matrix=M32), 7, 1)
            pl = describe(M32, cutoff, 100_000)
            assert list(dims) == pl["dims"], (M, cutoff, list(dims), pl["dims"])
