"""GPU parity: pair-list consumers that stay on the device — SearchConnectivity (connectivity.rs:8-38) and
unwrap_connectivity_dim (modify.rs:72-131) — vs the oracle.

What the reference defines (and what is compared exactly): the adjacency as a set per atom, the components, their
start atoms and the returned selections.  The unwrapped coordinates depend, in their last bits, on the spanning tree
the walk happens to follow (pair order = rayon's scheduling in the reference), so they are compared to a few f32 ulps
and through the property that matters: every contact is a direct (non-periodic) contact afterwards."""
import numpy as np
import pytest

from oracle import oracle_py as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mb():
    import molar_b200
    return molar_b200


def _chains(n_mol, n_per, box_len, seed, bond=0.15):
    rng = np.random.default_rng(seed)
    mols = []
    for _ in range(n_mol):
        start = rng.random(3) * box_len
        steps = rng.normal(size=(n_per, 3))
        steps *= bond / np.linalg.norm(steps, axis=1)[:, None]
        mols.append(start + np.cumsum(steps, axis=0))
    whole = np.concatenate(mols)
    return whole, np.mod(whole, box_len).astype(np.float32)


def test_connectivity_csr_vs_pair_list(mb):
    box = np.diag([6.0, 7.0, 8.0]).astype(np.float32)
    xyz = orc.synth_frame(20260, 0, 30_000, box, stray_permille=10)
    s = mb.System(xyz, box=box)
    pairs, dist = mb.distance_search(0.5, s(), dims=[True, True, True])
    row_ptr, cols = s.connectivity()
    assert row_ptr[-1] == 2 * len(pairs) == len(cols)
    p = np.asarray(pairs, dtype=np.int64)
    deg = np.bincount(np.concatenate([p[:, 0], p[:, 1]]), minlength=len(xyz))
    assert np.array_equal(np.diff(row_ptr.astype(np.int64)), deg)
    # same (atom, neighbour) multiset as the pair list taken in both directions
    rows = np.repeat(np.arange(len(xyz)), deg)
    got = np.sort(rows * len(xyz) + cols.astype(np.int64))
    want = np.sort(np.concatenate([p[:, 0] * len(xyz) + p[:, 1], p[:, 1] * len(xyz) + p[:, 0]]))
    assert np.array_equal(got, want)
    s.close()


def _rows_from_pairs(pairs, n):
    """SearchConnectivity::from_iter (connectivity.rs:19-35) of an (i, j) list: sorted (atom * n + neighbour) keys."""
    p = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    return np.sort(np.concatenate([p[:, 0] * n + p[:, 1], p[:, 1] * n + p[:, 0]]))


def _rows_from_csr(row_ptr, cols, n):
    deg = np.diff(row_ptr.astype(np.int64))
    assert deg.min() >= 0 and row_ptr[0] == 0 and row_ptr[-1] == len(cols)
    return np.sort(np.repeat(np.arange(n), deg) * n + cols.astype(np.int64))


TRIC_SMALL = np.array([[6.5, -0.8, -0.8], [0.0, 6.5, -0.8], [0.0, 0.0, 6.5]], np.float32)


@pytest.mark.parametrize("case", ["ortho_pbc", "tric_pbc_strays", "partial_pbc", "nonperiodic", "selection", "small_fallback",
                                  "dense_tiles"])
def test_search_connectivity_vs_oracle_pairs(mb, case):
    """Neighbour rows written by the search kernel (full shell, count pass + fill pass) == the adjacency the reference
    builds from ITS pair list (connectivity.rs:8-38 over distance_search.rs:892-954), compared with the oracle's pairs."""
    box, pbc, dims, n, cutoff, ids, stray = np.diag([6.0, 7.0, 8.0]).astype(np.float32), 7, [True] * 3, 30_000, 0.5, None, 0
    if case == "tric_pbc_strays":
        box, stray, cutoff = TRIC_SMALL, 10, 0.6
    elif case == "partial_pbc":
        pbc, dims, stray = 5, [True, False, True], 10
    elif case == "nonperiodic":
        pbc, dims = 0, None
    elif case == "selection":
        ids = np.sort(np.random.default_rng(5).choice(n, 20_000, replace=False)).astype(np.uint64)
    elif case == "small_fallback":
        n = 2_000
    elif case == "dense_tiles":
        # 60 atoms per home tile and long candidate streams: several home batches per tile, counter folds, batches
        # visited again after a full run table
        box, n, cutoff = np.diag([3.0, 3.0, 3.0]).astype(np.float32), 60_000, 0.45
    xyz = orc.synth_frame(20260, 3, n, box, stray_permille=stray)
    want, _, _ = orc.search_single(cutoff, xyz, ids, orc.Box(matrix=box) if pbc else None, pbc, 4)
    s = mb.System(xyz, box=box)
    sel = s() if ids is None else s(ids)
    row_ptr, cols = sel.search_connectivity(cutoff, dims=dims)
    assert len(cols) == 2 * len(want)
    assert np.array_equal(_rows_from_csr(row_ptr, cols, n), _rows_from_pairs(want, n))
    if case != "small_fallback":  # (small grids go through pair list -> CSR)
        with pytest.raises(mb._capi.MolarB200Error):
            mb._capi.check(s._lib.mb_fill_pairs(s._h, None, None))  # no pair list on the context
    s.close()


@pytest.mark.parametrize("n_mol,n_per,L", [(40, 60, 5.0), (3, 2000, 24.0), (500, 3, 5.0)])
def test_unwrap_connectivity_vs_oracle(mb, n_mol, n_per, L):
    # (3, 2000): long chains, deep walk (hundreds of levels); the box is large enough that a chain never touches
    # its own periodic image — otherwise the contact graph has loops around the box and "the" unwrapped position
    # is ambiguous by a lattice vector in the reference as well
    box = np.diag([L, L, L]).astype(np.float32)
    whole, wrapped = _chains(n_mol, n_per, L, seed=n_mol)
    want_xyz, want_roots, want_n = orc.unwrap_connectivity(0.2, wrapped, orc.Box(matrix=box))
    s = mb.System(wrapped, box=box)
    sel_all = s()
    sels = sel_all.unwrap_connectivity(0.2)
    got = s.coords()
    assert np.array_equal(sel_all.roots, want_roots)
    assert len(np.unique(sel_all.roots)) == want_n
    # returned selections: members of every component without its start atom, none for isolated atoms
    want_sets = [set(np.flatnonzero(want_roots == r)) - {r} for r in np.unique(want_roots)]
    want_sets = [w for w in want_sets if w]
    assert [set(int(i) for i in x.get_index()) for x in sels] == want_sets
    # start atoms do not move
    starts = np.unique(want_roots)
    assert np.array_equal(got[starts], wrapped[starts])
    # positions: equal to the oracle's walk up to f32 rounding along different spanning trees
    assert np.abs(got - want_xyz).max() < 2e-5
    # every bond of every chain is whole again
    d = np.linalg.norm(np.diff(got.reshape(n_mol, n_per, 3).astype(np.float64), axis=1), axis=2)
    assert d.max() < 0.15 + 1e-4
    s.close()


def test_unwrap_connectivity_selection_partial_dims_triclinic(mb):
    box = np.array([[5.0, -1.0, -0.8], [0.0, 5.2, -1.1], [0.0, 0.0, 4.9]], np.float32)
    whole, _ = _chains(30, 40, 4.0, seed=3)
    M = box.astype(np.float64)
    frac = np.linalg.solve(M, whole.T).T
    wrapped = (M @ (frac - np.floor(frac)).T).T.astype(np.float32)
    ids = np.arange(0, len(wrapped), 1, dtype=np.uint64)[200:1000]
    for dims, bits in (([True, True, True], 7), ([True, False, True], 5)):
        want_xyz, want_roots, want_n = orc.unwrap_connectivity(0.2, wrapped, orc.Box(matrix=box), ids, bits)
        s = mb.System(wrapped, box=box)
        sel = s(ids)
        sel.unwrap_connectivity(0.2, dims=dims)
        got = s.coords()
        assert np.array_equal(sel.roots, want_roots)
        assert np.array_equal(got[:200], wrapped[:200]) and np.array_equal(got[1000:], wrapped[1000:])
        assert np.abs(got - want_xyz).max() < 2e-5
        s.close()


def test_unwrap_needs_a_box(mb):
    s = mb.System(np.zeros((10, 3), np.float32))
    with pytest.raises(mb.MolarB200Error) as e:
        s().unwrap_connectivity(0.2)
    assert e.value.code == -4
    s.close()
