"""Host-side periodic-box tables, checked on the CPU through mb_box_describe (no GPU).

The minimum-image code of the reductions (shortest_vector_dev) no longer loops over the reference's up to 26 triclinic
corrections: it forms start.(ia + jb + kc) from three basis dot products for 13 (+v, -v) pairs, compares both signs
with one threshold per pair, and evaluates only the surviving corrections — by their index in the REFERENCE's list, so
that the reference's evaluation order is kept.  The table that maps pairs onto list indices is built on the host
(to_dev_box); these tests pin it against the oracle's restatement of build_tric_corrections (periodic_box.rs:25-66) and
check, by brute force over the reduced cell, that the pruning never drops a correction that could win."""
import ctypes as C

import numpy as np
import pytest

from molar_b200 import _capi
from oracle import oracle_py as orc

PAIR_IJK = [(0, 0, 1), (0, 1, -1), (0, 1, 0), (0, 1, 1), (1, -1, -1), (1, -1, 0), (1, -1, 1), (1, 0, -1), (1, 0, 0),
            (1, 0, 1), (1, 1, -1), (1, 1, 0), (1, 1, 1)]
BOXES = {
    "config3": np.array([[21.5, -2.7, -2.7], [0.0, 21.5, -2.7], [0.0, 0.0, 21.5]], np.float32),
    "mdtraj": np.array([[10.0, 4.0, -4.0], [0.0, 10.0, 0.0], [0.0, 0.0, 10.0]], np.float32),
    "dodecahedron": np.array([[7.0, 0.0, 3.5], [0.0, 7.0, 3.5], [0.0, 0.0, 4.9497]], np.float32),
    "ortho": np.diag([3.0, 4.0, 5.0]).astype(np.float32),
}


def describe(M):
    L = _capi.load()
    b9 = np.ascontiguousarray(np.asarray(M, np.float32).T.reshape(9))
    n = C.c_int(0)
    corr = np.zeros(78, np.float32)
    thr = np.zeros(13, np.float32)
    bits = np.zeros(26, np.uint32)
    _capi.check(L.mb_box_describe(b9.ctypes.data_as(_capi.f32p), C.byref(n), corr.ctypes.data_as(_capi.f32p),
                                  thr.ctypes.data_as(_capi.f32p), bits.ctypes.data_as(C.POINTER(C.c_uint32))))
    return n.value, corr.reshape(26, 3)[: n.value], thr, bits


@pytest.mark.parametrize("name", sorted(BOXES))
def test_pair_table_matches_the_reference_correction_list(name):
    M = BOXES[name]
    n, corr, thr, bits = describe(M)
    want = np.asarray(orc.Box(matrix=M).corrections, np.float32).reshape(-1, 3)
    assert n == len(want) and np.array_equal(corr, want)  # same vectors, same order as the oracle's restatement
    if n == 0:
        assert not bits.any()
        return
    a, b, c = M[:, 0], M[:, 1], M[:, 2]
    seen = 0
    for p, (i, j, k) in enumerate(PAIR_IJK):
        v = (np.float32(i) * a + np.float32(j) * b) + np.float32(k) * c
        for sg, bit in ((1.0, bits[2 * p]), (-1.0, bits[2 * p + 1])):
            if bit == 0:  # this combination is not in the reference's list (beyond its diagonal bound)
                assert not any(np.array_equal(np.float32(sg) * v, w) for w in corr)
                continue
            idx = int(bit).bit_length() - 1
            assert bit == 1 << idx and not (seen >> idx) & 1  # one bit, used once
            seen |= 1 << idx
            assert np.array_equal(corr[idx], np.float32(sg) * v)
            assert thr[p] == np.float32(-0.4995 * float((corr[idx].astype(np.float64) ** 2).sum()))
    assert seen == (1 << n) - 1  # every correction of the list has a pair slot


@pytest.mark.parametrize("name", ["config3", "mdtraj", "dodecahedron"])
def test_pruning_never_drops_a_winning_correction(name):
    """For points of the reduced cell, the corrections that actually shorten the vector (evaluated as the reference
    does, f32) must all be among the candidates the three-dot-product test lets through."""
    M = BOXES[name]
    n, corr, thr, bits = describe(M)
    rng = np.random.default_rng(7)
    f = (rng.random((20000, 3)) - 0.5).astype(np.float32)
    s = (f @ M.T).astype(np.float32)
    da, db, dc = s @ M[:, 0], s @ M[:, 1], s @ M[:, 2]
    cand = np.zeros(len(s), np.uint32)
    for p, (i, j, k) in enumerate(PAIR_IJK):
        d = (i * da + j * db + k * dc).astype(np.float32)
        cand |= np.where(d < thr[p], bits[2 * p], 0).astype(np.uint32)
        cand |= np.where(-d < thr[p], bits[2 * p + 1], 0).astype(np.uint32)
    base = (s.astype(np.float32) ** 2).sum(1)
    wins_any = 0
    for idx in range(n):
        moved = ((s + corr[idx]) ** 2).sum(1)
        wins = moved < base
        wins_any += int(wins.sum())
        assert not (wins & (((cand >> idx) & 1) == 0)).any(), f"correction {idx} wins but was pruned"
    assert wins_any > 0  # the test exercises real corrections on these boxes
