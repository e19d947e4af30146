"""Shared helpers for the GPU parity tests: CUDA path (through the C ABI) vs the CPU oracle."""
import numpy as np

from oracle import oracle_py as orc

SEED = 20260

# config-3 triclinic shape: non-positive shear => the reference grid is also exact
TRIC = np.array([[21.5, -2.7, -2.7], [0.0, 21.5, -2.7], [0.0, 0.0, 21.5]], np.float32)


def oracle_single(cutoff, xyz, ids=None, box=None, pbc=0, nthreads=4):
    b = orc.Box(matrix=box) if box is not None else None
    ij, d, dims = orc.search_single(cutoff, xyz, ids, b, pbc, nthreads)
    pairs, dist = orc.canonical_pairs(ij, d)
    return pairs, dist, dims


def gpu_canonical(pairs, dist):
    """GPU output is already a set; sort it the same way for comparison and assert it IS a set."""
    p, d = orc.canonical_pairs(pairs, dist)
    assert len(p) == len(pairs), "GPU pair list contains duplicates"
    return p, d


def assert_same_pairs(gp, gd, op, od):
    assert gp.shape == op.shape, f"pair count differs: gpu {len(gp)} vs oracle {len(op)}"
    assert np.array_equal(gp, op), "pair sets differ"
    # distances: same f32 expression on both sides -> bit-exact
    assert np.array_equal(gd, od), f"distances differ (max abs {np.abs(gd - od).max()})"
