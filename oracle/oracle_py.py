"""ctypes binding of the CPU oracle (oracle/libmolar_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (molar_b200) never imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmolar_oracle.so")

_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_u64p = C.POINTER(C.c_uint64)


def build(force=False):
    src = os.path.join(_HERE, "molar_oracle.cpp")
    if (force or not os.path.exists(_LIB_PATH)
            or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_box_from_matrix.restype = C.c_void_p
        L.orc_box_from_matrix.argtypes = [_f32p]
        L.orc_box_from_vectors_angles.restype = C.c_void_p
        L.orc_box_from_vectors_angles.argtypes = [C.c_float] * 6
        L.orc_box_free.argtypes = [C.c_void_p]
        L.orc_box_get.argtypes = [C.c_void_p, _f32p, _f32p]
        L.orc_box_corrections.restype = C.c_int
        L.orc_box_corrections.argtypes = [C.c_void_p, _f32p]
        L.orc_box_lab_extents.argtypes = [C.c_void_p, _f32p]
        L.orc_box_shortest_vector.argtypes = [C.c_void_p, _f32p, C.c_uint8, _f32p]
        L.orc_box_distance_squared.restype = C.c_float
        L.orc_box_distance_squared.argtypes = [C.c_void_p, _f32p, _f32p, C.c_uint8]
        L.orc_result_len.restype = C.c_size_t
        L.orc_result_len.argtypes = [C.c_void_p]
        L.orc_result_fill.argtypes = [C.c_void_p, _u64p, _f32p]
        L.orc_result_fill_ids.argtypes = [C.c_void_p, _u64p]
        L.orc_result_grid_dims.argtypes = [C.c_void_p, _u64p]
        L.orc_result_free.argtypes = [C.c_void_p]
        L.orc_search_single.restype = C.c_void_p
        L.orc_search_single.argtypes = [C.c_float, _f32p, _u64p, C.c_size_t, C.c_int]
        L.orc_search_single_pbc.restype = C.c_void_p
        L.orc_search_single_pbc.argtypes = [C.c_float, _f32p, _u64p, C.c_size_t, C.c_void_p, C.c_uint8, C.c_int]
        L.orc_search_single_pbc_checksum.restype = None
        L.orc_search_single_pbc_checksum.argtypes = [C.c_float, _f32p, _u64p, C.c_size_t, C.c_void_p, C.c_uint8, C.c_int,
                                                      _u64p, _u64p]
        L.orc_search_double.restype = C.c_void_p
        L.orc_search_double.argtypes = [C.c_float, _f32p, _u64p, C.c_size_t, _f32p, _u64p, C.c_size_t, C.c_int]
        L.orc_search_double_pbc.restype = C.c_void_p
        L.orc_search_double_pbc.argtypes = [C.c_float, _f32p, _u64p, C.c_size_t, _f32p, _u64p, C.c_size_t,
                                            C.c_void_p, C.c_uint8, C.c_int]
        L.orc_search_within.restype = C.c_void_p
        L.orc_search_within.argtypes = [C.c_float, _f32p, _u64p, C.c_size_t, _f32p, _u64p, C.c_size_t,
                                        _f32p, _f32p, C.c_int]
        L.orc_search_within_pbc.restype = C.c_void_p
        L.orc_search_within_pbc.argtypes = [C.c_float, _f32p, _u64p, C.c_size_t, _f32p, _u64p, C.c_size_t,
                                            C.c_void_p, C.c_uint8, C.c_int]
        L.orc_search_double_vdw.restype = C.c_void_p
        L.orc_search_double_vdw.argtypes = [_f32p, _u64p, C.c_size_t, _f32p, _f32p, _u64p, C.c_size_t, _f32p,
                                            C.c_void_p, C.c_uint8, C.c_int]
        L.orc_unwrap_connectivity.argtypes = [C.c_float, _f32p, _u64p, C.c_size_t, C.c_void_p, C.c_uint8, C.c_int,
                                              C.POINTER(C.c_int64)]
        L.orc_unwrap_connectivity.restype = C.c_int64
        L.orc_within_bounds.argtypes = [C.c_float, _f32p, _u64p, C.c_size_t, _f32p, _f32p]
        for suf, fp in (("f32", _f32p), ("f64", _f64p)):
            getattr(L, "orc_center_of_mass_" + suf).argtypes = [_f32p, _f32p, _u64p, C.c_size_t, fp]
            getattr(L, "orc_gyration_" + suf).argtypes = [_f32p, _f32p, _u64p, C.c_size_t, fp]
            getattr(L, "orc_rmsd_" + suf).argtypes = [_f32p, _u64p, C.c_size_t, _f32p, _u64p, C.c_size_t, fp]
            getattr(L, "orc_rmsd_mw_" + suf).argtypes = [_f32p, _f32p, _u64p, C.c_size_t, _f32p, _u64p,
                                                         C.c_size_t, fp]
            getattr(L, "orc_fit_transform_" + suf).argtypes = [_f32p, _f32p, _u64p, _f32p, _f32p, _u64p,
                                                               C.c_size_t, C.c_int, fp, fp]
        L.orc_apply_transform_f32.argtypes = [_f32p, _u64p, C.c_size_t, _f32p, _f32p]
        L.orc_apply_transform_f64.argtypes = [_f32p, _u64p, C.c_size_t, _f64p, _f64p, _f64p]
        L.orc_center_pbc.argtypes = [_f32p, _f32p, _u64p, C.c_size_t, C.c_void_p, C.c_uint8, C.c_int, _f64p]
        L.orc_center_of_geometry.argtypes = [_f32p, _u64p, C.c_size_t, C.c_int, _f64p]
        L.orc_center_of_geometry.restype = None
        L.orc_gyration_pbc.argtypes = [_f32p, _f32p, _u64p, C.c_size_t, C.c_void_p, C.c_int, _f64p]
        L.orc_inertia.argtypes = [_f32p, _f32p, _u64p, C.c_size_t, C.c_void_p, C.c_int, _f64p, _f64p, _f64p, _f64p]
        L.orc_synth_frame.argtypes = [C.c_uint64, C.c_uint64, C.c_size_t, _f32p, C.c_int, _f32p]
        L.orc_synth_masses.argtypes = [C.c_uint64, C.c_size_t, _f32p]
        _lib = L
    return _lib


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_f32p)


def _ids(ids):
    if ids is None:
        return None, None
    a = np.ascontiguousarray(ids, dtype=np.uint64)
    return a, a.ctypes.data_as(_u64p)


class Box:
    """PeriodicBox (molar/src/periodic_box.rs). matrix: 3x3, COLUMNS are the box vectors."""

    def __init__(self, matrix=None, vectors_angles=None):
        L = lib()
        if matrix is not None:
            m = np.asarray(matrix, dtype=np.float32).reshape(3, 3)
            m9 = np.ascontiguousarray(m.T.reshape(9))  # column-major storage
            self.h = L.orc_box_from_matrix(m9.ctypes.data_as(_f32p))
        else:
            self.h = L.orc_box_from_vectors_angles(*[C.c_float(float(np.float32(v))) for v in vectors_angles])
        if not self.h:
            raise ValueError("PeriodicBoxError")

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_box_free(self.h)
            self.h = None

    @property
    def matrix(self):
        m9 = np.zeros(9, np.float32)
        lib().orc_box_get(self.h, m9.ctypes.data_as(_f32p), None)
        return m9.reshape(3, 3).T.copy()

    @property
    def matrix9(self):
        m9 = np.zeros(9, np.float32)
        lib().orc_box_get(self.h, m9.ctypes.data_as(_f32p), None)
        return m9

    @property
    def inv(self):
        i9 = np.zeros(9, np.float32)
        lib().orc_box_get(self.h, None, i9.ctypes.data_as(_f32p))
        return i9.reshape(3, 3).T.copy()

    @property
    def corrections(self):
        n = lib().orc_box_corrections(self.h, None)
        out = np.zeros((n, 3), np.float32)
        if n:
            lib().orc_box_corrections(self.h, out.ctypes.data_as(_f32p))
        return out

    def lab_extents(self):
        out = np.zeros(3, np.float32)
        lib().orc_box_lab_extents(self.h, out.ctypes.data_as(_f32p))
        return out

    def shortest_vector(self, v, dims=7):
        a, p = _f32(v)
        out = np.zeros(3, np.float32)
        lib().orc_box_shortest_vector(self.h, p, dims, out.ctypes.data_as(_f32p))
        return out

    def distance_squared(self, p1, p2, dims=7):
        a, pa = _f32(p1)
        b, pb = _f32(p2)
        return float(lib().orc_box_distance_squared(self.h, pa, pb, dims))


def _take_pairs(h, with_dist=True):
    L = lib()
    n = L.orc_result_len(h)
    ij = np.zeros((n, 2), np.uint64)
    d = np.zeros(n, np.float32)
    dims = np.zeros(3, np.uint64)
    if n:
        L.orc_result_fill(h, ij.ctypes.data_as(_u64p), d.ctypes.data_as(_f32p))
    L.orc_result_grid_dims(h, dims.ctypes.data_as(_u64p))
    L.orc_result_free(h)
    return ij, d, dims


def _take_ids(h):
    L = lib()
    n = L.orc_result_len(h)
    ids = np.zeros(n, np.uint64)
    if n:
        L.orc_result_fill_ids(h, ids.ctypes.data_as(_u64p))
    L.orc_result_free(h)
    return ids


def search_single(cutoff, xyz, ids=None, box=None, pbc=0, nthreads=1):
    """Raw reference output (ij[n,2], d[n], grid dims)."""
    x, xp = _f32(xyz)
    i, ip = _ids(ids)
    n = len(i) if i is not None else x.size // 3
    if box is not None and pbc:
        h = lib().orc_search_single_pbc(cutoff, xp, ip, n, box.h, pbc, nthreads)
    else:
        h = lib().orc_search_single(cutoff, xp, ip, n, nthreads)
    return _take_pairs(h)


def mix64(x):
    """The pair hash of checksum_kernel / orc_search_single_pbc_checksum, vectorised (numpy uint64, wrapping)."""
    x = np.asarray(x, np.uint64).copy()
    with np.errstate(over="ignore"):
        x ^= x >> np.uint64(33)
        x *= np.uint64(0xff51afd7ed558ccd)
        x ^= x >> np.uint64(33)
        x *= np.uint64(0xc4ceb9fe1a85ec53)
        x ^= x >> np.uint64(33)
    return x


def pairs_checksum(ij):
    """(count, sum, xor) of mix64((min<<32)|max) over a pair list, as mb_pairs_checksum computes it."""
    ij = np.asarray(ij, np.uint64).reshape(-1, 2)
    lo, hi = ij.min(1), ij.max(1)
    h = mix64((lo << np.uint64(32)) | hi)
    with np.errstate(over="ignore"):
        s = int(h.sum(dtype=np.uint64)) if len(h) else 0
    x = int(np.bitwise_xor.reduce(h)) if len(h) else 0
    return len(h), s, x


def search_single_pbc_checksum(cutoff, xyz, box, pbc=7, ids=None, nthreads=1):
    """count, sum, xor, grid dims of distance_search_single_pbc without materialising the pair list."""
    x, xp = _f32(xyz)
    i, ip = _ids(ids)
    n = len(i) if i is not None else x.size // 3
    out = np.zeros(3, np.uint64)
    dims = np.zeros(3, np.uint64)
    lib().orc_search_single_pbc_checksum(cutoff, xp, ip, n, box.h, pbc, nthreads, out.ctypes.data_as(_u64p),
                                         dims.ctypes.data_as(_u64p))
    return int(out[0]), int(out[1]), int(out[2]), [int(d) for d in dims]


def search_double(cutoff, xyz1, ids1, xyz2, ids2, box=None, pbc=0, nthreads=1):
    x1, p1 = _f32(xyz1)
    x2, p2 = _f32(xyz2)
    i1, ip1 = _ids(ids1)
    i2, ip2 = _ids(ids2)
    n1 = len(i1) if i1 is not None else x1.size // 3
    n2 = len(i2) if i2 is not None else x2.size // 3
    if box is not None and pbc:
        h = lib().orc_search_double_pbc(cutoff, p1, ip1, n1, p2, ip2, n2, box.h, pbc, nthreads)
    else:
        h = lib().orc_search_double(cutoff, p1, ip1, n1, p2, ip2, n2, nthreads)
    return _take_pairs(h)


def search_double_vdw(xyz1, ids1, vdw1, xyz2, ids2, vdw2, box=None, pbc=0, nthreads=1):
    """Raw reference output with LOCAL indices (ij[n,2], d[n], grid dims)."""
    x1, p1 = _f32(xyz1)
    x2, p2 = _f32(xyz2)
    i1, ip1 = _ids(ids1)
    i2, ip2 = _ids(ids2)
    v1, vp1 = _f32(vdw1)
    v2, vp2 = _f32(vdw2)
    n1 = len(i1) if i1 is not None else x1.size // 3
    n2 = len(i2) if i2 is not None else x2.size // 3
    assert len(v1) == n1 and len(v2) == n2
    h = lib().orc_search_double_vdw(p1, ip1, n1, vp1, p2, ip2, n2, vp2, box.h if (box is not None and pbc) else None,
                                    pbc, nthreads)
    return _take_pairs(h)


def within_bounds(cutoff, xyz, ids=None):
    x, xp = _f32(xyz)
    i, ip = _ids(ids)
    n = len(i) if i is not None else x.size // 3
    lo = np.zeros(3, np.float32)
    up = np.zeros(3, np.float32)
    lib().orc_within_bounds(cutoff, xp, ip, n, lo.ctypes.data_as(_f32p), up.ctypes.data_as(_f32p))
    return lo, up


def search_within(cutoff, xyz1, ids1, xyz2, ids2, box=None, pbc=0, lower=None, upper=None, nthreads=1):
    """Raw id list (may repeat), as distance_search_within[_pbc] returns it."""
    x1, p1 = _f32(xyz1)
    x2, p2 = _f32(xyz2)
    i1, ip1 = _ids(ids1)
    i2, ip2 = _ids(ids2)
    n1 = len(i1) if i1 is not None else x1.size // 3
    n2 = len(i2) if i2 is not None else x2.size // 3
    if box is not None and pbc:
        h = lib().orc_search_within_pbc(cutoff, p1, ip1, n1, p2, ip2, n2, box.h, pbc, nthreads)
    else:
        lo, lp = _f32(lower)
        up, upp = _f32(upper)
        h = lib().orc_search_within(cutoff, p1, ip1, n1, p2, ip2, n2, lp, upp, nthreads)
    return _take_ids(h)


def canonical_pairs(ij, d=None):
    """Sorted unique set of (min,max) pairs [+ the min distance reported for each]."""
    ij = np.asarray(ij, dtype=np.uint64).reshape(-1, 2)
    lo = np.minimum(ij[:, 0], ij[:, 1])
    hi = np.maximum(ij[:, 0], ij[:, 1])
    key = (lo << np.uint64(32)) | hi
    if d is None:
        key = np.unique(key)
        return np.stack([key >> np.uint64(32), key & np.uint64(0xFFFFFFFF)], axis=1)
    d = np.asarray(d, dtype=np.float32)
    order = np.lexsort((d, key))
    key, d = key[order], d[order]
    first = np.ones(len(key), bool)
    first[1:] = key[1:] != key[:-1]
    key, d = key[first], d[first]
    return np.stack([key >> np.uint64(32), key & np.uint64(0xFFFFFFFF)], axis=1), d


def ordered_pairs(ij, d=None):
    """Sorted unique set of ORDERED (i,j) pairs (double search: i from set 1, j from set 2)."""
    ij = np.asarray(ij, dtype=np.uint64).reshape(-1, 2)
    key = (ij[:, 0] << np.uint64(32)) | ij[:, 1]
    if d is None:
        key = np.unique(key)
        return np.stack([key >> np.uint64(32), key & np.uint64(0xFFFFFFFF)], axis=1)
    d = np.asarray(d, dtype=np.float32)
    order = np.lexsort((d, key))
    key, d = key[order], d[order]
    first = np.ones(len(key), bool)
    first[1:] = key[1:] != key[:-1]
    key, d = key[first], d[first]
    return np.stack([key >> np.uint64(32), key & np.uint64(0xFFFFFFFF)], axis=1), d


def _measure(name, prec):
    return getattr(lib(), f"orc_{name}_{prec}")


def center_of_mass(xyz, masses, ids=None, prec="f64"):
    x, xp = _f32(xyz)
    m, mp = _f32(masses)
    i, ip = _ids(ids)
    n = len(i) if i is not None else x.size // 3
    out = np.zeros(3, np.float64 if prec == "f64" else np.float32)
    rc = _measure("center_of_mass", prec)(xp, mp, ip, n, out.ctypes.data_as(_f64p if prec == "f64" else _f32p))
    return rc, out


def gyration(xyz, masses, ids=None, prec="f64"):
    x, xp = _f32(xyz)
    m, mp = _f32(masses)
    i, ip = _ids(ids)
    n = len(i) if i is not None else x.size // 3
    out = np.zeros(1, np.float64 if prec == "f64" else np.float32)
    rc = _measure("gyration", prec)(xp, mp, ip, n, out.ctypes.data_as(_f64p if prec == "f64" else _f32p))
    return rc, float(out[0])


def unwrap_connectivity(cutoff, xyz, box, ids=None, dims=7, nthreads=4):
    """-> (unwrapped copy of xyz, roots[n]: start atom of each selected atom's walk, number of start atoms)"""
    x = np.ascontiguousarray(xyz, dtype=np.float32).copy()
    i, ip = _ids(ids)
    n = len(i) if i is not None else x.size // 3
    roots = np.zeros(n, np.int64)
    ns = lib().orc_unwrap_connectivity(cutoff, x.ctypes.data_as(_f32p), ip, n, box.h, dims, nthreads,
                                       roots.ctypes.data_as(C.POINTER(C.c_int64)))
    return x, roots, int(ns)


_PREC = {"f32": 0, "f64": 1, "mixed": 2}


def center_pbc(xyz, masses, box, dims=7, ids=None, prec="mixed"):
    """center_of_mass_pbc[_dims] (masses given) / center_of_geometry_pbc[_dims] (masses None)."""
    x, xp = _f32(xyz)
    mp = None
    if masses is not None:
        m, mp = _f32(masses)
    i, ip = _ids(ids)
    n = len(i) if i is not None else x.size // 3
    out = np.zeros(3, np.float64)
    rc = lib().orc_center_pbc(xp, mp, ip, n, box.h, dims, _PREC[prec], out.ctypes.data_as(_f64p))
    return rc, out


def center_of_geometry(xyz, ids=None, prec="f64"):
    x, xp = _f32(xyz)
    i, ip = _ids(ids)
    n = len(i) if i is not None else x.size // 3
    out = np.zeros(3, np.float64)
    lib().orc_center_of_geometry(xp, ip, n, _PREC[prec], out.ctypes.data_as(_f64p))
    return out


def gyration_pbc(xyz, masses, box, ids=None, prec="mixed"):
    x, xp = _f32(xyz)
    m, mp = _f32(masses)
    i, ip = _ids(ids)
    n = len(i) if i is not None else x.size // 3
    out = np.zeros(1, np.float64)
    rc = lib().orc_gyration_pbc(xp, mp, ip, n, box.h, _PREC[prec], out.ctypes.data_as(_f64p))
    return rc, float(out[0])


def inertia(xyz, masses, box=None, ids=None, prec="mixed"):
    """returns rc, tensor[3,3], moments[3], axes[3,3] (columns = axes), centre[3]"""
    x, xp = _f32(xyz)
    m, mp = _f32(masses)
    i, ip = _ids(ids)
    n = len(i) if i is not None else x.size // 3
    t = np.zeros(9, np.float64)
    mom = np.zeros(3, np.float64)
    ax = np.zeros(9, np.float64)
    c = np.zeros(3, np.float64)
    rc = lib().orc_inertia(xp, mp, ip, n, box.h if box is not None else None, _PREC[prec], t.ctypes.data_as(_f64p),
                           mom.ctypes.data_as(_f64p), ax.ctypes.data_as(_f64p), c.ctypes.data_as(_f64p))
    return rc, t.reshape(3, 3), mom, ax.reshape(3, 3).T.copy(), c


def rmsd(xyz1, ids1, xyz2, ids2, prec="f64", masses1=None):
    x1, p1 = _f32(xyz1)
    x2, p2 = _f32(xyz2)
    i1, ip1 = _ids(ids1)
    i2, ip2 = _ids(ids2)
    n1 = len(i1) if i1 is not None else x1.size // 3
    n2 = len(i2) if i2 is not None else x2.size // 3
    out = np.zeros(1, np.float64 if prec == "f64" else np.float32)
    op = out.ctypes.data_as(_f64p if prec == "f64" else _f32p)
    if masses1 is None:
        rc = _measure("rmsd", prec)(p1, ip1, n1, p2, ip2, n2, op)
    else:
        m, mp = _f32(masses1)
        rc = _measure("rmsd_mw", prec)(p1, mp, ip1, n1, p2, ip2, n2, op)
    return rc, float(out[0])


def fit_transform(xyz1, masses1, ids1, xyz2, masses2, ids2, at_origin=False, prec="f64"):
    x1, p1 = _f32(xyz1)
    x2, p2 = _f32(xyz2)
    m1, mp1 = _f32(masses1)
    m2, mp2 = _f32(masses2)
    i1, ip1 = _ids(ids1)
    i2, ip2 = _ids(ids2)
    n = len(i1) if i1 is not None else x1.size // 3
    dt = np.float64 if prec == "f64" else np.float32
    pt = _f64p if prec == "f64" else _f32p
    R9 = np.zeros(9, dt)
    t3 = np.zeros(3, dt)
    rc = _measure("fit_transform", prec)(p1, mp1, ip1, p2, mp2, ip2, n, int(at_origin),
                                          R9.ctypes.data_as(pt), t3.ctypes.data_as(pt))
    return rc, R9.reshape(3, 3).T.copy(), t3


def apply_transform_f64(xyz, ids, R, t):
    x, xp = _f32(xyz)
    i, ip = _ids(ids)
    n = len(i) if i is not None else x.size // 3
    R9 = np.ascontiguousarray(np.asarray(R, np.float64).T.reshape(9))
    t3 = np.ascontiguousarray(np.asarray(t, np.float64))
    out = np.zeros((n, 3), np.float64)
    lib().orc_apply_transform_f64(xp, ip, n, R9.ctypes.data_as(_f64p), t3.ctypes.data_as(_f64p),
                                  out.ctypes.data_as(_f64p))
    return out


def apply_transform_f32(xyz, ids, R, t):
    x = np.array(xyz, dtype=np.float32, copy=True)
    i, ip = _ids(ids)
    n = len(i) if i is not None else x.size // 3
    R9 = np.ascontiguousarray(np.asarray(R, np.float32).T.reshape(9))
    t3 = np.ascontiguousarray(np.asarray(t, np.float32))
    lib().orc_apply_transform_f32(x.ctypes.data_as(_f32p), ip, n, R9.ctypes.data_as(_f32p),
                                  t3.ctypes.data_as(_f32p))
    return x


def synth_frame(seed, frame, n_atoms, box_matrix, stray_permille=0):
    m = np.asarray(box_matrix, dtype=np.float32).reshape(3, 3)
    m9 = np.ascontiguousarray(m.T.reshape(9))
    out = np.zeros((n_atoms, 3), np.float32)
    lib().orc_synth_frame(seed, frame, n_atoms, m9.ctypes.data_as(_f32p), stray_permille,
                          out.ctypes.data_as(_f32p))
    return out


def synth_masses(seed, n_atoms):
    out = np.zeros(n_atoms, np.float32)
    lib().orc_synth_masses(seed, n_atoms, out.ctypes.data_as(_f32p))
    return out
