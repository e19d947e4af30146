"""Host-side mirror of the reference's Python interface for the hot path.

Same names, argument meaning and error behaviour as pymolar
(molar_python/src/lib.rs:137-376; Sel.com/gyration/apply_transform: molar_python/src/selection.rs:816-941):

    pairs, dist = distance_search(cutoff, sel1, sel2=None, dims=None)
    r = rmsd(sel1, sel2); rw = rmsd_mw(sel1, sel2)
    tr = fit_transform(sel1, sel2); sel1.apply_transform(tr)
    sel.com(); sel.gyration()

A `System` owns one device context (one per calling thread: Send, not Sync) holding the frame,
the mass column and the box; a `Sel` is a sorted array of global indices into it, exactly like
MolAR's Sel (providers.rs:45-48).  Everything is computed by libmolar_b200.so on the GPU.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import MolarB200Error, check, f32p, f64p, u64p, i64p


def _pbc_bits(dims):
    if dims is None:
        return 0
    if isinstance(dims, (int, np.integer)):
        return int(dims) & 7
    return (1 if dims[0] else 0) | (2 if dims[1] else 0) | (4 if dims[2] else 0)


class PeriodicBox:
    """matrix: 3x3, COLUMNS are the box vectors a,b,c (periodic_box.rs:9-13)."""

    def __init__(self, matrix):
        self.matrix = np.asarray(matrix, dtype=np.float32).reshape(3, 3)

    @property
    def colmajor9(self):
        return np.ascontiguousarray(self.matrix.T.reshape(9))

    def get_matrix(self):
        return self.matrix.copy()


class IsometryTransform:
    """p' = R p + t (nalgebra IsometryMatrix3, what fit_transform returns)."""

    def __init__(self, R, t):
        self.R = np.asarray(R, dtype=np.float64).reshape(3, 3)
        self.t = np.asarray(t, dtype=np.float64).reshape(3)


class System:
    def __init__(self, coords, masses=None, box=None, device=0, vdw=None):
        self._lib = _capi.load()
        self._h = self._lib.mb_open(device)
        if not self._h:
            raise MolarB200Error(_capi.MB_ERR_CUDA, _capi.last_error())
        self._n = 0
        self._version = 0        # bumped whenever the device frame changes
        self._frame2_of = None   # (other system id, its version) currently staged as frame2
        self.box = None
        self.set_state(coords, box)
        if masses is not None:
            self.set_masses(masses)
        self.vdw = None if vdw is None else np.ascontiguousarray(vdw, dtype=np.float32)  # per-atom vdW radii

    # -- lifetime --
    def close(self):
        if getattr(self, "_h", None):
            self._lib.mb_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return self._n

    # -- state (System::set_state, selection/system.rs:230-236) --
    def set_state(self, coords, box=None):
        xyz = np.ascontiguousarray(coords, dtype=np.float32).reshape(-1, 3)
        if box is not None and not isinstance(box, PeriodicBox):
            box = PeriodicBox(box)
        self.box = box
        b9 = box.colmajor9.ctypes.data_as(f32p) if box is not None else None
        check(self._lib.mb_set_frame(self._h, xyz.ctypes.data, xyz.shape[0], b9))
        self._n = xyz.shape[0]
        self._version += 1

    def set_masses(self, masses):
        m = np.ascontiguousarray(masses, dtype=np.float32).reshape(-1)
        if m.shape[0] != self._n:
            raise ValueError(f"masses: {m.shape[0]} != {self._n} atoms")
        check(self._lib.mb_set_masses(self._h, m.ctypes.data, m.shape[0]))

    def set_option(self, key, value):
        check(self._lib.mb_set_option(self._h, key.encode(), float(value)))
        if key == "with_dist":
            self._with_dist = bool(value)

    def coords(self):
        out = np.empty((self._n, 3), np.float32)
        check(self._lib.mb_get_frame(self._h, out.ctypes.data, self._n))
        return out

    # -- selections (System::__call__) --
    def __call__(self, arg=None):
        if arg is None:
            return Sel(self, None, self._n)
        if isinstance(arg, tuple) and len(arg) == 2:
            idx = np.arange(arg[0], arg[1] + 1, dtype=np.uint64)  # inclusive range, like pymolar
        else:
            idx = np.unique(np.asarray(arg, dtype=np.uint64))  # Sel indices are a sorted set
        if idx.size == 0:
            raise ValueError("empty selection")
        if int(idx[-1]) >= self._n:
            raise IndexError(f"index {int(idx[-1])} out of range ({self._n} atoms)")
        return Sel(self, idx, idx.size)

    def launch_count(self):
        return int(self._lib.mb_launch_count(self._h))

    def synchronize(self):
        check(self._lib.mb_synchronize(self._h))

    def fill_pairs_u32(self, out):
        """Last pair list as u32 pairs (canonical i<j) into `out` (uint32 [>=P, 2], ideally page-locked)."""
        check(self._lib.mb_fill_pairs_u32(self._h, out.ctypes.data, None))

    _REDUCE = {"com": (0, 3), "cog": (1, 3), "gyration": (2, 1), "com_gyration": (3, 4)}

    def reduce_many(self, selections, what="com", return_status=False):
        """COM / COG / gyration of MANY selections in one launch (mb_reduce_many) — the device form of the reference's
        rayon loop over split selections (selection.rs:318-322, selection/par_split.rs:100-125).
        selections: a list of index arrays, or a tuple (ids, offsets) in CSR form.  Returns [n_sel, width] float64."""
        if isinstance(selections, tuple):
            ids = np.ascontiguousarray(selections[0], dtype=np.uint64)
            offsets = np.ascontiguousarray(selections[1], dtype=np.uint64)
        else:
            arrs = [np.unique(np.asarray(a, dtype=np.uint64)) for a in selections]
            offsets = np.zeros(len(arrs) + 1, np.uint64)
            offsets[1:] = np.cumsum([len(a) for a in arrs])
            ids = np.concatenate(arrs) if arrs else np.zeros(0, np.uint64)
        code, width = self._REDUCE[what]
        n_sel = len(offsets) - 1
        out = np.empty((n_sel, width), np.float64)
        status = np.zeros(n_sel, np.int32)
        rc = self._lib.mb_reduce_many(self._h, ids.ctypes.data_as(u64p), offsets.ctypes.data_as(u64p), n_sel, code,
                                      out.ctypes.data_as(f64p), status.ctypes.data_as(C.POINTER(C.c_int)))
        if return_status:
            if rc < 0 and not status.any():
                check(rc)
            return out, status
        check(rc)
        return out

    def connectivity(self):
        """Adjacency of the last pair list as CSR (row_ptr[n+1], cols): SearchConnectivity (connectivity.rs:8-38)."""
        row_ptr = np.zeros(self._n + 1, np.uint64)
        nnz = check(self._lib.mb_connectivity(self._h, self._n, row_ptr.ctypes.data_as(u64p)))
        cols = np.zeros(nnz, np.uint64)
        if nnz:
            check(self._lib.mb_fill_connectivity(self._h, cols.ctypes.data_as(u64p)))
        return row_ptr, cols

    def connectivity_checksum(self):
        """(sum, xor) of mix64((min << 32) | max) over the adjacency entries with row < column and over those with
        row > column, computed on the device."""
        out = np.zeros(4, np.uint64)
        check(self._lib.mb_connectivity_checksum(self._h, out.ctypes.data_as(u64p)))
        return int(out[0]), int(out[1]), int(out[2]), int(out[3])


class Sel:
    def __init__(self, system, index, n):
        self.sys = system
        self.index = index  # None = all atoms
        self._n = int(n)

    def __len__(self):
        return self._n

    def _ids(self):
        return (self.index.ctypes.data_as(u64p) if self.index is not None else None), self._n

    def get_index(self):
        return np.arange(self._n, dtype=np.uint64) if self.index is None else self.index.copy()

    def com(self, dims=None):
        """Centre of mass; with periodic `dims` every atom is taken at its closest image next to the first
        atom (center_of_mass_pbc_dims, molar_python/src/selection.rs:816-829)."""
        out = np.zeros(3, np.float64)
        p, n = self._ids()
        bits = _pbc_bits(dims)
        if bits:
            check(self.sys._lib.mb_center_pbc(self.sys._h, p, n, 1, bits, out.ctypes.data_as(f64p)))
        else:
            check(self.sys._lib.mb_center_of_mass(self.sys._h, p, n, out.ctypes.data_as(f64p)))
        return out

    def cog(self, dims=None):
        """Centre of geometry, optionally periodic (selection.rs:845-858)."""
        out = np.zeros(3, np.float64)
        p, n = self._ids()
        bits = _pbc_bits(dims)
        if bits:
            check(self.sys._lib.mb_center_pbc(self.sys._h, p, n, 0, bits, out.ctypes.data_as(f64p)))
        else:
            check(self.sys._lib.mb_center_of_geometry(self.sys._h, p, n, out.ctypes.data_as(f64p)))
        return out

    def gyration(self, pbc=False):
        """Radius of gyration (selection.rs:935-941)."""
        out = C.c_double(0.0)
        p, n = self._ids()
        if pbc:
            check(self.sys._lib.mb_gyration_pbc(self.sys._h, p, n, C.byref(out)))
        else:
            check(self.sys._lib.mb_gyration(self.sys._h, p, n, C.byref(out)))
        return out.value

    def gyration_pbc(self):
        return self.gyration(pbc=True)

    def inertia(self, pbc=False):
        """(moments[3] ascending, axes[3,3] with the principal axes as columns) (selection.rs:1068-1084)."""
        mom = np.zeros(3, np.float64)
        ax = np.zeros(9, np.float64)
        p, n = self._ids()
        check(self.sys._lib.mb_inertia(self.sys._h, p, n, 1 if pbc else 0, mom.ctypes.data_as(f64p),
                                       ax.ctypes.data_as(f64p)))
        return mom, ax.reshape(3, 3).T.copy()

    def search_connectivity(self, cutoff, dims=None):
        """SearchConnectivity of distance_search_single[_pbc](cutoff, self) (connectivity.rs:8-38 over
        distance_search.rs:892-954) as CSR (row_ptr[n_atoms + 1], cols), the rows written by the search kernel itself:
        no pair list in between.  dims=None: non-periodic."""
        bits = 0 if dims is None else _pbc_bits(dims)
        p, n = self._ids()
        row_ptr = np.zeros(self.sys._n + 1, np.uint64)
        nnz = check(self.sys._lib.mb_search_connectivity(self.sys._h, cutoff, p, n, bits, row_ptr.ctypes.data_as(u64p)))
        cols = np.zeros(nnz, np.uint64)
        if nnz:
            check(self.sys._lib.mb_fill_connectivity(self.sys._h, cols.ctypes.data_as(u64p)))
        return row_ptr, cols

    def inertia_pbc(self):
        return self.inertia(pbc=True)

    def principal_transform(self, pbc=False):
        """Rigid transform onto the principal axes (selection.rs:866-873)."""
        R9 = np.zeros(9, np.float64)
        t3 = np.zeros(3, np.float64)
        p, n = self._ids()
        check(self.sys._lib.mb_principal_transform(self.sys._h, p, n, 1 if pbc else 0, R9.ctypes.data_as(f64p),
                                                   t3.ctypes.data_as(f64p)))
        return IsometryTransform(R9.reshape(3, 3).T.copy(), t3)

    def principal_transform_pbc(self):
        return self.principal_transform(pbc=True)

    def unwrap_connectivity(self, cutoff, dims=None):
        """Make every connected group of the selection whole (Modify::unwrap_connectivity_dim, modify.rs:72-131).
        Returns, like the reference, one selection per start atom holding the atoms reached from it (the start
        atom itself is not included and isolated atoms yield nothing)."""
        bits = 7 if dims is None else _pbc_bits(dims)
        p, n = self._ids()
        roots = np.zeros(n, np.int64)
        check(self.sys._lib.mb_unwrap_connectivity(self.sys._h, cutoff, p, n, bits, roots.ctypes.data_as(i64p)))
        self.sys._version += 1
        idx = self.get_index()
        out = []
        order = np.argsort(roots, kind="stable")
        bounds = np.flatnonzero(np.diff(roots[order])) + 1
        for grp in np.split(order, bounds):
            members = grp[grp != roots[grp[0]]]
            if members.size:
                out.append(Sel(self.sys, idx[members].astype(np.uint64), members.size))
        self.roots = roots
        return out

    def apply_transform(self, tr):
        R9 = np.ascontiguousarray(tr.R.T.reshape(9), dtype=np.float64)
        t3 = np.ascontiguousarray(tr.t, dtype=np.float64)
        p, n = self._ids()
        check(self.sys._lib.mb_apply_transform(self.sys._h, p, n, R9.ctypes.data_as(f64p), t3.ctypes.data_as(f64p)))
        self.sys._version += 1


def _same_or_frame2(sel1, sel2):
    """sel2 may live in another System (the reference structure): stage its frame as frame2."""
    if sel2.sys is sel1.sys:
        return 0
    tag = (id(sel2.sys), sel2.sys._version)
    if sel1.sys._frame2_of != tag:  # re-stage only when the other system's frame changed
        xyz2 = sel2.sys.coords()
        check(sel1.sys._lib.mb_set_frame2(sel1.sys._h, xyz2.ctypes.data, xyz2.shape[0]))
        sel1.sys._frame2_of = tag
    return 1


def distance_search(cutoff, data1, data2=None, dims=None):
    """pymolar.distance_search (molar_python/src/lib.rs:254-376): (pairs[N,2], distances[N]).

    The pair SET equals the reference's; order is unspecified (the reference's order is rayon's).
    Single-selection pairs are canonical i<j and de-duplicated."""
    s = data1.sys
    pbc = _pbc_bits(dims)
    if isinstance(cutoff, str):
        # pymolar: cutoff == "vdw" -> distance_search_double_vdw[_pbc] on the atoms' vdW radii, local
        # indices converted to global afterwards (molar_python/src/lib.rs:311-349)
        if cutoff != "vdw":
            raise TypeError(f"Unknown cutoff type {cutoff}")
        if data2 is None:
            raise NotImplementedError("VdW distance search is not yet supported for single selection")
        if pbc and s.box is None:
            raise MolarB200Error(_capi.MB_ERR_NO_PBC, "pbc operation without periodic box")
        v1 = np.ascontiguousarray(data1.sys.vdw[data1.get_index().astype(np.int64)], dtype=np.float32)
        v2 = np.ascontiguousarray(data2.sys.vdw[data2.get_index().astype(np.int64)], dtype=np.float32)
        use2 = _same_or_frame2(data1, data2)
        p1, n1 = data1._ids()
        p2, n2 = data2._ids()
        cnt = check(s._lib.mb_search_double_vdw(s._h, p1, n1, v1.ctypes.data_as(f32p), p2, n2,
                                                v2.ctypes.data_as(f32p), use2, pbc))
        pairs = np.empty((cnt, 2), np.uint64)
        dist = np.empty(cnt, np.float32)
        if cnt:
            check(s._lib.mb_fill_pairs(s._h, pairs.ctypes.data, dist.ctypes.data))
            pairs[:, 0] = data1.get_index()[pairs[:, 0].astype(np.int64)]
            pairs[:, 1] = data2.get_index()[pairs[:, 1].astype(np.int64)]
        return pairs, dist
    if pbc and s.box is None:
        raise MolarB200Error(_capi.MB_ERR_NO_PBC, "pbc operation without periodic box")
    p1, n1 = data1._ids()
    if data2 is None:
        cnt = check(s._lib.mb_search_single(s._h, cutoff, p1, n1, pbc))
    else:
        use2 = _same_or_frame2(data1, data2)
        p2, n2 = data2._ids()
        cnt = check(s._lib.mb_search_double(s._h, cutoff, p1, n1, p2, n2, use2, pbc))
    pairs = np.empty((cnt, 2), np.uint64)
    # option with_dist=0 (pairs-only kernel): the distances were not computed and None is returned in their place
    dist = np.empty(cnt, np.float32) if getattr(s, "_with_dist", True) else None
    if cnt:
        check(s._lib.mb_fill_pairs(s._h, pairs.ctypes.data, dist.ctypes.data if dist is not None else None))
    return pairs, dist


def within(cutoff, data1, data2, dims=None, lower=None, upper=None):
    """distance_search_within[_pbc] (distance_search.rs:519-598) as the `within` AST node calls it
    (selection/ast.rs:589-631): sorted unique ids of data1 atoms within cutoff of any data2 atom."""
    s = data1.sys
    pbc = _pbc_bits(dims)
    use2 = _same_or_frame2(data1, data2)
    p1, n1 = data1._ids()
    p2, n2 = data2._ids()
    lo = np.ascontiguousarray(lower, np.float32).ctypes.data_as(f32p) if lower is not None else None
    up = np.ascontiguousarray(upper, np.float32).ctypes.data_as(f32p) if upper is not None else None
    cnt = check(s._lib.mb_search_within(s._h, cutoff, p1, n1, p2, n2, use2, pbc, lo, up))
    out = np.empty(cnt, np.uint64)
    if cnt:
        check(s._lib.mb_fill_ids(s._h, out.ctypes.data))
    return out


def rmsd(sel1, sel2):
    """measure.rs:485-504; raises on size mismatch like MeasureError::Sizes."""
    s = sel1.sys
    use2 = _same_or_frame2(sel1, sel2)
    p1, n1 = sel1._ids()
    p2, n2 = sel2._ids()
    out = C.c_double(0.0)
    check(s._lib.mb_rmsd(s._h, p1, n1, p2, n2, use2, 0, C.byref(out)))
    return out.value


rmsd_py = rmsd


def rmsd_mw(sel1, sel2):
    s = sel1.sys
    use2 = _same_or_frame2(sel1, sel2)
    p1, n1 = sel1._ids()
    p2, n2 = sel2._ids()
    out = C.c_double(0.0)
    check(s._lib.mb_rmsd(s._h, p1, n1, p2, n2, use2, 1, C.byref(out)))
    return out.value


def fit_transform(sel1, sel2, at_origin=False):
    """Transform that best fits sel1 ONTO sel2 (measure.rs:507-535)."""
    s = sel1.sys
    use2 = _same_or_frame2(sel1, sel2)
    p1, n1 = sel1._ids()
    p2, n2 = sel2._ids()
    R9 = np.zeros(9, np.float64)
    t3 = np.zeros(3, np.float64)
    check(s._lib.mb_fit_transform(s._h, p1, n1, p2, n2, use2, int(at_origin), R9.ctypes.data_as(f64p),
                                  t3.ctypes.data_as(f64p)))
    return IsometryTransform(R9.reshape(3, 3).T.copy(), t3)


_TRAJ_FORMATS = {"dcd": 0, "xtc": 1}


def probe_trajectory(data, fmt):
    """(frames, atoms) of a DCD / XTC byte stream (headers only, on the host)."""
    buf = np.frombuffer(data, dtype=np.uint8)
    nf, na = C.c_size_t(0), C.c_size_t(0)
    check(_capi.load().mb_traj_probe(buf.ctypes.data, buf.size, _TRAJ_FORMATS[fmt], C.byref(nf), C.byref(na)))
    return nf.value, na.value


def load_trajectory(data, fmt, first_frame=0, n_frames=None, device=0):
    return Trajectory(device).load(data, fmt, first_frame, n_frames)


class Trajectory:
    """Device-resident block of frames: the GPU analogue of the per-frame loop that
    AnalysisTask::run drives (analysis_task.rs:113-280)."""

    def __init__(self, device=0):
        self._lib = _capi.load()
        self._h = self._lib.mb_open(device)
        if not self._h:
            raise MolarB200Error(_capi.MB_ERR_CUDA, _capi.last_error())
        self.n_frames = 0
        self.n_atoms = 0

    def close(self):
        if getattr(self, "_h", None):
            self._lib.mb_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, key, value):
        check(self._lib.mb_set_option(self._h, key.encode(), float(value)))

    def synth(self, seed, first_frame, n_frames, n_atoms, box, stray_permille=0, mass_seed=None):
        b = box if isinstance(box, PeriodicBox) else PeriodicBox(box)
        check(self._lib.mb_batch_synth(self._h, seed, first_frame, n_frames, n_atoms,
                                       b.colmajor9.ctypes.data_as(f32p), stray_permille))
        if mass_seed is not None:
            check(self._lib.mb_batch_synth_masses(self._h, mass_seed, n_atoms))
        self.n_frames, self.n_atoms = n_frames, n_atoms

    def upload(self, frames, box=None, masses=None):
        x = np.ascontiguousarray(frames, dtype=np.float32)
        nf, na = x.shape[0], x.shape[1]
        b9 = None
        if box is not None:
            b = box if isinstance(box, PeriodicBox) else PeriodicBox(box)
            b9 = b.colmajor9.ctypes.data_as(f32p)
        check(self._lib.mb_batch_upload(self._h, x.ctypes.data, nf, na, b9))
        if masses is not None:
            m = np.ascontiguousarray(masses, dtype=np.float32)
            check(self._lib.mb_set_masses(self._h, m.ctypes.data, m.shape[0]))
        self.n_frames, self.n_atoms = nf, na

    def load(self, data, fmt, first_frame=0, n_frames=None):
        """Decode frames of a DCD / XTC byte stream on the device into the resident batch
        (FileHandler::read_state for every frame, io/dcd_handler.rs:389-464, io/xtc_handler.rs:64-110)."""
        buf = np.frombuffer(data, dtype=np.uint8)
        code = _TRAJ_FORMATS[fmt]
        if n_frames is None:
            n_frames = probe_trajectory(data, fmt)[0] - first_frame
        boxes = np.zeros((n_frames, 9), np.float32)
        times = np.zeros(n_frames, np.float32)
        check(self._lib.mb_batch_load_traj(self._h, buf.ctypes.data, buf.size, code, first_frame, n_frames,
                                           boxes.ctypes.data_as(f32p), times.ctypes.data_as(f32p)))
        self.n_frames = n_frames
        self.n_atoms = probe_trajectory(data, fmt)[1]
        self.boxes = boxes.reshape(n_frames, 3, 3).transpose(0, 2, 1).copy()  # [f] = 3x3, columns = box vectors
        self.times = times
        return self

    def frames(self, f0=0, f1=None):
        f1 = self.n_frames if f1 is None else f1
        out = np.empty((f1 - f0, self.n_atoms, 3), np.float32)
        check(self._lib.mb_batch_download(self._h, f0, f1, out.ctypes.data))
        return out

    def frame(self, f):
        check(self._lib.mb_batch_select(self._h, f))
        out = np.empty((self.n_atoms, 3), np.float32)
        check(self._lib.mb_get_frame(self._h, out.ctypes.data, self.n_atoms))
        return out

    def search(self, cutoff, dims=(True, True, True), f0=0, f1=None, count_only=False, checksums=False):
        f1 = self.n_frames if f1 is None else f1
        counts = np.zeros(f1 - f0, np.int64)
        chk = np.zeros((f1 - f0, 2), np.uint64) if checksums else None
        check(self._lib.mb_batch_search(self._h, cutoff, _pbc_bits(dims), f0, f1, 1 if count_only else 0,
                                        counts.ctypes.data_as(i64p),
                                        chk.ctypes.data_as(u64p) if checksums else None))
        return (counts, chk) if checksums else counts

    def stream_search(self, frames, cutoff, box, dims=(True, True, True), count_only=False):
        """Neighbour search over HOST frames [F][N][3] (upload overlapped with the search of the previous chunk)."""
        x = frames if isinstance(frames, np.ndarray) and frames.dtype == np.float32 and frames.flags.c_contiguous \
            else np.ascontiguousarray(frames, dtype=np.float32)
        nf, na = x.shape[0], x.shape[1]
        b = box if isinstance(box, PeriodicBox) else PeriodicBox(box)
        counts = np.zeros(nf, np.int64)
        check(self._lib.mb_stream_search(self._h, cutoff, _pbc_bits(dims), x.ctypes.data, nf, na,
                                         b.colmajor9.ctypes.data_as(f32p), 1 if count_only else 0,
                                         counts.ctypes.data_as(i64p)))
        self.n_frames, self.n_atoms = min(nf, 8), na
        return counts

    def stream_fit(self, frames, masses):
        """Kabsch fit of every HOST frame onto the first one; returns the RMSD after the fit per frame."""
        x = frames if isinstance(frames, np.ndarray) and frames.dtype == np.float32 and frames.flags.c_contiguous \
            else np.ascontiguousarray(frames, dtype=np.float32)
        nf, na = x.shape[0], x.shape[1]
        if masses is not None:
            m = np.ascontiguousarray(masses, dtype=np.float32)
            check(self._lib.mb_set_masses(self._h, m.ctypes.data, m.shape[0]))
        out = np.zeros(nf, np.float64)
        check(self._lib.mb_stream_fit(self._h, x.ctypes.data, nf, na, out.ctypes.data_as(f64p)))
        return out

    def stream_pipeline(self, frames, cutoff, box, masses=None, dims=(True, True, True)):
        """COM + gyration + contact count per HOST frame -> [F,5] {com_x, com_y, com_z, rg, count}."""
        x = frames if isinstance(frames, np.ndarray) and frames.dtype == np.float32 and frames.flags.c_contiguous \
            else np.ascontiguousarray(frames, dtype=np.float32)
        nf, na = x.shape[0], x.shape[1]
        if masses is not None:
            m = np.ascontiguousarray(masses, dtype=np.float32)
            check(self._lib.mb_set_masses(self._h, m.ctypes.data, m.shape[0]))
        b = box if isinstance(box, PeriodicBox) else PeriodicBox(box)
        out = np.zeros((nf, 5), np.float64)
        check(self._lib.mb_stream_pipeline(self._h, cutoff, _pbc_bits(dims), x.ctypes.data, nf, na,
                                           b.colmajor9.ctypes.data_as(f32p), out.ctypes.data_as(f64p)))
        return out

    def last_pairs(self):
        """Pair list of the last frame searched (host copy)."""
        n = C.c_int64(0)
        self._lib.mb_pairs_device(self._h, C.byref(n))
        pairs = np.empty((n.value, 2), np.uint64)
        if n.value:
            check(self._lib.mb_fill_pairs(self._h, pairs.ctypes.data, None))
        return pairs

    def fit(self, ref_frame=0, f0=0, f1=None, superpose=True):
        f1 = self.n_frames if f1 is None else f1
        out = np.zeros(f1 - f0, np.float64)
        check(self._lib.mb_batch_fit(self._h, ref_frame, f0, f1, int(superpose), out.ctypes.data_as(f64p)))
        return out

    def pipeline(self, cutoff, dims=(True, True, True), f0=0, f1=None):
        f1 = self.n_frames if f1 is None else f1
        out = np.zeros((f1 - f0, 5), np.float64)
        check(self._lib.mb_batch_pipeline(self._h, cutoff, _pbc_bits(dims), f0, f1, out.ctypes.data_as(f64p)))
        return out

    def masses_host(self):
        out = np.empty(self.n_atoms, np.float32)
        check(self._lib.mb_get_masses(self._h, out.ctypes.data, self.n_atoms))
        return out

    def scalars_device(self):
        rows, rd = C.c_size_t(0), C.c_size_t(0)
        p = self._lib.mb_batch_scalars_device(self._h, C.byref(rows), C.byref(rd))
        return p, rows.value, rd.value

    def stream(self):
        return self._lib.mb_stream(self._h)

    def timer_record(self, slot):
        """CUDA event on the context stream (mb_timer_record); slots 0..7."""
        check(self._lib.mb_timer_record(self._h, slot))

    def timer_ms(self, slot_begin, slot_end):
        out = C.c_double(0.0)
        check(self._lib.mb_timer_elapsed_ms(self._h, slot_begin, slot_end, C.byref(out)))
        return out.value

    def stat(self, key):
        out = C.c_double(0.0)
        check(self._lib.mb_get_stat(self._h, key.encode(), C.byref(out)))
        return out.value

    def synchronize(self):
        check(self._lib.mb_synchronize(self._h))

    def launch_count(self):
        return int(self._lib.mb_launch_count(self._h))
