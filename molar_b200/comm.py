"""Frame sharding across GPUs: one process (rank) per GPU, frames in contiguous blocks, nothing exchanged until the
per-frame scalars are gathered — through the LIBRARY's NCCL entry points (mb_comm_init / mb_gather_scalars /
mb_comm_max, include/molar_b200.h), not through torch.distributed.  The per-frame loop being parallelised is
AnalysisTask::run (molar/src/analysis_task.rs:113-280).

The only host-side job is handing rank 0's 128-byte NCCL unique id to the other ranks of the node; with a launcher
that exports RANK / WORLD_SIZE / MASTER_PORT (torchrun, mpirun wrappers) that is one small file under /tmp.
"""
import ctypes as C
import os
import time

import numpy as np

from . import _capi

ID_BYTES = 128


def frame_block(rank, frames_per_rank):
    """Global frame indices [first, last) owned by `rank` (weak scaling: every rank owns frames_per_rank frames)."""
    return rank * frames_per_rank, (rank + 1) * frames_per_rank


def split_frames(n_frames, world):
    """Strong-scaling partition of n_frames into `world` contiguous blocks (ceil(F/G) per GPU)."""
    per = -(-n_frames // world)
    return [(min(r * per, n_frames), min((r + 1) * per, n_frames)) for r in range(world)]


def env_rank():
    """(rank, world, local_rank) from the launcher's environment (defaults: single process)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def _id_path():
    # all ranks of one launch share the launcher as parent; the port separates concurrent launches of one parent
    return os.path.join(os.environ.get("MOLAR_B200_RDV_DIR", "/tmp"),
                        "molar_b200_nccl_%d_%s" % (os.getppid(), os.environ.get("MASTER_PORT", "0")))


def exchange_unique_id(rank, world, timeout_s=120.0):
    """Rank 0 creates the NCCL unique id (mb_comm_unique_id) and publishes it; the others wait for it."""
    L = _capi.load()
    path = _id_path()
    if rank == 0:
        buf = (C.c_ubyte * ID_BYTES)()
        _capi.check(L.mb_comm_unique_id(buf))
        tmp = path + ".tmp%d" % os.getpid()
        with open(tmp, "wb") as f:
            f.write(bytes(buf))
        os.replace(tmp, path)  # atomic: readers see nothing or all 128 bytes
        return bytes(buf)
    t0 = time.time()
    while True:
        try:
            with open(path, "rb") as f:
                b = f.read()
            if len(b) == ID_BYTES:
                return b
        except FileNotFoundError:
            pass
        if time.time() - t0 > timeout_s:
            raise TimeoutError("no NCCL unique id at %s after %.0f s" % (path, timeout_s))
        time.sleep(0.02)


class Comm:
    """The communicator of one rank, bound to one library context (a System or Trajectory handle)."""

    def __init__(self, owner, rank=None, world=None, unique_id=None):
        self._lib = _capi.load()
        self._h = owner._h
        r, w, _ = env_rank()
        self.rank = r if rank is None else rank
        self.world = w if world is None else world
        if self.world > 1:
            uid = unique_id if unique_id is not None else exchange_unique_id(self.rank, self.world)
            buf = (C.c_ubyte * ID_BYTES).from_buffer_copy(uid)
            _capi.check(self._lib.mb_comm_init(self._h, self.rank, self.world, buf))
            self.barrier()
            if self.rank == 0 and unique_id is None:
                try:
                    os.unlink(_id_path())
                except OSError:
                    pass

    def info(self):
        r, w, v = C.c_int(0), C.c_int(0), C.c_int(0)
        _capi.check(self._lib.mb_comm_info(self._h, C.byref(r), C.byref(w), C.byref(v)))
        return r.value, w.value, v.value

    def gather(self, rows=None, n_rows=None, n_cols=None):
        """All-gather of per-frame scalar rows [F, C] float64 -> [world*F, C] in global frame order on every rank.
        rows=None gathers the rows the last batch_pipeline / batch_fit left on the device (no host round trip)."""
        if rows is not None:
            a = np.ascontiguousarray(rows, dtype=np.float64)
            if a.ndim == 1:
                a = a.reshape(-1, 1)
            n_rows, n_cols = a.shape
            src = a.ctypes.data
        else:
            src = None
        out = np.empty((self.world * n_rows, n_cols), np.float64)
        _capi.check(self._lib.mb_gather_scalars(self._h, src, n_rows, n_cols, out.ctypes.data))
        return out

    def max(self, values):
        a = np.ascontiguousarray(np.atleast_1d(values), dtype=np.float64).copy()
        _capi.check(self._lib.mb_comm_max(self._h, a.ctypes.data_as(_capi.f64p), a.size))
        return a

    def barrier(self):
        _capi.check(self._lib.mb_comm_barrier(self._h))

    def close(self):
        self._lib.mb_comm_destroy(self._h)


def pinned_empty(shape, dtype=np.float32):
    """numpy array in page-locked host memory (mb_host_alloc); keep the returned array alive while it is in use and
    release it with pinned_free(arr)."""
    L = _capi.load()
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = L.mb_host_alloc(max(n, 1))
    if not p:
        raise _capi.MolarB200Error(_capi.MB_ERR_CUDA, _capi.last_error())
    buf = (C.c_ubyte * n).from_address(p)
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
    _PINNED[arr.ctypes.data] = p
    return arr


_PINNED = {}


def pinned_free(arr):
    p = _PINNED.pop(arr.ctypes.data, None)
    if p:
        _capi.load().mb_host_free(p)
