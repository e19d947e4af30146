"""Frame sharding and the scalar gather used by bench.py for N > 1 (one process per GPU).

Frames are independent units: rank r owns the contiguous block of global frames
[r*F, (r+1)*F); there is no data-path collective, only one all-gather of the per-frame scalars
(NCCL on GPUs, gloo in the CPU tests).  torch is plumbing here, not the product.
"""
import numpy as np


def frame_block(rank, frames_per_rank):
    """Global frame indices owned by `rank` (weak scaling: every rank owns frames_per_rank)."""
    return rank * frames_per_rank, (rank + 1) * frames_per_rank


def gather_rows(local_rows, world, device=None, out=None, stage=None):
    """All-gather equally sized [F, C] float64 row blocks; returns [world*F, C] in global frame order."""
    import torch
    import torch.distributed as dist
    rows = np.ascontiguousarray(local_rows, dtype=np.float64)
    if rows.ndim == 1:
        rows = rows.reshape(-1, 1)
    t = torch.from_numpy(rows)
    if device is not None:
        if stage is not None:
            stage.copy_(t)
            t = stage.to(device, non_blocking=True)
        else:
            t = t.to(device)
    if world == 1:
        return t
    if out is None:
        out = torch.empty((world * rows.shape[0], rows.shape[1]), dtype=torch.float64, device=t.device)
    dist.all_gather_into_tensor(out, t)
    return out
