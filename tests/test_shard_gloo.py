"""N>1 host logic on CPU: world_size 2 over gloo.  Each rank computes the per-frame scalars of
its own frame block (molar_b200.comm.frame_block; with the oracle here, the GPU library on the box), the rows are
all-gathered and must equal the serial result in global frame order.  On the GPU box the gather is the library's
mb_gather_scalars (NCCL); here, without GPUs, the same row layout goes through a gloo all-gather that lives in this
test."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

F, N = 3, 400
BOX = np.diag([3.0, 3.5, 4.0]).astype(np.float32)


def _rows(first, last):
    from oracle import oracle_py as orc
    m = orc.synth_masses(20260, N)
    out = []
    for f in range(first, last):
        xyz = orc.synth_frame(20260, f, N, BOX)
        rc, com = orc.center_of_mass(xyz, m)
        rc, rg = orc.gyration(xyz, m)
        ij, d, dims = orc.search_single(1.2, xyz, None, orc.Box(matrix=BOX), 7, 1)
        out.append([com[0], com[1], com[2], rg, float(len(orc.canonical_pairs(ij)))])
    return np.asarray(out)


def _gather_rows(local_rows, world):
    """All-gather equally sized [F, C] float64 row blocks over gloo -> [world*F, C] in global frame order (the layout
    mb_gather_scalars produces)."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(local_rows, dtype=np.float64))
    out = torch.empty((world * t.shape[0], t.shape[1]), dtype=torch.float64)
    dist.all_gather_into_tensor(out, t)
    return out


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from molar_b200 import comm
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    f0, f1 = comm.frame_block(rank, F)
    g = _gather_rows(_rows(f0, f1), world)
    dist.barrier()
    if rank == 0:
        q.put(g.numpy().copy())
    dist.destroy_process_group()


def test_two_rank_gather_matches_serial():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got.shape == (2 * F, 5)
    assert np.array_equal(got, _rows(0, 2 * F))


def test_frame_blocks_partition():
    from molar_b200 import comm
    blocks = [comm.frame_block(r, 5) for r in range(4)]
    assert blocks == [(0, 5), (5, 10), (10, 15), (15, 20)]
    # strong-scaling split: contiguous, covering, ceil(F / G) per rank, empty blocks at the end when F < G
    for nf, world in [(10, 4), (3, 8), (16, 8), (1, 1), (0, 2)]:
        parts = comm.split_frames(nf, world)
        assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == nf
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        assert max(b - a for a, b in parts) == -(-nf // world) if nf else True
