"""N>1 host logic on CPU: world_size 2 over gloo.  Each rank computes the per-frame scalars of
its own frame block (with the oracle here, the GPU library on the box), the rows are all-gathered
and must equal the serial result in global frame order."""
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


def _worker(rank, world, port, q):
    import torch.distributed as dist
    import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    f0, f1 = shard.frame_block(rank, F)
    g = shard.gather_rows(_rows(f0, f1), world)
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
    import shard
    blocks = [shard.frame_block(r, 5) for r in range(4)]
    assert blocks == [(0, 5), (5, 10), (10, 15), (15, 20)]
