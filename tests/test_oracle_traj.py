"""Pin the oracle's trajectory readers (oracle/traj_oracle.py).

XTC: the reference decodes through the un-vendored crate `molly`; its own tests only check self-consistency
(io.rs:1050-1080).  The external pin is the reference's pair of fixtures protein.xtc / protein.trr — the same
trajectory compressed and uncompressed: the restated decoder must reproduce the TRR coordinates bit for bit.
DCD: the reference ships no .dcd fixture (PARITY UNPINNED by the reference); the reader is checked against
files written by the restated writer, in both byte orders, with and without unit cell and fixed atoms.
"""
import numpy as np
import pytest

from oracle import traj_oracle as T


def test_xtc_decoder_reproduces_the_trr_twin_bit_for_bit(golden_dir):
    g = np.load(f"{golden_dir}/protein_xtc_trr.npz")
    buf = g["xtc_bytes"].tobytes()
    frames = T.read_xtc(buf)
    assert [f["step"] for f in frames] == list(g["steps"])
    for f, want in zip(frames, g["trr_xyz"]):
        assert f["xyz"].dtype == np.float32 and f["xyz"].shape == (4295, 3)
        assert np.array_equal(f["xyz"], want)
        assert f["precision"] == 1000.0
        # orthorhombic box, columns = box vectors
        assert np.count_nonzero(f["box"] - np.diag(np.diag(f["box"]))) == 0 and f["box"][0, 0] > 5.0


def test_xtc_small_fixtures(golden_dir):
    g = np.load(f"{golden_dir}/small_xtc.npz")
    fr = T.read_xtc(g["benzene_xtc"].tobytes())
    assert len(fr) == 5 and fr[0]["xyz"].shape == (12, 3)
    # first carbon of benzene.pdb: 16.590 17.698 16.677 A
    assert np.allclose(fr[0]["xyz"][0], [1.659, 1.770, 1.668], atol=2e-3)
    # every coordinate is an integer multiple of 1/precision
    q = fr[1]["xyz"].astype(np.float64) * 1000.0
    assert np.abs(q - np.round(q)).max() < 1e-3
    fr = T.read_xtc(g["new_xtc"].tobytes())
    assert len(fr) >= 2 and fr[0]["xyz"].shape == (4295, 3) and np.isfinite(fr[-1]["xyz"]).all()


@pytest.mark.parametrize("big_endian", [False, True])
@pytest.mark.parametrize("extra", [True, False])
def test_dcd_roundtrip(big_endian, extra):
    rng = np.random.default_rng(1)
    frames = (rng.random((4, 301, 3)) * 8.0 - 1.0).astype(np.float32)
    cells = [np.array([80.0, 0.0, 70.0, 0.0, 0.0, 60.0])] * 4
    buf = T.write_dcd(frames, boxes=cells, big_endian=big_endian, charmm_extra=extra)
    got = T.read_dcd(buf)
    assert len(got) == 4
    for f in range(4):
        ang = (frames[f] * np.float32(10.0)).astype(np.float32)
        assert np.array_equal(got[f]["xyz"], ang * np.float32(0.1))  # `x as Float * 0.1`
        assert (got[f]["cell"] is not None) == extra
        assert got[f]["time"] == pytest.approx(0.5 * f)


def test_dcd_fixed_atoms():
    rng = np.random.default_rng(2)
    frames = (rng.random((3, 50, 3)) * 4.0).astype(np.float32)
    fixed = [0, 7, 8, 49]
    got = T.read_dcd(T.write_dcd(frames, fixed=fixed))
    free = np.setdiff1d(np.arange(50), fixed)
    for f in (1, 2):
        assert np.array_equal(got[f]["xyz"][fixed], got[0]["xyz"][fixed])
        ang = (frames[f] * np.float32(10.0)).astype(np.float32)
        assert np.array_equal(got[f]["xyz"][free], (ang * np.float32(0.1))[free])


def _waterlike(n_mol, box_len, seed):
    rng = np.random.default_rng(seed)
    o = rng.random((n_mol, 3)) * box_len
    w = np.repeat(o, 3, axis=0) + rng.normal(0.0, 0.03, (3 * n_mol, 3))
    iso = rng.random((n_mol // 4, 3)) * box_len
    return np.concatenate([w[: n_mol], iso, w[n_mol:]]).astype(np.float32)


def _quantised(xyz, precision=1000.0):
    q = np.rint(xyz.astype(np.float64) * precision).astype(np.int64)
    return (q.astype(np.float32) * (np.float32(1.0) / np.float32(precision))).astype(np.float32)


def test_synthetic_xtc_writer_roundtrip():
    xyz = _waterlike(1500, 6.0, 5)
    box = np.diag([6.0, 6.0, 6.0]).astype(np.float32)
    buf = T.write_xtc_frame(xyz, box, step=7, time=3.5) + T.write_xtc_frame(xyz[::-1].copy(), box, step=8, time=4.0)
    fr = T.read_xtc(buf)
    assert [f["step"] for f in fr] == [7, 8] and fr[1]["time"] == 4.0
    assert np.array_equal(fr[0]["xyz"], _quantised(xyz)) and np.array_equal(fr[1]["xyz"], _quantised(xyz[::-1]))
    assert len(buf) < 0.5 * xyz.nbytes * 2  # runs of small integers do compress
    # one coordinate range wider than 2^24 integers: the three fields are stored separately
    far = xyz.copy()
    far[3] = [20000.0, -5.0, 1.0]
    fr = T.read_xtc(T.write_xtc_frame(far, box))
    assert np.array_equal(fr[0]["xyz"], _quantised(far))
