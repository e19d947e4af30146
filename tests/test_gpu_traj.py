"""GPU parity: DCD / XTC byte streams decoded on the device (mb_batch_load_traj) vs the oracle readers and the
reference's protein.xtc / protein.trr twin fixture.  Bit-exact: the decode is integer / byte work plus one f32
multiplication per coordinate."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle_py as orc
from oracle import traj_oracle as T

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mb():
    import molar_b200
    return molar_b200


def _load(mb, buf, fmt, first=0, count=None):
    traj = mb.load_trajectory(buf, fmt, first_frame=first, n_frames=count)
    return traj


def test_xtc_reference_fixture_vs_trr(mb, golden_dir):
    g = np.load(f"{golden_dir}/protein_xtc_trr.npz")
    buf = g["xtc_bytes"].tobytes()
    assert mb.probe_trajectory(buf, "xtc") == (4, 4295)
    traj = _load(mb, buf, "xtc")
    xyz = traj.frames()
    assert xyz.shape == (4, 4295, 3) and xyz.dtype == np.float32
    assert np.array_equal(xyz, g["trr_xyz"])
    want = T.read_xtc(buf)
    for f in range(4):
        assert np.array_equal(traj.boxes[f], want[f]["box"])
        assert traj.times[f] == np.float32(want[f]["time"])
    # partial range
    part = _load(mb, buf, "xtc", first=2, count=2).frames()
    assert np.array_equal(part, g["trr_xyz"][2:4])
    # the batch is directly usable: contacts of frame 3 == oracle search on the decoded coordinates
    n = int(traj.search(0.6, f0=3, f1=4, count_only=True)[0])
    ij, d, dims = orc.search_single(0.6, g["trr_xyz"][3], None, orc.Box(matrix=want[3]["box"]), 7, 4)
    assert n == len(orc.canonical_pairs(ij, d)[0])
    traj.close()


def test_xtc_small_fixtures_vs_oracle(mb, golden_dir):
    g = np.load(f"{golden_dir}/small_xtc.npz")
    for key in ("benzene_xtc", "new_xtc"):
        buf = g[key].tobytes()
        want = T.read_xtc(buf)
        traj = _load(mb, buf, "xtc")
        got = traj.frames()
        assert got.shape[0] == len(want)
        for f, w in enumerate(want):
            assert np.array_equal(got[f], w["xyz"]), (key, f)
        traj.close()


def test_xtc_synthetic_runs_and_wide_fields(mb):
    """Synthetic frames from the oracle's writer: long runs, repeated run lengths (flag bit clear), isolated
    atoms, and a coordinate range above 2^24 (separate bit fields)."""
    from tests.test_oracle_traj import _waterlike
    xyz = _waterlike(12_000, 5.0, 6)
    box = np.diag([5.0, 5.0, 5.0]).astype(np.float32)
    far = xyz.copy()
    far[17] = [17000.0, 2.0, -3.0]
    buf = T.write_xtc_frame(xyz, box, step=1, time=1.0) + T.write_xtc_frame(far, box, step=2, time=2.0) + \
        T.write_xtc_frame(xyz[::-1].copy(), box, step=3, time=3.0, smallidx=30, max_small=3)
    want = T.read_xtc(buf)
    traj = _load(mb, buf, "xtc")
    got = traj.frames()
    for f in range(3):
        assert np.array_equal(got[f], want[f]["xyz"]), f
    assert list(traj.times) == [1.0, 2.0, 3.0]
    traj.close()


def test_xtc_tiny_system_is_stored_raw(mb):
    # natoms <= 9: uncompressed big-endian floats (xdr3dfcoord)
    import struct
    xyz = np.arange(21, dtype=np.float32).reshape(7, 3) * np.float32(0.37)
    box = np.diag([3.0, 4.0, 5.0]).astype(np.float32)
    frame = struct.pack(">iiif", 1995, 7, 5, 2.5) + box.T.astype(">f4").tobytes() + struct.pack(">i", 7) + \
        xyz.astype(">f4").tobytes()
    traj = _load(mb, frame * 2, "xtc")
    assert np.array_equal(traj.frames(), np.stack([xyz, xyz]))
    assert traj.times[1] == 2.5 and np.array_equal(traj.boxes[0], box)
    traj.close()


@pytest.mark.parametrize("big_endian", [False, True])
@pytest.mark.parametrize("extra", [True, False])
def test_dcd_vs_oracle(mb, big_endian, extra):
    rng = np.random.default_rng(3)
    n, nf = 100_003, 5
    frames = (rng.random((nf, n, 3)) * 9.0 - 1.0).astype(np.float32)
    cells = [np.array([90.0, 0.0, 80.0, 0.0, 0.0, 70.0 + f]) for f in range(nf)]
    buf = T.write_dcd(frames, boxes=cells, big_endian=big_endian, charmm_extra=extra)
    want = T.read_dcd(buf)
    assert mb.probe_trajectory(buf, "dcd") == (nf, n)
    traj = _load(mb, buf, "dcd")
    got = traj.frames()
    for f in range(nf):
        assert np.array_equal(got[f], want[f]["xyz"])
        assert traj.times[f] == np.float32(want[f]["time"])
        if extra:
            assert np.allclose(np.diag(traj.boxes[f]), [9.0, 8.0, 7.0 + 0.1 * f], rtol=1e-6)
        else:
            assert not traj.boxes[f].any()
    sub = _load(mb, buf, "dcd", first=1, count=3).frames()
    assert np.array_equal(sub, got[1:4])
    traj.close()


def test_dcd_triclinic_cell_and_fixed_atoms(mb):
    rng = np.random.default_rng(4)
    n, nf = 5000, 4
    frames = (rng.random((nf, n, 3)) * 4.0).astype(np.float32)
    fixed = np.sort(rng.choice(n, 700, replace=False))
    # cosines of (gamma, beta, alpha) = (80, 75, 70) degrees
    cell = np.array([50.0, np.cos(np.radians(80.0)), 45.0, np.cos(np.radians(75.0)), np.cos(np.radians(70.0)), 40.0])
    buf = T.write_dcd(frames, boxes=[cell] * nf, fixed=list(fixed))
    want = T.read_dcd(buf)
    for first in (0, 1):
        traj = _load(mb, buf, "dcd", first=first, count=nf - first)
        got = traj.frames()
        for f in range(nf - first):
            assert np.array_equal(got[f], want[first + f]["xyz"])
        b = orc.Box(vectors_angles=(5.0, 4.5, 4.0, np.degrees(np.arccos(cell[4])), np.degrees(np.arccos(cell[3])),
                                    np.degrees(np.arccos(cell[1]))))
        assert np.allclose(traj.boxes[0], b.matrix, rtol=1e-6, atol=1e-6)
        traj.close()


def test_errors(mb):
    with pytest.raises(mb.MolarB200Error):
        mb.probe_trajectory(b"\x00" * 200, "dcd")
    with pytest.raises(mb.MolarB200Error):
        mb.probe_trajectory(b"\x00" * 200, "xtc")
    buf = T.write_dcd(np.zeros((2, 10, 3), np.float32))
    with pytest.raises(mb.MolarB200Error):
        mb.load_trajectory(buf, "dcd", first_frame=1, n_frames=5)
    # a record marker that does not match: the reference fails with BadRecord
    bad = bytearray(buf)
    bad[-4:] = b"\x01\x02\x03\x04"
    with pytest.raises(mb.MolarB200Error):
        mb.load_trajectory(bytes(bad), "dcd")


def test_xtc_corrupt_streams_are_rejected_not_decoded_out_of_bounds(mb, golden_dir):
    """Trajectory files are untrusted input.  A truncated compressed block, flipped bits in the block, a run code that
    overruns the atom count or an invalid small-integer index must end in MolarB200Error (or decode to SOMETHING of the
    right shape) — never in a crash or an out-of-frame write; the context stays usable afterwards."""
    g = np.load(f"{golden_dir}/protein_xtc_trr.npz")
    good = bytearray(g["xtc_bytes"].tobytes())
    nf, na = mb.probe_trajectory(bytes(good), "xtc")
    frame_len = len(good) // nf
    rng = np.random.default_rng(5)
    outcomes = {"error": 0, "decoded": 0}
    cases = []
    for trial in range(24):
        b = bytearray(good)
        kind = trial % 4
        if kind == 0:      # flip bits inside the compressed block of the LAST frame (no slack behind the batch)
            for _ in range(1 + trial):
                pos = (nf - 1) * frame_len + 92 + int(rng.integers(0, frame_len - 100))
                b[pos] ^= 1 << int(rng.integers(0, 8))
        elif kind == 1:    # small-integer index out of range
            b[(nf - 1) * frame_len + 84: (nf - 1) * frame_len + 88] = int(rng.integers(0, 200)).to_bytes(4, "big")
        elif kind == 2:    # all-ones block: every flag set, maximal run codes
            s = (nf - 1) * frame_len + 92
            b[s: s + 4000] = b"\xff" * 4000
        else:              # declared block length larger than what is there (truncated file)
            b = b[: (nf - 1) * frame_len + 92 + int(rng.integers(10, frame_len - 200))]
        cases.append(bytes(b))
    for b in cases:
        try:
            n_ok = mb.probe_trajectory(b, "xtc")[0]
            traj = mb.load_trajectory(b, "xtc")
            xyz = traj.frames()
            assert xyz.shape == (n_ok, na, 3)
            traj.close()
            outcomes["decoded"] += 1
        except mb.MolarB200Error:
            outcomes["error"] += 1
    assert outcomes["error"] > 0
    # the library is still healthy: the intact file decodes bit-exactly afterwards
    traj = mb.load_trajectory(bytes(good), "xtc")
    assert np.array_equal(traj.frames(), g["trr_xyz"])
    traj.close()
