"""Host logic of the trajectory ingest that needs no GPU: mb_traj_probe parses DCD headers / record markers and scans
XTC frame headers exactly like the readers of the reference (io/dcd_handler.rs:204-300; xtc frame headers as molly
walks them).  Compared with the oracle readers on the reference's fixtures and on synthetic files."""
import struct

import numpy as np
import pytest

from molar_b200 import probe_trajectory, MolarB200Error
from oracle import traj_oracle as T


def test_probe_xtc_reference_fixtures(golden_dir):
    g = np.load(f"{golden_dir}/protein_xtc_trr.npz")
    buf = g["xtc_bytes"].tobytes()
    assert probe_trajectory(buf, "xtc") == (len(T.xtc_frame_offsets(buf)), 4295) == (4, 4295)
    s = np.load(f"{golden_dir}/small_xtc.npz")
    assert probe_trajectory(s["benzene_xtc"].tobytes(), "xtc") == (5, 12)
    assert probe_trajectory(s["new_xtc"].tobytes(), "xtc") == (10, 4295)
    # a truncated last frame is not counted (the reference stops at EOF inside a frame)
    assert probe_trajectory(buf[:-100], "xtc") == (3, 4295)
    assert probe_trajectory(buf[: len(buf) // 4 + 50], "xtc")[0] == 1


@pytest.mark.parametrize("big_endian", [False, True])
@pytest.mark.parametrize("extra", [True, False])
@pytest.mark.parametrize("fixed", [None, [0, 3, 4]])
def test_probe_dcd_synthetic(big_endian, extra, fixed):
    rng = np.random.default_rng(0)
    frames = rng.random((5, 37, 3)).astype(np.float32)
    buf = T.write_dcd(frames, big_endian=big_endian, charmm_extra=extra, fixed=fixed)
    assert probe_trajectory(buf, "dcd") == (5, 37)
    assert len(T.read_dcd(buf)) == 5
    # cut inside the last frame: four complete frames remain
    assert probe_trajectory(buf[:-10], "dcd") == (4, 37)


def test_probe_rejects_garbage():
    for fmt in ("dcd", "xtc"):
        with pytest.raises(MolarB200Error):
            probe_trajectory(b"\x00" * 256, fmt)
    good = T.write_dcd(np.zeros((1, 4, 3), np.float32))
    bad = bytearray(good)
    bad[4:8] = b"XXXX"  # not CORD
    with pytest.raises(MolarB200Error):
        probe_trajectory(bytes(bad), "dcd")
    # xtc with a changing atom count
    box = np.eye(3, dtype=np.float32)
    a = T.write_xtc_frame(np.zeros((12, 3), np.float32) + np.arange(12)[:, None] * 0.01, box)
    b = T.write_xtc_frame(np.zeros((13, 3), np.float32) + np.arange(13)[:, None] * 0.01, box)
    with pytest.raises(MolarB200Error):
        probe_trajectory(a + b, "xtc")
    # wrong magic in the second frame
    c = bytearray(a + a)
    c[len(a):len(a) + 4] = struct.pack(">i", 2023)
    with pytest.raises(MolarB200Error):
        probe_trajectory(bytes(c), "xtc")
