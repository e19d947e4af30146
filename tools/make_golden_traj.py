#!/usr/bin/env python3
"""Generate tests/golden/protein_xtc_trr.npz and tests/golden/small_xtc.npz from the reference's trajectory fixtures.

Runs ONLY in the build container (needs /root/reference).  The reference ships the same trajectory twice:
molar/tests/protein.xtc (compressed, 4001 frames, 4295 atoms) and molar/tests/protein.trr (uncompressed, every
10th step).  A few XTC frames are stored as raw bytes next to the matching TRR coordinates, which pins an XTC
decoder without any reference code.  The two small XTC fixtures (benzene.xtc: 12 atoms; new.xtc: 4295 atoms,
2 frames... up to 10) are stored whole.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import traj_oracle as T  # noqa: E402

REF = "/root/reference/molar/tests"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def main():
    xb = open(os.path.join(REF, "protein.xtc"), "rb").read()
    tb = open(os.path.join(REF, "protein.trr"), "rb").read()
    offs = T.xtc_frame_offsets(xb) + [len(xb)]
    steps, trr_xyz, chunks = [], [], []
    toff = 0
    for k in range(41):
        ft, toff = T.read_trr_frame(tb, toff)
        if k in (0, 1, 2, 40):
            s = ft["step"]
            steps.append(s)
            trr_xyz.append(ft["xyz"])
            chunks.append(xb[offs[s]:offs[s + 1]])
    np.savez_compressed(os.path.join(OUT, "protein_xtc_trr.npz"), steps=np.asarray(steps),
                        trr_xyz=np.stack(trr_xyz).astype(np.float32),
                        xtc_bytes=np.frombuffer(b"".join(chunks), np.uint8))
    small = {}
    for name in ("benzene.xtc", "new.xtc"):
        b = open(os.path.join(REF, name), "rb").read()
        o = T.xtc_frame_offsets(b)
        keep = o[min(len(o), 10) - 1 + 1] if len(o) > 10 else len(b)
        small[name.replace(".", "_")] = np.frombuffer(b[:keep], np.uint8)
    np.savez_compressed(os.path.join(OUT, "small_xtc.npz"), **small)
    for f in ("protein_xtc_trr.npz", "small_xtc.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
