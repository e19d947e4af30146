#!/usr/bin/env python3
"""tests/golden/protein_inertia.npz: the one known answer the reference holds for the measure path.

molar/src/selection.rs:198-213 (`test_inertia`) loads tests/protein.pdb, takes all 4295 atoms and prints the inertia
axes; the three axes it is expected to print are kept in the test as a comment (:209-211).  The fixture holds what the
hot path needs to reproduce them: coordinates parsed with the reference's rules (pdb_handler.rs:166-218: fixed
columns, f32, Angstrom -> nm as x * 0.1 in f32), masses from the element column through the reference's table
(pdb_handler.rs:196-200, periodic_table.rs:36-40), the CRYST1 box, and the three axes of the comment.
Runs only where /root/reference exists; the output is committed."""
import os

import numpy as np

REF = "/root/reference/molar/tests/protein.pdb"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "protein_inertia.npz")
# periodic_table.rs:36-40 (f32 table), the elements protein.pdb contains
ELEMENT_MASS = {"H": 1.00794, "C": 12.0107, "N": 14.0067, "O": 15.9994, "S": 32.065}
# selection.rs:209-211
REF_AXES = [[0.7308828830718994, 0.3332606256008148, -0.5956068634986877],
            [0.6804488301277161, -0.28815552592277527, 0.6737624406814575],
            [-0.05291106179356575, 0.8977214694023132, 0.4373747706413269]]


def main():
    xyz, m, box = [], [], None
    for line in open(REF):
        if line.startswith("ATOM  ") or line.startswith("HETATM"):
            xyz.append([np.float32(line[30:38]) * np.float32(0.1), np.float32(line[38:46]) * np.float32(0.1),
                        np.float32(line[46:54]) * np.float32(0.1)])
            m.append(np.float32(ELEMENT_MASS[line[76:78].strip().upper()]))
        elif line.startswith("CRYST1"):
            a, b, c = (np.float32(line[6:15]), np.float32(line[15:24]), np.float32(line[24:33]))
            assert [float(line[33:40]), float(line[40:47]), float(line[47:54])] == [90.0, 90.0, 90.0]
            box = np.diag([a * np.float32(0.1), b * np.float32(0.1), c * np.float32(0.1)]).astype(np.float32)
    np.savez_compressed(OUT, xyz=np.asarray(xyz, np.float32), masses=np.asarray(m, np.float32), box=box,
                        ref_axes=np.asarray(REF_AXES, np.float64))
    print(OUT, len(xyz), "atoms")


if __name__ == "__main__":
    main()
