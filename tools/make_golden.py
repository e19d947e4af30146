#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the reference's own fixtures and golden vectors.

Runs ONLY in the build container (needs /root/reference); the outputs are committed so that
nothing on the GPU box reads /root/reference.  Nothing is copied verbatim from the reference's
sources: the PDB files are parsed with the reference's rules (molar/src/io/pdb_handler.rs:166-218:
fixed columns, f32 parse, Angstrom -> nm via `x * 0.1` in f32) and reduced to the arrays the
hot path needs; the golden id vectors are the expected answers of the reference's own tests
(molar/tests/generated_vmd_tests.in:27,35 and generated_pteros_tests.in:21,27, included at
molar/src/selection/selection_expr.rs:302-310).
"""
import os
import re
import sys

import numpy as np

REF = "/root/reference/molar/tests"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def parse_pdb(path):
    xyz, resid = [], []
    box = None
    with open(path) as f:
        for line in f:
            if line.startswith("ATOM  ") or line.startswith("HETATM"):
                x = np.float32(line[30:38]) * np.float32(0.1)
                y = np.float32(line[38:46]) * np.float32(0.1)
                z = np.float32(line[46:54]) * np.float32(0.1)
                xyz.append((x, y, z))
                try:
                    resid.append(int(line[22:26]))
                except ValueError:
                    resid.append(0)
            elif line.startswith("CRYST1"):
                a, b, c = (np.float32(line[6:15]), np.float32(line[15:24]), np.float32(line[24:33]))
                al, be, ga = (np.float32(line[33:40]), np.float32(line[40:47]), np.float32(line[47:54]))
                box = (a * np.float32(0.1), b * np.float32(0.1), c * np.float32(0.1), al, be, ga)
            elif line.startswith("ENDMDL"):
                break
    return np.asarray(xyz, dtype=np.float32), np.asarray(resid, dtype=np.int32), box


def parse_golden(path):
    """-> {selection string: np.array of ids}"""
    txt = open(path).read()
    out = {}
    # VMD flavour: "a b c".split(" ") ... get_selection_index("sel")
    for m in re.finditer(r'"([0-9 ]+)"\s*\.split.*?get_selection_index2?\("([^"]+)"\)', txt, re.S):
        out[m.group(2)] = np.asarray([int(t) for t in m.group(1).split()], dtype=np.int64)
    # pteros flavour: vec![a, b, c]; ... get_selection_index2(\"sel\")
    for m in re.finditer(r'vec!\[([0-9, ]*)\];\s*assert_eq!\(get_selection_index2?\("([^"]+)"\)', txt, re.S):
        out[m.group(2)] = np.asarray([int(t) for t in m.group(1).replace(",", " ").split()], dtype=np.int64)
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    xyz, resid, box = parse_pdb(os.path.join(REF, "albumin.pdb"))
    assert xyz.shape == (76400, 3), xyz.shape
    assert box[3] == 90 and box[4] == 90 and box[5] == 90
    box9 = np.zeros(9, dtype=np.float32)  # column-major, orthorhombic
    box9[0], box9[4], box9[8] = box[0], box[1], box[2]
    vmd = parse_golden(os.path.join(REF, "generated_vmd_tests.in"))
    pt = parse_golden(os.path.join(REF, "generated_pteros_tests.in"))
    gold = {}
    gold.update(vmd)
    gold.update(pt)
    # inner selections as the reference evaluates them (resid keyword = exact match on resid)
    inner = {r: np.nonzero(resid == r)[0].astype(np.int64) for r in (10, 20, 555)}
    assert np.array_equal(inner[10], gold["resid 10"]), "resid parse disagrees with golden vector"
    assert np.array_equal(inner[555], gold["resid 555"])
    cases = {
        "within_0.5_resid10": ("within 0.5 of resid 10", 0.5, 10, 0),
        "within_0.3_resid20": ("within 0.3 of resid 20", 0.3, 20, 0),
        "within_0.5_resid555": ("within 0.5 of resid 555", 0.5, 555, 0),
        "within_0.5_pbc_resid555": ("within 0.5 pbc yyy of resid 555", 0.5, 555, 7),
    }
    save = dict(xyz=xyz, box9=box9)
    for key, (sel, cutoff, r, pbc) in cases.items():
        ans = gold[sel]
        save[key + "_answer"] = ans
        save[key + "_inner"] = inner[r]
        save[key + "_params"] = np.asarray([cutoff, pbc], dtype=np.float64)
        print(f"{sel!r}: {len(ans)} ids, inner {len(inner[r])}")
    np.savez_compressed(os.path.join(OUT, "albumin_within.npz"), **save)

    xyz2, _, box2 = parse_pdb(os.path.join(REF, "2lao.pdb"))
    assert xyz2.shape[0] == 1911, xyz2.shape
    b9 = np.zeros(9, dtype=np.float32)
    b9[0], b9[4], b9[8] = box2[0], box2[1], box2[2]
    np.savez_compressed(os.path.join(OUT, "2lao.npz"), xyz=xyz2, box9=b9)

    xyz3, _, box3 = parse_pdb(os.path.join(REF, "triclinic.pdb"))
    # keep a 6000-atom subsample (every 9th atom) + the CRYST1 record: enough to exercise the
    # reference's real triclinic box (60/60/90) without committing 56k atoms
    np.savez_compressed(os.path.join(OUT, "triclinic_sub.npz"), xyz=xyz3[::9].copy(),
                        cryst1=np.asarray(box3, dtype=np.float32))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    sys.exit(main())
