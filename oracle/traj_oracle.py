"""CPU restatement of the reference's trajectory readers.  TEST INFRASTRUCTURE ONLY (see oracle/README.md):
only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.

DCD   molar/src/io/dcd_handler.rs:204-300 (header), :389-464 (read_state), :172-200 (unit cell), :132-139,
      :202-226, :466-520 (writer, used here to make synthetic fixtures).
XTC   molar/src/io/xtc_handler.rs:64-110 hands the decompression to the third-party crate `molly`
      (molar/Cargo.toml:36: git dependency, version >= 0.6.1, NOT vendored under /root/reference).  What is
      restated is the published xdrfile algorithm (xdr3dfcoord / receivebits / receiveints / sizeofints) that
      molly implements.  PINNED: tests/golden/protein_xtc_trr.npz holds frames of the reference's own fixture
      tests/protein.xtc next to the same frames of its uncompressed twin tests/protein.trr; this decoder
      reproduces the TRR coordinates bit for bit (tests/test_oracle_traj.py).
TRR   plain XDR floats; read here only to pin the XTC decoder (io/trr_handler.rs).
Pure-Python bit loops: use on small cases (thousands of atoms, a few frames).
"""
import struct

import numpy as np

MAGICINTS = [0, 0, 0, 0, 0, 0, 0, 0, 0, 8, 10, 12, 16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406,
             512, 645, 812, 1024, 1290, 1625, 2048, 2580, 3250, 4096, 5060, 6501, 8192, 10321, 13003, 16384, 20642,
             26007, 32768, 41285, 52015, 65536, 82570, 104031, 131072, 165140, 208063, 262144, 330280, 416127, 524287,
             660561, 832255, 1048576, 1321122, 1664510, 2097152, 2642245, 3329021, 4194304, 5284491, 6658042, 8388607,
             10568983, 13316085, 16777216]
FIRSTIDX = 9


# ---------------------------------------------------------------------------------------------- XTC
def _sizeofint(size):
    num, nbits = 1, 0
    while size >= num and nbits < 32:
        nbits += 1
        num <<= 1
    return nbits


def _sizeofints(sizes):
    nbytes = 1
    b = [1] + [0] * 31
    for s in sizes:
        tmp, bc = 0, 0
        while bc < nbytes:
            tmp = b[bc] * s + tmp
            b[bc] = tmp & 0xFF
            tmp >>= 8
            bc += 1
        while tmp != 0:
            b[bc] = tmp & 0xFF
            tmp >>= 8
            bc += 1
        nbytes = bc
    num, nbits = 1, 0
    nbytes -= 1
    while b[nbytes] >= num:
        nbits += 1
        num *= 2
    return nbits + nbytes * 8


class _Bits:
    """receivebits: MSB-first bit reader"""

    def __init__(self, data):
        self.v = int.from_bytes(data, "big")
        self.total = len(data) * 8
        self.pos = 0

    def get(self, n):
        if n == 0:
            return 0
        if self.pos + n > self.total:
            raise ValueError("xtc: compressed block exhausted")
        r = (self.v >> (self.total - self.pos - n)) & ((1 << n) - 1)
        self.pos += n
        return r


def _receiveints(bs, nbits, sizes):
    """three integers packed as one mixed-radix number stored as bytes, first byte least significant"""
    by = []
    while nbits > 8:
        by.append(bs.get(8))
        nbits -= 8
    if nbits > 0:
        by.append(bs.get(nbits))
    while len(by) < 4:
        by.append(0)
    nums = [0, 0, 0]
    for i in (2, 1):
        num = 0
        for j in range(len(by) - 1, -1, -1):
            num = (num << 8) | by[j]
            p = num // sizes[i]
            by[j] = p
            num -= p * sizes[i]
        nums[i] = num
    nums[0] = by[0] | (by[1] << 8) | (by[2] << 16) | (by[3] << 24)
    return nums


def read_xtc_frame(buf, off=0):
    """-> (dict(step, time, box[3,3] columns = box vectors, xyz[n,3] f32, precision), next offset)"""
    magic, natoms, step = struct.unpack_from(">iii", buf, off)
    if magic != 1995:
        raise ValueError("xtc: bad magic")
    time, = struct.unpack_from(">f", buf, off + 12)
    stored = np.array(struct.unpack_from(">9f", buf, off + 16), np.float32)
    box = stored.reshape(3, 3).T.copy()  # stored rows are the box vectors; Matrix3::from_iterator is column-major
    lsize, = struct.unpack_from(">i", buf, off + 52)
    off += 56
    if lsize <= 9:
        xyz = np.array(struct.unpack_from(">%df" % (3 * lsize), buf, off), np.float32).reshape(-1, 3)
        return dict(step=step, time=time, box=box, xyz=xyz, precision=0.0), off + 12 * lsize
    prec, = struct.unpack_from(">f", buf, off)
    minint = struct.unpack_from(">3i", buf, off + 4)
    maxint = struct.unpack_from(">3i", buf, off + 16)
    smallidx, nbytes = struct.unpack_from(">ii", buf, off + 28)
    off += 36
    data = bytes(buf[off:off + nbytes])
    off += (nbytes + 3) // 4 * 4
    sizeint = [maxint[i] - minint[i] + 1 for i in range(3)]
    if any(s > 0xFFFFFF for s in sizeint):
        bitsizeint = [_sizeofint(s) for s in sizeint]
        bitsize = 0
    else:
        bitsize = _sizeofints(sizeint)
    smaller = MAGICINTS[max(FIRSTIDX, smallidx - 1)] // 2
    smallnum = MAGICINTS[smallidx] // 2
    sizesmall = [MAGICINTS[smallidx]] * 3
    bs = _Bits(data)
    out = []
    i = run = 0
    while i < lsize:
        this = [bs.get(bitsizeint[k]) for k in range(3)] if bitsize == 0 else _receiveints(bs, bitsize, sizeint)
        i += 1
        this = [this[k] + minint[k] for k in range(3)]
        prev = list(this)
        flag = bs.get(1)
        is_smaller = 0
        if flag == 1:
            run = bs.get(5)
            is_smaller = run % 3
            run -= is_smaller
            is_smaller -= 1
        if run > 0:
            for k in range(0, run, 3):
                t = _receiveints(bs, smallidx, sizesmall)
                i += 1
                t = [t[q] + prev[q] - smallnum for q in range(3)]
                if k == 0:
                    t, prev = prev, t  # first small atom goes in front of the full one
                    out.append(prev)
                else:
                    prev = list(t)
                out.append(t)
        else:
            out.append(this)
        smallidx += is_smaller
        if is_smaller < 0:
            smallnum = smaller
            smaller = MAGICINTS[smallidx - 1] // 2 if smallidx > FIRSTIDX else 0
        elif is_smaller > 0:
            smaller = smallnum
            smallnum = MAGICINTS[smallidx] // 2
        sizesmall = [MAGICINTS[smallidx]] * 3
    ints = np.array(out[:lsize], np.int32)
    inv = np.float32(1.0) / np.float32(prec)
    xyz = (ints.astype(np.float32) * inv).astype(np.float32)
    return dict(step=step, time=time, box=box, xyz=xyz, precision=prec), off


def xtc_frame_offsets(buf):
    offs, off = [], 0
    while off + 56 <= len(buf):
        natoms, = struct.unpack_from(">i", buf, off + 4)
        if natoms <= 9:
            size = 56 + 12 * natoms
        else:
            if off + 92 > len(buf):
                break
            nbytes, = struct.unpack_from(">i", buf, off + 88)
            size = 92 + (nbytes + 3) // 4 * 4
        if off + size > len(buf):
            break
        offs.append(off)
        off += size
    return offs


def read_xtc(buf, frames=None):
    offs = xtc_frame_offsets(buf)
    return [read_xtc_frame(buf, offs[f])[0] for f in (range(len(offs)) if frames is None else frames)]


# ---------------------------------------------------------------------------------------------- TRR
def read_trr_frame(buf, off=0):
    magic, _ = struct.unpack_from(">ii", buf, off)
    if magic != 1993:
        raise ValueError("trr: bad magic")
    slen, = struct.unpack_from(">i", buf, off + 8)
    off += 12 + (slen + 3) // 4 * 4
    (ir, e, box_size, vir, pres, top, sym, x_size, v_size, f_size, natoms, step, nre) = struct.unpack_from(">13i", buf, off)
    off += 52
    dbl = (box_size == 72) if box_size else (x_size == natoms * 24)
    t, lam = struct.unpack_from(">dd" if dbl else ">ff", buf, off)
    off += 16 if dbl else 8
    off += box_size + vir + pres
    x = np.frombuffer(buf, dtype=">f8" if dbl else ">f4", count=natoms * 3, offset=off).reshape(-1, 3)
    off += x_size + v_size + f_size
    return dict(step=step, time=t, xyz=x.astype(np.float32)), off


# ---------------------------------------------------------------------------------------------- DCD
def write_dcd(frames_nm, boxes=None, big_endian=False, charmm_extra=True, fixed=None):
    """frames_nm: [F][N][3] f32 (nm).  Follows the reference writer (CHARMM CORD header, one unit-cell record per
    frame) and adds what the reader must also cope with: big-endian files, no unit-cell block, fixed atoms
    (`fixed`: sorted 0-based indices; frames after the first then hold only the free atoms)."""
    e = ">" if big_endian else "<"
    frames_nm = np.asarray(frames_nm, np.float32)
    nf, n, _ = frames_nm.shape

    def rec(payload):
        return struct.pack(e + "i", len(payload)) + payload + struct.pack(e + "i", len(payload))

    hdr = bytearray(84)
    hdr[0:4] = b"CORD"
    hdr[4:8] = struct.pack(e + "i", nf)
    hdr[12:16] = struct.pack(e + "i", 1)
    nfixed = 0 if fixed is None else len(fixed)
    hdr[32:36] = struct.pack(e + "i", nfixed)
    hdr[36:40] = struct.pack(e + "f", 0.5)
    hdr[40:44] = struct.pack(e + "i", 1 if charmm_extra else 0)
    hdr[76:80] = struct.pack(e + "i", 24)
    title = bytearray(84)
    title[0:4] = struct.pack(e + "i", 1)
    out = [rec(bytes(hdr)), rec(bytes(title)), rec(struct.pack(e + "i", n))]
    free = None
    if nfixed:
        free = np.setdiff1d(np.arange(n), np.asarray(fixed))
        out.append(rec(struct.pack(e + "%di" % len(free), *[int(v) + 1 for v in free])))
    for f in range(nf):
        if charmm_extra:
            cell = np.zeros(6, np.float64)
            if boxes is not None and boxes[f] is not None:
                cell[:] = boxes[f]  # [A, cos(gamma), B, cos(beta), cos(alpha), C] in Angstrom / cosines
            out.append(rec(cell.astype(e + "f8").tobytes()))
        ang = (frames_nm[f] * np.float32(10.0)).astype(np.float32)  # (p.x * 10.0) as f32  (dcd_handler.rs:486-490)
        sel = ang if (free is None or f == 0) else ang[free]
        for d in range(3):
            out.append(rec(np.ascontiguousarray(sel[:, d]).astype(e + "f4").tobytes()))
    return b"".join(out)


def read_dcd(buf):
    """-> list of dict(xyz[n,3] f32 nm, cell (6 f64 or None), time)"""
    le, be = struct.unpack_from("<I", buf, 0)[0], struct.unpack_from(">I", buf, 0)[0]
    if le == 84:
        e = "<"
    elif be == 84:
        e = ">"
    else:
        raise ValueError("dcd: bad magic")
    h = buf[4:88]
    if h[0:4] != b"CORD":
        raise ValueError("dcd: bad magic")
    istart, nsavc = struct.unpack_from(e + "ii", h, 8)
    nfixed = max(struct.unpack_from(e + "i", h, 32)[0], 0)
    is_charmm = struct.unpack_from(e + "i", h, 76)[0] != 0
    extra = struct.unpack_from(e + "i", h, 40)[0] != 0
    fourd = struct.unpack_from(e + "i", h, 44)[0] != 0
    delta = struct.unpack_from(e + "f", h, 36)[0] if is_charmm else np.float32(struct.unpack_from(e + "d", h, 36)[0])
    off = 92

    def record(o):
        n, = struct.unpack_from(e + "I", buf, o)
        end, = struct.unpack_from(e + "I", buf, o + 4 + n)
        if end != n:
            raise ValueError("dcd: bad record")
        return o + 4, n, o + 8 + n

    _, _, off = record(off)
    p, n, off = record(off)
    natoms = max(struct.unpack_from(e + "i", buf, p)[0], 0)
    free = None
    if nfixed:
        p, n, off = record(off)
        free = np.maximum(np.frombuffer(buf, e + "u4", natoms - nfixed, p).astype(np.int64) - 1, 0)
    frames, fixed_coords, cur = [], None, 0
    while off < len(buf):
        try:
            nread = natoms - nfixed if (nfixed and cur > 0) else natoms
            cell = None
            if extra:
                p, n, off = record(off)
                cell = np.frombuffer(buf, e + "f8", 6, p).astype(np.float64)
            blocks = []
            for _ in range(3):
                p, n, off = record(off)
                if n != nread * 4:
                    raise ValueError("dcd: unexpected record length")
                blocks.append(np.frombuffer(buf, e + "f4", nread, p).astype(np.float32))
            if fourd:
                _, _, off = record(off)
        except struct.error:
            break
        scaled = [b * np.float32(0.1) for b in blocks]  # x as Float * 0.1
        if nfixed == 0 or cur == 0:
            xyz = np.stack(scaled, axis=1).astype(np.float32)
        else:
            xyz = fixed_coords.copy()
            xyz[free] = np.stack(scaled, axis=1)
        if cur == 0 and nfixed:
            fixed_coords = xyz.copy()
        time = np.float32(istart + cur * nsavc) * np.float32(delta)
        frames.append(dict(xyz=xyz, cell=cell, time=float(time)))
        cur += 1
    return frames


# ---------------------------------------------------------------------------------------------- XTC writer
class _BitWriter:
    def __init__(self):
        self.out = bytearray()
        self.acc = 0
        self.n = 0

    def put(self, nbits, value):
        if nbits == 0:
            return
        self.acc = (self.acc << nbits) | (value & ((1 << nbits) - 1))
        self.n += nbits
        while self.n >= 8:
            self.n -= 8
            self.out.append((self.acc >> self.n) & 0xFF)
        self.acc &= (1 << self.n) - 1

    def finish(self):
        if self.n:
            self.out.append((self.acc << (8 - self.n)) & 0xFF)
            self.acc = self.n = 0
        return bytes(self.out)


def _sendints(bw, nbits, sizes, nums):
    """inverse of _receiveints: mixed-radix number, bytes least significant first"""
    v = (nums[0] * sizes[1] + nums[1]) * sizes[2] + nums[2]
    full, rem = divmod(nbits, 8)
    for j in range(full):
        bw.put(8, (v >> (8 * j)) & 0xFF)
    if rem:
        bw.put(rem, (v >> (8 * full)) & ((1 << rem) - 1))


def write_xtc_frame(xyz_nm, box, step=0, time=0.0, precision=1000.0, smallidx=24, max_small=8):
    """A VALID xtc frame for synthetic data (test / bench input).  Not the xdrfile compressor: the small-integer
    width stays fixed (`smallidx`), runs are formed greedily wherever consecutive atoms are close enough, and
    the flag bit is left clear when the run length repeats — every branch of the decoder is exercised except
    the adaptive change of the small-integer width."""
    xyz = np.asarray(xyz_nm, np.float32)
    n = len(xyz)
    box = np.asarray(box, np.float32).reshape(3, 3)
    head = struct.pack(">iiif", 1995, n, step, time) + box.T.astype(">f4").tobytes() + struct.pack(">i", n)
    if n <= 9:
        return head + xyz.astype(">f4").tobytes()
    ints = np.rint(xyz.astype(np.float64) * precision).astype(np.int64)
    minint, maxint = ints.min(0), ints.max(0)
    sizeint = [int(maxint[k] - minint[k] + 1) for k in range(3)]
    if any(s > 0xFFFFFF for s in sizeint):
        bitsizeint, bitsize = [_sizeofint(s) for s in sizeint], 0
    else:
        bitsize = _sizeofints(sizeint)
    ss = MAGICINTS[smallidx]
    smallnum = ss // 2
    rel = (ints - minint).tolist()
    a = ints.tolist()

    def close(p, q):
        return all(0 <= a[p][k] - a[q][k] + smallnum < ss for k in range(3))

    bw = _BitWriter()
    i, prev_run = 0, 0
    while i < n:
        # atoms i.. : [S1, F, S2, S3, ...] on output; F is stored first
        nsmall = 0
        if i + 1 < n and close(i, i + 1):
            nsmall = 1
            last = i
            j = i + 2
            while j < n and nsmall < max_small and close(j, last):
                last = j
                nsmall += 1
                j += 1
        f = i + 1 if nsmall else i
        if bitsize == 0:
            for k in range(3):
                bw.put(bitsizeint[k], rel[f][k])
        else:
            _sendints(bw, bitsize, sizeint, rel[f])
        run = 3 * nsmall
        if run == prev_run:
            bw.put(1, 0)
        else:
            bw.put(1, 1)
            bw.put(5, run + 1)  # run % 3 == 1: the width of the small integers stays as it is
        prev_run = run
        if nsmall:
            _sendints(bw, smallidx, [ss] * 3, [a[i][k] - a[f][k] + smallnum for k in range(3)])
            last = i
            for j in range(i + 2, i + 1 + nsmall):
                _sendints(bw, smallidx, [ss] * 3, [a[j][k] - a[last][k] + smallnum for k in range(3)])
                last = j
        i += 1 + nsmall
    data = bw.finish()
    pad = (-len(data)) % 4
    return (head + struct.pack(">f", precision) + struct.pack(">3i", *[int(v) for v in minint]) +
            struct.pack(">3i", *[int(v) for v in maxint]) + struct.pack(">ii", smallidx, len(data)) + data + b"\0" * pad)
