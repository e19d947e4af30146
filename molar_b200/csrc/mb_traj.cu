// mb_traj.cu — trajectory ingest: DCD and XTC byte streams decoded ON THE DEVICE into the resident
// frame batch ([F][N][3] f32, the layout every mb_batch_* call works on).
//
// Replaces the reference's per-frame readers on the IO thread:
//   DCD  molar/src/io/dcd_handler.rs:204-300 (header), :389-450 (read_state: x[],y[],z[] f32 records,
//        either endianness, A -> nm as `x as Float * 0.1`, optional fixed-atom frames), :172-200 (unit cell)
//   XTC  molar/src/io/xtc_handler.rs:64-110, which delegates the decompression to the third-party crate
//        `molly` (Cargo.toml:36, git dependency, not in the reference tree); what is restated here is the
//        published xdrfile algorithm molly implements (xdr3dfcoord: mixed-radix packed integers,
//        run-length coded small displacements, adaptive small-integer width).
// Only record markers and headers are read on the host (a few bytes per frame); coordinates never are.
//
// XTC on a GPU.  The bit stream of a frame is a chain of GROUPS (one full-width atom, a flag bit, an
// optional 5-bit run code, run/3 small atoms); where a group starts depends on every flag before it,
// what it contains does not.  So the decode is split:
//   xtc_scan_kernel    one warp per frame walks ONLY the flag / run bits (32 groups per step while the flag
//                      stays clear) and records, per group, the bit offset, the first atom index, the
//                      small-integer width and the run length;
//   xtc_decode_kernel  one thread per group unpacks the mixed-radix integers (64/128-bit arithmetic
//                      instead of the byte-wise long division of the serial code), applies the delta
//                      chain inside the group and writes nm coordinates.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "mb_common.cuh"

namespace mb {

// ---------------------------------------------------------------------------------------------
// host helpers
// ---------------------------------------------------------------------------------------------
static inline uint32_t rd_u32(const uint8_t* p, bool big) {
    return big ? ((uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3])
               : ((uint32_t)p[3] << 24 | (uint32_t)p[2] << 16 | (uint32_t)p[1] << 8 | p[0]);
}
static inline float rd_f32(const uint8_t* p, bool big) {
    uint32_t u = rd_u32(p, big);
    float f;
    memcpy(&f, &u, 4);
    return f;
}
static inline double rd_f64(const uint8_t* p, bool big) {
    uint64_t u = 0;
    for (int i = 0; i < 8; ++i) u |= (uint64_t)p[big ? 7 - i : i] << (8 * i);
    double d;
    memcpy(&d, &u, 8);
    return d;
}

// PeriodicBox::from_vectors_angles (periodic_box.rs:188-235), f32; false on error.  m9: column-major.
static bool box_from_vectors_angles(float a, float b, float c, float alpha, float beta, float gamma, float m9[9]) {
    float m[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    if (a == 0.0f || b == 0.0f || c == 0.0f) return false;
    if (alpha < 60.0f || beta < 60.0f || gamma < 60.0f) return false;
    m[0][0] = a;
    if (alpha != 90.0f || beta != 90.0f || gamma != 90.0f) {
        const float rads_per_deg = 3.14159265358979323846f / 180.0f;  // Rust to_radians(): self * (PI / 180)
        const float cosa = alpha != 90.0f ? std::cos(alpha * rads_per_deg) : 0.0f;
        const float cosb = beta != 90.0f ? std::cos(beta * rads_per_deg) : 0.0f;
        float sing = 1.0f, cosg = 0.0f;
        if (gamma != 90.0f) {
            sing = std::sin(gamma * rads_per_deg);
            cosg = std::cos(gamma * rads_per_deg);
        }
        m[0][1] = b * cosg;
        m[1][1] = b * sing;
        m[0][2] = c * cosb;
        m[1][2] = c * (cosa - cosb * cosg) / sing;
        m[2][2] = std::sqrt(c * c - std::pow(m[0][2], 2.0f) - std::pow(m[1][2], 2.0f));
    } else {
        m[1][1] = b;
        m[2][2] = c;
    }
    for (int col = 0; col < 3; ++col)
        for (int row = 0; row < 3; ++row) m9[col * 3 + row] = m[row][col];
    return true;
}

// parse_unit_cell (dcd_handler.rs:172-200): [A, cos(gamma)|gamma, B, cos(beta)|beta, cos(alpha)|alpha, C]
static bool dcd_unit_cell(const double cell[6], float m9[9]) {
    const double a = cell[0], b = cell[2], c = cell[5];
    if (a == 0.0) return false;
    const double rad2deg = 180.0 / 3.14159265358979323846;
    double alpha, beta, gamma;
    if (std::fabs(cell[4]) <= 1.0) {
        alpha = std::acos(cell[4]) * rad2deg;
        beta = std::acos(cell[3]) * rad2deg;
        gamma = std::acos(cell[1]) * rad2deg;
    } else {
        alpha = cell[4];
        beta = cell[3];
        gamma = cell[1];
    }
    return box_from_vectors_angles((float)(a * 0.1), (float)(b * 0.1), (float)(c * 0.1), (float)alpha, (float)beta,
                                   (float)gamma, m9);
}

struct DcdInfo {
    bool swap = false;  // file is big-endian
    size_t n_atoms = 0, n_fixed = 0, n_free = 0;
    bool extra = false, fourd = false;
    std::vector<uint32_t> free_idx;  // 0-based
    size_t frame_offset = 0;
    size_t first_size = 0, next_size = 0;
    size_t n_frames = 0;  // complete frames present in the buffer
    int istart = 0, nsavc = 0;
    float delta = 0.0f;
};

static int dcd_parse(const uint8_t* b, size_t nb, DcdInfo& I) {
    if (nb < 92) return fail(MB_ERR_ARG, "dcd: buffer too small for a header");
    const uint32_t le = rd_u32(b, false), be = rd_u32(b, true);
    if (le == 84) I.swap = false;
    else if (be == 84) I.swap = true;
    else return fail(MB_ERR_ARG, "dcd: bad magic (first record length is not 84)");
    const bool big = I.swap;
    const uint8_t* h = b + 4;
    if (rd_u32(b + 88, big) != 84) return fail(MB_ERR_ARG, "dcd: bad header record");
    if (memcmp(h, "CORD", 4) != 0) return fail(MB_ERR_ARG, "dcd: bad magic (no CORD)");
    const int32_t nfix = (int32_t)rd_u32(h + 32, big);
    I.n_fixed = nfix > 0 ? (size_t)nfix : 0;
    I.extra = rd_u32(h + 40, big) != 0;
    I.fourd = rd_u32(h + 44, big) != 0;
    I.istart = (int32_t)rd_u32(h + 8, big);
    I.nsavc = (int32_t)rd_u32(h + 12, big);
    const bool is_charmm = rd_u32(h + 76, big) != 0;
    I.delta = is_charmm ? rd_f32(h + 36, big) : (float)rd_f64(h + 36, big);  // dcd_handler.rs:236-240
    size_t off = 92;
    auto skip_record = [&](size_t* len_out, size_t* payload) -> bool {
        if (off + 4 > nb) return false;
        const size_t len = rd_u32(b + off, big);
        if (off + 8 + len > nb) return false;
        if (rd_u32(b + off + 4 + len, big) != len) return false;
        if (len_out) *len_out = len;
        if (payload) *payload = off + 4;
        off += 8 + len;
        return true;
    };
    if (!skip_record(nullptr, nullptr)) return fail(MB_ERR_ARG, "dcd: bad title record");
    size_t len = 0, pay = 0;
    if (!skip_record(&len, &pay) || len < 4) return fail(MB_ERR_ARG, "dcd: bad atom-count record");
    const int32_t na = (int32_t)rd_u32(b + pay, big);
    I.n_atoms = na > 0 ? (size_t)na : 0;
    if (I.n_atoms == 0) return fail(MB_ERR_ARG, "dcd: no atoms");
    if (I.n_fixed > I.n_atoms) return fail(MB_ERR_ARG, "dcd: more fixed atoms than atoms");
    I.n_free = I.n_atoms - I.n_fixed;
    if (I.n_fixed > 0) {
        if (!skip_record(&len, &pay)) return fail(MB_ERR_ARG, "dcd: bad free-atom index record");
        const size_t cnt = std::min(len / 4, I.n_free);
        if (cnt < I.n_free) return fail(MB_ERR_ARG, "dcd: free-atom index record too short");
        I.free_idx.resize(cnt);
        for (size_t k = 0; k < cnt; ++k) {
            const uint32_t v = rd_u32(b + pay + 4 * k, big);
            I.free_idx[k] = v > 0 ? v - 1 : 0;  // saturating_sub(1)
            if (I.free_idx[k] >= I.n_atoms) return fail(MB_ERR_ARG, "dcd: free-atom index out of range");
        }
    }
    I.frame_offset = off;
    const size_t ex = I.extra ? 56 : 0;
    I.first_size = ex + 3 * (I.n_atoms * 4 + 8) + (I.fourd ? I.n_atoms * 4 + 8 : 0);
    I.next_size = I.n_fixed ? ex + 3 * (I.n_free * 4 + 8) + (I.fourd ? I.n_free * 4 + 8 : 0) : I.first_size;
    I.n_frames = 0;
    if (nb >= off + I.first_size) I.n_frames = 1 + (nb - off - I.first_size) / I.next_size;
    return MB_OK;
}

static inline size_t dcd_frame_off(const DcdInfo& I, size_t f) {
    return f == 0 ? I.frame_offset : I.frame_offset + I.first_size + (f - 1) * I.next_size;
}

// record markers of one frame (the reference fails with BadRecord on any mismatch, dcd_handler.rs:84-130)
static int dcd_check_frame(const uint8_t* b, const DcdInfo& I, size_t f) {
    const bool big = I.swap;
    size_t off = dcd_frame_off(I, f);
    const size_t n = (f == 0 || I.n_fixed == 0) ? I.n_atoms : I.n_free;
    if (I.extra) {
        if (rd_u32(b + off, big) != 48 || rd_u32(b + off + 52, big) != 48)
            return fail(MB_ERR_ARG, "dcd: frame %zu: bad unit-cell record", f);
        off += 56;
    }
    for (int blk = 0; blk < 3; ++blk) {
        if (rd_u32(b + off, big) != n * 4 || rd_u32(b + off + 4 + n * 4, big) != n * 4)
            return fail(MB_ERR_ARG, "dcd: frame %zu: unexpected coordinate record length", f);
        off += n * 4 + 8;
    }
    return MB_OK;
}

// ---------------------------------------------------------------------------------------------
// DCD on the device: x[], y[], z[] blocks -> [atom][3], byte swap, * 0.1
// ---------------------------------------------------------------------------------------------
struct DcdDev {
    const uint32_t* raw;  // bytes of the loaded frames, starting at the first one (4-byte aligned)
    unsigned long long first_words, next_words;  // frame sizes in 32-bit words
    int first_is_full;    // the first loaded frame is frame 0 of the file (full atom set)
    int n_atoms, n_free, n_fixed, extra_words, swap;
    const int* slot;      // fixed-atom files: atom -> index in the free-atom blocks, or -1
    const float* fixed;   // fixed-atom files: decoded frame 0 of the file
    float* out;
};

__device__ __forceinline__ float dcd_val(uint32_t w, int swap) {
    if (swap) w = __byte_perm(w, 0, 0x0123);
    return xmul(__uint_as_float(w), 0.1f);  // `x as Float * 0.1` (dcd_handler.rs:425)
}

__global__ void __launch_bounds__(256) dcd_unpack_kernel(const __grid_constant__ DcdDev D) {
    __shared__ float s[768];
    const unsigned f = blockIdx.y;
    const int a0 = blockIdx.x * 256, t = threadIdx.x;
    const bool full = (D.n_fixed == 0) || (f == 0 && D.first_is_full);
    const unsigned long long fw = f == 0 ? 0ull : D.first_words + (unsigned long long)(f - 1) * D.next_words;
    const int nblk = full ? D.n_atoms : D.n_free;  // entries per coordinate block of this frame
    const uint32_t* xb = D.raw + fw + D.extra_words + 1;
    const uint32_t* yb = xb + nblk + 2;
    const uint32_t* zb = yb + nblk + 2;
    const int a = a0 + t;
    if (a < D.n_atoms) {
        float x, y, z;
        const int k = full ? a : D.slot[a];
        if (k >= 0) {
            x = dcd_val(__ldg(xb + k), D.swap);
            y = dcd_val(__ldg(yb + k), D.swap);
            z = dcd_val(__ldg(zb + k), D.swap);
        } else {  // fixed atom: coordinates of frame 0, already in nm (dcd_handler.rs:437-440)
            x = D.fixed[3 * (size_t)a];
            y = D.fixed[3 * (size_t)a + 1];
            z = D.fixed[3 * (size_t)a + 2];
        }
        s[3 * t] = x;
        s[3 * t + 1] = y;
        s[3 * t + 2] = z;
    }
    __syncthreads();
    const int cnt = min(256, D.n_atoms - a0) * 3;
    float* o = D.out + ((size_t)f * D.n_atoms + a0) * 3;
    for (int i = t; i < cnt; i += 256) o[i] = s[i];
}

// ---------------------------------------------------------------------------------------------
// XTC on the device
// ---------------------------------------------------------------------------------------------
__constant__ int c_magicints[73] = {
    0, 0, 0, 0, 0, 0, 0, 0, 0, 8, 10, 12, 16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512, 645,
    812, 1024, 1290, 1625, 2048, 2580, 3250, 4096, 5060, 6501, 8192, 10321, 13003, 16384, 20642, 26007, 32768, 41285,
    52015, 65536, 82570, 104031, 131072, 165140, 208063, 262144, 330280, 416127, 524287, 660561, 832255, 1048576,
    1321122, 1664510, 2097152, 2642245, 3329021, 4194304, 5284491, 6658042, 8388607, 10568983, 13316085, 16777216};
constexpr int XTC_FIRSTIDX = 9;
constexpr int XTC_LASTIDX = 73;

struct XtcFrame {
    unsigned long long data_off;   // byte offset of the compressed block (or of the raw floats) in the device buffer
    unsigned long long group_off;  // first entry of this frame in the group arrays
    unsigned nbytes;
    int natoms;
    int raw;                       // natoms <= 9: uncompressed big-endian floats
    int minint[3];
    unsigned sizeint[3];
    int bitsize;                   // 0: three separate fields of bitsizeint[] bits
    int bitsizeint[3];
    int smallidx;
    float inv_precision;
};

// MSB-first bit reader (xdrfile receivebits); n <= 32.  The buffer is padded, reading one byte past is safe.
__device__ __forceinline__ unsigned xtc_bits(const uint8_t* __restrict__ d, unsigned long long bp, int n) {
    unsigned v = 0;
    while (n > 0) {
        const unsigned byte = d[bp >> 3];
        const int avail = 8 - (int)(bp & 7);
        const int take = min(avail, n);
        v = (v << take) | ((byte >> (avail - take)) & ((1u << take) - 1u));
        bp += take;
        n -= take;
    }
    return v;
}

// xdrfile receiveints for three integers: the field is read as BYTES, first byte least significant (the last,
// partial byte most significant), and the resulting number is split by two divisions.
__device__ __forceinline__ void xtc_unpack3(const uint8_t* __restrict__ d, unsigned long long bp, int nbits, unsigned s1,
                                            unsigned s2, int out[3]) {
    if (nbits <= 64) {
        unsigned long long v = 0;
        int shift = 0;
        while (nbits > 8) {
            v |= (unsigned long long)xtc_bits(d, bp, 8) << shift;
            bp += 8;
            shift += 8;
            nbits -= 8;
        }
        if (nbits > 0) v |= (unsigned long long)xtc_bits(d, bp, nbits) << shift;
        if (v <= 0xffffffffull) {
            unsigned w = (unsigned)v;
            out[2] = (int)(w % s2);
            w /= s2;
            out[1] = (int)(w % s1);
            out[0] = (int)(w / s1);
        } else {
            out[2] = (int)(v % s2);
            v /= s2;
            out[1] = (int)(v % s1);
            out[0] = (int)(unsigned)(v / s1);
        }
    } else {
        unsigned __int128 v = 0;
        int shift = 0;
        while (nbits > 8) {
            v |= (unsigned __int128)xtc_bits(d, bp, 8) << shift;
            bp += 8;
            shift += 8;
            nbits -= 8;
        }
        if (nbits > 0) v |= (unsigned __int128)xtc_bits(d, bp, nbits) << shift;
        out[2] = (int)(unsigned)(v % s2);
        v /= s2;
        out[1] = (int)(unsigned)(v % s1);
        out[0] = (int)(unsigned)(v / s1);
    }
}

// One WARP per frame: group table.  The walk is serial in principle (a group's length depends on its own flag
// bit), but as long as the flag stays clear every group has the same length — the steady state of solvent,
// where the run length repeats — so the 32 lanes test the flag bits of the next 32 groups under that assumption
// and the warp advances over the longest confirmed prefix at once; a set flag (new run code, possibly a new
// small-integer width) is then handled as one serial step.  A step is ~0.33 us of dependent instructions (staging
// the stream in shared memory did not change that: it is not a memory-latency chain), so one frame of 1M atoms takes
// ~12 ms on its own — throughput comes from the frames in flight: one warp each, thousands fit on the chip.
constexpr int XTC_SCAN_THREADS = 128;
__global__ void __launch_bounds__(XTC_SCAN_THREADS) xtc_scan_kernel(const uint8_t* __restrict__ raw,
                                                                    const XtcFrame* __restrict__ frames, int nf,
                                                                    unsigned* __restrict__ g_bit,
                                                                    unsigned* __restrict__ g_atom,
                                                                    unsigned short* __restrict__ g_meta,
                                                                    unsigned* __restrict__ g_count,
                                                                    int* __restrict__ status) {
    const int f = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    const unsigned lane = threadIdx.x & 31u;
    if (f >= nf) return;
    const XtcFrame F = frames[f];
    if (F.raw) {
        if (lane == 0) g_count[f] = 0;
        return;
    }
    const uint8_t* d = raw + F.data_off;
    const unsigned long long total_bits = (unsigned long long)F.nbytes * 8ull;
    const unsigned fullbits = (unsigned)(F.bitsize ? F.bitsize : F.bitsizeint[0] + F.bitsizeint[1] + F.bitsizeint[2]);
    unsigned long long bp = 0;
    int i = 0, nsmall = 0, smallidx = F.smallidx;  // nsmall = run / 3: small triples behind the full one of a group
    unsigned ng = 0;
    bool bad = false;
    // The walk is one dependent chain per frame — every set flag changes the layout of what follows — so its speed is
    // the latency of one step.  The step is kept short: the flag AND the run code behind it come from one 16-bit window
    // (two independent byte loads; the lane that finds its flag set already holds the new run code, which reaches the
    // others by a shuffle instead of a second dependent load), the divisions are gone (run / 3 as a multiply, the
    // group count of the tail only computed in the tail), and the prefetch is issued once per 1 KB of stream.
    unsigned long long pf = 0, pf_blk = ~0ull;
    while (i < F.natoms) {
        if (smallidx < XTC_FIRSTIDX || smallidx >= XTC_LASTIDX) {
            bad = true;
            break;
        }
        const unsigned long long cur = bp >> 3;
        if ((cur >> 10) != pf_blk) {
            pf_blk = cur >> 10;
            while (pf < cur + 8192ull && pf < F.nbytes) {
                const unsigned long long a = pf + (unsigned long long)lane * 128ull;
                if (a < F.nbytes) asm volatile("prefetch.global.L2 [%0];" ::"l"(d + a));
                pf += 4096ull;
            }
            const unsigned long long a1 = cur + 256ull + (unsigned long long)lane * 128ull;
            if (lane < 16u && a1 < F.nbytes) asm volatile("prefetch.global.L1 [%0];" ::"l"(d + a1));
        }
        const int per = 1 + nsmall;
        const unsigned len0 = fullbits + 1u + (unsigned)(nsmall * smallidx);  // length of a group whose flag is clear
        const int remaining = F.natoms - i;
        const int left = remaining >= 32 * per ? 32 : (remaining + per - 1) / per;
        const unsigned long long p = bp + (unsigned long long)lane * len0, fp = p + fullbits;
        const bool ok = (int)lane < left && fp < total_bits;
        // bits fp .. fp + 5: flag, then the 5-bit run code (meaningful when the flag is set); the stream is padded,
        // so the window may be read past its end — whether the code is really there is checked below
        unsigned six = 0x20u | 0x100u;  // no group here: reads as "flag set, no valid run code"
        if (ok) {
            const unsigned long long by = fp >> 3;
            const unsigned w16 = ((unsigned)d[by] << 8) | (unsigned)d[by + 1];
            six = (w16 >> (10u - (unsigned)(fp & 7ull))) & 0x3fu;
            if (fp + 6 > total_bits) six |= 0x100u;
        }
        const unsigned clear = __ballot_sync(0xffffffffu, (six & 0x20u) == 0u);
        const int nz = clear == 0xffffffffu ? 32 : __ffs((int)~clear) - 1;
        if ((int)lane < nz) {
            const size_t e = F.group_off + ng + lane;
            g_bit[e] = (unsigned)p;
            g_atom[e] = (unsigned)(i + (int)lane * per);
            g_meta[e] = (unsigned short)(smallidx | ((3 * nsmall) << 8));
        }
        const unsigned nxt = __shfl_sync(0xffffffffu, six, nz & 31);
        bp += (unsigned long long)nz * len0;
        i += nz * per;
        ng += (unsigned)nz;
        if (nz < 32 && i < F.natoms) {
            // the next group has its flag set: new run code (and possibly a new small-integer width)
            if (nxt & 0x100u) {
                bad = true;
                break;
            }
            const int code = (int)(nxt & 31u);
            const int ns = (code * 11) >> 5;  // code / 3 for code < 32
            const int is_smaller = code - 3 * ns - 1;
            if (lane == 0) {
                const size_t e = F.group_off + ng;
                g_bit[e] = (unsigned)bp;
                g_atom[e] = (unsigned)i;
                g_meta[e] = (unsigned short)(smallidx | ((3 * ns) << 8));
            }
            ++ng;
            i += 1 + ns;
            bp += fullbits + 6ull + (unsigned long long)ns * (unsigned)smallidx;
            smallidx += is_smaller;
            nsmall = ns;
        }
    }
    if (lane == 0) {
        if (bad || i > F.natoms) atomicExch(status, f + 1);  // corrupt stream
        g_count[f] = ng;
    }
}

// one thread per group (blockIdx.y = frame)
__global__ void __launch_bounds__(128) xtc_decode_kernel(const uint8_t* __restrict__ raw,
                                                         const XtcFrame* __restrict__ frames,
                                                         const unsigned* __restrict__ g_bit,
                                                         const unsigned* __restrict__ g_atom,
                                                         const unsigned short* __restrict__ g_meta,
                                                         const unsigned* __restrict__ g_count, float* __restrict__ out,
                                                         int natoms) {
    const int f = blockIdx.y;
    const XtcFrame& F = frames[f];
    float* o = out + (size_t)f * natoms * 3;
    if (F.raw) {
        // natoms <= 9: plain big-endian floats
        for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < natoms * 3; k += gridDim.x * blockDim.x) {
            const uint8_t* p = raw + F.data_off + 4 * (size_t)k;
            o[k] = __uint_as_float((unsigned)p[0] << 24 | (unsigned)p[1] << 16 | (unsigned)p[2] << 8 | p[3]);
        }
        return;
    }
    // The stream is untrusted input and this kernel runs before the scan's verdict is read back, so it is safe on
    // its own: the group table never exceeds natoms entries (one group holds at least one atom), a group whose bits
    // would run past the end of the compressed block or whose small-integer width is invalid is skipped, and every
    // store is clamped to the frame.  (A corrupt stream is still reported through `status` by the scan.)
    const unsigned ng = min(g_count[f], (unsigned)natoms);
    const uint8_t* d = raw + F.data_off;
    const float inv = F.inv_precision;
    const unsigned long long total_bits = (unsigned long long)F.nbytes * 8ull;
    const unsigned fullbits = (unsigned)(F.bitsize ? F.bitsize : F.bitsizeint[0] + F.bitsizeint[1] + F.bitsizeint[2]);
    auto put = [&](int at, const int v[3]) {
        if (at >= 0 && at < natoms) {
            o[3 * (size_t)at] = (float)v[0] * inv;
            o[3 * (size_t)at + 1] = (float)v[1] * inv;
            o[3 * (size_t)at + 2] = (float)v[2] * inv;
        }
    };
    for (unsigned g = blockIdx.x * blockDim.x + threadIdx.x; g < ng; g += gridDim.x * blockDim.x) {
        const size_t e = F.group_off + g;
        unsigned long long bp = g_bit[e];
        int i = (int)g_atom[e];
        const int smallidx = g_meta[e] & 0xff, run = g_meta[e] >> 8;
        if (run > 0 && (smallidx < XTC_FIRSTIDX || smallidx >= XTC_LASTIDX || c_magicints[smallidx] == 0)) continue;
        if (bp + fullbits + 1ull + (unsigned long long)(run / 3) * (unsigned)smallidx > total_bits) continue;
        int cur[3];
        if (F.bitsize == 0) {
            cur[0] = (int)xtc_bits(d, bp, F.bitsizeint[0]);
            cur[1] = (int)xtc_bits(d, bp + F.bitsizeint[0], F.bitsizeint[1]);
            cur[2] = (int)xtc_bits(d, bp + F.bitsizeint[0] + F.bitsizeint[1], F.bitsizeint[2]);
            bp += F.bitsizeint[0] + F.bitsizeint[1] + F.bitsizeint[2];
        } else {
            xtc_unpack3(d, bp, F.bitsize, F.sizeint[1], F.sizeint[2], cur);
            bp += F.bitsize;
        }
        cur[0] += F.minint[0];
        cur[1] += F.minint[1];
        cur[2] += F.minint[2];
        // flag bit, and the 5-bit run code that follows it when the flag is set
        bp += xtc_bits(d, bp, 1) ? 6 : 1;
        if (run > 0) {
            const unsigned ss = (unsigned)c_magicints[smallidx];
            const int smallnum = (int)(ss / 2);
            int prev[3] = {cur[0], cur[1], cur[2]};
            for (int k = 0; k < run; k += 3) {
                int t[3];
                xtc_unpack3(d, bp, smallidx, ss, ss, t);
                bp += smallidx;
                t[0] += prev[0] - smallnum;
                t[1] += prev[1] - smallnum;
                t[2] += prev[2] - smallnum;
                if (k == 0) {
                    // the first small atom is stored BEFORE the full one (water: O H H -> H O H on disk)
                    put(i, t);
                    ++i;
                    put(i, cur);
                    ++i;
                    // after the swap the delta chain continues from the small atom
                    prev[0] = t[0];
                    prev[1] = t[1];
                    prev[2] = t[2];
                } else {
                    prev[0] = t[0];
                    prev[1] = t[1];
                    prev[2] = t[2];
                    put(i, t);
                    ++i;
                }
            }
        } else {
            put(i, cur);
        }
    }
}

// xdrfile sizeofint / sizeofints
static int xtc_sizeofint(unsigned size) {
    unsigned long long num = 1;
    int nbits = 0;
    while (size >= num && nbits < 32) {
        nbits++;
        num <<= 1;
    }
    return nbits;
}
static int xtc_sizeofints(const unsigned sizes[3]) {
    unsigned bytes[32];
    unsigned nbytes = 1, bytecnt, tmp;
    bytes[0] = 1;
    int nbits = 0;
    for (int i = 0; i < 3; ++i) {
        tmp = 0;
        for (bytecnt = 0; bytecnt < nbytes; bytecnt++) {
            unsigned long long t = (unsigned long long)bytes[bytecnt] * sizes[i] + tmp;
            bytes[bytecnt] = (unsigned)(t & 0xff);
            tmp = (unsigned)(t >> 8);
        }
        while (tmp != 0) {
            bytes[bytecnt++] = tmp & 0xff;
            tmp >>= 8;
        }
        nbytes = bytecnt;
    }
    unsigned num = 1;
    nbytes--;
    while (bytes[nbytes] >= num) {
        nbits++;
        num *= 2;
    }
    return nbits + (int)nbytes * 8;
}

struct XtcHostFrame {
    size_t off;       // frame start in the file
    size_t size;      // bytes of the whole frame
    XtcFrame dev;     // data_off relative to the file start (rebased when uploaded)
    float box[9];     // as stored: row-major, rows = box vectors
    float time;
    int step;
};

static int xtc_parse(const uint8_t* b, size_t nb, std::vector<XtcHostFrame>& frames, size_t* natoms_out) {
    size_t off = 0;
    size_t natoms0 = 0;
    while (off + 56 <= nb) {
        XtcHostFrame H;
        memset(&H, 0, sizeof(H));
        const uint32_t magic = rd_u32(b + off, true);
        if (magic != 1995) {
            if (frames.empty()) return fail(MB_ERR_ARG, "xtc: bad magic %u (only the 1995 format is supported)", magic);
            return fail(MB_ERR_ARG, "xtc: bad magic in frame %zu", frames.size());
        }
        const int natoms = (int)rd_u32(b + off + 4, true);
        if (natoms <= 0) return fail(MB_ERR_ARG, "xtc: frame %zu: no atoms", frames.size());
        H.step = (int)rd_u32(b + off + 8, true);
        H.time = rd_f32(b + off + 12, true);
        for (int k = 0; k < 9; ++k) H.box[k] = rd_f32(b + off + 16 + 4 * k, true);
        const int lsize = (int)rd_u32(b + off + 52, true);
        if (lsize != natoms) return fail(MB_ERR_ARG, "xtc: frame %zu: atom count mismatch", frames.size());
        if (frames.empty()) natoms0 = (size_t)natoms;
        else if ((size_t)natoms != natoms0) return fail(MB_ERR_ARG, "xtc: frame %zu: atom count changes", frames.size());
        H.off = off;
        H.dev.natoms = natoms;
        if (natoms <= 9) {
            H.size = 56 + 12 * (size_t)natoms;
            if (off + H.size > nb) break;
            H.dev.raw = 1;
            H.dev.data_off = off + 56;
        } else {
            if (off + 92 > nb) break;
            const float precision = rd_f32(b + off + 56, true);
            unsigned sizeint[3];
            for (int k = 0; k < 3; ++k) {
                H.dev.minint[k] = (int)rd_u32(b + off + 60 + 4 * k, true);
                const int mx = (int)rd_u32(b + off + 72 + 4 * k, true);
                sizeint[k] = (unsigned)(mx - H.dev.minint[k] + 1);
                H.dev.sizeint[k] = sizeint[k];
            }
            if (sizeint[0] == 0 || sizeint[1] == 0 || sizeint[2] == 0)
                return fail(MB_ERR_ARG, "xtc: frame %zu: empty coordinate range (maxint < minint)", frames.size());
            if (!(precision > 0.0f) || !std::isfinite(precision))
                return fail(MB_ERR_ARG, "xtc: frame %zu: bad precision", frames.size());
            H.dev.smallidx = (int)rd_u32(b + off + 84, true);
            H.dev.nbytes = rd_u32(b + off + 88, true);
            H.size = 92 + ((size_t)H.dev.nbytes + 3) / 4 * 4;
            if (off + H.size > nb) break;
            if (H.dev.smallidx < XTC_FIRSTIDX || H.dev.smallidx >= XTC_LASTIDX)
                return fail(MB_ERR_ARG, "xtc: frame %zu: bad small-integer index", frames.size());
            if ((sizeint[0] | sizeint[1] | sizeint[2]) > 0xffffffu) {
                for (int k = 0; k < 3; ++k) H.dev.bitsizeint[k] = xtc_sizeofint(sizeint[k]);
                H.dev.bitsize = 0;
            } else {
                H.dev.bitsize = xtc_sizeofints(sizeint);
            }
            H.dev.inv_precision = 1.0f / precision;
            H.dev.data_off = off + 92;
        }
        frames.push_back(H);
        off += H.size;
    }
    if (frames.empty()) return fail(MB_ERR_ARG, "xtc: no complete frame in the buffer");
    *natoms_out = natoms0;
    return MB_OK;
}

// XTC stores the box as three ROW vectors; the reference fills a column-major Matrix3 from that iterator
// (xtc_handler.rs:99), so the stored floats ARE the column-major matrix with columns = box vectors.
static void xtc_box_colmajor(const float stored[9], float m9[9]) { memcpy(m9, stored, 9 * sizeof(float)); }

// CUDA-event stopwatch on the context stream (mb_get_stat "traj_*")
struct EvTimer {
    cudaEvent_t a = nullptr, b = nullptr;
    cudaStream_t st;
    explicit EvTimer(cudaStream_t s) : st(s) {
        cudaEventCreate(&a);
        cudaEventCreate(&b);
    }
    ~EvTimer() {
        if (a) cudaEventDestroy(a);
        if (b) cudaEventDestroy(b);
    }
    void start() { cudaEventRecord(a, st); }
    void stop() { cudaEventRecord(b, st); }
    double ms() {  // after the stream has been synchronised
        float t = 0.f;
        return cudaEventElapsedTime(&t, a, b) == cudaSuccess ? (double)t : 0.0;
    }
};

static int finish_batch(Ctx& c, size_t n_frames, size_t n_atoms, const float* box9_first) {
    if (box9_first) {
        if (host_box_from_colmajor(box9_first, &c.box) == MB_OK) c.has_box = true;
        else c.has_box = false;
    } else {
        c.has_box = false;
    }
    c.batch_frames = n_frames;
    c.batch_atoms = n_atoms;
    c.d_xyz = c.batch.as<float>();
    c.n_atoms = n_atoms;
    return MB_OK;
}

}  // namespace mb

using namespace mb;

extern "C" {

int mb_traj_probe(const void* bytes, size_t n_bytes, int format, size_t* n_frames, size_t* n_atoms) {
    if (!bytes || !n_frames || !n_atoms) return fail(MB_ERR_ARG, "null argument");
    const uint8_t* b = static_cast<const uint8_t*>(bytes);
    if (format == MB_TRAJ_DCD) {
        DcdInfo I;
        MB_TRY(dcd_parse(b, n_bytes, I));
        *n_frames = I.n_frames;
        *n_atoms = I.n_atoms;
        return MB_OK;
    }
    if (format == MB_TRAJ_XTC) {
        std::vector<XtcHostFrame> fr;
        size_t na = 0;
        MB_TRY(xtc_parse(b, n_bytes, fr, &na));
        *n_frames = fr.size();
        *n_atoms = na;
        return MB_OK;
    }
    return fail(MB_ERR_ARG, "unknown trajectory format %d", format);
}

int mb_batch_load_traj(MbCtx* h, const void* bytes, size_t n_bytes, int format, size_t first_frame, size_t n_frames,
                       float* boxes9_out, float* times_out) {
    if (!h || !bytes) return fail(MB_ERR_ARG, "null argument");
    if (n_frames == 0) return fail(MB_ERR_ARG, "mb_batch_load_traj: no frames requested");
    Ctx& c = h->c;
    MB_CUDA(cudaSetDevice(c.device));
    const uint8_t* b = static_cast<const uint8_t*>(bytes);

    if (format == MB_TRAJ_DCD) {
        DcdInfo I;
        MB_TRY(dcd_parse(b, n_bytes, I));
        if (first_frame + n_frames > I.n_frames)
            return fail(MB_ERR_ARG, "dcd: frames [%zu,%zu) requested, %zu present", first_frame, first_frame + n_frames, I.n_frames);
        if (I.n_atoms > 0x7fffffffull) return fail(MB_ERR_ARG, "dcd: too many atoms");
        for (size_t f = first_frame; f < first_frame + n_frames; ++f) MB_TRY(dcd_check_frame(b, I, f));
        float first_box[9];
        bool have_first_box = false;
        for (size_t f = 0; f < n_frames; ++f) {
            float m9[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            bool ok = false;
            if (I.extra) {
                const uint8_t* p = b + dcd_frame_off(I, first_frame + f) + 4;
                double cell[6];
                for (int k = 0; k < 6; ++k) cell[k] = rd_f64(p + 8 * k, I.swap);
                ok = dcd_unit_cell(cell, m9);
            }
            if (f == 0 && ok) {
                memcpy(first_box, m9, sizeof(m9));
                have_first_box = true;
            }
            if (boxes9_out) memcpy(boxes9_out + 9 * f, m9, sizeof(m9));
            // (istart + cur_frame * nsavc) as Float * delta   (dcd_handler.rs:461)
            if (times_out) times_out[f] = (float)(I.istart + (int)(first_frame + f) * I.nsavc) * I.delta;
        }
        const size_t start = dcd_frame_off(I, first_frame), end = dcd_frame_off(I, first_frame + n_frames);
        MB_TRY(c.traj_raw.reserve(end - start + 16));
        MB_TRY(c.batch.reserve(n_frames * I.n_atoms * 3 * sizeof(float)));
        EvTimer t_h2d(c.stream), t_dec(c.stream);
        t_h2d.start();
        MB_CUDA(cudaMemcpyAsync(c.traj_raw.p, b + start, end - start, cudaMemcpyHostToDevice, c.stream));
        t_h2d.stop();
        c.traj_raw_bytes = (double)(end - start);
        DcdDev D;
        memset(&D, 0, sizeof(D));
        D.raw = c.traj_raw.as<uint32_t>();
        D.first_words = (first_frame == 0 ? I.first_size : I.next_size) / 4;
        D.next_words = I.next_size / 4;
        D.first_is_full = first_frame == 0;
        D.n_atoms = (int)I.n_atoms;
        D.n_free = (int)I.n_free;
        D.n_fixed = (int)I.n_fixed;
        D.extra_words = I.extra ? 14 : 0;
        D.swap = I.swap ? 1 : 0;
        D.out = c.batch.as<float>();
        const unsigned bx = (unsigned)((I.n_atoms + 255) / 256);
        if (I.n_fixed > 0) {
            // atom -> slot in the free-atom blocks; fixed atoms come from frame 0 of the FILE
            std::vector<int> slot(I.n_atoms, -1);
            for (size_t k = 0; k < I.free_idx.size(); ++k) slot[I.free_idx[k]] = (int)k;
            MB_TRY(c.traj_aux.reserve(I.n_atoms * sizeof(int) + I.n_atoms * 3 * sizeof(float) + I.first_size + 64));
            char* aux = static_cast<char*>(c.traj_aux.p);
            int* d_slot = reinterpret_cast<int*>(aux);
            float* d_fixed = reinterpret_cast<float*>(aux + I.n_atoms * sizeof(int));
            MB_CUDA(cudaMemcpyAsync(d_slot, slot.data(), I.n_atoms * sizeof(int), cudaMemcpyHostToDevice, c.stream));
            MB_CUDA(cudaStreamSynchronize(c.stream));  // `slot` goes out of scope below
            D.slot = d_slot;
            if (first_frame == 0) {
                D.fixed = D.out;  // frame 0 is decoded by the same launch: do it first, on its own
                DcdDev D0 = D;
                dcd_unpack_kernel<<<dim3(bx, 1), 256, 0, c.stream>>>(D0);
                c.launches++;
            } else {
                // decode frame 0 of the file into scratch
                uint32_t* d_raw0 = reinterpret_cast<uint32_t*>(aux + I.n_atoms * sizeof(int) + I.n_atoms * 3 * sizeof(float));
                MB_TRY(dcd_check_frame(b, I, 0));
                MB_CUDA(cudaMemcpyAsync(d_raw0, b + I.frame_offset, I.first_size, cudaMemcpyHostToDevice, c.stream));
                DcdDev D0 = D;
                D0.raw = d_raw0;
                D0.first_words = I.first_size / 4;
                D0.first_is_full = 1;
                D0.out = d_fixed;
                D0.fixed = d_fixed;
                dcd_unpack_kernel<<<dim3(bx, 1), 256, 0, c.stream>>>(D0);
                c.launches++;
                D.fixed = d_fixed;
            }
        }
        t_dec.start();
        if (n_frames > 65535) {
            // grid.y limit: launch in slabs
            for (size_t f0 = 0; f0 < n_frames; f0 += 65535) {
                const size_t nf = std::min<size_t>(65535, n_frames - f0);
                DcdDev Ds = D;
                if (f0 > 0) {
                    Ds.raw = D.raw + D.first_words + (f0 - 1) * D.next_words;
                    Ds.first_words = D.next_words;
                    Ds.first_is_full = 0;
                    Ds.out = D.out + f0 * I.n_atoms * 3;
                }
                dcd_unpack_kernel<<<dim3(bx, (unsigned)nf), 256, 0, c.stream>>>(Ds);
                c.launches++;
            }
        } else {
            dcd_unpack_kernel<<<dim3(bx, (unsigned)n_frames), 256, 0, c.stream>>>(D);
            c.launches++;
        }
        t_dec.stop();
        MB_CUDA(cudaGetLastError());
        MB_CUDA(cudaStreamSynchronize(c.stream));
        c.traj_h2d_ms = t_h2d.ms();
        c.traj_decode_ms = t_dec.ms();
        c.traj_scan_ms = 0.0;
        return finish_batch(c, n_frames, I.n_atoms, have_first_box ? first_box : nullptr);
    }

    if (format == MB_TRAJ_XTC) {
        std::vector<XtcHostFrame> fr;
        size_t na = 0;
        MB_TRY(xtc_parse(b, n_bytes, fr, &na));
        if (first_frame + n_frames > fr.size())
            return fail(MB_ERR_ARG, "xtc: frames [%zu,%zu) requested, %zu present", first_frame, first_frame + n_frames, fr.size());
        const size_t start = fr[first_frame].off;
        const size_t end = fr[first_frame + n_frames - 1].off + fr[first_frame + n_frames - 1].size;
        MB_TRY(c.traj_raw.reserve(end - start + 64));
        EvTimer t_h2d(c.stream);
        t_h2d.start();
        MB_CUDA(cudaMemcpyAsync(c.traj_raw.p, b + start, end - start, cudaMemcpyHostToDevice, c.stream));
        t_h2d.stop();
        c.traj_raw_bytes = (double)(end - start);
        c.traj_decode_ms = c.traj_scan_ms = 0.0;
        MB_CUDA(cudaMemsetAsync(static_cast<char*>(c.traj_raw.p) + (end - start), 0, 64, c.stream));
        MB_TRY(c.batch.reserve(n_frames * na * 3 * sizeof(float)));
        // frames are decoded in passes that bound the group tables (<= one group per atom)
        const size_t pass_frames = std::max<size_t>(1, std::min<size_t>(std::min<size_t>(n_frames, 65535), ((size_t)1 << 27) / na));
        const size_t max_groups = pass_frames * na;
        const size_t desc_bytes = (pass_frames * sizeof(XtcFrame) + 255) / 256 * 256;
        const size_t cnt_bytes = (pass_frames * sizeof(unsigned) + 255) / 256 * 256;
        MB_TRY(c.traj_aux.reserve(256 + desc_bytes + cnt_bytes + max_groups * (4 + 4 + 2) + 64));
        char* aux = static_cast<char*>(c.traj_aux.p);
        int* d_status = reinterpret_cast<int*>(aux);
        XtcFrame* d_frames = reinterpret_cast<XtcFrame*>(aux + 256);
        unsigned* d_count = reinterpret_cast<unsigned*>(aux + 256 + desc_bytes);
        unsigned* d_bit = reinterpret_cast<unsigned*>(aux + 256 + desc_bytes + cnt_bytes);
        unsigned* d_atom = d_bit + max_groups;
        unsigned short* d_meta = reinterpret_cast<unsigned short*>(d_atom + max_groups);
        MB_CUDA(cudaMemsetAsync(d_status, 0, sizeof(int), c.stream));
        std::vector<XtcFrame> desc(pass_frames);
        for (size_t f0 = 0; f0 < n_frames; f0 += pass_frames) {
            const size_t nf = std::min(pass_frames, n_frames - f0);
            for (size_t k = 0; k < nf; ++k) {
                desc[k] = fr[first_frame + f0 + k].dev;
                desc[k].data_off -= start;
                desc[k].group_off = k * na;
            }
            MB_CUDA(cudaMemcpyAsync(d_frames, desc.data(), nf * sizeof(XtcFrame), cudaMemcpyHostToDevice, c.stream));
            MB_CUDA(cudaStreamSynchronize(c.stream));  // `desc` is reused by the next pass
            EvTimer t_scan(c.stream), t_all(c.stream);
            t_all.start();
            t_scan.start();
            xtc_scan_kernel<<<(unsigned)((nf * 32 + XTC_SCAN_THREADS - 1) / XTC_SCAN_THREADS), XTC_SCAN_THREADS, 0, c.stream>>>(c.traj_raw.as<uint8_t>(), d_frames, (int)nf,
                                                                             d_bit, d_atom, d_meta, d_count, d_status);
            t_scan.stop();
            const unsigned gx = (unsigned)std::max<size_t>(1, std::min<size_t>((na + 127) / 128, 1024));
            xtc_decode_kernel<<<dim3(gx, (unsigned)nf), 128, 0, c.stream>>>(c.traj_raw.as<uint8_t>(), d_frames, d_bit, d_atom,
                                                                            d_meta, d_count,
                                                                            c.batch.as<float>() + f0 * na * 3, (int)na);
            t_all.stop();
            c.launches += 2;
            MB_CUDA(cudaGetLastError());
            MB_CUDA(cudaStreamSynchronize(c.stream));
            c.traj_decode_ms += t_all.ms();
            c.traj_scan_ms += t_scan.ms();
        }
        c.traj_h2d_ms = t_h2d.ms();
        int status = 0;
        MB_CUDA(cudaMemcpyAsync(&status, d_status, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
        MB_CUDA(cudaStreamSynchronize(c.stream));
        if (status != 0) return fail(MB_ERR_ARG, "xtc: corrupt compressed block in frame %zu", first_frame + (size_t)status - 1);
        float first_box[9];
        for (size_t f = 0; f < n_frames; ++f) {
            float m9[9];
            xtc_box_colmajor(fr[first_frame + f].box, m9);
            if (f == 0) memcpy(first_box, m9, sizeof(m9));
            if (boxes9_out) memcpy(boxes9_out + 9 * f, m9, sizeof(m9));
            if (times_out) times_out[f] = fr[first_frame + f].time;
        }
        return finish_batch(c, n_frames, na, first_box);
    }
    return fail(MB_ERR_ARG, "unknown trajectory format %d", format);
}

/* copy frames [f0,f1) of the resident batch to the host (n_atoms x 3 f32 each) */
int mb_batch_download(MbCtx* h, size_t f0, size_t f1, float* xyz_out) {
    if (!h || !xyz_out) return fail(MB_ERR_ARG, "null argument");
    Ctx& c = h->c;
    if (!c.batch.p || f1 > c.batch_frames || f0 >= f1) return fail(MB_ERR_ARG, "mb_batch_download: bad frame range");
    MB_CUDA(cudaSetDevice(c.device));
    const size_t fb = c.batch_atoms * 3 * sizeof(float);
    MB_CUDA(cudaMemcpyAsync(xyz_out, c.batch.as<char>() + f0 * fb, (f1 - f0) * fb, cudaMemcpyDeviceToHost, c.stream));
    MB_CUDA(cudaStreamSynchronize(c.stream));
    return MB_OK;
}

}  // extern "C"
