// mb_search.cu — cell-list neighbour search on sm_100a.
//
// Replaces distance_search_{single,double,within}[_pbc] (molar/src/distance_search.rs:519-954).
// The traversal is NOT the reference's (serial grid of Vec<Vec<..>> + rayon over a cell-pair plan);
// what is kept bit-for-bit is the DECISION the reference takes for every atom pair:
//   * which reference grid cell each atom falls in          (populate / populate_pbc, :120-210)
//   * whether the two cells are MASK-adjacent and with which wrapped dims (search_plan, :217-269)
//   * d2 <= cutoff*cutoff with d2 from the direct difference (un-wrapped cell pair) or from
//     PeriodicBox::distance_squared restricted to the wrapped dims (:485-489, periodic_box.rs:286-318)
// evaluated with unfused f32 arithmetic in the reference's operation order.
//
// Data layout in HBM (per search):
//   tmp4   [n]        float4  {eff.x, eff.y, eff.z, bits(global id)}  binned, unsorted
//   sorted4[n]        float4  same records sorted by FINE cell (x fastest) — one 16-B record per
//                             atom so a neighbour run along x is one contiguous, 128-B-coalesced stream
//   cell_start[nc+1]  u32     exclusive scan of fine-cell populations
//   pairs  [P]        uint2   canonical (i<j) global ids, bump-allocated in 8-KB warp flushes
//   dists  [P]        f32     sqrt(d2) (optional)
//   neighbour-list modes (SearchConnectivity, connectivity.rs:8-38, without a pair list in between):
//   deg / row_ptr [N+1] u32   neighbours per atom (kernel mode 4, full shell) and their exclusive scan
//   cols   [2P]       u32     every atom's row, written at its own cursor (kernel mode 5)
// The fine grid is the reference grid subdivided k[d] times per dimension in FRACTIONAL space, so a
// fine cell lies inside exactly one reference cell and the (adjacent?, wrapped dims) decision is
// uniform per (home fine cell, neighbour run) — nothing but the distance test is left per pair.
#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cstdlib>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "mb_common.cuh"

namespace mb {

constexpr unsigned DROPPED = 0xFFFFFFFFu;
constexpr int MAX_REACH = 7;
constexpr int MAX_ROWS = (2 * MAX_REACH + 1) * (2 * MAX_REACH + 1);
#ifndef MB_STAGE_CAP
#define MB_STAGE_CAP 768
#endif
constexpr int STAGE_CAP = MB_STAGE_CAP;  // pairs staged per warp before a flush (>= 8 homes x 64 candidates of a dense pass)
constexpr int STAGE_ROOM = STAGE_CAP + 2;  // entries allocated per warp: a carried-over odd pair + a full pass of STAGE_CAP (keeps 16-byte alignment)
#ifndef MB_SEARCH_WARPS
#define MB_SEARCH_WARPS 8
#endif
#ifndef MB_SEARCH_MIN_CTAS
#define MB_SEARCH_MIN_CTAS 3
#endif
#ifndef MB_SEARCH_PIPELINE
#define MB_SEARCH_PIPELINE 0
#endif
constexpr int SEARCH_WARPS = MB_SEARCH_WARPS;  // warps per CTA; MB_SEARCH_MIN_CTAS CTAs per SM bound the registers (tuning builds only)
constexpr int CNT_STRIDE = 4;  // u64 words per search: [0] pairs, [1] tile work counter, [2] distance tests, [3] spare
constexpr int PAIR_TAIL = 8192;  // spare pair slots behind pair_cap: the odd last pair of each warp (bulk flushes move even counts)
constexpr int PAD_CANDS = 64;  // finite far-away records behind each sorted array (read by idle lanes of the candidate stream)

struct GridSpec {
    int periodic_variant;  // 1: populate_pbc binning, 0: populate (bounds) binning
    unsigned pbc;          // PbcDims bits
    int dims[3];           // reference Grid::dims
    int k[3];              // subdivision
    unsigned kmagic[3];    // floor(2^32 / k) + 1 (unused for k == 1)
    int hx;                // home tile = hx consecutive fine cells along x (hx divides k[0])
    int fd[3];             // fine dims
    float lower[3];        // non-periodic variant: bounds
    float dim_sz[3];
    DevBox box;
};

struct NbrRow {
    signed char dy, dz, dxlo, dxhi;
};

struct SearchParams {
    const float4* sorted;        // home atoms (set 1), sorted by fine cell
    const unsigned* cell_start;
    const float4* sortedB;       // candidate atoms (set 2); same arrays as above for a single-set search
    const unsigned* cell_startB;
    int two_sets;                // 1: rows cover the full shell, no self run, pairs are (set-1 id, set-2 id)
    unsigned char* flags;        // MODE 3 (`within`): flags[global id of a set-1 atom] = 1 if it has a neighbour
    GridSpec g;
    float rc2;
    float band;            // direct tests: |d2f - rc2| below this is re-evaluated with the reference's exact expression
    float rc2_lo, rc2_hi;  // band around cutoff^2 outside which the shifted-image filter is decisive
    int fast_pbc;          // 1: wrapped cell pairs may use the filter (see plan_cells)
    int nrows;
    int dx_min, dx_max;  // smallest dxlo / largest dxhi over all rows (x reach of the whole table)
    NbrRow rows[MAX_ROWS];
    uint2* pairs;
    float* dists;
    unsigned long long pair_cap;
    unsigned long long* counter;  // [0] pairs found, [1] work counter (as u64), [2] distance tests (MODE 2)
    unsigned long long n_sortedB;  // atoms in the candidate set
    float one;                     // 1.0f, opaque to ptxas: multiplier of the exact packed sums (see d2_pair_fma)
    // vdW search (distance_search.rs:767-879): radii per selected atom of the home / candidate set, indexed by the LOCAL
    // id stored in .w of the records; NULL for a plain cutoff search
    const float* vdwA;
    const float* vdwB;
    // neighbour-list modes (4: degrees, 5: rows): nl_deg[global id] = number of neighbours (MODE 4) / next free slot of
    // the atom's row (MODE 5); nl_cols = the rows, back to back in id order
    unsigned* nl_deg;
    unsigned* nl_cols;
};

// ---------------------------------------------------------------------------------------------
// binning
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int cell_from_float(float v, int dim) {
    // Rust `(x.floor() as usize).clamp(0, dim-1)`: saturating cast, NaN -> 0
    float f = floorf(v);
    if (!(f > 0.0f)) return 0;
    if (f >= (float)dim) return dim - 1;
    return (int)f;
}
__device__ __forceinline__ int subcell(float u, int c, int k) {
    if (k == 1) return 0;
    float t = xsub(u, (float)c);  // in [0,1) up to rounding
    int s = (int)floorf(t * (float)k);
    return min(max(s, 0), k - 1);
}

// One thread per selected atom: reference cell (exact), effective position (exact), fine cell.
// Fine-cell populations are counted with warp-aggregated atomics (one atomic per distinct cell
// per warp), whose return value doubles as the atom's rank inside its cell for the scatter.
__global__ void __launch_bounds__(256) bin_atoms_kernel(const float* __restrict__ xyz,
                                                        const unsigned long long* __restrict__ ids, int n,
                                                        GridSpec g, float4* __restrict__ out4,
                                                        unsigned* __restrict__ cellid, unsigned* __restrict__ rank,
                                                        unsigned* __restrict__ cell_count,
                                                        unsigned long long* __restrict__ refcell, int local_ids) {
    int kidx = blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = kidx < n;
    unsigned cell = DROPPED;
    float ex = 0, ey = 0, ez = 0;
    unsigned gid = 0;
    int loc[3] = {0, 0, 0};
    if (valid) {
        gid = ids ? (unsigned)ids[kidx] : (unsigned)kidx;
        float p[3] = {xyz[3 * (size_t)gid], xyz[3 * (size_t)gid + 1], xyz[3 * (size_t)gid + 2]};
        ex = p[0];
        ey = p[1];
        ez = p[2];
        int sub[3] = {0, 0, 0};
        bool skip = false;
        if (g.periodic_variant) {
            float rel[3];
            xmatvec(g.box.inv, p[0], p[1], p[2], rel[0], rel[1], rel[2]);  // to_box_coords (:156)
            bool correct = true;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                if (rel[d] < 0.0f || rel[d] >= 1.0f) {
                    if (!((g.pbc >> d) & 1u)) skip = true;  // non-periodic dim out of bounds (:163-165)
                    else correct = false;
                    break;  // the reference stops scanning at the first offending dim (:165,168)
                }
            }
            if (!skip) {
                if (!correct) {
#pragma unroll
                    for (int d = 0; d < 3; ++d)
                        if ((g.pbc >> d) & 1u) {
                            float r = xsub(rel[d], truncf(rel[d]));  // fract() (:186)
                            if (r < 0.0f) r = xadd(1.0f, r);         // (:187-189)
                            rel[d] = r;
                        }
                    xmatvec(g.box.m, rel[0], rel[1], rel[2], ex, ey, ez);  // wrapped_pos (:196)
                }
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    float u = xmul(rel[d], (float)g.dims[d]);
                    loc[d] = cell_from_float(u, g.dims[d]);  // (:176-177,191-192)
                    sub[d] = subcell(u, loc[d], g.k[d]);
                }
            }
        } else {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                // (dims * (pos - lower) / dim_sz).floor() as isize  (:133)
                float u = xdiv(xmul((float)g.dims[d], xsub(p[d], g.lower[d])), g.dim_sz[d]);
                float f = floorf(u);
                if (!(f >= 0.0f) || f >= (float)g.dims[d]) {
                    // NaN casts to 0 in Rust: keep that corner exact
                    if (f != f) { loc[d] = 0; sub[d] = 0; continue; }
                    skip = true;
                    break;
                }
                loc[d] = (int)f;
                sub[d] = subcell(u, loc[d], g.k[d]);
            }
        }
        // Atoms with a non-finite coordinate: the reference bins them (NaN casts to cell 0) but no comparison
        // with them is ever true, so they are in no pair.  They are left out here, which keeps NaNs away from the
        // distance loop (it reads a sign bit instead of comparing).
        if (!(isfinite(p[0]) && isfinite(p[1]) && isfinite(p[2]))) skip = true;
        if (!skip) {
            int fx = loc[0] * g.k[0] + sub[0], fy = loc[1] * g.k[1] + sub[1], fz = loc[2] * g.k[2] + sub[2];
            cell = (unsigned)(fx + g.fd[0] * (fy + g.fd[1] * fz));
        }
    }
    if (cell_count) {
        unsigned lane = threadIdx.x & 31u;
        unsigned peers = __match_any_sync(0xffffffffu, cell);
        unsigned leader = __ffs(peers) - 1;
        unsigned base = 0;
        if (lane == leader && cell != DROPPED) base = atomicAdd(&cell_count[cell], __popc(peers));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (valid) rank[kidx] = base + __popc(peers & ((1u << lane) - 1u));
    }
    if (valid) {
        // vdW searches report LOCAL indices (distance_search.rs:791-792): tag with the position in the selection
        out4[kidx] = make_float4(ex, ey, ez, __uint_as_float(local_ids ? (unsigned)kidx : gid));
        cellid[kidx] = cell;
        if (refcell)
            refcell[kidx] = cell == DROPPED ? ~0ull
                                            : ((unsigned long long)loc[0] | ((unsigned long long)loc[1] << 21) |
                                               ((unsigned long long)loc[2] << 42));
    }
}

__global__ void __launch_bounds__(256) scatter_kernel(const float4* __restrict__ in4,
                                                      const unsigned* __restrict__ cellid,
                                                      const unsigned* __restrict__ rank,
                                                      const unsigned* __restrict__ cell_start, int n,
                                                      float4* __restrict__ sorted) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    // far-away finite records behind the array: what idle lanes of the candidate stream read (PAD_CANDS entries)
    if (k < PAD_CANDS) sorted[n + k] = make_float4(-1.0e18f, -1.0e18f, -1.0e18f, 0.f);
    if (k >= n) return;
    unsigned c = cellid[k];
    if (c == DROPPED) return;
    sorted[cell_start[c] + rank[k]] = in4[k];
}

// ---- exclusive scan of u32 (cell populations) : tile sums -> tile offsets -> apply ----------
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ unsigned block_exclusive_scan(unsigned v, unsigned* total, unsigned* smem) {
    unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    unsigned inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (unsigned)o) inc += t;
    }
    if (lane == 31) smem[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        unsigned w = lane < (SCAN_THREADS / 32) ? smem[lane] : 0;
        unsigned wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= (unsigned)o) wi += t;
        }
        if (lane < (SCAN_THREADS / 32)) smem[lane] = wi - w;
        if (lane == 31) smem[32] = wi;
    }
    __syncthreads();
    unsigned res = smem[wid] + inc - v;
    *total = smem[32];
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums_kernel(const unsigned* __restrict__ in, int n,
                                                                      unsigned* __restrict__ tile_sums) {
    __shared__ unsigned smem[33];
    int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i)
        if (base + i < n) s += in[base + i];
    unsigned total;
    block_exclusive_scan(s, &total, smem);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}
__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_offsets_kernel(unsigned* __restrict__ tile_sums,
                                                                         int ntiles) {
    __shared__ unsigned smem[33];
    unsigned carry = 0;
    for (int b = 0; b < ntiles; b += SCAN_THREADS) {
        int i = b + threadIdx.x;
        unsigned v = i < ntiles ? tile_sums[i] : 0;
        unsigned total;
        unsigned ex = block_exclusive_scan(v, &total, smem);
        if (i < ntiles) tile_sums[i] = carry + ex;
        carry += total;
    }
}
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const unsigned* __restrict__ in, int n,
                                                                  const unsigned* __restrict__ tile_off,
                                                                  unsigned* __restrict__ out /* n+1 */) {
    __shared__ unsigned smem[33];
    int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    unsigned v[SCAN_ITEMS];
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = base + i < n ? in[base + i] : 0;
        s += v[i];
    }
    unsigned total;
    unsigned ex = block_exclusive_scan(s, &total, smem) + tile_off[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) out[base + i] = ex;
        ex += v[i];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == SCAN_THREADS - 1) out[n] = ex;
}

int exclusive_scan_u32(Ctx* c, const unsigned* in, int n, unsigned* out) {
    int ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    MB_TRY(c->scan_tmp.reserve((size_t)ntiles * sizeof(unsigned)));
    unsigned* ts = c->scan_tmp.as<unsigned>();
    scan_tile_sums_kernel<<<ntiles, SCAN_THREADS, 0, c->stream>>>(in, n, ts);
    scan_tile_offsets_kernel<<<1, SCAN_THREADS, 0, c->stream>>>(ts, ntiles);
    scan_apply_kernel<<<ntiles, SCAN_THREADS, 0, c->stream>>>(in, n, ts, out);
    c->launches += 3;
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

// ---------------------------------------------------------------------------------------------
// pair emission: per-warp staging in shared memory, flushed with ONE global atomic per ~1k pairs
// and fully coalesced 8-byte stores.
// ---------------------------------------------------------------------------------------------
// explicit shared-window accesses (32-bit addresses): generic pointers to shared memory cost the
// compiler a window-base computation at every use site
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sts64(unsigned addr, unsigned a, unsigned b) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void sts32f(unsigned addr, float a) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(a) : "memory");
}
__device__ __forceinline__ uint2 lds64(unsigned addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ unsigned lds32(unsigned addr) {
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ float lds32f(unsigned addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}

// stage_sa / staged_sa: shared-window addresses of this warp's pair / distance staging buffers.
// MODE 0: staged entries are final (home id, candidate id) pairs, so a flush is a plain copy (one LDS.64 + one STG.64
// per 32 pairs, straight-line predicated chunks) and the buffer survives home batches and tiles.
// MODE 1: a staged entry is (j, candidate id) with j the slot of the home atom in ws.home: the home id is looked up
// here, 32 pairs per instruction — so the buffer must be flushed before the home batch changes.
template <int Q>
__device__ __forceinline__ void flush_chunk_copy(int r, unsigned sa, uint2* d) {
    asm volatile(
        "{ .reg .pred p; .reg .b32 a, b;\n"
        "  setp.gt.s32 p, %0, %1;\n"
        "  @p ld.shared.v2.u32 {a, b}, [%2+%3];\n"
        "  @p st.global.v2.u32 [%4+%3], {a, b};\n"
        "}"
        :
        : "r"(r), "n"(Q * 32), "r"(sa), "n"(Q * 256), "l"(d)
        : "memory");
}

template <int MODE>
__device__ __forceinline__ void warp_flush(unsigned stage_sa, unsigned staged_sa, unsigned home_sa, int& stage_n,
                                           const SearchParams& P, unsigned lane) {
    constexpr bool DIST = MODE == 1;
    static_assert(STAGE_CAP % 128 == 0 && STAGE_CAP >= 512, "a pass of 8 homes x 64 (or 4 x 128) candidates must fit; MODE 1 flushes rounds of 128");
    __syncwarp();
    if (stage_n == 0) return;
    const int n = stage_n;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(P.counter, (unsigned long long)n);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base + (unsigned long long)n <= P.pair_cap) {
        if (MODE == 0) {
            // four chunks per round of a real loop (fully unrolled, ptxas hoists all sixteen loads and spills)
            uint2* d = P.pairs + base + lane;
            unsigned sa = stage_sa + lane * 8u;
            int r = n - (int)lane;  // entry q*32 + lane exists iff r > q*32
            asm volatile("" : "+r"(sa), "+r"(r));
#pragma unroll 1
            for (int q = 0; q < n; q += 128) {
                flush_chunk_copy<0>(r, sa, d);
                flush_chunk_copy<1>(r, sa, d);
                flush_chunk_copy<2>(r, sa, d);
                flush_chunk_copy<3>(r, sa, d);
                r -= 128;
                sa += 1024u;
                d += 128;
            }
        } else {
            uint2* d = P.pairs + base + lane;
            unsigned sa = stage_sa + lane * 8u;
            int r = n - (int)lane;  // entry q*32 + lane exists iff r > q*32
            asm volatile("" : "+r"(sa), "+r"(r));  // keep both in registers (ptxas otherwise re-derives them from %tid per chunk)
#pragma unroll
            for (int q = 0; q < STAGE_CAP / 32; ++q) {
                if (r > q * 32) {
                    uint2 v = lds64(sa + (unsigned)q * 256u);  // (home slot, candidate id)
                    v.x = lds32(home_sa + 16u * v.x + 12u);
                    d[q * 32] = v;
                }
            }
        }
        if (DIST) {
            float* dd = P.dists + base + lane;
            const unsigned sd = staged_sa + lane * 4u;
            const int r = n - (int)lane;
#pragma unroll
            for (int q = 0; q < STAGE_CAP / 32; ++q)
                if (r > q * 32) dd[q * 32] = lds32f(sd + (unsigned)q * 128u);
        }
    }
    __syncwarp();
    stage_n = 0;
}

// MODE 0 flush through the async proxy: the staged pairs are final, so the whole buffer leaves with ONE bulk copy
// (cp.async.bulk shared -> global, SASS UBLKCP) issued by lane 0 instead of an LDS.64 + STG.64 per 32 pairs.  Bulk
// copies move multiples of 16 bytes between 16-byte aligned addresses, so every flush moves an EVEN number of pairs
// (the global pair counter stays even) and an odd last entry is carried over to slot 0.  The entries written by the
// other lanes reach the async proxy through fence.proxy.async + __syncwarp; the buffer is reused after
// wait_group.read.  Returns the number of entries left in the buffer (0 or 1).
__device__ __forceinline__ int warp_flush_bulk(unsigned stage_sa, int n, const SearchParams& P, unsigned lane) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    const int ne = n & ~1;
    if (lane == 0) {
        uint2 carry = make_uint2(0u, 0u);
        if (n & 1) carry = lds64(stage_sa + 8u * (unsigned)(n - 1));
        if (ne) {
            const unsigned long long base = atomicAdd(P.counter, (unsigned long long)ne);
            if (base + (unsigned long long)ne <= P.pair_cap) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(P.pairs + base),
                             "r"(stage_sa), "r"(ne * 8)
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
        }
        if (n & 1) sts64(stage_sa, carry.x, carry.y);
    }
    __syncwarp();
    return n & 1;
}

// The odd pairs the warps were left with at the end of a MODE 0 search sit behind pair_cap (slots [pair_cap,
// pair_cap + counter[3])): append them to the list and fold their number into the pair count.
__global__ void merge_pair_tail_kernel(uint2* __restrict__ pairs, unsigned long long pair_cap,
                                       unsigned long long* __restrict__ counter) {
    const unsigned long long m = counter[0], t = counter[3];
    if (m + t <= pair_cap)
        for (unsigned long long k = threadIdx.x; k < t; k += blockDim.x) pairs[m + k] = pairs[pair_cap + k];
    __syncthreads();
    if (threadIdx.x == 0) {
        counter[0] = m + t;
        counter[3] = 0;
    }
}

// ---- packed f32x2 helpers (sm_100: FADD2 / FMUL2 process two neighbour atoms per lane) ----------
// sub.rn / mul.rn are IEEE per element, so each half rounds exactly like the scalar op.  The sums
// of squares stay SCALAR __fadd_rn: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (even
// with -fmad=false), which would break bit-exactness; it never contracts into a scalar FADD.
__device__ __forceinline__ unsigned long long pk2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
// same, but never rematerialised: used for the loop-invariant neighbour coordinates (ptxas otherwise
// re-packs the register pair in front of every FADD2 of the inner loop)
__device__ __forceinline__ unsigned long long pk2_once(float lo, float hi) {
    unsigned long long r;
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(unsigned long long u, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(u));
}
__device__ __forceinline__ unsigned long long sub2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ int bfind32(unsigned m) {
    int r;
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(m));
    return r;
}
// squared distances of two neighbours (packed) to one home atom: (dx*dx + dy*dy) + dz*dz, unfused
__device__ __forceinline__ void d2_pair(unsigned long long nx, unsigned long long ny, unsigned long long nz,
                                        const float4& h, float& d0, float& d1) {
    unsigned long long dx = sub2(nx, pk2(h.x, h.x)), dy = sub2(ny, pk2(h.y, h.y)), dz = sub2(nz, pk2(h.z, h.z));
    unsigned long long xx = mul2(dx, dx), yy = mul2(dy, dy), zz = mul2(dz, dz);
    float x0, x1, y0, y1, z0, z1;
    upk2(xx, x0, x1);
    upk2(yy, y0, y1);
    upk2(zz, z0, z1);
    d0 = xadd(xadd(x0, y0), z0);
    d1 = xadd(xadd(x1, y1), z1);
}

__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
// rc2 - ((dx*dx + dy*dy) + dz*dz) for two neighbours, every operation rounded on its own, all packed.
// The sums are fma(s, 1.0, t) = round(s + t): the same number as an add, but a PACKED add fed by a packed
// multiply is contracted by ptxas into FFMA2 (one rounding instead of two), while a fused multiply-add whose
// multiplier is a run-time 1.0 is left alone.  The sign bit of each half answers d2 <= rc2 (no NaNs reach
// this: non-finite atoms are dropped at binning and the padding is finite): clear = within the cutoff.
__device__ __forceinline__ unsigned long long rc2_minus_d2(unsigned long long nx, unsigned long long ny,
                                                           unsigned long long nz, const float4& h,
                                                           unsigned long long one2, unsigned long long rc22) {
    const unsigned long long dx = sub2(nx, pk2(h.x, h.x)), dy = sub2(ny, pk2(h.y, h.y)), dz = sub2(nz, pk2(h.z, h.z));
    const unsigned long long xx = mul2(dx, dx), yy = mul2(dy, dy), zz = mul2(dz, dz);
    const unsigned long long s = fma2(fma2(xx, one2, yy), one2, zz);
    return sub2(rc22, s);
}

// Filter form of the same test: d2f - rc2 with d2f accumulated by three fused multiply-adds that start from -rc2
// (6 packed instructions instead of 9).  d2f is NOT the reference's number — the reference rounds every product and every
// sum — but both approximate S = dx^2 + dy^2 + dz^2 of the SAME f32 differences: |d2_ref - S| <= 3.01 u S and
// |t - (S - rc2)| <= 2 u max(S, rc2) + u |t| (u = 2^-24), so |t| > 10.1 u rc2 implies sign(t) = sign(d2_ref - rc2)
// (S > 2 rc2 is a miss for both whatever t).  The caller keeps min |t| of a step and re-runs the step with
// rc2_minus_d2 when it is below band = 16 u rc2 (SearchParams::band); a few thousand tests per 10^9 are.
__device__ __forceinline__ unsigned long long d2f_minus_rc2(unsigned long long nx, unsigned long long ny,
                                                            unsigned long long nz, const float4& h,
                                                            unsigned long long nrc22) {
    const unsigned long long dx = sub2(nx, pk2(h.x, h.x)), dy = sub2(ny, pk2(h.y, h.y)), dz = sub2(nz, pk2(h.z, h.z));
    return fma2(dz, dz, fma2(dy, dy, fma2(dx, dx, nrc22)));
}
// running minimum of |t| over the tests of a step (one FMNMX3 with |.| operand modifiers)
__device__ __forceinline__ float min3abs(float m, float a, float b) {
    float r;
    asm("{ .reg .f32 x, y;\n  abs.f32 x, %2;\n  abs.f32 y, %3;\n  min.f32 %0, %1, x, y; }" : "=f"(r) : "f"(m), "f"(a), "f"(b));
    return r;
}

// vdW variant: c2 - ((dx*dx + dy*dy) + dz*dz) with c2 = the pair's own squared cutoff (already rounded).  The final
// subtraction is fma(s, -1.0, c2) with a run-time -1.0: one rounding of the exact difference, and nothing ptxas could
// contract the product c*c into.
__device__ __forceinline__ unsigned long long cut2_minus_d2(unsigned long long nx, unsigned long long ny,
                                                            unsigned long long nz, const float4& h,
                                                            unsigned long long one2, unsigned long long mone2,
                                                            unsigned long long c2) {
    const unsigned long long dx = sub2(nx, pk2(h.x, h.x)), dy = sub2(ny, pk2(h.y, h.y)), dz = sub2(nz, pk2(h.z, h.z));
    const unsigned long long xx = mul2(dx, dx), yy = mul2(dy, dy), zz = mul2(dz, dz);
    const unsigned long long s = fma2(fma2(xx, one2, yy), one2, zz);
    return fma2(s, mone2, c2);
}

// Out-of-line copy of the exact periodic distance for the rare paths of the cell kernel (band
// resolution, distance output of wrapped pairs): keeps the hot loop inside the instruction cache.
__device__ __noinline__ float d2_pbc_call(const DevBox& bx, float ax, float ay, float az, float bxx, float byy,
                                          float bzz, unsigned w) {
    return d2_pbc(bx, ax, ay, az, bxx, byy, bzz, w);
}

// x / k for the small subdivision factors (k <= 8): magic = floor(2^32 / k) + 1, exact for x < 2^32 / k
__device__ __forceinline__ int div_k(int x, int k, unsigned magic) { return k == 1 ? x : (int)__umulhi((unsigned)x, magic); }

// Is reference cell `cn` MASK-adjacent to `ch` along one dim, and is that adjacency a wrapped one?
// Valid when every periodic dim has >= 3 reference cells (the kernel's precondition), where each
// unordered adjacent cell pair appears exactly once in search_plan with a unique wrapped flag.
__device__ __forceinline__ bool ref_adjacent(int ch, int cn, int dim, bool periodic, unsigned& wbit,
                                             unsigned& sbit) {
    int diff = abs(cn - ch);
    wbit = 0;
    sbit = cn > ch ? 1u : 0u;  // wrapped: is the neighbour cell the one at the high end?
    if (diff <= 1) return true;
    if (periodic && diff == dim - 1) {
        wbit = 1;
        return true;
    }
    return false;
}

// Runs of one neighbour row (dy,dz,[dxlo,dxhi]) of home fine cell (fx,fy,fz): contiguous x-ranges of
// fine cells that lie in reference cells MASK-adjacent to the home reference cell, each with a
// uniform wrapped-dims flag.  emit(first_cell, last_cell, flag) with flag = w | sgn << 3.
// y / z part of a neighbour row: wrapped fine coordinates, MASK adjacency of the reference cells and the wrapped /
// sign flags of the two dims.  Returns false when the row does not exist for this home tile.
__device__ __forceinline__ bool row_yz(const GridSpec& g, int fy, int fz, int cy, int cz, NbrRow row, int& row_base,
                                       unsigned& fyz) {
    const int fdy = g.fd[1], fdz = g.fd[2];
    const bool pery = g.pbc & 2u, perz = g.pbc & 4u;
    int ny = fy + row.dy, nz = fz + row.dz;
    if (ny < 0) { if (!pery) return false; ny += fdy; } else if (ny >= fdy) { if (!pery) return false; ny -= fdy; }
    if (nz < 0) { if (!perz) return false; nz += fdz; } else if (nz >= fdz) { if (!perz) return false; nz -= fdz; }
    unsigned wy, wz, sy, sz;
    if (!ref_adjacent(cy, div_k(ny, g.k[1], g.kmagic[1]), g.dims[1], pery, wy, sy)) return false;
    if (!ref_adjacent(cz, div_k(nz, g.k[2], g.kmagic[2]), g.dims[2], perz, wz, sz)) return false;
    row_base = (nz * fdy + ny) * g.fd[0];
    const unsigned w = (wy << 1) | (wz << 2);
    fyz = w | ((((sy << 1) | (sz << 2)) & w) << 3);
    return true;
}

template <class F>
__device__ __forceinline__ void gen_row_runs(const GridSpec& g, int fx, int fy, int fz, int cx, int cy, int cz,
                                             NbrRow row, F&& emit) {
    const int fdx = g.fd[0], fdy = g.fd[1], fdz = g.fd[2];
    const bool perx = g.pbc & 1u, pery = g.pbc & 2u, perz = g.pbc & 4u;
    int ny = fy + row.dy, nz = fz + row.dz;
    if (ny < 0) { if (!pery) return; ny += fdy; } else if (ny >= fdy) { if (!pery) return; ny -= fdy; }
    if (nz < 0) { if (!perz) return; nz += fdz; } else if (nz >= fdz) { if (!perz) return; nz -= fdz; }
    unsigned wy, wz, sy, sz;
    if (!ref_adjacent(cy, div_k(ny, g.k[1], g.kmagic[1]), g.dims[1], pery, wy, sy)) return;
    if (!ref_adjacent(cz, div_k(nz, g.k[2], g.kmagic[2]), g.dims[2], perz, wz, sz)) return;
    const int row_base = (nz * fdy + ny) * fdx;
    // x range, possibly split by the periodic boundary into <= 2 raw segments
    int xa = fx + row.dxlo, xb = fx + row.dxhi;
    int seg_lo0 = 0, seg_hi0 = -1, seg_lo1 = 0, seg_hi1 = -1;
    if (xa < 0) {
        if (perx) { seg_lo0 = xa + fdx; seg_hi0 = min(xb, -1) + fdx; }
        xa = 0;
    }
    if (xb >= fdx) {
        if (perx) { seg_lo1 = max(xa, fdx) - fdx; seg_hi1 = xb - fdx; }
        xb = fdx - 1;
    }
    const int kx = g.k[0];
#pragma unroll 1
    for (int sg = 0; sg < 3; ++sg) {
        const int lo = sg == 0 ? seg_lo0 : (sg == 1 ? seg_lo1 : xa);
        const int hi = sg == 0 ? seg_hi0 : (sg == 1 ? seg_hi1 : xb);
        if (lo > hi) continue;
        int run_lo = -1, run_hi = -1;
        unsigned run_f = 0;
        const int c_first = div_k(lo, kx, g.kmagic[0]), c_last = div_k(hi, kx, g.kmagic[0]);
#pragma unroll 1
        for (int cxn = c_first; cxn <= c_last + 1; ++cxn) {  // one sentinel iteration flushes the last run
            unsigned wx = 0, sx = 0;
            const bool adj = cxn <= c_last && ref_adjacent(cx, cxn, g.dims[0], perx, wx, sx);
            const int a = max(lo, cxn * kx), b = min(hi, cxn * kx + kx - 1);
            const unsigned w = wx | (wy << 1) | (wz << 2);
            const unsigned f = w | ((((sx | (sy << 1) | (sz << 2))) & w) << 3);
            if (adj && run_lo >= 0 && f == run_f) {
                run_hi = b;
                continue;
            }
            if (run_lo >= 0) emit(row_base + run_lo, row_base + run_hi, run_f);
            run_lo = -1;
            if (adj) {
                run_lo = a;
                run_hi = b;
                run_f = f;
            }
        }
    }
}

constexpr int MAX_RUNS = 224;        // >= 1 + 32 rows * 6 runs
constexpr int RUN_BITMAP_WORDS = 96;  // 3072 stream positions: the run of a candidate is found by a popcount
constexpr unsigned RUN_SELF = 0x40u;  // flag bit: the home cell itself (pairs counted once by index order)

// per-warp shared memory
struct __align__(16) WarpShared {
    float4 home[32];                     // home batch (position + id bits), padded with far-away finite points
    // One record per run of the candidate stream: x = first atom (index into the sorted array) MINUS the run's stream
    // position, so that a candidate's address is x + its stream position; y = flags (w | sgn << 3 | RUN_SELF) in the
    // low byte, the run's stream position above it.  Entry nr is the sentinel behind the last run: it maps stream
    // positions >= T onto the padding records behind the sorted array.
    uint2 rec[MAX_RUNS + 1];
    unsigned rbits[RUN_BITMAP_WORDS + 2];  // bit p set <=> a run starts at stream position p (p < 32 * RUN_BITMAP_WORDS)
    unsigned hid[32];                    // ids of the home batch, compact: the emission loop reads them as broadcast words
    float hvdw[32];                      // vdW search: radii of the home batch
};

// Phase A of the pair kernel: lanes work on different neighbour rows of the home tile and append their runs
// (contiguous ranges of the sorted candidate array) to the warp's run table, direct runs of a batch of 32 rows
// first, wrapped runs after them; `first` adds the self run in front.  Fills rows from row0 on until all rows are
// in or the table (MAXR entries) is full; returns the number of runs (nr) and of candidates (T) and advances row0.
template <int MAXR>
__device__ __forceinline__ void fill_run_table(const SearchParams& P, uint2* __restrict__ rec, int fx, int fy, int fz,
                                               int cx, int cy, int cz, unsigned hs, unsigned he, bool first,
                                               unsigned lane, int& row0, unsigned& nr, unsigned& T,
                                               unsigned* __restrict__ rbits) {
    const GridSpec& g = P.g;
    nr = 0;
    T = 0;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < RUN_BITMAP_WORDS / 32; ++k) rbits[k * 32 + lane] = 0u;
    if (lane < 2) rbits[RUN_BITMAP_WORDS + lane] = 0u;
    __syncwarp();
    auto mark = [&](unsigned pos) {
        if (pos < 32u * RUN_BITMAP_WORDS) atomicOr(&rbits[pos >> 5], 1u << (pos & 31u));
    };
    if (first && !P.two_sets) {
        if (lane == 0) {
            rec[0] = make_uint2(hs, RUN_SELF);
            mark(0u);
        }
        nr = 1;
        T = he - hs;
    }
    // Tiles whose whole x reach stays inside the grid AND inside the reference cells next to the home one (most of
    // them): every row is ONE run along x with no x flags, so the general run generator (periodic split, one
    // adjacency decision per reference cell) is not needed — two cell_start loads per row.
    const bool fast_x = fx + P.dx_min >= 0 && fx + P.dx_max < g.fd[0] &&
                        div_k(fx + P.dx_min, g.k[0], g.kmagic[0]) >= cx - 1 &&
                        div_k(fx + P.dx_max, g.k[0], g.kmagic[0]) <= cx + 1;
#pragma unroll 1
    while (row0 < P.nrows) {
        const int ri = row0 + (int)lane;
        NbrRow row = P.rows[min(ri, P.nrows - 1)];
        // pass 1: count runs and atoms per class (direct / wrapped)
        unsigned nd = 0, nw = 0, ld = 0, lw = 0;
        unsigned fs = 0, flen = 0, ff = 0;  // fast path: the row's single run
        if (fast_x) {
            int row_base;
            if (ri < P.nrows && row_yz(g, fy, fz, cy, cz, row, row_base, ff)) {
                fs = P.cell_startB[row_base + fx + row.dxlo];
                flen = P.cell_startB[row_base + fx + row.dxhi + 1] - fs;
                if (flen) {
                    if (ff & 7u) { nw = 1; lw = flen; } else { nd = 1; ld = flen; }
                }
            }
        } else if (ri < P.nrows)
            gen_row_runs(g, fx, fy, fz, cx, cy, cz, row, [&](int c0, int c1, unsigned f) {
                unsigned len = P.cell_startB[c1 + 1] - P.cell_startB[c0];
                if (len) {
                    if (f & 7u) { ++nw; lw += len; } else { ++nd; ld += len; }
                }
            });
        // warp scans: counts packed (direct low 16, wrapped high 16), lengths separately
        unsigned cn = nd | (nw << 16), cni = cn, ldi = ld, lwi = lw;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned t0 = __shfl_up_sync(0xffffffffu, cni, o), t1 = __shfl_up_sync(0xffffffffu, ldi, o),
                     t2 = __shfl_up_sync(0xffffffffu, lwi, o);
            if (lane >= (unsigned)o) { cni += t0; ldi += t1; lwi += t2; }
        }
        const unsigned tot_c = __shfl_sync(0xffffffffu, cni, 31);
        const unsigned tot_d = tot_c & 0xffffu, tot_w = tot_c >> 16;
        const unsigned tot_ld = __shfl_sync(0xffffffffu, ldi, 31), tot_lw = __shfl_sync(0xffffffffu, lwi, 31);
        if (nr + tot_d + tot_w > (unsigned)MAXR) break;  // table full: process it first
        // pass 2: write this lane's runs (direct runs of the batch first, then wrapped ones)
        unsigned sd = nr + (cni & 0xffffu) - nd, sw = nr + tot_d + (cni >> 16) - nw;
        unsigned pd = T + ldi - ld, pw = T + tot_ld + lwi - lw;
        if (fast_x) {
            if (flen) {
                const unsigned at = (ff & 7u) ? sw : sd, pos = (ff & 7u) ? pw : pd;
                rec[at] = make_uint2(fs - pos, ff | (pos << 8));
                mark(pos);
            }
        } else if (ri < P.nrows)
            gen_row_runs(g, fx, fy, fz, cx, cy, cz, row, [&](int c0, int c1, unsigned f) {
                unsigned s = P.cell_startB[c0], len = P.cell_startB[c1 + 1] - s;
                if (len) {
                    if (f & 7u) {
                        rec[sw] = make_uint2(s - pw, f | (pw << 8));
                        mark(pw);
                        ++sw; pw += len;
                    } else {
                        rec[sd] = make_uint2(s - pd, f | (pd << 8));
                        mark(pd);
                        ++sd; pd += len;
                    }
                }
            });
        nr += tot_d + tot_w;
        T += tot_ld + tot_lw;
        row0 += 32;
    }
    if (lane == 0) {
        rec[nr] = make_uint2((unsigned)P.n_sortedB - T, T << 8);  // sentinel: positions >= T read the padding records
        mark(T);
    }
    __syncwarp();
}

// inclusive warp scan step: v += (value of lane - o) for lanes >= o; the shuffle's own predicate says whether the
// source lane exists, so no lane compare is needed
__device__ __forceinline__ void scan_step(int& v, int o) {
    asm volatile(
        "{ .reg .pred p; .reg .b32 t;\n"
        "  shfl.sync.up.b32 t|p, %0, %1, 0, 0xffffffff;\n"
        "  @p add.s32 %0, %0, t; }"
        : "+r"(v)
        : "r"(o));
}

// Emission of one group of four home slots (bits 0-3 of the masks): the candidates of this lane that hit them are staged
// as final (home id, candidate id) pairs at the lane's running position.  Per (home, candidate half): LOP3 (bit -> value
// and predicate), one predicated STS.64 and the position bump as (bit << (3 - J)) + position — no predicated add, no bit
// search — plus the MOV that puts the home id next to the candidate id.  (Two STS.32 instead of MOV + STS.64 issue
// fewer instructions but run into the shared-memory store rate: 1.41 ms against 1.29 ms, mio_throttle 1.6.)
#define MB_EMIT_ONE(J, E, ID, H)                                                     \
    "  and.b32 t, " E ", " #J ";\n  setp.ne.u32 p, t, 0;\n"                          \
    "  @p st.shared.v2.u32 [%0], {" H ", " ID "};\n"
__device__ __forceinline__ void emit_group(unsigned e0, unsigned e1, unsigned id0, unsigned id1, unsigned& sp, uint4 h) {
    asm volatile(
        "{ .reg .pred p; .reg .b32 t;\n"
        MB_EMIT_ONE(1, "%1", "%3", "%5") "  shl.b32 t, t, 3;\n  add.u32 %0, %0, t;\n"
        MB_EMIT_ONE(1, "%2", "%4", "%5") "  shl.b32 t, t, 3;\n  add.u32 %0, %0, t;\n"
        MB_EMIT_ONE(2, "%1", "%3", "%6") "  shl.b32 t, t, 2;\n  add.u32 %0, %0, t;\n"
        MB_EMIT_ONE(2, "%2", "%4", "%6") "  shl.b32 t, t, 2;\n  add.u32 %0, %0, t;\n"
        MB_EMIT_ONE(4, "%1", "%3", "%7") "  shl.b32 t, t, 1;\n  add.u32 %0, %0, t;\n"
        MB_EMIT_ONE(4, "%2", "%4", "%7") "  shl.b32 t, t, 1;\n  add.u32 %0, %0, t;\n"
        MB_EMIT_ONE(8, "%1", "%3", "%8") "  add.u32 %0, %0, t;\n"
        MB_EMIT_ONE(8, "%2", "%4", "%8") "  add.u32 %0, %0, t;\n"
        "}"
        : "+r"(sp)
        : "r"(e0), "r"(e1), "r"(id0), "r"(id1), "r"(h.x), "r"(h.y), "r"(h.z), "r"(h.w)
        : "memory");
}
// Same for a fused double step (four candidates per lane): a lane's entries are home-major, candidates 0..3 within a home.
#define MB_EMIT_Q(J, SH, H)                                                                        \
    MB_EMIT_ONE(J, "%1", "%5", H) SH MB_EMIT_ONE(J, "%2", "%6", H) SH MB_EMIT_ONE(J, "%3", "%7", H) SH \
        MB_EMIT_ONE(J, "%4", "%8", H) SH
template <int NHOME>  // homes of the group that exist (the last group of a batch may hold fewer than four)
__device__ __forceinline__ void emit_group4(unsigned e0, unsigned e1, unsigned e2, unsigned e3, unsigned id0, unsigned id1,
                                            unsigned id2, unsigned id3, unsigned& sp, uint4 h) {
    if (NHOME == 4)
        asm volatile(
            "{ .reg .pred p; .reg .b32 t;\n"
            MB_EMIT_Q(1, "  shl.b32 t, t, 3;\n  add.u32 %0, %0, t;\n", "%9")
            MB_EMIT_Q(2, "  shl.b32 t, t, 2;\n  add.u32 %0, %0, t;\n", "%10")
            MB_EMIT_Q(4, "  shl.b32 t, t, 1;\n  add.u32 %0, %0, t;\n", "%11")
            MB_EMIT_Q(8, "  add.u32 %0, %0, t;\n", "%12")
            "}"
            : "+r"(sp)
            : "r"(e0), "r"(e1), "r"(e2), "r"(e3), "r"(id0), "r"(id1), "r"(id2), "r"(id3), "r"(h.x), "r"(h.y), "r"(h.z), "r"(h.w)
            : "memory");
    else if (NHOME == 3)
        asm volatile(
            "{ .reg .pred p; .reg .b32 t;\n"
            MB_EMIT_Q(1, "  shl.b32 t, t, 3;\n  add.u32 %0, %0, t;\n", "%9")
            MB_EMIT_Q(2, "  shl.b32 t, t, 2;\n  add.u32 %0, %0, t;\n", "%10")
            MB_EMIT_Q(4, "  shl.b32 t, t, 1;\n  add.u32 %0, %0, t;\n", "%11")
            "}"
            : "+r"(sp)
            : "r"(e0), "r"(e1), "r"(e2), "r"(e3), "r"(id0), "r"(id1), "r"(id2), "r"(id3), "r"(h.x), "r"(h.y), "r"(h.z), "r"(h.w)
            : "memory");
    else if (NHOME == 2)
        asm volatile(
            "{ .reg .pred p; .reg .b32 t;\n"
            MB_EMIT_Q(1, "  shl.b32 t, t, 3;\n  add.u32 %0, %0, t;\n", "%9")
            MB_EMIT_Q(2, "  shl.b32 t, t, 2;\n  add.u32 %0, %0, t;\n", "%10")
            "}"
            : "+r"(sp)
            : "r"(e0), "r"(e1), "r"(e2), "r"(e3), "r"(id0), "r"(id1), "r"(id2), "r"(id3), "r"(h.x), "r"(h.y), "r"(h.z), "r"(h.w)
            : "memory");
    else
        asm volatile(
            "{ .reg .pred p; .reg .b32 t;\n"
            MB_EMIT_Q(1, "  shl.b32 t, t, 3;\n  add.u32 %0, %0, t;\n", "%9")
            "}"
            : "+r"(sp)
            : "r"(e0), "r"(e1), "r"(e2), "r"(e3), "r"(id0), "r"(id1), "r"(id2), "r"(id3), "r"(h.x), "r"(h.y), "r"(h.z), "r"(h.w)
            : "memory");
}
#undef MB_EMIT_Q
#undef MB_EMIT_ONE
__device__ __forceinline__ uint4 lds128u(unsigned addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}

// ---- neighbour-list modes: per-home-atom hit counters kept BIT-SLICED in every lane ---------------------------------
// Lane l looks at its own candidates; bit j of plane k is bit k of "how many of my candidates hit home slot j".  Adding
// the (up to four) hit masks of a step is a carry-save compression to a 3-bit sliced number and one ripple through the
// seven planes — about 18 LOP3 for 128 tests x 32 homes, independent of the number of home atoms.  Seven planes hold
// 127: the planes are folded into the per-home totals (lane j keeps home j's) every 31 steps.
constexpr int VC_PLANES = 7;
constexpr unsigned VC_MAX_STEPS = 31;  // 31 steps x <= 4 hits per step and lane <= 127
__device__ __forceinline__ void vcount_add(unsigned (&p)[VC_PLANES], unsigned b1, unsigned b2, unsigned b4) {
    unsigned cy = p[0] & b1;
    p[0] ^= b1;
    unsigned t = p[1] ^ b2 ^ cy;
    cy = (p[1] & b2) | (p[1] & cy) | (b2 & cy);
    p[1] = t;
    t = p[2] ^ b4 ^ cy;
    cy = (p[2] & b4) | (p[2] & cy) | (b4 & cy);
    p[2] = t;
#pragma unroll
    for (int k = 3; k < VC_PLANES; ++k) {
        t = p[k] & cy;
        p[k] ^= cy;
        cy = t;
    }
}
__device__ __forceinline__ void vcount_add4(unsigned (&p)[VC_PLANES], unsigned m0, unsigned m1, unsigned m2, unsigned m3) {
    const unsigned s = m0 ^ m1 ^ m2, c = (m0 & m1) | (m0 & m2) | (m1 & m2);
    const unsigned s2 = s ^ m3, c2 = s & m3;
    vcount_add(p, s2, c ^ c2, c & c2);
}
__device__ __forceinline__ void vcount_fold(unsigned (&p)[VC_PLANES], unsigned& deg, int nh, unsigned lane) {
#pragma unroll 1
    for (int j = 0; j < nh; ++j) {
        unsigned v = 0;
#pragma unroll
        for (int k = 0; k < VC_PLANES; ++k) v |= ((p[k] >> j) & 1u) << k;
        v = __reduce_add_sync(0xffffffffu, v);
        if (lane == (unsigned)j) deg += v;
    }
#pragma unroll
    for (int k = 0; k < VC_PLANES; ++k) p[k] = 0u;
}
// the bit of home slot (candidate index - first home index) when the candidate IS one of the home atoms
__device__ __forceinline__ unsigned self_slot_bit(unsigned cand_index, unsigned hb) {
    const unsigned d = cand_index - hb;
    return d < 32u ? (1u << d) : 0u;
}
// MODE 5: the hits of home slot j (one ballot per candidate quarter) go to the atom's row at the running cursor, in
// lane order; lane j owns the cursor of home slot j
template <int NQ>
__device__ __forceinline__ void rows_emit(const unsigned (&m)[NQ], const unsigned (&id)[NQ], unsigned& cur, unsigned lane,
                                          unsigned* __restrict__ cols) {
    unsigned any = m[0];
#pragma unroll
    for (int q = 1; q < NQ; ++q) any |= m[q];
    any = __reduce_or_sync(0xffffffffu, any);
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll 1
    while (any) {
        const int j = __ffs(any) - 1;
        any &= any - 1u;
        unsigned base = __shfl_sync(0xffffffffu, cur, j);
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const bool hit = (m[q] >> j) & 1u;
            const unsigned b = __ballot_sync(0xffffffffu, hit);
            if (hit) cols[base + __popc(b & lt)] = id[q];
            base += __popc(b);
        }
        if (lane == (unsigned)j) cur = base;
    }
}

// One warp per home tile = hx consecutive fine cells along x (dynamic work counter).
//  Phase A  lanes work on different neighbour rows in parallel and build a table of runs: contiguous
//           ranges of the sorted atom array (self cell, then direct runs, then wrapped runs).
//  Phase B  the concatenation of the runs is one stream of candidate atoms; it is consumed 64 atoms
//           per step (two per lane, two coalesced float4 loads), tested against the home atoms that
//           are broadcast from shared memory (one LDS.128 per home atom), hits are kept as per-lane
//           bit masks and expanded into the staging buffer afterwards.
// MODE: 0 pairs, 1 pairs + distances, 2 count only, 3 `within` flags (two sets), 4 / 5 neighbour list of ONE set over the
// full shell (two_sets = 1 with sortedB == sorted): 4 counts the neighbours of every atom, 5 writes the rows
//
// Staging entries.  MODE 1: (home slot j, candidate id).  MODE 0: (remaining hit bits of the candidate, candidate id):
// the home slot is the LOWEST set bit, decoded in the flush 32 entries per instruction — the expansion loop then
// needs no bit search at all: store, clear the lowest bit (x & (x - 1)), repeat, both candidates of a lane in the
// same iteration.
// VDW: the cutoff of a pair is (vdw1[i] + vdw2[j]) + EPSILON instead of one number for all (two-set searches only).
template <int MODE, bool VDW = false>
__global__ void __launch_bounds__(SEARCH_WARPS * 32, MB_SEARCH_MIN_CTAS) search_cells_kernel(const __grid_constant__ SearchParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    WarpShared& ws = reinterpret_cast<WarpShared*>(smem_raw)[wid];
    uint2* stage = (MODE == 0 || MODE == 1)
                       ? reinterpret_cast<uint2*>(smem_raw + SEARCH_WARPS * sizeof(WarpShared)) + wid * STAGE_ROOM
                       : nullptr;
    float* stage_d = MODE == 1 ? reinterpret_cast<float*>(smem_raw + SEARCH_WARPS * (sizeof(WarpShared) +
                                                                                    STAGE_ROOM * sizeof(uint2))) +
                                     wid * STAGE_CAP
                               : nullptr;
    const float4* __restrict__ home = ws.home;
    const unsigned home_sa = smem_addr(ws.home), rec_sa = smem_addr(ws.rec), rbits_sa = smem_addr(ws.rbits), hid_sa = smem_addr(ws.hid);
    const unsigned stage_sa = stage ? smem_addr(stage) : 0u, staged_sa = stage_d ? smem_addr(stage_d) : 0u;
    int stage_n = 0;
    unsigned long long count = 0;
    unsigned ntests = 0;  // MODE 2: distance tests this lane evaluated (FP32-pipe roofline of the count-only search)
    unsigned vplane[VC_PLANES] = {0u, 0u, 0u, 0u, 0u, 0u, 0u};  // MODE 4: bit-sliced hit counters of the home batch
    unsigned vsteps = 0;
    const GridSpec& g = P.g;
    const int fdx = g.fd[0], fdy = g.fd[1], fdz = g.fd[2];
    const int hx = g.hx, tdx = fdx / hx;  // tiles per x-row
    const unsigned ntiles = (unsigned)(tdx * fdy * fdz);
    const float rc2 = P.rc2;
    // padding of unused home slots / candidate lanes: finite and 2e18 apart, so a padded test is never a hit and
    // never produces a NaN (the direct path reads the SIGN of rc2 - d2)
    const float pad_home = 1.0e18f, pad_cand = -1.0e18f;
    const float finf = __int_as_float(0x7f800000);
    const unsigned le_mask = (2u << lane) - 1u;

    for (;;) {
        unsigned tile = 0;
        if (lane == 0) tile = (unsigned)atomicAdd(P.counter + 1, 1ull);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= ntiles) break;
        const int fx = (int)(tile % (unsigned)tdx) * hx, fy = (int)((tile / (unsigned)tdx) % (unsigned)fdy),
                  fz = (int)(tile / (unsigned)(tdx * fdy));
        const unsigned cell0 = (unsigned)(fx + fdx * (fy + fdy * fz));
        const unsigned hs = P.cell_start[cell0], he = P.cell_start[cell0 + hx];
        if (hs == he) continue;
        // the tile lies inside one reference cell (hx divides k[0])
        const int cx = div_k(fx, g.k[0], g.kmagic[0]), cy = div_k(fy, g.k[1], g.kmagic[1]),
                  cz = div_k(fz, g.k[2], g.kmagic[2]);

        int row0 = 0;       // next neighbour row to put into the table
        bool first = true;  // the first table starts with the self run
#pragma unroll 1
        do {
            // ---------------- Phase A: fill the run table ----------------
            unsigned nr, T;
            fill_run_table<MAX_RUNS>(P, ws.rec, fx, fy, fz, cx, cy, cz, hs, he, first, lane, row0, nr, T, ws.rbits);
            const bool use_bits = T < 32u * RUN_BITMAP_WORDS;  // every run start AND the sentinel are in the bitmap

            // ---------------- Phase B: consume the stream ----------------
#pragma unroll 1
            for (unsigned hb = hs; hb < he; hb += 32) {
                const int nh = min(32u, he - hb);
                __syncwarp();
                {
                    const float4 hrec = (lane < (unsigned)nh) ? __ldg(&P.sorted[hb + lane]) : make_float4(pad_home, pad_home, pad_home, 0.f);
                    ws.home[lane] = hrec;
                    if (MODE == 0) ws.hid[lane] = __float_as_uint(hrec.w);
                    if (VDW) ws.hvdw[lane] = (lane < (unsigned)nh) ? __ldg(&P.vdwA[__float_as_uint(hrec.w)]) : 0.f;
                }
                __syncwarp();
                const unsigned valid_slots = nh >= 32 ? 0xffffffffu : ((1u << nh) - 1u);
                unsigned cur0 = 0, cur1 = 0;
                unsigned runs_before = 0;  // run starts at stream positions < c0 (bitmap path)
                unsigned hit_any = 0;      // MODE 3: bit j = home atom j has a set-2 atom within the cutoff
                // MODE 4: neighbours of home atom `lane` found in this visit; MODE 5: next free slot of its row (a batch
                // can be visited again when the run table overflowed: the cursor lives in global memory in between;
                // only this lane ever touches it)
                unsigned nl_val = 0;
                const unsigned nl_gid = __float_as_uint(ws.home[lane].w);
                if (MODE == 5 && lane < (unsigned)nh) nl_val = __ldcg(P.nl_deg + nl_gid);
                // Candidates of stream positions [c, c + 64): run index of a position = (run starts at positions <= it)
                // - 1: two broadcast words and a popcount instead of a per-lane search.  Positions >= T fall into the
                // sentinel run, whose records are the far-away padding behind the sorted array: no bounds checks.
                auto fetch = [&](unsigned c, float4& q0, float4& q1, unsigned& g0, unsigned& g1, unsigned& b0i, unsigned& b1i) {
                    const uint2 bw = lds64(rbits_sa + (c >> 5) * 4u);
                    const unsigned pw0 = __popc(bw.x);
                    const unsigned k0 = runs_before + __popc(bw.x & le_mask) - 1u;
                    const unsigned k1 = runs_before + pw0 + __popc(bw.y & le_mask) - 1u;
                    runs_before += pw0 + __popc(bw.y);
                    const uint2 r0 = lds64(rec_sa + k0 * 8u), r1 = lds64(rec_sa + k1 * 8u);
                    b0i = r0.x + c + lane;
                    b1i = r1.x + c + lane + 32u;
                    q0 = __ldg(&P.sortedB[b0i]);
                    q1 = __ldg(&P.sortedB[b1i]);
                    g0 = r0.y;
                    g1 = r1.y;
                };
#if MB_SEARCH_PIPELINE
                // software pipeline: the loads of step s + 1 are issued before the tests of step s (needs the registers
                // of a 3-CTA/SM build)
                float4 n0n = make_float4(0.f, 0.f, 0.f, 0.f), n1n = n0n;
                unsigned f0n = 0, f1n = 0, a0n = 0, a1n = 0;
                if (use_bits && T > 0) fetch(0u, n0n, n1n, f0n, f1n, a0n, a1n);
#endif
#pragma unroll 1
                for (unsigned c0 = 0; c0 < T; c0 += 64) {
                    // ---- fused double step (pairs-only and count-only kernels): 128 candidates, four per lane, when
                    // neither half has a wrapped candidate.  One LDS.128 per home atom serves 128 tests, and the cursor,
                    // the step flags, the prefix scan and the emission loop's own overhead are paid once per 128
                    // candidates.  Mixed steps, steps with a test inside the band and the last odd step of a stream take
                    // the single-step path below.
                    if ((MODE == 0 || MODE == 2 || MODE >= 4) && !VDW && use_bits && c0 + 64 < T) {
                        const unsigned rb_save = runs_before;
                        float4 q0, q1, q2, q3;
                        unsigned g0, g1, g2, g3, i0, i1, i2, i3;
                        fetch(c0, q0, q1, g0, g1, i0, i1);
                        fetch(c0 + 64, q2, q3, g2, g3, i2, i3);
                        const unsigned fl = __reduce_or_sync(0xffffffffu, (g0 | g1 | g2 | g3) & (7u | RUN_SELF));
                        bool fused = (fl & 7u) == 0u;
                        unsigned m0 = 0, m1 = 0, m2 = 0, m3 = 0;
                        if (fused) {
                            asm volatile("prefetch.global.L1 [%0];" ::"l"(P.sortedB + i0 + 128));
                            asm volatile("prefetch.global.L1 [%0];" ::"l"(P.sortedB + i1 + 128));
                            asm volatile("prefetch.global.L1 [%0];" ::"l"(P.sortedB + i2 + 128));
                            asm volatile("prefetch.global.L1 [%0];" ::"l"(P.sortedB + i3 + 128));
                            const unsigned long long zz2 = pk2(0.f, 0.f);
                            const unsigned long long nxa = add2(pk2(q0.x, q1.x), zz2), nya = add2(pk2(q0.y, q1.y), zz2),
                                                     nza = add2(pk2(q0.z, q1.z), zz2);
                            const unsigned long long nxb = add2(pk2(q2.x, q3.x), zz2), nyb = add2(pk2(q2.y, q3.y), zz2),
                                                     nzb = add2(pk2(q2.z, q3.z), zz2);
                            const unsigned long long nrc22 = pk2_once(-rc2, -rc2);
                            float tmin = 3.0e38f;
                            // homes from the top down (each test shifts its sign bit in, so home j ends up at bit j): the
                            // nh mod 4 homes on top one by one, the rest four per round — no padded slots are tested
                            auto test_home = [&](int j) {
                                const float4 h = home[j];
                                float t0, t1, t2, t3;
                                upk2(d2f_minus_rc2(nxa, nya, nza, h, nrc22), t0, t1);
                                upk2(d2f_minus_rc2(nxb, nyb, nzb, h, nrc22), t2, t3);
                                m0 = __funnelshift_l(__float_as_uint(t0), m0, 1);
                                m1 = __funnelshift_l(__float_as_uint(t1), m1, 1);
                                m2 = __funnelshift_l(__float_as_uint(t2), m2, 1);
                                m3 = __funnelshift_l(__float_as_uint(t3), m3, 1);
                                tmin = min3abs(min3abs(tmin, t0, t1), t2, t3);
                            };
                            int top = nh;
#pragma unroll 1
                            for (int r = nh & 3; r > 0; --r) test_home(--top);
#pragma unroll 1
                            for (int gj = top - 4; gj >= 0; gj -= 4) {
#pragma unroll
                                for (int jj = 3; jj >= 0; --jj) test_home(gj + jj);
                            }
                            fused = !__any_sync(0xffffffffu, tmin <= P.band);
                        }
                        if (fused) {
                            m0 &= valid_slots;  // sign set = within
                            m1 &= valid_slots;
                            m2 &= valid_slots;
                            m3 &= valid_slots;
                            if (fl & RUN_SELF) {
                                // home cell against itself: keep (home j, atom a) only for a > hb + j
                                auto self_mask = [&](unsigned g, unsigned ai) {
                                    if (!(g & RUN_SELF)) return 0xffffffffu;
                                    const int lim = min(max((int)ai - (int)hb, 0), 32);
                                    return lim >= 32 ? 0xffffffffu : ((1u << lim) - 1u);
                                };
                                m0 &= self_mask(g0, i0);
                                m1 &= self_mask(g1, i1);
                                m2 &= self_mask(g2, i2);
                                m3 &= self_mask(g3, i3);
                            }
                            if (MODE >= 4) {
                                // the candidate stream of the full shell contains the home atoms themselves
                                m0 &= ~self_slot_bit(i0, hb);
                                m1 &= ~self_slot_bit(i1, hb);
                                m2 &= ~self_slot_bit(i2, hb);
                                m3 &= ~self_slot_bit(i3, hb);
                                c0 += 64;
                                if (MODE == 4) {
                                    if (vsteps == VC_MAX_STEPS) {
                                        vcount_fold(vplane, nl_val, nh, lane);
                                        vsteps = 0;
                                    }
                                    vcount_add4(vplane, m0, m1, m2, m3);
                                    ++vsteps;
                                } else {
                                    const unsigned mm[4] = {m0, m1, m2, m3};
                                    const unsigned ii[4] = {__float_as_uint(q0.w), __float_as_uint(q1.w),
                                                            __float_as_uint(q2.w), __float_as_uint(q3.w)};
                                    rows_emit<4>(mm, ii, nl_val, lane, P.nl_cols);
                                }
                                continue;
                            }
                            const int cnt = __popc(m0) + __popc(m1) + __popc(m2) + __popc(m3);
                            c0 += 64;  // the loop's own increment adds the other half
                            if (MODE == 2) {
                                count += cnt;
                                ntests += 4u * (unsigned)nh;
                                continue;
                            }
                            int inc = cnt;
                            scan_step(inc, 1);
                            scan_step(inc, 2);
                            scan_step(inc, 4);
                            scan_step(inc, 8);
                            scan_step(inc, 16);
                            const int tot = __shfl_sync(0xffffffffu, inc, 31);
                            if (tot == 0) continue;
                            const unsigned id0 = __float_as_uint(q0.w), id1 = __float_as_uint(q1.w),
                                           id2 = __float_as_uint(q2.w), id3 = __float_as_uint(q3.w);
                            // one pass over all homes, or passes of four homes (<= 4 x 128 entries) for a dense step
                            const bool single = tot <= STAGE_CAP;
                            const int hstep = single ? 32 : 4;
#pragma unroll 1
                            for (int h0 = 0; h0 < nh; h0 += hstep) {
                                unsigned e0 = m0, e1 = m1, e2 = m2, e3 = m3;
                                int off = inc - cnt, need = tot;
                                if (!single) {
                                    e0 = (m0 >> h0) & 0xfu;
                                    e1 = (m1 >> h0) & 0xfu;
                                    e2 = (m2 >> h0) & 0xfu;
                                    e3 = (m3 >> h0) & 0xfu;
                                    int v = __popc(e0) + __popc(e1) + __popc(e2) + __popc(e3), vi = v;
                                    scan_step(vi, 1);
                                    scan_step(vi, 2);
                                    scan_step(vi, 4);
                                    scan_step(vi, 8);
                                    scan_step(vi, 16);
                                    need = __shfl_sync(0xffffffffu, vi, 31);
                                    off = vi - v;
                                    if (need == 0) continue;
                                }
                                if (stage_n + need > STAGE_CAP + 1) stage_n = warp_flush_bulk(stage_sa, stage_n, P, lane);
                                unsigned sp = stage_sa + 8u * (unsigned)(stage_n + off);
                                unsigned ha = hid_sa + 4u * (unsigned)h0;
                                const int hend = min(nh, h0 + hstep);
                                uint4 hqa = lds128u(ha), hqb = lds128u(ha + 16u);
                                // full groups of four homes, two per round; then the last, partial group (1-3 homes)
                                const int hfull = h0 + ((hend - h0) & ~3);
                                bool odd = false;  // is the next group's id quad in hqb?
#pragma unroll 1
                                for (int gj = h0; gj < hfull; gj += 8) {
                                    emit_group4<4>(e0, e1, e2, e3, id0, id1, id2, id3, sp, hqa);
                                    hqa = lds128u(ha + 32u);
                                    e0 >>= 4; e1 >>= 4; e2 >>= 4; e3 >>= 4;
                                    if (gj + 4 >= hfull) {
                                        odd = true;
                                        break;
                                    }
                                    emit_group4<4>(e0, e1, e2, e3, id0, id1, id2, id3, sp, hqb);
                                    hqb = lds128u(ha + 48u);
                                    e0 >>= 4; e1 >>= 4; e2 >>= 4; e3 >>= 4;
                                    ha += 32u;
                                }
                                const int hrem = hend - hfull;
                                if (hrem) {
                                    const uint4 hq = odd ? hqb : hqa;
                                    if (hrem == 3) emit_group4<3>(e0, e1, e2, e3, id0, id1, id2, id3, sp, hq);
                                    else if (hrem == 2) emit_group4<2>(e0, e1, e2, e3, id0, id1, id2, id3, sp, hq);
                                    else emit_group4<1>(e0, e1, e2, e3, id0, id1, id2, id3, sp, hq);
                                }
                                stage_n += need;
                            }
                            continue;
                        }
                        runs_before = rb_save;  // not fused: the single-step path fetches its 64 candidates again
                    }
                    const unsigned p0 = c0 + lane, p1 = p0 + 32;
                    float4 n0, n1;
                    unsigned f0, f1, a0i, a1i;
                    if (use_bits) {
#if MB_SEARCH_PIPELINE
                        n0 = n0n; n1 = n1n; f0 = f0n; f1 = f1n; a0i = a0n; a1i = a1n;
                        if (c0 + 64 < T) fetch(c0 + 64, n0n, n1n, f0n, f1n, a0n, a1n);
#else
                        fetch(c0, n0, n1, f0, f1, a0i, a1i);
#endif
                    } else {
                        // very long streams (dense systems): per-lane forward search over the run positions
                        n0 = make_float4(pad_cand, pad_cand, pad_cand, 0.f);
                        n1 = n0;
                        f0 = f1 = a0i = a1i = 0;
                        if (p0 < T) {
                            while ((ws.rec[cur0 + 1].y >> 8) <= p0) ++cur0;
                            const uint2 r0 = ws.rec[cur0];
                            a0i = r0.x + p0;
                            f0 = r0.y;
                            n0 = __ldg(&P.sortedB[a0i]);
                        }
                        if (p1 < T) {
                            cur1 = max(cur1, cur0);
                            while ((ws.rec[cur1 + 1].y >> 8) <= p1) ++cur1;
                            const uint2 r1 = ws.rec[cur1];
                            a1i = r1.x + p1;
                            f1 = r1.y;
                            n1 = __ldg(&P.sortedB[a1i]);
                        }
                    }
                    unsigned m0 = 0, m1 = 0;
                    float vc0 = 0.f, vc1 = 0.f;  // vdW radii of the two candidates (padding records carry local id 0)
                    if (VDW) {
                        vc0 = __ldg(&P.vdwB[__float_as_uint(n0.w)]);
                        vc1 = __ldg(&P.vdwB[__float_as_uint(n1.w)]);
                    }
                    // one warp-wide OR of the flags decides the path of the step: wrapped candidates (bits 0-2) need
                    // the periodic distance, self-run candidates (RUN_SELF) the index-order filter
                    const unsigned step_flags = __reduce_or_sync(0xffffffffu, (f0 | f1) & (7u | RUN_SELF));
                    const bool any_wrapped = (step_flags & 7u) != 0u;
                    // the next 64 stream positions usually continue the same runs: start pulling those lines into L1
#if !MB_SEARCH_PIPELINE
                    if (use_bits) {
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(P.sortedB + a0i + 64));
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(P.sortedB + a1i + 64));
                    }
#endif
                    if (VDW && any_wrapped) {
                        // vdW search, step with wrapped cell pairs: every test with the reference's own expression
                        // (direct difference or PeriodicBox::distance_squared) against the pair's own cutoff
                        for (int j = 0; j < nh; ++j) {
                            const float4 hh = home[j];
                            const float hv = ws.hvdw[j];
                            const float cut0 = xadd(xadd(hv, vc0), FLT_EPSILON), cut1 = xadd(xadd(hv, vc1), FLT_EPSILON);
                            const float d0 = (f0 & 7u) ? d2_pbc_call(P.g.box, hh.x, hh.y, hh.z, n0.x, n0.y, n0.z, f0 & 7u)
                                                       : d2_direct(hh.x, hh.y, hh.z, n0.x, n0.y, n0.z);
                            const float d1 = (f1 & 7u) ? d2_pbc_call(P.g.box, hh.x, hh.y, hh.z, n1.x, n1.y, n1.z, f1 & 7u)
                                                       : d2_direct(hh.x, hh.y, hh.z, n1.x, n1.y, n1.z);
                            if (d0 <= xmul(cut0, cut0)) m0 |= 1u << j;
                            if (d1 <= xmul(cut1, cut1)) m1 |= 1u << j;
                        }
                    } else if (!any_wrapped) {
                        // ---- all 64 candidates come from un-wrapped cell pairs: direct difference ----
                        // x + 0 is exact: the FADD2 only serves to give each packed coordinate its own aligned
                        // register pair (ptxas otherwise re-packs the halves in front of every use)
                        const unsigned long long zz2 = pk2(0.f, 0.f);
                        const unsigned long long nx = add2(pk2(n0.x, n1.x), zz2), ny = add2(pk2(n0.y, n1.y), zz2),
                                                 nz = add2(pk2(n0.z, n1.z), zz2);
                        // Home slots are visited from the top down and each test pushes its "not within" bit
                        // (the sign of rc2 - d2) into the accumulators with one funnel shift, so that slot j ends
                        // up at bit j.
                        const unsigned long long one2 = pk2_once(P.one, P.one), rc22 = pk2_once(rc2, rc2);
                        const unsigned long long mone2 = pk2(-P.one, -P.one), vc2 = pk2(vc0, vc1),
                                                 eps2 = pk2(FLT_EPSILON, FLT_EPSILON);
                        const int nh4 = (nh + 3) & ~3;
                        unsigned a0 = 0, a1 = 0;
                        bool exact = VDW;
                        if (!VDW) {
                            // filter pass: fused d2f - rc2, sign SET = within; min |t| says whether any test of the
                            // step is too close to the cutoff for the filter to be trusted
                            const unsigned long long nrc22 = pk2_once(-rc2, -rc2);
                            float tmin = 3.0e38f;
#pragma unroll 1
                            for (int gj = nh4 - 4; gj >= 0; gj -= 4) {
#pragma unroll
                                for (int jj = 3; jj >= 0; --jj) {
                                    float t0, t1;
                                    upk2(d2f_minus_rc2(nx, ny, nz, home[gj + jj], nrc22), t0, t1);
                                    a0 = __funnelshift_l(__float_as_uint(t0), a0, 1);
                                    a1 = __funnelshift_l(__float_as_uint(t1), a1, 1);
                                    tmin = min3abs(tmin, t0, t1);
                                }
                            }
                            exact = __any_sync(0xffffffffu, tmin <= P.band);
                            a0 = ~a0;  // the exact loop below produces "not within" bits
                            a1 = ~a1;
                        }
                        if (exact) {
                        a0 = a1 = 0;
#pragma unroll 1
                        for (int gj = nh4 - 4; gj >= 0; gj -= 4) {
#pragma unroll
                            for (int jj = 3; jj >= 0; --jj) {
                                unsigned long long t;
                                if (VDW) {
                                    // cutoff = (vdw1[i] + vdw2[j]) + EPSILON, squared (distance_search.rs:392-393): packed,
                                    // every operation rounded on its own; c*c - d2 as fma(d2, -1.0, c*c) with a run-time
                                    // -1.0 so that ptxas cannot contract the product into the subtraction
                                    const float hv = ws.hvdw[gj + jj];
                                    const unsigned long long cc = add2(add2(pk2(hv, hv), vc2), eps2);
                                    t = cut2_minus_d2(nx, ny, nz, home[gj + jj], one2, mone2, mul2(cc, cc));
                                } else {
                                    t = rc2_minus_d2(nx, ny, nz, home[gj + jj], one2, rc22);
                                }
                                float t0, t1;
                                upk2(t, t0, t1);
                                a0 = __funnelshift_l(__float_as_uint(t0), a0, 1);
                                a1 = __funnelshift_l(__float_as_uint(t1), a1, 1);
                            }
                        }
                        }
                        m0 = ~a0 & valid_slots;
                        m1 = ~a1 & valid_slots;
                        // home cell against itself (first stream positions): keep (home j, atom a) only for a > hb + j
                        if (step_flags & RUN_SELF) {
                            if (f0 & RUN_SELF) {
                                const int lim = min(max((int)a0i - (int)hb, 0), 32);
                                m0 &= lim >= 32 ? 0xffffffffu : ((1u << lim) - 1u);
                            }
                            if (f1 & RUN_SELF) {
                                const int lim = min(max((int)a1i - (int)hb, 0), 32);
                                m1 &= lim >= 32 ? 0xffffffffu : ((1u << lim) - 1u);
                            }
                        }
                    } else {
                        // ---- mixed step: wrapped cell pairs (and possibly the self cell).
                        // Wrapped candidates are tested on the lattice-shifted image; outside the band
                        // [lo,hi] around cutoff^2 that test is decisive, inside it (and always, when the
                        // filter's preconditions do not hold: lo=-1, hi=inf) the reference's exact
                        // PeriodicBox::distance_squared decides.
                        float sx0 = 0.f, sy0 = 0.f, sz0 = 0.f, sx1 = 0.f, sy1 = 0.f, sz1 = 0.f;
#pragma unroll
                        for (int d = 0; d < 3; ++d) {
                            const float bx = P.g.box.m[d], by = P.g.box.m[3 + d], bz = P.g.box.m[6 + d];
                            if ((f0 >> d) & 1u) {
                                const float sg = ((f0 >> (3 + d)) & 1u) ? -1.0f : 1.0f;
                                sx0 += sg * bx; sy0 += sg * by; sz0 += sg * bz;
                            }
                            if ((f1 >> d) & 1u) {
                                const float sg = ((f1 >> (3 + d)) & 1u) ? -1.0f : 1.0f;
                                sx1 += sg * bx; sy1 += sg * by; sz1 += sg * bz;
                            }
                        }
                        const bool w0 = (f0 & 7u) != 0u, w1 = (f1 & 7u) != 0u;
                        const float lo0 = w0 ? (P.fast_pbc ? P.rc2_lo : -1.0f) : rc2, hi0 = w0 ? (P.fast_pbc ? P.rc2_hi : finf) : rc2;
                        const float lo1 = w1 ? (P.fast_pbc ? P.rc2_lo : -1.0f) : rc2, hi1 = w1 ? (P.fast_pbc ? P.rc2_hi : finf) : rc2;
                        const unsigned long long nx = add2(pk2(n0.x, n1.x), pk2(sx0, sx1)),
                                                 ny = add2(pk2(n0.y, n1.y), pk2(sy0, sy1)),
                                                 nz = add2(pk2(n0.z, n1.z), pk2(sz0, sz1));
                        unsigned b0 = 0, b1 = 0;  // possible hits (superset of the certain ones in m0/m1)
#pragma unroll 1
                        for (int gj = 0; gj < nh; gj += 4) {
                            unsigned l0 = 0, l1 = 0, u0 = 0, u1 = 0;
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                float d0, d1;
                                d2_pair(nx, ny, nz, home[gj + jj], d0, d1);
                                if (d0 <= lo0) l0 |= 1u << jj;
                                if (d1 <= lo1) l1 |= 1u << jj;
                                if (d0 <= hi0) u0 |= 1u << jj;
                                if (d1 <= hi1) u1 |= 1u << jj;
                            }
                            m0 |= l0 << gj; m1 |= l1 << gj;
                            b0 |= u0 << gj; b1 |= u1 << gj;
                        }
                        // unused home slots hold finite padding: with the unbounded band of the exact mode (hi = inf)
                        // they would qualify for a re-evaluation, so they are masked out here
                        m0 &= valid_slots;
                        m1 &= valid_slots;
                        b0 = (b0 & valid_slots) ^ m0;
                        b1 = (b1 & valid_slots) ^ m1;
                        while (b0) {
                            const int j = bfind32(b0);
                            b0 ^= 1u << j;
                            const float4 h = home[j];
                            if (d2_pbc_call(P.g.box, h.x, h.y, h.z, n0.x, n0.y, n0.z, f0 & 7u) <= rc2) m0 |= 1u << j;
                        }
                        while (b1) {
                            const int j = bfind32(b1);
                            b1 ^= 1u << j;
                            const float4 h = home[j];
                            if (d2_pbc_call(P.g.box, h.x, h.y, h.z, n1.x, n1.y, n1.z, f1 & 7u) <= rc2) m1 |= 1u << j;
                        }
                        // home cell against itself: keep (home j, atom a) only for a > hb + j
                        if (f0 & RUN_SELF) {
                            const int lim = min(max((int)a0i - (int)hb, 0), 32);
                            m0 &= lim >= 32 ? 0xffffffffu : ((1u << lim) - 1u);
                        }
                        if (f1 & RUN_SELF) {
                            const int lim = min(max((int)a1i - (int)hb, 0), 32);
                            m1 &= lim >= 32 ? 0xffffffffu : ((1u << lim) - 1u);
                        }
                    }
                    if (MODE == 3) {
                        hit_any |= m0 | m1;
                        continue;
                    }
                    if (MODE >= 4) {
                        m0 &= ~self_slot_bit(a0i, hb);
                        m1 &= ~self_slot_bit(a1i, hb);
                        if (MODE == 4) {
                            if (vsteps == VC_MAX_STEPS) {
                                vcount_fold(vplane, nl_val, nh, lane);
                                vsteps = 0;
                            }
                            vcount_add(vplane, m0 ^ m1, m0 & m1, 0u);
                            ++vsteps;
                        } else {
                            const unsigned mm[2] = {m0, m1};
                            const unsigned ii[2] = {__float_as_uint(n0.w), __float_as_uint(n1.w)};
                            rows_emit<2>(mm, ii, nl_val, lane, P.nl_cols);
                        }
                        continue;
                    }
                    const int c0n = __popc(m0), c1n = __popc(m1);
                    if (MODE == 2) {
                        count += c0n + c1n;
                        ntests += 2u * (unsigned)((nh + 3) & ~3);  // per lane: two candidates x the home slots visited
                        continue;
                    }
                    // inclusive scan of the hits per lane (both candidates together: a lane's second candidate is
                    // staged right behind its first)
                    const int cnt = c0n + c1n;
                    int inc = cnt;
                    scan_step(inc, 1);
                    scan_step(inc, 2);
                    scan_step(inc, 4);
                    scan_step(inc, 8);
                    scan_step(inc, 16);
                    const int tot = __shfl_sync(0xffffffffu, inc, 31);
                    if (tot == 0) continue;
                    const unsigned id0 = __float_as_uint(n0.w), id1 = __float_as_uint(n1.w);
                    if (MODE == 0) {
                        // Stage the hits as final pairs, a lane's entries in home order.  Normally one pass over all
                        // homes; a step that found more pairs than the buffer holds goes through passes of eight homes
                        // (<= 8 x 64 entries), each with its own prefix sums.
                        const bool single = tot <= STAGE_CAP;
                        const int hstep = single ? 32 : 8;
#pragma unroll 1
                        for (int h0 = 0; h0 < nh; h0 += hstep) {
                            unsigned e0 = m0, e1 = m1;
                            int off = inc - cnt, need = tot;  // exclusive prefix of this lane, entries of this pass
                            if (!single) {
                                e0 = (m0 >> h0) & 0xffu;
                                e1 = (m1 >> h0) & 0xffu;
                                int v = __popc(e0) + __popc(e1), vi = v;
                                scan_step(vi, 1);
                                scan_step(vi, 2);
                                scan_step(vi, 4);
                                scan_step(vi, 8);
                                scan_step(vi, 16);
                                need = __shfl_sync(0xffffffffu, vi, 31);
                                off = vi - v;
                                if (need == 0) continue;
                            }
                            if (stage_n + need > STAGE_CAP + 1) stage_n = warp_flush_bulk(stage_sa, stage_n, P, lane);
                            unsigned sp = stage_sa + 8u * (unsigned)(stage_n + off);
                            unsigned ha = hid_sa + 4u * (unsigned)h0;
                            const int hend = min(nh, h0 + hstep);
                            // two groups per round, their ids loaded one round ahead into alternating registers (the
                            // reads past the batch run into hvdw: harmless)
                            uint4 hqa = lds128u(ha), hqb = lds128u(ha + 16u);
#pragma unroll 1
                            for (int gj = h0; gj < hend; gj += 8) {
                                emit_group(e0, e1, id0, id1, sp, hqa);
                                hqa = lds128u(ha + 32u);
                                e0 >>= 4;
                                e1 >>= 4;
                                if (gj + 4 >= hend) break;
                                emit_group(e0, e1, id0, id1, sp, hqb);
                                hqb = lds128u(ha + 48u);
                                e0 >>= 4;
                                e1 >>= 4;
                                ha += 32u;
                            }
                            stage_n += need;
                        }
                        continue;
                    }
                    // MODE 1: expand the masks into the staging buffer as (home slot, candidate id) + distance.  Normally
                    // one pass; a step that found more pairs than the buffer holds goes through four (one candidate per
                    // lane x one 16-lane group: <= 512).
                    const int npass = tot <= STAGE_CAP ? 1 : 4;
#pragma unroll 1
                    for (int pass = 0; pass < npass; ++pass) {
                        unsigned e0 = m0, e1 = m1;
                        int off = inc - cnt, need = tot;  // exclusive prefix of this lane, entries of this pass
                        if (npass != 1) {
                            const bool act = (int)(lane >> 4) == (pass & 1);
                            e0 = (act && pass < 2) ? m0 : 0u;
                            e1 = (act && pass >= 2) ? m1 : 0u;
                            int v = __popc(e0) + __popc(e1), vi = v;
                            scan_step(vi, 1);
                            scan_step(vi, 2);
                            scan_step(vi, 4);
                            scan_step(vi, 8);
                            scan_step(vi, 16);
                            need = __shfl_sync(0xffffffffu, vi, 31);
                            off = vi - v;
                        }
                        if (stage_n + need > STAGE_CAP) warp_flush<MODE>(stage_sa, staged_sa, home_sa, stage_n, P, lane);
                        unsigned sp0 = stage_sa + 8u * (unsigned)(stage_n + off);
                        unsigned sp1 = sp0 + 8u * (unsigned)__popc(e0);
                        {
                            unsigned dp0 = staged_sa + 4u * (unsigned)(stage_n + off);
                            unsigned dp1 = dp0 + 4u * (unsigned)__popc(e0);
                            while (e0) {
                                const int j = bfind32(e0);
                                e0 ^= 1u << j;
                                sts64(sp0, (unsigned)j, id0);
                                sp0 += 8u;
                                const float4 h = home[j];
                                float d2 = (f0 & 7u) ? d2_pbc_call(P.g.box, h.x, h.y, h.z, n0.x, n0.y, n0.z, f0 & 7u)
                                                     : d2_direct(h.x, h.y, h.z, n0.x, n0.y, n0.z);
                                sts32f(dp0, __fsqrt_rn(d2));
                                dp0 += 4u;
                            }
                            while (e1) {
                                const int j = bfind32(e1);
                                e1 ^= 1u << j;
                                sts64(sp1, (unsigned)j, id1);
                                sp1 += 8u;
                                const float4 h = home[j];
                                float d2 = (f1 & 7u) ? d2_pbc_call(P.g.box, h.x, h.y, h.z, n1.x, n1.y, n1.z, f1 & 7u)
                                                     : d2_direct(h.x, h.y, h.z, n1.x, n1.y, n1.z);
                                sts32f(dp1, __fsqrt_rn(d2));
                                dp1 += 4u;
                            }
                        }
                        stage_n += need;
                    }
                }
                if (MODE == 3) {
                    hit_any = __reduce_or_sync(0xffffffffu, hit_any);
                    if (lane < (unsigned)nh && ((hit_any >> lane) & 1u)) P.flags[__float_as_uint(home[lane].w)] = 1;
                }
                if (MODE == 4) {
                    vcount_fold(vplane, nl_val, nh, lane);
                    vsteps = 0;
                    if (lane < (unsigned)nh && nl_val) atomicAdd(P.nl_deg + nl_gid, nl_val);
                    count += nl_val;  // u64 total: the host checks it against the 32-bit row offsets
                }
                if (MODE == 5 && lane < (unsigned)nh) __stcg(P.nl_deg + nl_gid, nl_val);
                // MODE 1: staged entries name home atoms by their slot in ws.home: write them out before it changes
                if (MODE == 1) warp_flush<MODE>(stage_sa, staged_sa, home_sa, stage_n, P, lane);
            }
            first = false;
        } while (row0 < P.nrows);
    }
    if (MODE == 0) {
        // last flush: the even part as a bulk copy, an odd last pair into the tail behind pair_cap (merge_pair_tail_kernel)
        stage_n = warp_flush_bulk(stage_sa, stage_n, P, lane);
        if (lane == 0) {
            if (stage_n) {
                const unsigned long long k = atomicAdd(P.counter + 3, 1ull);
                if (k < (unsigned long long)PAIR_TAIL) P.pairs[P.pair_cap + k] = lds64(stage_sa);
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    }
    if (MODE == 2 || MODE == 4) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) count += __shfl_xor_sync(0xffffffffu, count, o);
        if (lane == 0 && count) atomicAdd(P.counter, count);
        if (lane == 0 && ntests) atomicAdd(P.counter + 2, 32ull * ntests);  // every lane evaluated the same number
    }
}


// ---------------------------------------------------------------------------------------------
// general all-pairs kernel: any grid (degenerate dims, partial PBC), two sets, `within`.
// Exact for every case the reference handles, O(n1*n2): used for small systems and as the
// route for double/within searches in this round.
// ---------------------------------------------------------------------------------------------
struct BruteParams {
    const float4* a4;
    const unsigned long long* aref;
    int na;
    const float4* b4;
    const unsigned long long* bref;
    int nb;
    int mode;  // 0 single (i<j over the same array), 1 double (ordered i,j), 2 within (flag set-1 atoms)
    int with_dist;
    int dims[3];
    unsigned pbc;
    int periodic_variant;
    float rc2;
    DevBox box;
    uint2* pairs;
    float* dists;
    unsigned long long pair_cap;
    unsigned long long* counter;
    unsigned char* flags;
    int count_only;
    const float* avdw;  // vdW mode: radii per selected atom of set A / B (indexed by the local id in .w), else NULL
    const float* bvdw;
};

// Set of wrapped-dims flags under which search_plan lists the (unordered) reference cell pair
// {ca, cb}; the pair is a hit if ANY of them passes (the reference then simply emits it more than
// once).  Handles dims[d] in {1,2}, where the same cell pair is reached both directly and wrapped.
__device__ __forceinline__ bool general_pair_test(const BruteParams& P, float4 a, unsigned long long ra, float4 b,
                                                  unsigned long long rb, float rc2, float& d2min) {
    unsigned opt0 = 0, opt1 = 0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        int ca = (int)((ra >> (21 * d)) & 0x1FFFFFull), cb = (int)((rb >> (21 * d)) & 0x1FFFFFull);
        int diff = abs(ca - cb);
        int dim = P.dims[d];
        bool per = P.periodic_variant && ((P.pbc >> d) & 1u);
        bool o0 = diff <= 1;
        bool o1 = per && (dim == 1 ? diff == 0 : (dim == 2 ? diff == 1 : diff == dim - 1));
        if (!o0 && !o1) return false;
        opt0 |= (o0 ? 1u : 0u) << d;
        opt1 |= (o1 ? 1u : 0u) << d;
    }
    bool hit = false;
    d2min = FLT_MAX;
    for (unsigned w = 0; w < 8; ++w) {
        // w is allowed if every set bit is in opt1 and every clear bit is in opt0
        if ((w & ~opt1) || ((~w & 7u) & ~opt0)) continue;
        float d2 = w ? d2_pbc(P.box, a.x, a.y, a.z, b.x, b.y, b.z, w) : d2_direct(a.x, a.y, a.z, b.x, b.y, b.z);
        if (d2 <= rc2) {
            hit = true;
            d2min = fminf(d2min, d2);
        }
    }
    return hit;
}

constexpr int BRUTE_THREADS = 256;
constexpr int BRUTE_TILE = 256;

__global__ void __launch_bounds__(BRUTE_THREADS) brute_kernel(const __grid_constant__ BruteParams P) {
    __shared__ float4 tb4[BRUTE_TILE];
    __shared__ unsigned long long tbr[BRUTE_TILE];
    __shared__ uint2 stage_all[BRUTE_THREADS / 32][256];
    __shared__ float stage_d_all[BRUTE_THREADS / 32][256];
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    uint2* stage = stage_all[wid];
    float* stage_d = stage_d_all[wid];
    int stage_n = 0;
    unsigned long long count = 0;

    const int i = blockIdx.x * BRUTE_THREADS + threadIdx.x;
    const bool ivalid = i < P.na;
    float4 a = ivalid ? P.a4[i] : make_float4(0, 0, 0, 0);
    unsigned long long ra = ivalid ? P.aref[i] : ~0ull;
    const bool alive = ivalid && ra != ~0ull;
    bool found = false;
    const float va = (P.avdw && ivalid) ? P.avdw[__float_as_uint(a.w)] : 0.f;

    // blockIdx.y strides over tiles of set B
    for (int t0 = blockIdx.y * BRUTE_TILE; t0 < P.nb; t0 += gridDim.y * BRUTE_TILE) {
        if (P.mode == 0 && t0 + BRUTE_TILE <= blockIdx.x * BRUTE_THREADS) continue;  // j > i only
        __syncthreads();
        int j = t0 + threadIdx.x;
        if (j < P.nb) {
            tb4[threadIdx.x] = P.b4[j];
            tbr[threadIdx.x] = P.bref[j];
        } else {
            tbr[threadIdx.x] = ~0ull;
        }
        __syncthreads();
        const int tn = min(BRUTE_TILE, P.nb - t0);
        for (int jj = 0; jj < tn; ++jj) {
            unsigned long long rb = tbr[jj];
            bool hit = false;
            float d2 = 0.f;
            if (alive && rb != ~0ull && !(P.mode == 0 && t0 + jj <= i)) {
                float rc2 = P.rc2;
                if (P.avdw) {
                    // cutoff = vdw1[i] + vdw2[j] + EPSILON ; d2 <= cutoff*cutoff  (distance_search.rs:392-393,423-425)
                    const float cut = xadd(xadd(va, P.bvdw[__float_as_uint(tb4[jj].w)]), FLT_EPSILON);
                    rc2 = xmul(cut, cut);
                }
                hit = general_pair_test(P, a, ra, tb4[jj], rb, rc2, d2);
            }
            if (P.mode == 2) {
                found |= hit;
                continue;
            }
            unsigned bal = __ballot_sync(0xffffffffu, hit);
            if (!bal) continue;
            int total = __popc(bal);
            if (P.count_only) {
                if (lane == 0) count += total;
                continue;
            }
            if (stage_n + total > 256) {
                __syncwarp();
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(P.counter, (unsigned long long)stage_n);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (base + stage_n <= P.pair_cap)
                    for (int q = lane; q < stage_n; q += 32) {
                        P.pairs[base + q] = stage[q];
                        if (P.with_dist) P.dists[base + q] = stage_d[q];
                    }
                __syncwarp();
                stage_n = 0;
            }
            if (hit) {
                int off = stage_n + __popc(bal & ((1u << lane) - 1u));
                unsigned ia = __float_as_uint(a.w), ib = __float_as_uint(tb4[jj].w);
                stage[off] = P.mode == 0 ? make_uint2(min(ia, ib), max(ia, ib)) : make_uint2(ia, ib);
                stage_d[off] = __fsqrt_rn(d2);
            }
            stage_n += total;
        }
    }
    if (P.mode == 2) {
        if (found) P.flags[i] = 1;
        return;
    }
    if (P.count_only) {
        if (lane == 0 && count) atomicAdd(P.counter, count);
        return;
    }
    __syncwarp();
    if (stage_n) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(P.counter, (unsigned long long)stage_n);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base + stage_n <= P.pair_cap)
            for (int q = lane; q < stage_n; q += 32) {
                P.pairs[base + q] = stage[q];
                if (P.with_dist) P.dists[base + q] = stage_d[q];
            }
    }
}

// ---- small helpers ---------------------------------------------------------------------------
// min/max over a selection; init value parameterised (0 for compute_min_max, distance_search.rs:602-616;
// +-FLT_MAX for Measure::min_max, measure.rs:22-36)
__global__ void __launch_bounds__(256) minmax_kernel(const float* __restrict__ xyz,
                                                     const unsigned long long* __restrict__ ids, int n,
                                                     float init_lo, float init_hi, float* __restrict__ out6) {
    float lo[3] = {init_lo, init_lo, init_lo}, hi[3] = {init_hi, init_hi, init_hi};
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        size_t id = ids ? (size_t)ids[k] : (size_t)k;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            float v = xyz[3 * id + d];
            if (v < lo[d]) lo[d] = v;
            if (v > hi[d]) hi[d] = v;
        }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
            hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            // float atomic min/max through the ordered-int trick
            int li = __float_as_int(lo[d]), hi_i = __float_as_int(hi[d]);
            if (li >= 0) atomicMin((int*)&out6[d], li); else atomicMax((unsigned*)&out6[d], (unsigned)li);
            if (hi_i >= 0) atomicMax((int*)&out6[3 + d], hi_i); else atomicMin((unsigned*)&out6[3 + d], (unsigned)hi_i);
        }
    }
}

// by_gid: flags are indexed by GLOBAL atom id (cell kernel) instead of by position in the selection
__global__ void compact_flags_kernel(const unsigned char* __restrict__ flags, const unsigned* __restrict__ pos,
                                     const unsigned long long* __restrict__ ids, int n, int by_gid,
                                     unsigned long long* __restrict__ out) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    unsigned long long gid = ids ? ids[k] : (unsigned long long)k;
    if (flags[by_gid ? gid : (unsigned long long)k]) out[pos[k]] = gid;
}
__global__ void flags_to_u32_kernel(const unsigned char* __restrict__ flags, const unsigned long long* __restrict__ ids,
                                    int n, int by_gid, unsigned* __restrict__ out) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = flags[by_gid ? (ids ? ids[k] : (unsigned long long)k) : (unsigned long long)k];
}

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}
__global__ void __launch_bounds__(256) checksum_kernel(const uint2* __restrict__ pairs, unsigned long long n,
                                                       int canonical, unsigned long long* __restrict__ out2) {
    unsigned long long s = 0, x = 0;
    for (unsigned long long k = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; k < n;
         k += (unsigned long long)gridDim.x * blockDim.x) {
        uint2 p = pairs[k];
        if (canonical && p.x > p.y) {  // single-set pairs are stored unordered: hash (min,max)
            unsigned t = p.x;
            p.x = p.y;
            p.y = t;
        }
        unsigned long long h = mix64(((unsigned long long)p.x << 32) | p.y);
        s += h;
        x ^= h;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        x ^= __shfl_xor_sync(0xffffffffu, x, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&out2[0], s);
        atomicXor(&out2[1], x);
    }
}

// ---------------------------------------------------------------------------------------------
// host side: grid planning
// ---------------------------------------------------------------------------------------------
static inline int dims_from_extent(float ext, float cutoff) {
    // clamp_min((extents[d] / cutoff).floor() as usize, 1)   (distance_search.rs:103-110)
    float q = std::floor(ext / cutoff);
    if (!(q >= 1.0f)) return 1;  // negative, NaN -> saturating cast gives 0 -> clamp to 1
    if (q > 2.0e6f) return 2000000;
    return (int)q;
}

// minimum of |M (delta + u)|^2 over u in [-1,1]^3 (distance between two equal parallelepiped cells
// offset by the integer vector delta): enumerate the 27 active sets of the box-constrained QP.
// `hx`: the home side spans hx cells along x (cells [0,hx)), the neighbour side is cell delta.
static double min_cell_dist2(const double M[3][3], const int delta[3], int hx = 1) {
    double best = 1e300;
    const double ulo[3] = {-(double)hx, -1.0, -1.0};
    for (int cfg = 0; cfg < 27; ++cfg) {
        int st[3] = {cfg % 3, (cfg / 3) % 3, cfg / 9};  // 0: free, 1: lower bound, 2: upper bound
        double fixed[3] = {0, 0, 0};
        int freev[3], nf = 0;
        for (int d = 0; d < 3; ++d) {
            if (st[d] == 0) freev[nf++] = d;
            else fixed[d] = (double)delta[d] + (st[d] == 1 ? ulo[d] : 1.0);
        }
        // r0 = M * (fixed part + delta for free vars)
        double base[3] = {0, 0, 0};
        for (int d = 0; d < 3; ++d) {
            double coef = st[d] == 0 ? (double)delta[d] : fixed[d];
            for (int r = 0; r < 3; ++r) base[r] += M[r][d] * coef;
        }
        // minimise |base + sum_f M[:,f] u_f|^2 over free u (normal equations, nf<=3)
        double A[3][3] = {{0}}, rhs[3] = {0, 0, 0}, u[3] = {0, 0, 0};
        for (int a = 0; a < nf; ++a) {
            for (int b = 0; b < nf; ++b)
                for (int r = 0; r < 3; ++r) A[a][b] += M[r][freev[a]] * M[r][freev[b]];
            for (int r = 0; r < 3; ++r) rhs[a] -= M[r][freev[a]] * base[r];
        }
        bool ok = true;
        if (nf > 0) {
            // Gaussian elimination with partial pivoting
            double aug[3][4];
            for (int a = 0; a < nf; ++a) {
                for (int b = 0; b < nf; ++b) aug[a][b] = A[a][b];
                aug[a][nf] = rhs[a];
            }
            for (int col = 0; col < nf && ok; ++col) {
                int piv = col;
                for (int r = col + 1; r < nf; ++r)
                    if (std::fabs(aug[r][col]) > std::fabs(aug[piv][col])) piv = r;
                if (std::fabs(aug[piv][col]) < 1e-300) { ok = false; break; }
                for (int q = 0; q <= nf; ++q) std::swap(aug[col][q], aug[piv][q]);
                for (int r = 0; r < nf; ++r) {
                    if (r == col) continue;
                    double f = aug[r][col] / aug[col][col];
                    for (int q = col; q <= nf; ++q) aug[r][q] -= f * aug[col][q];
                }
            }
            if (ok)
                for (int a = 0; a < nf; ++a) {
                    u[a] = aug[a][nf] / aug[a][a];
                    if (u[a] < ulo[freev[a]] - 1e-12 || u[a] > 1.0 + 1e-12) ok = false;
                }
        }
        if (!ok) continue;
        double v[3] = {base[0], base[1], base[2]};
        for (int a = 0; a < nf; ++a)
            for (int r = 0; r < 3; ++r) v[r] += M[r][freev[a]] * u[a];
        double d2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
        best = std::min(best, d2);
    }
    return best;
}

struct Plan {
    GridSpec g;
    bool use_cells;
    bool full_shell;  // two-set search: rows cover both half-spaces and the home row entirely
    int fast_pbc;
    float rc2_lo, rc2_hi;
    int nrows;
    NbrRow rows[MAX_ROWS];
    size_t ncells;
    double volume;
};

// Decide subdivision and neighbour-offset table for the fast cell kernel; fall back to the
// general all-pairs kernel when the reference grid is degenerate.
static void plan_cells(const Ctx* c, Plan& pl, float cutoff, size_t n) {
    GridSpec& g = pl.g;
    pl.use_cells = false;
    for (int d = 0; d < 3; ++d) {
        g.k[d] = 1;
        g.kmagic[d] = 0;
        g.fd[d] = g.dims[d];
    }
    g.hx = 1;
    pl.ncells = (size_t)g.dims[0] * g.dims[1] * g.dims[2];
    if (c->opt_force_brute) return;
    if (n < 4096) return;  // all-pairs is cheaper than five launches
    for (int d = 0; d < 3; ++d)
        if (g.periodic_variant && ((g.pbc >> d) & 1u) && g.dims[d] < 3) return;
    // reference-cell lattice in lab space (columns = cell edge vectors)
    double Mref[3][3];
    if (g.periodic_variant) {
        for (int r = 0; r < 3; ++r)
            for (int col = 0; col < 3; ++col) Mref[r][col] = (double)g.box.m[r * 3 + col] / g.dims[col];
    } else {
        for (int r = 0; r < 3; ++r)
            for (int col = 0; col < 3; ++col) Mref[r][col] = r == col ? (double)g.dim_sz[col] / g.dims[col] : 0.0;
    }
    auto thickness = [&](const double M[3][3], double t[3]) {
        // perpendicular width of the slab spanned by the other two edges = 1/|row d of M^-1|
        double det = M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
                     M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
        for (int d = 0; d < 3; ++d) {
            int a = (d + 1) % 3, b = (d + 2) % 3;
            double cr[3] = {M[1][a] * M[2][b] - M[2][a] * M[1][b], M[2][a] * M[0][b] - M[0][a] * M[2][b],
                            M[0][a] * M[1][b] - M[1][a] * M[0][b]};
            double nn = std::sqrt(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]);
            t[d] = nn > 0 ? std::fabs(det) / nn : 0.0;
        }
        return std::fabs(det);
    };
    double tref[3];
    double vcell = thickness(Mref, tref);
    pl.volume = vcell * pl.ncells;
    if (!(vcell > 0)) return;
    const double rc = (double)cutoff;
    const double rc_cull = rc * (1.0 + 1e-4) + 1e-6;
    int k[3];
    for (int d = 0; d < 3; ++d) {
        if (c->opt_subdiv > 0) k[d] = c->opt_subdiv;
        // tile edge ~0.4-0.55 cutoff in y and z, ~0.6-0.85 cutoff along x (the tile is sliced hx times along x, so a
        // longer tile costs no hit rate; swept on B200 with the fused 128-candidate step: k_x 3 -> 2 and hx 3 -> 6
        // take the 1M-atom frame from 1.016 to 0.98 ms)
        else k[d] = (int)std::floor(tref[d] / ((d == 0 ? 0.63 : 0.42) * rc) + 0.5);
        k[d] = std::max(1, std::min(k[d], 8));
    }
    // keep enough atoms per fine cell for the home loop to amortise its per-run overhead
    if (c->opt_subdiv <= 0) {
        for (;;) {
            double cells = (double)pl.ncells * k[0] * k[1] * k[2];
            if ((double)n / cells >= c->opt_atoms_per_cell || (k[0] == 1 && k[1] == 1 && k[2] == 1)) break;
            int dmax = 0;
            for (int d = 1; d < 3; ++d)
                if (k[d] > k[dmax] || (k[d] == k[dmax] && tref[d] / k[d] < tref[dmax] / k[dmax])) dmax = d;
            if (k[dmax] == 1) break;
            --k[dmax];
        }
    }
    int hx_auto = 6;  // x slices per home tile (swept on B200)
    // optional per-dimension overrides and the x slicing of the home tile
    if (c->opt_subdiv_xyz[0] > 0) k[0] = std::min(c->opt_subdiv_xyz[0], 8);
    if (c->opt_subdiv_xyz[1] > 0) k[1] = std::min(c->opt_subdiv_xyz[1], 8);
    if (c->opt_subdiv_xyz[2] > 0) k[2] = std::min(c->opt_subdiv_xyz[2], 8);
    int hx = c->opt_slice_x > 0 ? std::min(c->opt_slice_x, 8) : hx_auto;
    for (int attempt = 0; attempt < 16; ++attempt) {
        // fine-cell lattice: tiles of k[] per reference cell, each tile sliced hx times along x
        const int kf[3] = {k[0] * hx, k[1], k[2]};
        double Mf[3][3];
        for (int r = 0; r < 3; ++r)
            for (int col = 0; col < 3; ++col) Mf[r][col] = Mref[r][col] / kf[col];
        double tf[3];
        thickness(Mf, tf);
        int R[3];
        bool ok = true;
        for (int d = 0; d < 3; ++d) {
            R[d] = (int)std::floor(rc_cull / tf[d]) + 1;
            // never need to look past the adjacent reference cells
            R[d] = std::min(R[d], 2 * kf[d] - 1);
            if (R[d] < 1) R[d] = 1;
            int fdd = g.dims[d] * kf[d];
            bool per = g.periodic_variant && ((g.pbc >> d) & 1u);
            if (d > 0 && R[d] > MAX_REACH) ok = false;
            if (d == 0 && R[d] + hx > 100) ok = false;
            if (per && fdd < 2 * (R[d] + (d == 0 ? hx : 0)) + 1) ok = false;
        }
        size_t ncells = (size_t)g.dims[0] * kf[0] * g.dims[1] * kf[1] * g.dims[2] * kf[2];
        if (ncells > ((size_t)1 << 27)) ok = false;
        if (!ok) {
            if (hx > 1) {
                --hx;
                continue;
            }
            int dmax = 0;
            for (int d = 1; d < 3; ++d)
                if (k[d] > k[dmax]) dmax = d;
            if (k[dmax] == 1) return;  // cannot satisfy: general kernel
            --k[dmax];
            continue;
        }
        // Neighbour offset table: rows (dy,dz) with the contiguous range of fine cells [dxlo,dxhi]
        // (relative to the first cell of the home tile) that can hold an atom within the cutoff of
        // an atom of the home tile.  Only the positive half-space of offsets is kept (dz>0, or dz==0
        // and dy>0, or the home row itself to the right of the tile): every unordered pair of atoms
        // in different tiles is then visited from exactly one side, with or without periodic
        // wrapping (offsets are unique modulo the grid because fd >= 2(R+hx)+1).
        int nrows = 0;
        const bool full = pl.full_shell;
        for (int dz = (full ? -R[2] : 0); dz <= R[2]; ++dz)
            for (int dy = ((dz == 0 && !full) ? 0 : -R[1]); dy <= R[1]; ++dy) {
                int lo = 127, hi = -128;
                for (int dx = -R[0]; dx <= hx - 1 + R[0]; ++dx) {
                    int delta[3] = {dx, dy, dz};
                    if (min_cell_dist2(Mf, delta, hx) <= rc_cull * rc_cull) {
                        lo = std::min(lo, dx);
                        hi = std::max(hi, dx);
                    }
                }
                if (dz == 0 && dy == 0 && !full) lo = std::max(lo, hx);
                if (lo <= hi) {
                    pl.rows[nrows].dy = (signed char)dy;
                    pl.rows[nrows].dz = (signed char)dz;
                    pl.rows[nrows].dxlo = (signed char)lo;
                    pl.rows[nrows].dxhi = (signed char)hi;
                    ++nrows;
                }
            }
        pl.nrows = nrows;
        for (int d = 0; d < 3; ++d) {
            g.k[d] = kf[d];
            g.kmagic[d] = (unsigned)(0x100000000ull / (unsigned long long)kf[d]) + 1u;
            g.fd[d] = g.dims[d] * kf[d];
        }
        g.hx = hx;
        pl.ncells = ncells;
        pl.use_cells = true;
        // Shifted-image filter for wrapped cell pairs (process_run KIND 2).  Sound when
        //  (a) every periodic dim has >= 4 reference cells, so the reference's round() of the
        //      fractional difference of two atoms in wrapped-adjacent cells is always +-1, and
        //  (b) every periodic dim's perpendicular box width exceeds 2.01*cutoff, so no OTHER image
        //      (including the triclinic corrections, periodic_box.rs:299-317) can be within cutoff.
        // Then d2 from the shifted difference and the reference's d2 approximate the same real
        // number; |difference| <= eps (bound below), and only tests inside [rc2-eps, rc2+eps] are
        // re-evaluated with the reference's exact arithmetic.
        pl.fast_pbc = 0;
        pl.rc2_lo = pl.rc2_hi = cutoff * cutoff;
        if (g.periodic_variant && !c->opt_exact_pbc) {
            bool ok2 = true;
            double Mb[3][3], tb[3];
            for (int r = 0; r < 3; ++r)
                for (int col = 0; col < 3; ++col) Mb[r][col] = g.box.m[r * 3 + col];
            thickness(Mb, tb);
            double lsum = 0, kappa = 1.0;
            for (int r = 0; r < 3; ++r) {
                double rowsum = 0;
                for (int col = 0; col < 3; ++col) {
                    lsum += std::fabs((double)g.box.m[r * 3 + col]);
                    double acc = 0;
                    for (int q = 0; q < 3; ++q) acc += std::fabs((double)g.box.m[r * 3 + q]) * std::fabs((double)g.box.inv[q * 3 + col]);
                    rowsum += acc;
                }
                kappa = std::max(kappa, rowsum);
            }
            for (int d = 0; d < 3; ++d)
                if ((g.pbc >> d) & 1u) {
                    if (g.dims[d] < 4) ok2 = false;
                    if (!(tb[d] > 2.01 * rc)) ok2 = false;
                }
            const double u = 5.9604644775390625e-08;
            double E = 64.0 * u * kappa * (lsum + rc);
            double eps = 2.0 * rc * E + E * E;
            double rc2d = (double)(cutoff * cutoff);
            if (!(eps < 0.004 * rc2d)) ok2 = false;
            if (ok2) {
                pl.fast_pbc = 1;
                pl.rc2_lo = (float)(rc2d - eps) * (1.0f - 2e-7f);
                pl.rc2_hi = (float)(rc2d + eps) * (1.0f + 2e-7f);
            }
        }
        return;
    }
}

// plan_cells solves ~10^4 tiny QPs: memoise per (box, cutoff, pbc, n, options)
struct PlanKey {
    float cutoff;
    unsigned pbc;
    size_t n;
    float m[9];
    int subdiv, brute, exact_pbc, sx, sy, sz, slice, full;
    double apc;
};
struct PlanCache {
    bool valid = false;
    PlanKey key;
    Plan plan;
};
void free_plan_cache(Ctx* c) {
    delete static_cast<PlanCache*>(c->plan_cache);
    c->plan_cache = nullptr;
}

static int upload_ids(Ctx* c, DevBuf& buf, const uint64_t* ids, size_t n, const unsigned long long** out) {
    *out = nullptr;
    if (!ids) return MB_OK;
    MB_TRY(buf.reserve(n * sizeof(uint64_t)));
    MB_CUDA(cudaMemcpyAsync(buf.p, ids, n * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
    *out = buf.as<unsigned long long>();
    return MB_OK;
}

static int check_sel(const Ctx*, const uint64_t* ids, size_t n, size_t n_atoms, const char* what) {
    return validate_sel(ids, n, n_atoms, what);
}

// bounds of a selection on the device -> host (one small sync)
static int device_minmax(Ctx* c, const float* xyz, const unsigned long long* d_ids, size_t n, float init_lo,
                         float init_hi, float lo[3], float hi[3]) {
    MB_TRY(c->counters.reserve(256));
    float* d6 = c->counters.as<float>() + 32;
    float init[6] = {init_lo, init_lo, init_lo, init_hi, init_hi, init_hi};
    MB_CUDA(cudaMemcpyAsync(d6, init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
    int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)c->sm_count * 8);
    minmax_kernel<<<blocks, 256, 0, c->stream>>>(xyz, d_ids, (int)n, init_lo, init_hi, d6);
    c->launches++;
    float out[6];
    MB_CUDA(cudaMemcpyAsync(out, d6, sizeof(out), cudaMemcpyDeviceToHost, c->stream));
    MB_CUDA(cudaStreamSynchronize(c->stream));
    for (int d = 0; d < 3; ++d) {
        lo[d] = out[d];
        hi[d] = out[3 + d];
    }
    return MB_OK;
}

static int make_grid_pbc(const Ctx* c, float cutoff, uint8_t pbc, GridSpec& g) {
    if (!c->has_box) return fail(MB_ERR_NO_PBC, "periodic search requested but the frame has no box");
    float ext[3];
    host_box_lab_extents(c->box, ext);
    g.periodic_variant = 1;
    g.pbc = pbc;
    for (int d = 0; d < 3; ++d) {
        g.dims[d] = dims_from_extent(ext[d], cutoff);
        g.lower[d] = 0;
        g.dim_sz[d] = 1;
    }
    g.box = to_dev_box(c->box);
    return MB_OK;
}

// periodic-variant plan for the context's current box, memoised
static int get_plan_pbc(Ctx* c, float cutoff, uint8_t pbc, size_t n, Plan& out, bool full_shell = false) {
    if (!c->has_box) return fail(MB_ERR_NO_PBC, "periodic search requested but the frame has no box");
    PlanKey k;
    memset(&k, 0, sizeof(k));
    k.cutoff = cutoff;
    k.pbc = pbc;
    k.n = n;
    for (int r = 0; r < 3; ++r)
        for (int col = 0; col < 3; ++col) k.m[r * 3 + col] = c->box.m[r][col];
    k.subdiv = c->opt_subdiv;
    k.brute = c->opt_force_brute;
    k.exact_pbc = c->opt_exact_pbc;
    k.sx = c->opt_subdiv_xyz[0];
    k.sy = c->opt_subdiv_xyz[1];
    k.sz = c->opt_subdiv_xyz[2];
    k.slice = c->opt_slice_x;
    k.full = full_shell ? 1 : 0;
    k.apc = c->opt_atoms_per_cell;
    PlanCache* pc = static_cast<PlanCache*>(c->plan_cache);
    if (!pc) {
        pc = new PlanCache;
        c->plan_cache = pc;
    }
    if (pc->valid && memcmp(&pc->key, &k, sizeof(k)) == 0) {
        out = pc->plan;
        return MB_OK;
    }
    memset(&out, 0, sizeof(out));
    out.full_shell = full_shell;
    MB_TRY(make_grid_pbc(c, cutoff, pbc, out.g));
    plan_cells(c, out, cutoff, n);
    pc->key = k;
    pc->plan = out;
    pc->valid = true;
    return MB_OK;
}

static void make_grid_bounds(float cutoff, const float lo[3], const float hi[3], GridSpec& g) {
    g.periodic_variant = 0;
    g.pbc = 0;
    memset(&g.box, 0, sizeof(g.box));
    for (int d = 0; d < 3; ++d) {
        g.lower[d] = lo[d];
        g.dim_sz[d] = hi[d] - lo[d];
        g.dims[d] = dims_from_extent(g.dim_sz[d], cutoff);
    }
}

// pads by (-cutoff - EPSILON, cutoff + EPSILON)   (distance_search.rs:633-634,643-644)
static void pad_bounds(float cutoff, float lo[3], float hi[3]) {
    float dl = -cutoff - FLT_EPSILON, du = cutoff + FLT_EPSILON;
    for (int d = 0; d < 3; ++d) {
        lo[d] = lo[d] + dl;
        hi[d] = hi[d] + du;
    }
}

static int ensure_pair_capacity(Ctx* c, size_t want, bool with_dist) {
    if (want < 1024) want = 1024;
    if (want > c->pair_cap || !c->pairs.p) {
        MB_TRY(c->pairs.reserve((want + PAIR_TAIL) * sizeof(uint2)));
        c->pair_cap = c->pairs.cap / sizeof(uint2) - PAIR_TAIL;
    }
    if (with_dist) MB_TRY(c->dists.reserve(c->pair_cap * sizeof(float)));
    return MB_OK;
}

static int bin_set(Ctx* c, const float* xyz, const unsigned long long* d_ids, size_t n, const GridSpec& g,
                   DevBuf& tmp4, DevBuf& cellid, DevBuf* rank, unsigned* cell_count, DevBuf* refcell,
                   int local_ids = 0) {
    MB_TRY(tmp4.reserve(n * sizeof(float4)));
    MB_TRY(cellid.reserve(n * sizeof(unsigned)));
    if (rank) MB_TRY(rank->reserve(n * sizeof(unsigned)));
    if (refcell) MB_TRY(refcell->reserve(n * sizeof(unsigned long long)));
    int blocks = (int)((n + 255) / 256);
    bin_atoms_kernel<<<blocks, 256, 0, c->stream>>>(xyz, d_ids, (int)n, g, tmp4.as<float4>(), cellid.as<unsigned>(),
                                                   rank ? rank->as<unsigned>() : nullptr, cell_count,
                                                   refcell ? refcell->as<unsigned long long>() : nullptr, local_ids);
    c->launches++;
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

template <int MODE>
static int launch_search_cells(Ctx* c, const SearchParams& P) {
    size_t smem = SEARCH_WARPS * sizeof(WarpShared);
    if (MODE == 0 || MODE == 1) smem += SEARCH_WARPS * STAGE_ROOM * sizeof(uint2);
    if (MODE == 1) smem += SEARCH_WARPS * STAGE_CAP * sizeof(float);
    int per_sm = 1;
    constexpr bool CAN_VDW = MODE == 0 || MODE == 1;
    const bool vdw = CAN_VDW && P.vdwA != nullptr;
    auto kern = vdw ? search_cells_kernel<MODE, CAN_VDW> : search_cells_kernel<MODE, false>;
    MB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SEARCH_WARPS * 32, smem));
    if (per_sm < 1) per_sm = 1;
    int blocks = c->sm_count * per_sm;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (c->opt_profile) {
        MB_CUDA(cudaEventCreate(&e0));
        MB_CUDA(cudaEventCreate(&e1));
        MB_CUDA(cudaEventRecord(e0, c->stream));
    }
    if (blocks * SEARCH_WARPS > PAIR_TAIL) blocks = PAIR_TAIL / SEARCH_WARPS;  // one tail slot per warp
    kern<<<blocks, SEARCH_WARPS * 32, smem, c->stream>>>(P);
    c->launches++;
    if (c->opt_profile) {
        MB_CUDA(cudaEventRecord(e1, c->stream));
        c->prof_events.push_back(e0);
        c->prof_events.push_back(e1);
    }
    if (MODE == 0) {
        merge_pair_tail_kernel<<<1, 256, 0, c->stream>>>(P.pairs, P.pair_cap, P.counter);
        c->launches++;
    }
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

// Enqueue the whole cell-path search for the frame at `xyz` (no host sync).  Counter block layout
// at d_counter: [0] pairs found, [1] work counter.
int enqueue_cells_search(Ctx* c, const float* xyz, const unsigned long long* d_ids, size_t n, const Plan& pl,
                         float cutoff, int mode, unsigned long long* d_counter) {
    const GridSpec& g = pl.g;
    MB_TRY(c->cell_count.reserve((pl.ncells + 1) * sizeof(unsigned)));
    MB_TRY(c->cell_start.reserve((pl.ncells + 2) * sizeof(unsigned)));
    MB_TRY(c->sorted4.reserve((n + PAD_CANDS) * sizeof(float4)));
    MB_CUDA(cudaMemsetAsync(c->cell_count.p, 0, (pl.ncells + 1) * sizeof(unsigned), c->stream));
    MB_CUDA(cudaMemsetAsync(d_counter, 0, CNT_STRIDE * sizeof(unsigned long long), c->stream));
    MB_TRY(bin_set(c, xyz, d_ids, n, g, c->tmp4a, c->cellid_a, &c->rank_a, c->cell_count.as<unsigned>(), nullptr));
    MB_TRY(exclusive_scan_u32(c, c->cell_count.as<unsigned>(), (int)pl.ncells, c->cell_start.as<unsigned>()));
    scatter_kernel<<<(int)((n + 255) / 256), 256, 0, c->stream>>>(c->tmp4a.as<float4>(), c->cellid_a.as<unsigned>(),
                                                                 c->rank_a.as<unsigned>(), c->cell_start.as<unsigned>(),
                                                                 (int)n, c->sorted4.as<float4>());
    c->launches++;
    SearchParams P;
    P.sorted = c->sorted4.as<float4>();
    P.cell_start = c->cell_start.as<unsigned>();
    P.sortedB = P.sorted;
    P.cell_startB = P.cell_start;
    P.two_sets = 0;
    P.flags = nullptr;
    P.vdwA = P.vdwB = nullptr;
    P.g = g;
    P.rc2 = cutoff * cutoff;
    P.band = 16.0f * 5.9604645e-8f * P.rc2;
    P.one = 1.0f;
    P.rc2_lo = pl.rc2_lo;
    P.rc2_hi = pl.rc2_hi;
    P.fast_pbc = pl.fast_pbc;
    P.nrows = pl.nrows;
    memcpy(P.rows, pl.rows, sizeof(NbrRow) * pl.nrows);
    P.dx_min = 127;
    P.dx_max = -128;
    for (int r = 0; r < pl.nrows; ++r) {
        P.dx_min = std::min(P.dx_min, (int)pl.rows[r].dxlo);
        P.dx_max = std::max(P.dx_max, (int)pl.rows[r].dxhi);
    }
    P.pairs = c->pairs.as<uint2>();
    P.dists = c->dists.as<float>();
    P.pair_cap = c->pair_cap;
    P.counter = d_counter;
    P.n_sortedB = n;
    if (mode == 2) return launch_search_cells<2>(c, P);
    if (mode == 1) return launch_search_cells<1>(c, P);
    return launch_search_cells<0>(c, P);
}

// Two-set variant (double / within): set 1 supplies the home tiles, set 2 the candidate stream, both
// binned on the same fine grid; the neighbour table covers the full shell.  kmode 0/1 pairs, 3 flags.
static int enqueue_cells_search2(Ctx* c, const float* xyz1, const unsigned long long* d_ids1, size_t n1,
                                 const float* xyz2, const unsigned long long* d_ids2, size_t n2, const Plan& pl,
                                 float cutoff, int kmode, unsigned long long* d_counter,
                                 const float* d_vdw1 = nullptr, const float* d_vdw2 = nullptr) {
    const GridSpec& g = pl.g;
    const size_t cb = (pl.ncells + 2) * sizeof(unsigned);
    MB_TRY(c->cell_count.reserve(cb));
    MB_TRY(c->cell_start.reserve(cb));
    MB_TRY(c->cell_count_b.reserve(cb));
    MB_TRY(c->cell_start_b.reserve(cb));
    MB_TRY(c->sorted4.reserve((n1 + PAD_CANDS) * sizeof(float4)));
    MB_TRY(c->sorted4_b.reserve((n2 + PAD_CANDS) * sizeof(float4)));
    MB_CUDA(cudaMemsetAsync(c->cell_count.p, 0, (pl.ncells + 1) * sizeof(unsigned), c->stream));
    MB_CUDA(cudaMemsetAsync(c->cell_count_b.p, 0, (pl.ncells + 1) * sizeof(unsigned), c->stream));
    MB_CUDA(cudaMemsetAsync(d_counter, 0, CNT_STRIDE * sizeof(unsigned long long), c->stream));
    const int local_ids = d_vdw1 ? 1 : 0;  // vdW searches report LOCAL indices, which also index the radii
    MB_TRY(bin_set(c, xyz1, d_ids1, n1, g, c->tmp4a, c->cellid_a, &c->rank_a, c->cell_count.as<unsigned>(), nullptr, local_ids));
    MB_TRY(bin_set(c, xyz2, d_ids2, n2, g, c->tmp4b, c->cellid_b, &c->rank_b, c->cell_count_b.as<unsigned>(), nullptr, local_ids));
    MB_TRY(exclusive_scan_u32(c, c->cell_count.as<unsigned>(), (int)pl.ncells, c->cell_start.as<unsigned>()));
    MB_TRY(exclusive_scan_u32(c, c->cell_count_b.as<unsigned>(), (int)pl.ncells, c->cell_start_b.as<unsigned>()));
    scatter_kernel<<<(int)((n1 + 255) / 256), 256, 0, c->stream>>>(c->tmp4a.as<float4>(), c->cellid_a.as<unsigned>(),
                                                                  c->rank_a.as<unsigned>(), c->cell_start.as<unsigned>(),
                                                                  (int)n1, c->sorted4.as<float4>());
    scatter_kernel<<<(int)((n2 + 255) / 256), 256, 0, c->stream>>>(c->tmp4b.as<float4>(), c->cellid_b.as<unsigned>(),
                                                                  c->rank_b.as<unsigned>(), c->cell_start_b.as<unsigned>(),
                                                                  (int)n2, c->sorted4_b.as<float4>());
    c->launches += 2;
    SearchParams P;
    P.sorted = c->sorted4.as<float4>();
    P.cell_start = c->cell_start.as<unsigned>();
    P.sortedB = c->sorted4_b.as<float4>();
    P.cell_startB = c->cell_start_b.as<unsigned>();
    P.two_sets = 1;
    P.flags = c->flags.as<unsigned char>();
    P.vdwA = d_vdw1;
    P.vdwB = d_vdw2;
    P.g = g;
    P.rc2 = cutoff * cutoff;
    P.band = 16.0f * 5.9604645e-8f * P.rc2;
    P.one = 1.0f;
    P.rc2_lo = pl.rc2_lo;
    P.rc2_hi = pl.rc2_hi;
    P.fast_pbc = pl.fast_pbc;
    P.nrows = pl.nrows;
    memcpy(P.rows, pl.rows, sizeof(NbrRow) * pl.nrows);
    P.dx_min = 127;
    P.dx_max = -128;
    for (int r = 0; r < pl.nrows; ++r) {
        P.dx_min = std::min(P.dx_min, (int)pl.rows[r].dxlo);
        P.dx_max = std::max(P.dx_max, (int)pl.rows[r].dxhi);
    }
    P.pairs = c->pairs.as<uint2>();
    P.dists = c->dists.as<float>();
    P.pair_cap = c->pair_cap;
    P.counter = d_counter;
    P.n_sortedB = n2;
    if (kmode == 3) return launch_search_cells<3>(c, P);
    if (kmode == 1) return launch_search_cells<1>(c, P);
    return launch_search_cells<0>(c, P);
}

static int enqueue_brute(Ctx* c, const Plan& pl, float cutoff, int mode, int with_dist, int count_only, size_t na,
                         size_t nb, const float4* a4, const unsigned long long* aref, const float4* b4,
                         const unsigned long long* bref, unsigned long long* d_counter,
                         const float* avdw = nullptr, const float* bvdw = nullptr) {
    BruteParams P;
    P.avdw = avdw;
    P.bvdw = bvdw;
    P.a4 = a4;
    P.aref = aref;
    P.na = (int)na;
    P.b4 = b4;
    P.bref = bref;
    P.nb = (int)nb;
    P.mode = mode;
    P.with_dist = with_dist;
    for (int d = 0; d < 3; ++d) P.dims[d] = pl.g.dims[d];
    P.pbc = pl.g.pbc;
    P.periodic_variant = pl.g.periodic_variant;
    P.rc2 = cutoff * cutoff;
    P.box = pl.g.box;
    P.pairs = c->pairs.as<uint2>();
    P.dists = c->dists.as<float>();
    P.pair_cap = c->pair_cap;
    P.counter = d_counter;
    P.flags = c->flags.as<unsigned char>();
    P.count_only = count_only;
    int bx = (int)((na + BRUTE_THREADS - 1) / BRUTE_THREADS);
    int tiles = (int)((nb + BRUTE_TILE - 1) / BRUTE_TILE);
    int by = std::max(1, std::min(tiles, (c->sm_count * 8 + bx - 1) / bx));
    by = std::min(by, 65535);
    brute_kernel<<<dim3(bx, by), BRUTE_THREADS, 0, c->stream>>>(P);
    c->launches++;
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

static size_t estimate_pairs(const Plan& pl, size_t n1, size_t n2, float cutoff, bool single) {
    double vol = pl.volume > 0 ? pl.volume : 0;
    if (!(vol > 0)) {
        // bounding volume of the reference grid
        if (pl.g.periodic_variant) {
            const float* m = pl.g.box.m;
            vol = std::fabs((double)m[0] * (m[4] * m[8] - m[5] * m[7]) - (double)m[1] * (m[3] * m[8] - m[5] * m[6]) +
                            (double)m[2] * (m[3] * m[7] - m[4] * m[6]));
        } else {
            vol = (double)pl.g.dim_sz[0] * pl.g.dim_sz[1] * pl.g.dim_sz[2];
        }
    }
    double sphere = 4.18879 * (double)cutoff * cutoff * cutoff;
    double est = vol > 0 ? (double)n1 * (double)n2 * sphere / vol : (double)n1 * n2;
    if (single) est *= 0.5;
    double maxp = single ? 0.5 * (double)n1 * (double)(n1 - 1) : (double)n1 * (double)n2;
    est = std::min(est * 1.15 + 4096.0, maxp + 16.0);
    return (size_t)est;
}

// mode: 0 pairs(+dist per option), 2 count only
int search_single_impl(Ctx* c, float cutoff, const uint64_t* ids, size_t n, uint8_t pbc, int mode,
                       int64_t* count_out) {
    if (!c->d_xyz) return fail(MB_ERR_STATE, "no frame set");
    if (!(cutoff > 0.0f)) return fail(MB_ERR_ARG, "cutoff must be positive");
    MB_CUDA(cudaSetDevice(c->device));
    MB_TRY(check_sel(c, ids, n, c->n_atoms, "search_single"));
    if (n > 0x7fffffffull) return fail(MB_ERR_ARG, "selection too large");
    const unsigned long long* d_ids;
    MB_TRY(upload_ids(c, c->ids1, ids, n, &d_ids));
    MB_TRY(c->counters.reserve(256));
    unsigned long long* d_counter = c->counters.as<unsigned long long>();

    Plan pl;
    memset(&pl, 0, sizeof(pl));
    if (pbc) {
        MB_TRY(get_plan_pbc(c, cutoff, pbc, n, pl));
    } else {
        float lo[3], hi[3];
        MB_TRY(device_minmax(c, c->d_xyz, d_ids, n, 0.0f, 0.0f, lo, hi));  // compute_min_max starts at 0 (:603-604)
        pad_bounds(cutoff, lo, hi);
        make_grid_bounds(cutoff, lo, hi, pl.g);
        plan_cells(c, pl, cutoff, n);
    }
    const bool with_dist = mode != 2 && c->opt_with_dist;
    const int kmode = mode == 2 ? 2 : (with_dist ? 1 : 0);
    if (mode != 2) MB_TRY(ensure_pair_capacity(c, std::max(estimate_pairs(pl, n, n, cutoff, true), c->pair_cap), with_dist));

    unsigned long long found = 0;
    for (int attempt = 0; attempt < 3; ++attempt) {
        if (pl.use_cells) {
            MB_TRY(enqueue_cells_search(c, c->d_xyz, d_ids, n, pl, cutoff, kmode, d_counter));
        } else {
            MB_CUDA(cudaMemsetAsync(d_counter, 0, CNT_STRIDE * sizeof(unsigned long long), c->stream));
            MB_TRY(bin_set(c, c->d_xyz, d_ids, n, pl.g, c->tmp4a, c->cellid_a, nullptr, nullptr, &c->refcell_a));
            MB_TRY(enqueue_brute(c, pl, cutoff, 0, with_dist, mode == 2, n, n, c->tmp4a.as<float4>(),
                                 c->refcell_a.as<unsigned long long>(), c->tmp4a.as<float4>(),
                                 c->refcell_a.as<unsigned long long>(), d_counter));
        }
        MB_CUDA(cudaMemcpyAsync(&found, d_counter, sizeof(found), cudaMemcpyDeviceToHost, c->stream));
        MB_CUDA(cudaStreamSynchronize(c->stream));
        c->harvest_profile();
        if (mode == 2 || found <= c->pair_cap) break;
        MB_TRY(ensure_pair_capacity(c, (size_t)found + 1024, with_dist));  // exact size known now: rerun
    }
    c->last.kind = mode == 2 ? 4 : 1;
    c->last.count = (int64_t)found;
    c->last.has_dist = with_dist;
    for (int d = 0; d < 3; ++d) c->last.grid_dims[d] = pl.g.dims[d];
    *count_out = (int64_t)found;
    return MB_OK;
}

// Neighbour list of ONE selection written by the search itself (SearchConnectivity, connectivity.rs:8-38, without a
// pair list in between): the cell kernel runs over the full neighbour shell twice — MODE 4 counts the neighbours of
// every atom (bit-sliced per-lane counters, one atomic per home atom and visit), an exclusive scan turns the counts into
// row starts, MODE 5 writes every atom's row at its own cursor (the rows of a home tile belong to one warp: no atomics).
// Layout in c->conn_tmp: deg / cursors [n_index + 2] | row_ptr [n_index + 2] | extra_bytes of caller scratch; rows in
// c->conn_cols.  *done = false (nothing computed) when the grid is not one the cell kernel handles.
int neighbor_rows_cells(Ctx* c, float cutoff, const uint64_t* ids, size_t n, uint8_t pbc, size_t n_index,
                        size_t extra_bytes, bool* done) {
    *done = false;
    if (!c->d_xyz) return fail(MB_ERR_STATE, "no frame set");
    if (!(cutoff > 0.0f)) return fail(MB_ERR_ARG, "cutoff must be positive");
    if (n_index < c->n_atoms || n_index > 0x7fffffffull) return fail(MB_ERR_ARG, "connectivity: bad index space");
    MB_CUDA(cudaSetDevice(c->device));
    MB_TRY(check_sel(c, ids, n, c->n_atoms, "search_connectivity"));
    if (n > 0x7fffffffull) return fail(MB_ERR_ARG, "selection too large");
    const unsigned long long* d_ids;
    MB_TRY(upload_ids(c, c->ids1, ids, n, &d_ids));
    MB_TRY(c->counters.reserve(256));
    unsigned long long* d_counter = c->counters.as<unsigned long long>();
    Plan pl;
    memset(&pl, 0, sizeof(pl));
    if (pbc) {
        MB_TRY(get_plan_pbc(c, cutoff, pbc, n, pl, true));
    } else {
        float lo[3], hi[3];
        MB_TRY(device_minmax(c, c->d_xyz, d_ids, n, 0.0f, 0.0f, lo, hi));  // compute_min_max starts at 0 (:603-604)
        pad_bounds(cutoff, lo, hi);
        make_grid_bounds(cutoff, lo, hi, pl.g);
        pl.full_shell = true;
        plan_cells(c, pl, cutoff, n);
    }
    if (!pl.use_cells) return MB_OK;
    const GridSpec& g = pl.g;
    MB_TRY(c->cell_count.reserve((pl.ncells + 1) * sizeof(unsigned)));
    MB_TRY(c->cell_start.reserve((pl.ncells + 2) * sizeof(unsigned)));
    MB_TRY(c->sorted4.reserve((n + PAD_CANDS) * sizeof(float4)));
    const size_t words = 2 * (n_index + 2);
    MB_TRY(c->conn_tmp.reserve(words * sizeof(unsigned) + extra_bytes));
    unsigned* deg = c->conn_tmp.as<unsigned>();
    unsigned* rp = deg + (n_index + 2);
    MB_CUDA(cudaMemsetAsync(c->cell_count.p, 0, (pl.ncells + 1) * sizeof(unsigned), c->stream));
    MB_CUDA(cudaMemsetAsync(d_counter, 0, CNT_STRIDE * sizeof(unsigned long long), c->stream));
    MB_CUDA(cudaMemsetAsync(deg, 0, (n_index + 2) * sizeof(unsigned), c->stream));
    MB_TRY(bin_set(c, c->d_xyz, d_ids, n, g, c->tmp4a, c->cellid_a, &c->rank_a, c->cell_count.as<unsigned>(), nullptr));
    MB_TRY(exclusive_scan_u32(c, c->cell_count.as<unsigned>(), (int)pl.ncells, c->cell_start.as<unsigned>()));
    scatter_kernel<<<(int)((n + 255) / 256), 256, 0, c->stream>>>(c->tmp4a.as<float4>(), c->cellid_a.as<unsigned>(),
                                                                 c->rank_a.as<unsigned>(), c->cell_start.as<unsigned>(),
                                                                 (int)n, c->sorted4.as<float4>());
    c->launches++;
    SearchParams P;
    memset(&P, 0, sizeof(P));
    P.sorted = c->sorted4.as<float4>();
    P.cell_start = c->cell_start.as<unsigned>();
    P.sortedB = P.sorted;
    P.cell_startB = P.cell_start;
    P.two_sets = 1;  // full shell, no self run: the home atoms meet themselves in the stream and are masked out there
    P.g = g;
    P.rc2 = cutoff * cutoff;
    P.band = 16.0f * 5.9604645e-8f * P.rc2;
    P.one = 1.0f;
    P.rc2_lo = pl.rc2_lo;
    P.rc2_hi = pl.rc2_hi;
    P.fast_pbc = pl.fast_pbc;
    P.nrows = pl.nrows;
    memcpy(P.rows, pl.rows, sizeof(NbrRow) * pl.nrows);
    P.dx_min = 127;
    P.dx_max = -128;
    for (int r = 0; r < pl.nrows; ++r) {
        P.dx_min = std::min(P.dx_min, (int)pl.rows[r].dxlo);
        P.dx_max = std::max(P.dx_max, (int)pl.rows[r].dxhi);
    }
    P.counter = d_counter;
    P.n_sortedB = n;
    P.nl_deg = deg;
    MB_TRY(launch_search_cells<4>(c, P));
    MB_TRY(exclusive_scan_u32(c, deg, (int)n_index, rp));
    MB_CUDA(cudaMemcpyAsync(deg, rp, n_index * sizeof(unsigned), cudaMemcpyDeviceToDevice, c->stream));  // row cursors
    unsigned total = 0;
    unsigned long long total64 = 0;
    MB_CUDA(cudaMemcpyAsync(&total, rp + n_index, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
    MB_CUDA(cudaMemcpyAsync(&total64, d_counter, sizeof(total64), cudaMemcpyDeviceToHost, c->stream));
    MB_CUDA(cudaMemsetAsync(d_counter, 0, CNT_STRIDE * sizeof(unsigned long long), c->stream));  // tile work counter
    MB_CUDA(cudaStreamSynchronize(c->stream));
    if (total64 > 0xfffffff0ull) return fail(MB_ERR_ARG, "connectivity: more than 2^32 entries");
    if (total64 != total) return fail(MB_ERR_STATE, "connectivity: degree count and row offsets disagree");
    MB_TRY(c->conn_cols.reserve(((size_t)total + 1) * sizeof(unsigned)));
    P.nl_cols = c->conn_cols.as<unsigned>();
    if (total) MB_TRY(launch_search_cells<5>(c, P));
    c->harvest_profile();
    c->conn_n = n_index;
    c->conn_nnz = total;
    c->last.kind = 4;  // no pair list on the context
    c->last.count = (int64_t)(total / 2);
    c->last.has_dist = false;
    for (int d = 0; d < 3; ++d) c->last.grid_dims[d] = pl.g.dims[d];
    *done = true;
    return MB_OK;
}

// double / within: general kernel over two binned sets
static int search_two_sets(Ctx* c, float cutoff, const uint64_t* ids1, size_t n1, const uint64_t* ids2, size_t n2,
                           int use_frame2, uint8_t pbc, int within, const float* lower3, const float* upper3,
                           int64_t* count_out, const float* vdw1 = nullptr, const float* vdw2 = nullptr) {
    if (!c->d_xyz) return fail(MB_ERR_STATE, "no frame set");
    const bool vdw = vdw1 && vdw2;
    if (vdw) {
        // grid cutoff = max(vdw1) + max(vdw2) + EPSILON   (distance_search.rs:781-783,845-847)
        if (n1 == 0 || n2 == 0) return fail(MB_ERR_ARG, "vdw search: empty selection");
        float m1 = vdw1[0], m2 = vdw2[0];
        for (size_t k = 1; k < n1; ++k) m1 = std::fmax(m1, vdw1[k]);
        for (size_t k = 1; k < n2; ++k) m2 = std::fmax(m2, vdw2[k]);
        cutoff = (m1 + m2) + FLT_EPSILON;
    }
    if (!(cutoff > 0.0f)) return fail(MB_ERR_ARG, "cutoff must be positive");
    MB_CUDA(cudaSetDevice(c->device));
    const float* xyz2 = use_frame2 ? c->xyz2.as<float>() : c->d_xyz;
    size_t natoms2 = use_frame2 ? c->n_atoms2 : c->n_atoms;
    if (use_frame2 && !c->xyz2.p) return fail(MB_ERR_STATE, "frame2 not set");
    MB_TRY(check_sel(c, ids1, n1, c->n_atoms, "search set 1"));
    MB_TRY(check_sel(c, ids2, n2, natoms2, "search set 2"));
    const unsigned long long *d_ids1, *d_ids2;
    MB_TRY(upload_ids(c, c->ids1, ids1, n1, &d_ids1));
    MB_TRY(upload_ids(c, c->ids2, ids2, n2, &d_ids2));
    MB_TRY(c->counters.reserve(256));
    unsigned long long* d_counter = c->counters.as<unsigned long long>();

    Plan pl;
    memset(&pl, 0, sizeof(pl));
    if (pbc) {
        MB_TRY(make_grid_pbc(c, cutoff, pbc, pl.g));
    } else {
        float lo[3], hi[3];
        if (within && lower3 && upper3) {
            for (int d = 0; d < 3; ++d) {
                lo[d] = lower3[d];
                hi[d] = upper3[d];
            }
        } else if (within) {
            // Measure::min_max of set 1 (measure.rs:22-36), padded as the `within` AST node does
            MB_TRY(device_minmax(c, c->d_xyz, d_ids1, n1, FLT_MAX, -FLT_MAX, lo, hi));
            pad_bounds(cutoff, lo, hi);
        } else {
            float lo2[3], hi2[3];
            MB_TRY(device_minmax(c, c->d_xyz, d_ids1, n1, 0.0f, 0.0f, lo, hi));
            MB_TRY(device_minmax(c, xyz2, d_ids2, n2, 0.0f, 0.0f, lo2, hi2));
            for (int d = 0; d < 3; ++d) {
                lo[d] = std::fmin(lo[d], lo2[d]);
                hi[d] = std::fmax(hi[d], hi2[d]);
            }
            pad_bounds(cutoff, lo, hi);
        }
        make_grid_bounds(cutoff, lo, hi, pl.g);
    }
    const bool with_dist = !within && c->opt_with_dist;
    // Cell path when the all-pairs product is large and the grid is not degenerate; the general
    // all-pairs kernel otherwise (small sets, 1-2 reference cells in a periodic dim).
    pl.full_shell = true;
    bool cells = false;
    if ((double)n1 * (double)n2 > c->opt_two_set_cells_min && n1 <= 0x7fffffffull && n2 <= 0x7fffffffull) {
        if (pbc) {
            MB_TRY(get_plan_pbc(c, cutoff, pbc, std::max(n1, (size_t)4096), pl, true));
        } else {
            plan_cells(c, pl, cutoff, std::max(n1, (size_t)4096));
        }
        cells = pl.use_cells;
    }
    if (!cells)
        for (int d = 0; d < 3; ++d) {
            pl.g.k[d] = 1;
            pl.g.fd[d] = pl.g.dims[d];
            pl.g.hx = 1;
        }
    const float *d_vdw1 = nullptr, *d_vdw2 = nullptr;
    if (vdw) {
        MB_TRY(c->vdw_a.reserve(n1 * sizeof(float)));
        MB_TRY(c->vdw_b.reserve(n2 * sizeof(float)));
        MB_CUDA(cudaMemcpyAsync(c->vdw_a.p, vdw1, n1 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        MB_CUDA(cudaMemcpyAsync(c->vdw_b.p, vdw2, n2 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        d_vdw1 = c->vdw_a.as<float>();
        d_vdw2 = c->vdw_b.as<float>();
    }
    if (!cells) {
        MB_TRY(bin_set(c, c->d_xyz, d_ids1, n1, pl.g, c->tmp4a, c->cellid_a, nullptr, nullptr, &c->refcell_a, vdw ? 1 : 0));
        MB_TRY(bin_set(c, xyz2, d_ids2, n2, pl.g, c->tmp4b, c->cellid_b, nullptr, nullptr, &c->refcell_b, vdw ? 1 : 0));
    }
    unsigned long long found = 0;
    if (within) {
        // flags: by position in set 1 (all-pairs kernel) or by global atom id (cell kernel)
        const size_t nflags = cells ? c->n_atoms : n1;
        MB_TRY(c->flags.reserve(nflags + 16));
        MB_CUDA(cudaMemsetAsync(c->flags.p, 0, nflags, c->stream));
        if (cells) {
            MB_TRY(enqueue_cells_search2(c, c->d_xyz, d_ids1, n1, xyz2, d_ids2, n2, pl, cutoff, 3, d_counter));
        } else {
            MB_TRY(enqueue_brute(c, pl, cutoff, 2, 0, 0, n1, n2, c->tmp4a.as<float4>(),
                                 c->refcell_a.as<unsigned long long>(), c->tmp4b.as<float4>(),
                                 c->refcell_b.as<unsigned long long>(), d_counter));
        }
        // ordered compaction of the flagged ids (ids are sorted, so the output is sorted and unique);
        // the cell arrays of set 2 are free again and serve as scan scratch
        MB_TRY(c->cell_count_b.reserve((n1 + 1) * sizeof(unsigned)));
        MB_TRY(c->cell_start_b.reserve((n1 + 2) * sizeof(unsigned)));
        flags_to_u32_kernel<<<(int)((n1 + 255) / 256), 256, 0, c->stream>>>(c->flags.as<unsigned char>(), d_ids1, (int)n1,
                                                                           cells ? 1 : 0, c->cell_count_b.as<unsigned>());
        c->launches++;
        MB_TRY(exclusive_scan_u32(c, c->cell_count_b.as<unsigned>(), (int)n1, c->cell_start_b.as<unsigned>()));
        MB_TRY(c->out_ids.reserve((n1 + 1) * sizeof(unsigned long long)));
        compact_flags_kernel<<<(int)((n1 + 255) / 256), 256, 0, c->stream>>>(
            c->flags.as<unsigned char>(), c->cell_start_b.as<unsigned>(), d_ids1, (int)n1, cells ? 1 : 0,
            c->out_ids.as<unsigned long long>());
        c->launches++;
        unsigned total = 0;
        MB_CUDA(cudaMemcpyAsync(&total, c->cell_start_b.as<unsigned>() + n1, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
        MB_CUDA(cudaStreamSynchronize(c->stream));
        c->harvest_profile();
        found = total;
        c->last.kind = 3;
    } else {
        MB_TRY(ensure_pair_capacity(c, std::max(estimate_pairs(pl, n1, n2, cutoff, false), c->pair_cap), with_dist));
        for (int attempt = 0; attempt < 3; ++attempt) {
            if (cells) {
                MB_TRY(enqueue_cells_search2(c, c->d_xyz, d_ids1, n1, xyz2, d_ids2, n2, pl, cutoff, with_dist ? 1 : 0,
                                             d_counter, d_vdw1, d_vdw2));
            } else {
                MB_CUDA(cudaMemsetAsync(d_counter, 0, CNT_STRIDE * sizeof(unsigned long long), c->stream));
                MB_TRY(enqueue_brute(c, pl, cutoff, 1, with_dist, 0, n1, n2, c->tmp4a.as<float4>(),
                                     c->refcell_a.as<unsigned long long>(), c->tmp4b.as<float4>(),
                                     c->refcell_b.as<unsigned long long>(), d_counter, d_vdw1, d_vdw2));
            }
            MB_CUDA(cudaMemcpyAsync(&found, d_counter, sizeof(found), cudaMemcpyDeviceToHost, c->stream));
            MB_CUDA(cudaStreamSynchronize(c->stream));
            c->harvest_profile();
            if (found <= c->pair_cap) break;
            MB_TRY(ensure_pair_capacity(c, (size_t)found + 1024, with_dist));
        }
        c->last.kind = 2;
    }
    c->last.count = (int64_t)found;
    c->last.has_dist = with_dist;
    for (int d = 0; d < 3; ++d) c->last.grid_dims[d] = pl.g.dims[d];
    *count_out = (int64_t)found;
    return MB_OK;
}

// ---- batch entry (device-resident frames, PBC variant; no host sync per frame) ---------------
static void swap_slot(Ctx* c, SearchSlot& sl) {
    std::swap(c->tmp4a, sl.tmp4a);
    std::swap(c->cellid_a, sl.cellid_a);
    std::swap(c->rank_a, sl.rank_a);
    std::swap(c->cell_count, sl.cell_count);
    std::swap(c->cell_start, sl.cell_start);
    std::swap(c->sorted4, sl.sorted4);
    std::swap(c->scan_tmp, sl.scan_tmp);
    std::swap(c->pairs, sl.pairs);
    std::swap(c->dists, sl.dists);
    std::swap(c->pair_cap, sl.pair_cap);
    std::swap(c->stream, sl.stream);
}
// make slot `s` (0 = the context's own) the one the members refer to
static void install_slot(Ctx* c, int s) {
    if (c->installed_slot == s) return;
    if (c->installed_slot != 0) swap_slot(c, c->alt[c->installed_slot - 1]);
    if (s != 0) swap_slot(c, c->alt[s - 1]);
    c->installed_slot = s;
}

int batch_search_impl(Ctx* c, float cutoff, uint8_t pbc, size_t f0, size_t f1, int mode, int64_t* counts,
                      uint64_t* checksums2) {
    if (!c->batch.p || f1 > c->batch_frames || f0 >= f1) return fail(MB_ERR_ARG, "batch_search: bad frame range");
    if (!pbc || !c->has_box) return fail(MB_ERR_NO_PBC, "batch_search needs a periodic box");
    MB_CUDA(cudaSetDevice(c->device));
    const size_t n = c->batch_atoms, nf = f1 - f0;
    Plan pl;
    MB_TRY(get_plan_pbc(c, cutoff, pbc, n, pl));
    if (!pl.use_cells) return fail(MB_ERR_ARG, "batch_search: grid is degenerate for the cell kernel (dims %d %d %d)",
                                   pl.g.dims[0], pl.g.dims[1], pl.g.dims[2]);
    const int kmode = mode == 1 ? 2 : 0;
    // Frames alternate over NS slots (scratch + pair buffer + stream each): the bin/scan/scatter of
    // frame f+1 overlaps the search of frame f, and for small frames the tail of one search kernel
    // (fewer home tiles than resident warps) is back-filled by the next frame's.
    const int NS = c->opt_batch_streams > 0 ? std::min(c->opt_batch_streams, 3) : (n <= 300000 ? 3 : 2);
    install_slot(c, 0);
    cudaStream_t main_stream = c->stream;
    for (int s = 1; s < NS; ++s)
        if (!c->alt[s - 1].stream) MB_CUDA(cudaStreamCreateWithFlags(&c->alt[s - 1].stream, cudaStreamNonBlocking));
    if (!c->aux_event) MB_CUDA(cudaEventCreateWithFlags(&c->aux_event, cudaEventDisableTiming));
    const size_t want = kmode != 2 ? estimate_pairs(pl, n, n, cutoff, true) : 0;
    if (kmode != 2)
        for (int s = 0; s < NS; ++s) {
            install_slot(c, s);
            MB_TRY(ensure_pair_capacity(c, std::max(want, c->pair_cap), false));
        }
    install_slot(c, 0);
    // per-frame counters at stride CNT_STRIDE (pairs, work, tests, -) ; checksums after them
    MB_TRY(c->batch_tmp.reserve(nf * (CNT_STRIDE + 2) * sizeof(unsigned long long)));
    unsigned long long* d_cnt = c->batch_tmp.as<unsigned long long>();
    unsigned long long* d_chk = d_cnt + CNT_STRIDE * nf;
    if (checksums2) MB_CUDA(cudaMemsetAsync(d_chk, 0, nf * 2 * sizeof(unsigned long long), main_stream));
    std::vector<unsigned long long> h_cnt(CNT_STRIDE * nf);
    const bool dbg = getenv("MB_DEBUG_TIMING") != nullptr;
    for (int attempt = 0; attempt < 2; ++attempt) {
        auto t_a = std::chrono::steady_clock::now();
        // alternates start after everything already queued on the context stream
        MB_CUDA(cudaEventRecord(c->aux_event, main_stream));
        for (int s = 1; s < NS; ++s) MB_CUDA(cudaStreamWaitEvent(c->alt[s - 1].stream, c->aux_event, 0));
        for (size_t f = 0; f < nf; ++f) {
            install_slot(c, (int)(f % NS));
            const float* xyz = c->batch.as<float>() + (f0 + f) * n * 3;
            int rc = enqueue_cells_search(c, xyz, nullptr, n, pl, cutoff, kmode, d_cnt + CNT_STRIDE * f);
            if (rc < 0) {
                install_slot(c, 0);
                return rc;
            }
        }
        install_slot(c, 0);
        for (int s = 1; s < NS; ++s) {
            MB_CUDA(cudaEventRecord(c->aux_event, c->alt[s - 1].stream));
            MB_CUDA(cudaStreamWaitEvent(main_stream, c->aux_event, 0));
        }
        auto t_b = std::chrono::steady_clock::now();
        MB_CUDA(cudaMemcpyAsync(h_cnt.data(), d_cnt, CNT_STRIDE * nf * sizeof(unsigned long long), cudaMemcpyDeviceToHost, main_stream));
        MB_CUDA(cudaStreamSynchronize(main_stream));
        auto t_c = std::chrono::steady_clock::now();
        c->harvest_profile();
        if (dbg)
            fprintf(stderr, "[mb] batch_search attempt %d: enqueue %.3f ms, wait %.3f ms, frames %zu, cap %zu, ncells %zu k=(%d,%d,%d) hx=%d rows=%d streams=%d\n",
                    attempt, std::chrono::duration<double, std::milli>(t_b - t_a).count(),
                    std::chrono::duration<double, std::milli>(t_c - t_b).count(), nf, c->pair_cap, pl.ncells,
                    pl.g.k[0], pl.g.k[1], pl.g.k[2], pl.g.hx, pl.nrows, NS);
        unsigned long long mx = 0;
        for (size_t f = 0; f < nf; ++f) mx = std::max(mx, h_cnt[CNT_STRIDE * f]);
        size_t min_cap = c->pair_cap;
        for (int s = 1; s < NS; ++s) min_cap = std::min(min_cap, c->alt[s - 1].pair_cap);
        if (kmode == 2 || mx <= min_cap) break;
        for (int s = 0; s < NS; ++s) {
            install_slot(c, s);
            MB_TRY(ensure_pair_capacity(c, (size_t)mx + mx / 16 + 1024, false));
        }
        install_slot(c, 0);
    }
    if (counts)
        for (size_t f = 0; f < nf; ++f) counts[f] = (int64_t)h_cnt[CNT_STRIDE * f];
    if (checksums2 && kmode != 2) {
        // verification path (not the timed one): redo frame by frame on the context stream so the
        // checksum kernel knows the pair count
        for (size_t f = 0; f < nf; ++f) {
            const float* xyz = c->batch.as<float>() + (f0 + f) * n * 3;
            MB_TRY(enqueue_cells_search(c, xyz, nullptr, n, pl, cutoff, kmode, d_cnt + CNT_STRIDE * f));
            unsigned long long cnt = h_cnt[CNT_STRIDE * f];
            int blocks = (int)std::min<unsigned long long>((cnt + 255) / 256 + 1, (unsigned long long)c->sm_count * 16);
            checksum_kernel<<<blocks, 256, 0, c->stream>>>(c->pairs.as<uint2>(), cnt, 1, d_chk + 2 * f);
            c->launches++;
        }
        MB_CUDA(cudaMemcpyAsync(checksums2, d_chk, nf * 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
        MB_CUDA(cudaStreamSynchronize(c->stream));
    } else if (kmode != 2) {
        // the pair list of the LAST frame must be the one mb_fill_pairs / mb_pairs_device see
        const int sl = (int)((nf - 1) % NS);
        if (sl != 0) {
            std::swap(c->pairs, c->alt[sl - 1].pairs);
            std::swap(c->dists, c->alt[sl - 1].dists);
            std::swap(c->pair_cap, c->alt[sl - 1].pair_cap);
        }
    }
    c->last.kind = kmode == 2 ? 4 : 1;
    c->last.count = (int64_t)h_cnt[CNT_STRIDE * (nf - 1)];
    if (kmode == 2) {
        double t = 0;
        for (size_t f = 0; f < nf; ++f) t += (double)h_cnt[CNT_STRIDE * f + 2];
        c->last_tests_per_frame = t / (double)nf;
    }
    c->last.has_dist = false;
    for (int d = 0; d < 3; ++d) c->last.grid_dims[d] = pl.g.dims[d];
    return MB_OK;
}

// count-only search of one device frame, result left in a device counter (used by the pipeline)
int enqueue_count_frame(Ctx* c, const float* xyz, size_t n, float cutoff, uint8_t pbc, unsigned long long* d_counter2) {
    Plan pl;
    MB_TRY(get_plan_pbc(c, cutoff, pbc, n, pl));
    if (!pl.use_cells) return fail(MB_ERR_ARG, "pipeline: grid is degenerate for the cell kernel");
    return enqueue_cells_search(c, xyz, nullptr, n, pl, cutoff, 2, d_counter2);
}

}  // namespace mb

using namespace mb;

extern "C" {

int64_t mb_search_single(MbCtx* h, float cutoff, const uint64_t* ids, size_t n, uint8_t pbc_dims) {
    if (!h) return fail(MB_ERR_ARG, "null context");
    int64_t cnt = 0;
    int rc = search_single_impl(&h->c, cutoff, ids, n, pbc_dims & 7, 0, &cnt);
    return rc < 0 ? rc : cnt;
}

int64_t mb_count_single(MbCtx* h, float cutoff, const uint64_t* ids, size_t n, uint8_t pbc_dims) {
    if (!h) return fail(MB_ERR_ARG, "null context");
    int64_t cnt = 0;
    int rc = search_single_impl(&h->c, cutoff, ids, n, pbc_dims & 7, 2, &cnt);
    return rc < 0 ? rc : cnt;
}

int64_t mb_search_double(MbCtx* h, float cutoff, const uint64_t* ids1, size_t n1, const uint64_t* ids2, size_t n2,
                         int use_frame2, uint8_t pbc_dims) {
    if (!h) return fail(MB_ERR_ARG, "null context");
    int64_t cnt = 0;
    int rc = search_two_sets(&h->c, cutoff, ids1, n1, ids2, n2, use_frame2, pbc_dims & 7, 0, nullptr, nullptr, &cnt);
    return rc < 0 ? rc : cnt;
}

int64_t mb_search_double_vdw(MbCtx* h, const uint64_t* ids1, size_t n1, const float* vdw1, const uint64_t* ids2,
                             size_t n2, const float* vdw2, int use_frame2, uint8_t pbc_dims) {
    if (!h || !vdw1 || !vdw2) return fail(MB_ERR_ARG, "null argument");
    int64_t cnt = 0;
    int rc = search_two_sets(&h->c, 1.0f, ids1, n1, ids2, n2, use_frame2, pbc_dims & 7, 0, nullptr, nullptr, &cnt, vdw1,
                             vdw2);
    return rc < 0 ? rc : cnt;
}

int64_t mb_search_within(MbCtx* h, float cutoff, const uint64_t* ids1, size_t n1, const uint64_t* ids2, size_t n2,
                         int use_frame2, uint8_t pbc_dims, const float* lower3, const float* upper3) {
    if (!h) return fail(MB_ERR_ARG, "null context");
    int64_t cnt = 0;
    int rc = search_two_sets(&h->c, cutoff, ids1, n1, ids2, n2, use_frame2, pbc_dims & 7, 1, lower3, upper3, &cnt);
    return rc < 0 ? rc : cnt;
}

// Host only (no device needed): the traversal plan the cell kernels would use for this box / cutoff / selection
// size — reference grid, subdivision, neighbour-row table, filter band.  Exists so that the planning logic can be
// tested exhaustively on the CPU (tests/test_plan_host.py).
int mb_plan_describe(const float* box9_colmajor, float cutoff, uint8_t pbc_dims, size_t n, int full_shell, int out_int[16],
                     signed char* rows4_out, float out_band[2]) {
    if (!box9_colmajor || !out_int) return fail(MB_ERR_ARG, "null argument");
    if (!(cutoff > 0.0f)) return fail(MB_ERR_ARG, "cutoff must be positive");
    Ctx c;  // plain host state: no CUDA call is made on this path
    MB_TRY(host_box_from_colmajor(box9_colmajor, &c.box));
    c.has_box = true;
    // tuning studies (tools/model_steps.py): "subdiv_x=3,subdiv_y=4,subdiv_z=4,slice_x=3,atoms_per_cell=6"
    if (const char* e = getenv("MOLAR_B200_PLAN_OPTS")) {
        int v;
        double dv;
        if (const char* q = strstr(e, "subdiv_x=")) if (sscanf(q + 9, "%d", &v) == 1) c.opt_subdiv_xyz[0] = v;
        if (const char* q = strstr(e, "subdiv_y=")) if (sscanf(q + 9, "%d", &v) == 1) c.opt_subdiv_xyz[1] = v;
        if (const char* q = strstr(e, "subdiv_z=")) if (sscanf(q + 9, "%d", &v) == 1) c.opt_subdiv_xyz[2] = v;
        if (const char* q = strstr(e, "slice_x=")) if (sscanf(q + 8, "%d", &v) == 1) c.opt_slice_x = v;
        if (const char* q = strstr(e, "atoms_per_cell=")) if (sscanf(q + 15, "%lf", &dv) == 1) c.opt_atoms_per_cell = dv;
    }
    Plan pl;
    memset(&pl, 0, sizeof(pl));
    pl.full_shell = full_shell != 0;
    MB_TRY(make_grid_pbc(&c, cutoff, pbc_dims & 7, pl.g));
    plan_cells(&c, pl, cutoff, n);
    out_int[0] = pl.use_cells ? 1 : 0;
    for (int d = 0; d < 3; ++d) {
        out_int[1 + d] = pl.g.dims[d];
        out_int[4 + d] = pl.g.k[d];
        out_int[8 + d] = pl.g.fd[d];
    }
    out_int[7] = pl.g.hx;
    out_int[11] = pl.nrows;
    out_int[12] = pl.fast_pbc;
    out_int[13] = MAX_ROWS;
    out_int[14] = out_int[15] = 0;
    if (rows4_out)
        for (int r = 0; r < pl.nrows; ++r) {
            rows4_out[4 * r] = pl.rows[r].dy;
            rows4_out[4 * r + 1] = pl.rows[r].dz;
            rows4_out[4 * r + 2] = pl.rows[r].dxlo;
            rows4_out[4 * r + 3] = pl.rows[r].dxhi;
        }
    if (out_band) {
        out_band[0] = pl.rc2_lo;
        out_band[1] = pl.rc2_hi;
    }
    return MB_OK;
}

__global__ void widen_pairs_kernel(const uint2* __restrict__ in, unsigned long long n, ulonglong2* __restrict__ out) {
    unsigned long long k = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (k < n) {
        uint2 p = in[k];
        out[k] = make_ulonglong2(p.x, p.y);
    }
}

int mb_fill_pairs(MbCtx* h, uint64_t* ij, float* dist) {
    if (!h) return fail(MB_ERR_ARG, "null context");
    Ctx& c = h->c;
    if (c.last.kind != 1 && c.last.kind != 2) return fail(MB_ERR_STATE, "no pair list on this context");
    MB_CUDA(cudaSetDevice(c.device));
    size_t P = (size_t)c.last.count;
    if (P == 0) return MB_OK;
    if (ij) {
        // widen u32 -> usize on the host side of the copy, in pinned chunks
        const size_t chunk = (size_t)1 << 22;
        MB_TRY(c.pinned_reserve(chunk * sizeof(uint2)));
        uint2* hp = static_cast<uint2*>(c.h_pinned);
        for (size_t off = 0; off < P; off += chunk) {
            size_t m = std::min(chunk, P - off);
            MB_CUDA(cudaMemcpyAsync(hp, c.pairs.as<uint2>() + off, m * sizeof(uint2), cudaMemcpyDeviceToHost, c.stream));
            MB_CUDA(cudaStreamSynchronize(c.stream));
            if (c.last.kind == 1) {  // single search: canonical i < j (the device list is unordered)
                for (size_t k = 0; k < m; ++k) {
                    unsigned a = hp[k].x, b = hp[k].y;
                    ij[2 * (off + k)] = a < b ? a : b;
                    ij[2 * (off + k) + 1] = a < b ? b : a;
                }
            } else {
                for (size_t k = 0; k < m; ++k) {
                    ij[2 * (off + k)] = hp[k].x;
                    ij[2 * (off + k) + 1] = hp[k].y;
                }
            }
        }
    }
    if (dist) {
        if (!c.last.has_dist) return fail(MB_ERR_STATE, "distances were not computed (option with_dist=0)");
        MB_CUDA(cudaMemcpyAsync(dist, c.dists.p, P * sizeof(float), cudaMemcpyDeviceToHost, c.stream));
        MB_CUDA(cudaStreamSynchronize(c.stream));
    }
    return MB_OK;
}

__global__ void canonical_pairs_kernel(uint2* __restrict__ pairs, unsigned long long n) {
    unsigned long long k = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (k < n) {
        uint2 p = pairs[k];
        if (p.x > p.y) pairs[k] = make_uint2(p.y, p.x);
    }
}

// The pair list as the device holds it: u32 x 2 per pair (8 B — half the PCIe bytes of the usize form), single-set
// pairs made canonical (i < j) in place first.  ij32 should be page-locked for the copy to run at full PCIe speed.
int mb_fill_pairs_u32(MbCtx* h, uint32_t* ij32, float* dist) {
    if (!h) return fail(MB_ERR_ARG, "null context");
    Ctx& c = h->c;
    if (c.last.kind != 1 && c.last.kind != 2) return fail(MB_ERR_STATE, "no pair list on this context");
    MB_CUDA(cudaSetDevice(c.device));
    const size_t P = (size_t)c.last.count;
    if (P == 0) return MB_OK;
    if (ij32) {
        if (c.last.kind == 1) {
            canonical_pairs_kernel<<<(unsigned)((P + 255) / 256), 256, 0, c.stream>>>(c.pairs.as<uint2>(), P);
            c.launches++;
            MB_CUDA(cudaGetLastError());
        }
        MB_CUDA(cudaMemcpyAsync(ij32, c.pairs.p, P * sizeof(uint2), cudaMemcpyDeviceToHost, c.stream));
    }
    if (dist) {
        if (!c.last.has_dist) return fail(MB_ERR_STATE, "distances were not computed (option with_dist=0)");
        MB_CUDA(cudaMemcpyAsync(dist, c.dists.p, P * sizeof(float), cudaMemcpyDeviceToHost, c.stream));
    }
    MB_CUDA(cudaStreamSynchronize(c.stream));
    return MB_OK;
}

int mb_fill_ids(MbCtx* h, uint64_t* ids) {
    if (!h || !ids) return fail(MB_ERR_ARG, "null argument");
    Ctx& c = h->c;
    if (c.last.kind != 3) return fail(MB_ERR_STATE, "no id list on this context");
    MB_CUDA(cudaSetDevice(c.device));
    if (c.last.count == 0) return MB_OK;
    MB_CUDA(cudaMemcpyAsync(ids, c.out_ids.p, (size_t)c.last.count * sizeof(uint64_t), cudaMemcpyDeviceToHost, c.stream));
    MB_CUDA(cudaStreamSynchronize(c.stream));
    return MB_OK;
}

const void* mb_pairs_device(MbCtx* h, int64_t* n_pairs) {
    if (!h || (h->c.last.kind != 1 && h->c.last.kind != 2)) {
        if (n_pairs) *n_pairs = 0;
        return nullptr;
    }
    if (n_pairs) *n_pairs = h->c.last.count;
    return h->c.pairs.p;
}

int mb_pairs_checksum(MbCtx* h, uint64_t out2[2]) {
    if (!h || !out2) return fail(MB_ERR_ARG, "null argument");
    Ctx& c = h->c;
    if (c.last.kind != 1 && c.last.kind != 2) return fail(MB_ERR_STATE, "no pair list on this context");
    MB_CUDA(cudaSetDevice(c.device));
    MB_TRY(c.counters.reserve(256));
    unsigned long long* d2 = c.counters.as<unsigned long long>() + 8;
    MB_CUDA(cudaMemsetAsync(d2, 0, 2 * sizeof(unsigned long long), c.stream));
    unsigned long long cnt = (unsigned long long)c.last.count;
    int blocks = (int)std::min<unsigned long long>((cnt + 255) / 256 + 1, (unsigned long long)c.sm_count * 16);
    checksum_kernel<<<blocks, 256, 0, c.stream>>>(c.pairs.as<uint2>(), cnt, c.last.kind == 1 ? 1 : 0, d2);
    c.launches++;
    MB_CUDA(cudaMemcpyAsync(out2, d2, 2 * sizeof(uint64_t), cudaMemcpyDeviceToHost, c.stream));
    MB_CUDA(cudaStreamSynchronize(c.stream));
    return MB_OK;
}

int mb_last_grid_dims(MbCtx* h, uint64_t dims3[3]) {
    if (!h || !dims3) return fail(MB_ERR_ARG, "null argument");
    for (int d = 0; d < 3; ++d) dims3[d] = h->c.last.grid_dims[d];
    return MB_OK;
}

}  // extern "C"
