// mb_common.cuh — shared declarations of libmolar_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include <string>
#include <vector>

#include "../../include/molar_b200.h"

namespace mb {

// ---------------------------------------------------------------------------------------------
// error plumbing: thread-local message, never throw across the ABI
// (precedent: molar_gromacs/gromacs/wrapper.cpp:32,139-158)
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);

#define MB_CUDA(call)                                                                         \
    do {                                                                                      \
        cudaError_t _e = (call);                                                              \
        if (_e != cudaSuccess)                                                                \
            return ::mb::fail(MB_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, \
                              cudaGetErrorString(_e));                                        \
    } while (0)

#define MB_TRY(expr)              \
    do {                          \
        int _rc = (expr);         \
        if (_rc < 0) return _rc;  \
    } while (0)

// ---------------------------------------------------------------------------------------------
// PeriodicBox on the host, f32, in the reference's evaluation order
// (molar/src/periodic_box.rs:25-66,156-176,369-375).  Row-major m[r][c]; columns = a,b,c.
// ---------------------------------------------------------------------------------------------
struct HostBox {
    float m[3][3];
    float inv[3][3];
    int ncorr;
    float corr[26][3];
};
// returns MB_OK or MB_ERR_BOX
int host_box_from_colmajor(const float* m9, HostBox* out);
void host_box_lab_extents(const HostBox& b, float out[3]);

// Box as kernels see it (passed by value inside kernel parameter structs).
struct DevBox {
    float m[9];    // row-major
    float inv[9];  // row-major
    int ncorr;
    float corr[26 * 3];
    float corr_thr[26];  // -0.4995 |s|^2 per correction: start.s above it => the correction cannot win (pruned)
    float rin2;          // (0.499 min |s|)^2: |start|^2 below it => no correction can win
    // The 26 lattice combinations as 13 pairs (+v, -v), v = i a + j b + k c in the order of PAIR_IJK (mb_common.cuh):
    // start.v is a sum of the three basis products, start.(-v) its negative, and both share one threshold.
    // pair_bit[2p] / [2p + 1]: bit (1 << index in corr[]) of +v / -v, 0 when the reference's list does not hold it.
    float pair_thr[13];
    unsigned pair_bit[26];
};
DevBox to_dev_box(const HostBox& b);

// ---------------------------------------------------------------------------------------------
// grow-only device buffer
// ---------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);  // keeps contents only if no reallocation was needed
    void release();
    template <class T>
    T* as() const { return static_cast<T*>(p); }
};

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
struct SearchResult {
    int kind = 0;  // 0 none, 1 pairs (single), 2 pairs (double), 3 ids (within), 4 count only
    int64_t count = 0;
    bool has_dist = false;
    uint64_t grid_dims[3] = {0, 0, 0};
};

// One set of per-frame search scratch + output + the stream it is used on.  The context's own
// members (tmp4a, ..., pairs, stream) are "slot 0"; batch_search installs the alternates by swapping
// them in, so consecutive frames are pre-processed and searched on different streams.
struct SearchSlot {
    DevBuf tmp4a, cellid_a, rank_a, cell_count, cell_start, sorted4, scan_tmp, pairs, dists;
    size_t pair_cap = 0;
    cudaStream_t stream = nullptr;
};

struct Ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t aux_stream[4] = {nullptr, nullptr, nullptr, nullptr};  // batch_fit alternates frame groups over these
    cudaEvent_t aux_event = nullptr;
    int sm_count = 148;
    uint64_t launches = 0;

    // current frame
    const float* d_xyz = nullptr;  // points into xyz_own, the batch, or adopted memory
    size_t n_atoms = 0;
    DevBuf xyz_own;
    bool has_box = false;
    HostBox box;
    // second coordinate set
    DevBuf xyz2;
    size_t n_atoms2 = 0;
    // masses
    DevBuf masses;
    size_t n_masses = 0;

    // selections uploaded for the current call
    DevBuf ids1, ids2;
    // pinned staging
    void* h_pinned = nullptr;
    size_t h_pinned_cap = 0;
    int pinned_reserve(size_t bytes);
    // Small results (a centre, a tensor, a transform) are written by the finishing thread of a reduction kernel straight
    // into mapped page-locked host memory: the call ends with the stream synchronisation, without a D2H copy behind it.
    double* h_res = nullptr;  // host view (64 doubles)
    double* d_res = nullptr;  // device view of the same memory
    int host_results();

    // search scratch
    DevBuf tmp4a, tmp4b;    // binned atoms (float4: eff pos + id bits), unsorted
    DevBuf cellid_a, cellid_b, rank_a;
    DevBuf vdw_a, vdw_b;          // vdW radii of the two selections (vdW search)
    DevBuf refcell_a, refcell_b;  // packed reference cell coords (u64) for the general kernel
    DevBuf sorted4;         // atoms sorted by fine cell
    DevBuf cell_count, cell_start, scan_tmp;
    DevBuf rank_b, cell_count_b, cell_start_b, sorted4_b;  // second set of a two-set cell search
    DevBuf pairs, dists, flags, out_ids;
    DevBuf counters;        // small block of device counters / results
    DevBuf reduce_tmp;      // per-block partials
    DevBuf pbc_tmp;         // ticket + results + partials of the periodic reductions (mb_measure_pbc.cu)
    SearchResult last;
    double last_tests_per_frame = 0.0;  // distance tests per frame of the last count-only batch search
    size_t pair_cap = 0;
    SearchSlot alt[2];        // alternate slots for batch_search (streams created lazily)
    int installed_slot = 0;   // which slot currently lives in the members above

    // options
    int opt_subdiv = 0;       // 0 auto, else forced k for all dims
    int opt_subdiv_xyz[3] = {0, 0, 0};  // per-dimension override of the tile subdivision
    int opt_slice_x = 0;      // home tile = this many fine cells along x (0 = automatic)
    int opt_force_brute = 0;  // force the general all-pairs kernel
    double opt_atoms_per_cell = 8.0;  // minimum mean population of a home tile
    int opt_with_dist = 1;
    double opt_two_set_cells_min = 5.0e7;  // two-set searches use the cell kernel when n1*n2 exceeds this
    int opt_batch_streams = 0;  // streams (slots) batch_search alternates frames over; 0 = automatic
    int opt_fused_fit = 0;  // batch_fit: 0 two kernels per frame group, 1/2 TMA-staged single-pass kernels, 3 persistent kernel with an L2-served lagging second pass
    int opt_fit_lag = 0;    // fused_fit = 3: frames (of a team) between pass 1 and pass 2 (0 = automatic)
    int opt_fit_group = 0;  // two-kernel batch_fit: frames per group (0 = automatic)
    int opt_fit_streams = 0;  // two-kernel batch_fit: streams the groups alternate over (0 = automatic, <= 4)
    int opt_fit_teams = 0;  // fused_fit = 3: teams of CTAs working on different frames at once (0 = automatic)
    int opt_exact_pbc = 0;    // 1: wrapped cell pairs always use the exact PeriodicBox path (no filter)
    int opt_profile = 0;      // record CUDA events around every search-kernel launch
    std::vector<cudaEvent_t> prof_events;  // begin/end pairs not yet harvested
    double prof_search_ms = 0.0;
    uint64_t prof_search_launches = 0;
    int harvest_profile();

    // batch (device-resident trajectory)
    DevBuf batch;
    size_t batch_frames = 0, batch_atoms = 0;
    bool batch_has_box = false;
    DevBuf batch_scalars;  // rows of doubles
    size_t batch_rows = 0, batch_row_doubles = 0;
    DevBuf batch_tmp, batch_ref, pipe_tmp;
    DevBuf conn_tmp, conn_cols;  // pair-list consumers: CSR adjacency / union-find / BFS scratch (mb_connect.cu)
    size_t conn_n = 0, conn_nnz = 0;
    DevBuf traj_raw, traj_aux;  // trajectory ingest: raw file bytes / decode tables (mb_traj.cu)
    double traj_h2d_ms = 0.0, traj_decode_ms = 0.0, traj_scan_ms = 0.0, traj_raw_bytes = 0.0;  // last load, CUDA events

    // memoised search plan (owned by mb_search.cu)
    void* plan_cache = nullptr;

    // multi-GPU (mb_comm.cu): NCCL communicator of the frame-sharded ranks, staging for the scalar gather
    void* nccl_comm = nullptr;
    int comm_rank = 0, comm_world = 1;
    DevBuf comm_send, comm_recv;
    DevBuf many_tmp;  // mb_reduce_many: ids | offsets | results | status
    cudaEvent_t timer_ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaStream_t copy_stream = nullptr;  // uploads of mb_stream_* (own stream: never queues behind the work streams)
};

// ---------------------------------------------------------------------------------------------
// exact-arithmetic device helpers.  Every product and sum on a decision path must round on
// its own (Rust does not contract to FMA): use the _rn intrinsics, which nvcc never fuses.
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ float xmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float xadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float xsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float xdiv(float a, float b) { return __fdiv_rn(a, b); }

// nalgebra gemv order: ((M_i0*v0) + M_i1*v1) + M_i2*v2   (periodic_box.rs:340-360)
__device__ __forceinline__ void xmatvec(const float* __restrict__ M, float v0, float v1, float v2,
                                        float& r0, float& r1, float& r2) {
    r0 = xadd(xadd(xmul(M[0], v0), xmul(M[1], v1)), xmul(M[2], v2));
    r1 = xadd(xadd(xmul(M[3], v0), xmul(M[4], v1)), xmul(M[5], v2));
    r2 = xadd(xadd(xmul(M[6], v0), xmul(M[7], v1)), xmul(M[8], v2));
}
// Vector3::norm_squared: (x*x + y*y) + z*z
__device__ __forceinline__ float xnorm2(float x, float y, float z) {
    return xadd(xadd(xmul(x, x), xmul(y, y)), xmul(z, z));
}
// (pos2 - pos1).norm_squared()   (distance_search.rs:446,460,488,507)
__device__ __forceinline__ float d2_direct(float ax, float ay, float az, float bx, float by, float bz) {
    return xnorm2(xsub(bx, ax), xsub(by, ay), xsub(bz, az));
}
// PeriodicBox::distance_squared(p1,p2,dims)   (periodic_box.rs:286-318,379-381)
__device__ __forceinline__ float d2_pbc(const DevBox& bx, float ax, float ay, float az, float bxx, float byy,
                                        float bzz, unsigned w) {
    float v0 = xsub(bxx, ax), v1 = xsub(byy, ay), v2 = xsub(bzz, az);
    float f0, f1, f2;
    xmatvec(bx.inv, v0, v1, v2, f0, f1, f2);
    if (w & 1u) f0 = xsub(f0, roundf(f0));
    if (w & 2u) f1 = xsub(f1, roundf(f1));
    if (w & 4u) f2 = xsub(f2, roundf(f2));
    float s0, s1, s2;
    xmatvec(bx.m, f0, f1, f2, s0, s1, s2);
    float best2 = xnorm2(s0, s1, s2);
    if (bx.ncorr == 0 || w != 7u) return best2;
    // (no pruning here: this copy is inlined into the pair kernel's rare wrapped path, where the extra live values
    //  cost the hot loop registers; the pruned loop lives in shortest_vector_dev, used by the reductions)
    for (int c = 0; c < bx.ncorr; ++c) {
        float n2 = xnorm2(xadd(s0, bx.corr[3 * c]), xadd(s1, bx.corr[3 * c + 1]), xadd(s2, bx.corr[3 * c + 2]));
        if (n2 < best2) best2 = n2;
    }
    return best2;
}
// PeriodicBox::shortest_vector_dims (periodic_box.rs:286-318), f32, unfused, nalgebra gemv order
__device__ __forceinline__ void shortest_vector_dev(const DevBox& bx, float v0, float v1, float v2, unsigned w,
                                                    float& o0, float& o1, float& o2) {
    float f0, f1, f2;
    xmatvec(bx.inv, v0, v1, v2, f0, f1, f2);
    if (w & 1u) f0 = xsub(f0, roundf(f0));
    if (w & 2u) f1 = xsub(f1, roundf(f1));
    if (w & 4u) f2 = xsub(f2, roundf(f2));
    float s0, s1, s2;
    xmatvec(bx.m, f0, f1, f2, s0, s1, s2);
    o0 = s0;
    o1 = s1;
    o2 = s2;
    if (bx.ncorr == 0 || w != 7u) return;
    float best2 = xnorm2(s0, s1, s2);
    // The reference evaluates all (up to 26) triclinic corrections (periodic_box.rs:299-317).  A correction v can beat
    // `start` only if 2 start.v + |v|^2 < 0.  Candidates are found from THREE dot products (start with the box vectors):
    // start.(i a + j b + k c) is a sum of them, its negative serves the opposite correction, and a correction whose
    // real improvement is negative by more than 1e-3 |v|^2 — a thousand times any f32 rounding involved — is dropped.
    // The surviving ones (a bit mask, usually empty) are evaluated with the reference's exact expression against the
    // running best in the reference's order, which gives the reference's result.
    const float da = fmaf(s0, bx.m[0], fmaf(s1, bx.m[3], s2 * bx.m[6]));
    const float db = fmaf(s0, bx.m[1], fmaf(s1, bx.m[4], s2 * bx.m[7]));
    const float dc = fmaf(s0, bx.m[2], fmaf(s1, bx.m[5], s2 * bx.m[8]));
    unsigned cand = 0u;
#define MB_PAIR(P, EXPR)                                   \
    {                                                      \
        const float d = (EXPR);                            \
        if (d < bx.pair_thr[P]) cand |= bx.pair_bit[2 * P];      \
        if (-d < bx.pair_thr[P]) cand |= bx.pair_bit[2 * P + 1]; \
    }
    MB_PAIR(0, dc)                // ( 0, 0, 1)
    MB_PAIR(1, db - dc)           // ( 0, 1,-1)
    MB_PAIR(2, db)                // ( 0, 1, 0)
    MB_PAIR(3, db + dc)           // ( 0, 1, 1)
    MB_PAIR(4, (da - db) - dc)    // ( 1,-1,-1)
    MB_PAIR(5, da - db)           // ( 1,-1, 0)
    MB_PAIR(6, (da - db) + dc)    // ( 1,-1, 1)
    MB_PAIR(7, da - dc)           // ( 1, 0,-1)
    MB_PAIR(8, da)                // ( 1, 0, 0)
    MB_PAIR(9, da + dc)           // ( 1, 0, 1)
    MB_PAIR(10, (da + db) - dc)   // ( 1, 1,-1)
    MB_PAIR(11, da + db)          // ( 1, 1, 0)
    MB_PAIR(12, (da + db) + dc)   // ( 1, 1, 1)
#undef MB_PAIR
    while (cand) {
        const int c = __ffs(cand) - 1;
        cand &= cand - 1u;
        const float c0 = xadd(s0, bx.corr[3 * c]), c1 = xadd(s1, bx.corr[3 * c + 1]), c2 = xadd(s2, bx.corr[3 * c + 2]);
        const float n2 = xnorm2(c0, c1, c2);
        if (n2 < best2) {
            best2 = n2;
            o0 = c0;
            o1 = c1;
            o2 = c2;
        }
    }
}
#endif  // __CUDACC__

// Selection check shared by every entry point that takes ids: non-empty, at most 2^31-1 entries, and — when ids are
// given — STRICTLY INCREASING and < n_atoms (a Sel's index vector is a sorted set, providers.rs:45-48; kernels rely
// on it for bounds and for the lower_bound in unwrap_connectivity).  One O(n) host pass, cheap next to the H2D copy.
int validate_sel(const uint64_t* ids, size_t n, size_t n_atoms, const char* what);

// implemented in the respective translation units
int search_single_impl(Ctx* c, float cutoff, const uint64_t* ids, size_t n, uint8_t pbc, int mode,
                       int64_t* count_out);
int batch_search_impl(Ctx* c, float cutoff, uint8_t pbc, size_t f0, size_t f1, int mode, int64_t* counts,
                      uint64_t* checksums2);
int enqueue_count_frame(Ctx* c, const float* xyz, size_t n, float cutoff, uint8_t pbc,
                        unsigned long long* d_counter2);
// neighbour list of one selection as CSR rows written by the search kernel itself (mb_search.cu)
int neighbor_rows_cells(Ctx* c, float cutoff, const uint64_t* ids, size_t n, uint8_t pbc, size_t n_index,
                        size_t extra_bytes, bool* done);
void free_plan_cache(Ctx* c);
void comm_destroy(Ctx* c);  // mb_comm.cu
// exclusive scan of n u32 on the context stream, out[n] = total (mb_search.cu)
int exclusive_scan_u32(Ctx* c, const unsigned* in, int n, unsigned* out);
int batch_fit_impl(Ctx* c, size_t ref_frame, size_t f0, size_t f1, int superpose, double* rmsd_out);
int enqueue_batch_moments(Ctx* c, size_t f0, size_t f1, double* d_rows8, double* partials, unsigned* tickets,
                          int nb);
}  // namespace mb

struct MbCtx {
    mb::Ctx c;
};
