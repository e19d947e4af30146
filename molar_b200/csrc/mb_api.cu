// mb_api.cu — context, error plumbing, frame/mass residency (host side of the C ABI).
#include <cmath>
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "mb_common.cuh"

namespace mb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

// ---- PeriodicBox::from_matrix (periodic_box.rs:156-176), host f32, unfused -------------------
// This TU is compiled with -Xcompiler -ffp-contract=off.
static inline float nsq3(float x, float y, float z) {
    float a = x * x, b = y * y, c = z * z;
    return (a + b) + c;
}

int host_box_from_colmajor(const float* m9, HostBox* out) {
    HostBox& B = *out;
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) B.m[r][c] = m9[c * 3 + r];
    for (int c = 0; c < 3; ++c)
        if (std::sqrt(nsq3(B.m[0][c], B.m[1][c], B.m[2][c])) == 0.0f)
            return fail(MB_ERR_BOX, "zero length box vector");
    // nalgebra try_inverse 3x3: cofactors / determinant
    float m11 = B.m[0][0], m12 = B.m[0][1], m13 = B.m[0][2];
    float m21 = B.m[1][0], m22 = B.m[1][1], m23 = B.m[1][2];
    float m31 = B.m[2][0], m32 = B.m[2][1], m33 = B.m[2][2];
    float minor_m12_m23 = m22 * m33 - m32 * m23;
    float minor_m11_m23 = m21 * m33 - m31 * m23;
    float minor_m11_m22 = m21 * m32 - m31 * m22;
    float det = m11 * minor_m12_m23 - m12 * minor_m11_m23 + m13 * minor_m11_m22;
    if (det == 0.0f) return fail(MB_ERR_BOX, "box matrix inverse failed");
    B.inv[0][0] = minor_m12_m23 / det;
    B.inv[0][1] = (m13 * m32 - m33 * m12) / det;
    B.inv[0][2] = (m12 * m23 - m22 * m13) / det;
    B.inv[1][0] = -minor_m11_m23 / det;
    B.inv[1][1] = (m11 * m33 - m31 * m13) / det;
    B.inv[1][2] = (m13 * m21 - m23 * m11) / det;
    B.inv[2][0] = minor_m11_m22 / det;
    B.inv[2][1] = (m12 * m31 - m32 * m11) / det;
    B.inv[2][2] = (m11 * m22 - m21 * m12) / det;
    // build_tric_corrections (periodic_box.rs:25-66)
    B.ncorr = 0;
    if (B.m[0][1] == 0 && B.m[0][2] == 0 && B.m[1][0] == 0 && B.m[1][2] == 0 && B.m[2][0] == 0 && B.m[2][1] == 0)
        return MB_OK;
    float a[3] = {B.m[0][0], B.m[1][0], B.m[2][0]};
    float b[3] = {B.m[0][1], B.m[1][1], B.m[2][1]};
    float c[3] = {B.m[0][2], B.m[1][2], B.m[2][2]};
    auto nrm = [&](float sa, float sb, float sc, bool neg_a) {
        float v[3];
        for (int d = 0; d < 3; ++d) {
            float aa = neg_a ? -a[d] : a[d];
            float t = sb > 0 ? aa + b[d] : aa - b[d];
            v[d] = sc > 0 ? t + c[d] : t - c[d];
        }
        (void)sa;
        return std::sqrt(nsq3(v[0], v[1], v[2]));
    };
    float n1 = nrm(1, 1, 1, false), n2 = nrm(1, 1, -1, false), n3 = nrm(1, -1, 1, false), n4 = nrm(-1, 1, 1, true);
    float half_diag = 0.5f * std::fmax(std::fmax(std::fmax(n1, n2), n3), n4);
    float two_hd = 2.0f * half_diag;
    float bound2 = two_hd * two_hd;
    for (int i = -1; i <= 1; ++i)
        for (int j = -1; j <= 1; ++j)
            for (int k = -1; k <= 1; ++k) {
                if (i == 0 && j == 0 && k == 0) continue;
                float s[3];
                for (int d = 0; d < 3; ++d) {
                    float ia = (float)i * a[d], jb = (float)j * b[d], kc = (float)k * c[d];
                    s[d] = (ia + jb) + kc;
                }
                if (nsq3(s[0], s[1], s[2]) < bound2) {
                    for (int d = 0; d < 3; ++d) B.corr[B.ncorr][d] = s[d];
                    ++B.ncorr;
                }
            }
    return MB_OK;
}

void host_box_lab_extents(const HostBox& b, float out[3]) {  // periodic_box.rs:369-375
    for (int r = 0; r < 3; ++r) out[r] = (b.m[r][0] + b.m[r][1]) + b.m[r][2];
}

DevBox to_dev_box(const HostBox& b) {
    DevBox d;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            d.m[r * 3 + c] = b.m[r][c];
            d.inv[r * 3 + c] = b.inv[r][c];
        }
    d.ncorr = b.ncorr;
    for (int i = 0; i < 26; ++i)
        for (int k = 0; k < 3; ++k) d.corr[3 * i + k] = i < b.ncorr ? b.corr[i][k] : 0.0f;
    // Pruning data for the triclinic correction loop (mb_common.cuh, corrections_best): correction s can only give a
    // shorter vector than `start` if 2 start.s + |s|^2 < 0.  corr_thr[i] = -0.4995 |s_i|^2: start.s above it means the
    // real improvement is negative by more than 1e-3 |s|^2 — a thousand times the f32 rounding of the reference's
    // own comparison — so the correction is skipped without evaluating it.  rin2 = (0.499 min |s|)^2: inside that
    // sphere no correction at all can win.
    double smin2 = 1e300;
    for (int i = 0; i < 26; ++i) {
        double s2 = 0;
        for (int k = 0; k < 3; ++k) s2 += (double)d.corr[3 * i + k] * (double)d.corr[3 * i + k];
        d.corr_thr[i] = i < b.ncorr ? (float)(-0.4995 * s2) : 0.0f;
        if (i < b.ncorr) smin2 = std::min(smin2, s2);
    }
    d.rin2 = b.ncorr > 0 ? (float)(0.499 * 0.499 * smin2) : 0.0f;
    // the same corrections as 13 (+v, -v) pairs over the lattice coefficients (shortest_vector_dev): the reference's
    // list is generated for i, j, k = -1..1 in that nesting, minus (0,0,0), minus the ones beyond the diagonal bound —
    // replay the generation to find the index of every combination
    static const int PAIR_IJK[13][3] = {{0, 0, 1}, {0, 1, -1}, {0, 1, 0}, {0, 1, 1}, {1, -1, -1}, {1, -1, 0}, {1, -1, 1},
                                        {1, 0, -1}, {1, 0, 0}, {1, 0, 1}, {1, 1, -1}, {1, 1, 0}, {1, 1, 1}};
    for (int p = 0; p < 13; ++p) {
        d.pair_thr[p] = 0.0f;
        d.pair_bit[2 * p] = d.pair_bit[2 * p + 1] = 0u;
    }
    if (b.ncorr > 0) {
        const float a[3] = {b.m[0][0], b.m[1][0], b.m[2][0]}, bb[3] = {b.m[0][1], b.m[1][1], b.m[2][1]},
                    cc[3] = {b.m[0][2], b.m[1][2], b.m[2][2]};
        for (int i = -1; i <= 1; ++i)
            for (int j = -1; j <= 1; ++j)
                for (int k = -1; k <= 1; ++k) {
                    if (i == 0 && j == 0 && k == 0) continue;
                    float s[3];
                    for (int q = 0; q < 3; ++q) s[q] = ((float)i * a[q] + (float)j * bb[q]) + (float)k * cc[q];
                    int idx = -1;
                    for (int t = 0; t < b.ncorr; ++t)
                        if (b.corr[t][0] == s[0] && b.corr[t][1] == s[1] && b.corr[t][2] == s[2]) {
                            idx = t;
                            break;
                        }
                    if (idx < 0) continue;
                    for (int p = 0; p < 13; ++p) {
                        const int sg = (PAIR_IJK[p][0] == i && PAIR_IJK[p][1] == j && PAIR_IJK[p][2] == k)      ? 0
                                       : (PAIR_IJK[p][0] == -i && PAIR_IJK[p][1] == -j && PAIR_IJK[p][2] == -k) ? 1
                                                                                                                : -1;
                        if (sg < 0) continue;
                        d.pair_bit[2 * p + sg] = 1u << idx;
                        d.pair_thr[p] = d.corr_thr[idx];
                    }
                }
    }
    return d;
}

int DevBuf::reserve(size_t bytes) {
    if (bytes <= cap) return MB_OK;
    if (p) {
        cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        p = nullptr;
        return fail(MB_ERR_CUDA, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    }
    cap = want;
    return MB_OK;
}
void DevBuf::release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}

// Fold finished begin/end event pairs into the running totals (call after a stream sync).
int Ctx::harvest_profile() {
    for (size_t i = 0; i + 1 < prof_events.size(); i += 2) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, prof_events[i], prof_events[i + 1]) == cudaSuccess) {
            prof_search_ms += ms;
            prof_search_launches++;
        }
        cudaEventDestroy(prof_events[i]);
        cudaEventDestroy(prof_events[i + 1]);
    }
    prof_events.clear();
    return MB_OK;
}

int Ctx::host_results() {
    if (h_res) return MB_OK;
    void* hp = nullptr;
    cudaError_t e = cudaHostAlloc(&hp, 64 * sizeof(double), cudaHostAllocMapped);
    if (e != cudaSuccess) return fail(MB_ERR_CUDA, "cudaHostAlloc(mapped results) failed: %s", cudaGetErrorString(e));
    void* dp = nullptr;
    e = cudaHostGetDevicePointer(&dp, hp, 0);
    if (e != cudaSuccess) {
        cudaFreeHost(hp);
        return fail(MB_ERR_CUDA, "cudaHostGetDevicePointer failed: %s", cudaGetErrorString(e));
    }
    h_res = static_cast<double*>(hp);
    d_res = static_cast<double*>(dp);
    return MB_OK;
}

int Ctx::pinned_reserve(size_t bytes) {
    if (bytes <= h_pinned_cap) return MB_OK;
    if (h_pinned) cudaFreeHost(h_pinned);
    h_pinned = nullptr;
    h_pinned_cap = 0;
    size_t want = bytes + bytes / 4 + 4096;
    cudaError_t e = cudaMallocHost(&h_pinned, want);
    if (e != cudaSuccess) return fail(MB_ERR_CUDA, "cudaMallocHost(%zu) failed: %s", want, cudaGetErrorString(e));
    h_pinned_cap = want;
    return MB_OK;
}

int validate_sel(const uint64_t* ids, size_t n, size_t n_atoms, const char* what) {
    if (n == 0) return fail(MB_ERR_ARG, "%s: empty selection", what);
    if (n > 0x7fffffffull) return fail(MB_ERR_ARG, "%s: selection too large", what);
    if (!ids) {
        if (n > n_atoms) return fail(MB_ERR_ARG, "%s: identity selection of %zu > %zu atoms", what, n, n_atoms);
        return MB_OK;
    }
    if (ids[0] >= n_atoms) return fail(MB_ERR_ARG, "%s: index %llu out of range (%zu atoms)", what, (unsigned long long)ids[0], n_atoms);
    uint64_t bad = 0;  // branch-free scan: any non-increasing step sets it
    for (size_t k = 1; k < n; ++k) bad |= (uint64_t)(ids[k] <= ids[k - 1]);
    if (bad) return fail(MB_ERR_ARG, "%s: selection indices must be strictly increasing (a sorted set)", what);
    if (ids[n - 1] >= n_atoms)
        return fail(MB_ERR_ARG, "%s: index %llu out of range (%zu atoms)", what, (unsigned long long)ids[n - 1], n_atoms);
    return MB_OK;
}

}  // namespace mb

using namespace mb;

extern "C" {

const char* mb_last_error(void) { return g_err; }
int mb_abi_version(void) { return 1; }

MbCtx* mb_open(int device) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error("no CUDA device available (%s); libmolar_b200 has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return nullptr;
    }
    if (device < 0 || device >= ndev) {
        set_error("device %d out of range (%d devices)", device, ndev);
        return nullptr;
    }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        set_error("cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
        return nullptr;
    }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    if (prop.major < 10) {
        set_error("device %d is sm_%d%d; libmolar_b200 is built for sm_100a only", device, prop.major, prop.minor);
        return nullptr;
    }
    MbCtx* h = new MbCtx;
    h->c.device = device;
    h->c.sm_count = prop.multiProcessorCount;
    e = cudaStreamCreateWithFlags(&h->c.stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        set_error("cudaStreamCreate: %s", cudaGetErrorString(e));
        delete h;
        return nullptr;
    }
    return h;
}

void mb_close(MbCtx* h) {
    if (!h) return;
    Ctx& c = h->c;
    cudaSetDevice(c.device);
    cudaStreamSynchronize(c.stream);
    c.harvest_profile();
    DevBuf* bufs[] = {&c.xyz_own, &c.xyz2, &c.masses, &c.ids1, &c.ids2, &c.tmp4a, &c.tmp4b, &c.cellid_a,
                      &c.cellid_b, &c.rank_a, &c.refcell_a, &c.refcell_b, &c.sorted4, &c.cell_count,
                      &c.cell_start, &c.scan_tmp, &c.pairs, &c.dists, &c.flags, &c.out_ids, &c.counters,
                      &c.reduce_tmp, &c.batch, &c.batch_scalars, &c.batch_tmp, &c.batch_ref, &c.pipe_tmp, &c.vdw_a, &c.vdw_b, &c.rank_b, &c.cell_count_b, &c.cell_start_b, &c.sorted4_b, &c.pbc_tmp, &c.traj_raw, &c.traj_aux, &c.conn_tmp, &c.conn_cols};
    for (DevBuf* b : bufs) b->release();
    for (SearchSlot& sl : c.alt) {
        DevBuf* sb[] = {&sl.tmp4a, &sl.cellid_a, &sl.rank_a, &sl.cell_count, &sl.cell_start, &sl.sorted4, &sl.scan_tmp,
                        &sl.pairs, &sl.dists};
        for (DevBuf* b : sb) b->release();
        if (sl.stream) cudaStreamDestroy(sl.stream);
    }
    free_plan_cache(&c);
    comm_destroy(&c);
    c.comm_send.release();
    c.comm_recv.release();
    c.many_tmp.release();
    for (cudaEvent_t& e : c.timer_ev)
        if (e) cudaEventDestroy(e);
    if (c.copy_stream) cudaStreamDestroy(c.copy_stream);
    if (c.h_pinned) cudaFreeHost(c.h_pinned);
    if (c.h_res) cudaFreeHost(c.h_res);
    for (int i = 0; i < 4; ++i)
        if (c.aux_stream[i]) cudaStreamDestroy(c.aux_stream[i]);
    if (c.aux_event) cudaEventDestroy(c.aux_event);
    cudaStreamDestroy(c.stream);
    delete h;
}

void* mb_stream(MbCtx* h) { return h ? (void*)h->c.stream : nullptr; }

int mb_synchronize(MbCtx* h) {
    if (!h) return fail(MB_ERR_ARG, "null context");
    MB_CUDA(cudaSetDevice(h->c.device));
    MB_CUDA(cudaStreamSynchronize(h->c.stream));
    return MB_OK;
}

int mb_set_option(MbCtx* h, const char* key, double value) {
    if (!h || !key) return fail(MB_ERR_ARG, "null argument");
    Ctx& c = h->c;
    if (!strcmp(key, "subdiv")) c.opt_subdiv = (int)value;
    else if (!strcmp(key, "subdiv_x")) c.opt_subdiv_xyz[0] = (int)value;
    else if (!strcmp(key, "subdiv_y")) c.opt_subdiv_xyz[1] = (int)value;
    else if (!strcmp(key, "subdiv_z")) c.opt_subdiv_xyz[2] = (int)value;
    else if (!strcmp(key, "slice_x")) c.opt_slice_x = (int)value;
    else if (!strcmp(key, "force_brute")) c.opt_force_brute = (int)value;
    else if (!strcmp(key, "atoms_per_cell")) c.opt_atoms_per_cell = value;
    else if (!strcmp(key, "with_dist")) c.opt_with_dist = (int)value;
    else if (!strcmp(key, "exact_pbc")) c.opt_exact_pbc = (int)value;
    else if (!strcmp(key, "fused_fit")) c.opt_fused_fit = (int)value;
    else if (!strcmp(key, "fit_lag")) c.opt_fit_lag = (int)value;
    else if (!strcmp(key, "fit_teams")) c.opt_fit_teams = (int)value;
    else if (!strcmp(key, "fit_group")) c.opt_fit_group = (int)value;
    else if (!strcmp(key, "fit_streams")) c.opt_fit_streams = (int)value;
    else if (!strcmp(key, "batch_streams")) c.opt_batch_streams = (int)value;
    else if (!strcmp(key, "two_set_cells_min")) c.opt_two_set_cells_min = value;
    else if (!strcmp(key, "profile")) {
        c.opt_profile = (int)value;
        c.prof_search_ms = 0.0;
        c.prof_search_launches = 0;
    }
    else return fail(MB_ERR_ARG, "unknown option '%s'", key);
    return MB_OK;
}

// Host only (no device needed): the periodic-box tables the kernels use — the reference's triclinic corrections
// (build_tric_corrections, periodic_box.rs:25-66) and the 13 (+v, -v) pair table with its thresholds that
// shortest_vector_dev prunes them with.  Exists so that the table logic can be tested on the CPU.
int mb_box_describe(const float* box9_colmajor, int* ncorr_out, float corr78_out[78], float pair_thr13_out[13],
                    uint32_t pair_bit26_out[26]) {
    if (!box9_colmajor || !ncorr_out) return fail(MB_ERR_ARG, "null argument");
    HostBox hb;
    MB_TRY(host_box_from_colmajor(box9_colmajor, &hb));
    const DevBox d = to_dev_box(hb);
    *ncorr_out = d.ncorr;
    if (corr78_out) memcpy(corr78_out, d.corr, sizeof(d.corr));
    if (pair_thr13_out) memcpy(pair_thr13_out, d.pair_thr, sizeof(d.pair_thr));
    if (pair_bit26_out) memcpy(pair_bit26_out, d.pair_bit, sizeof(d.pair_bit));
    return MB_OK;
}

static int set_box(Ctx& c, const float* box9) {
    if (box9) {
        MB_TRY(host_box_from_colmajor(box9, &c.box));
        c.has_box = true;
    } else {
        c.has_box = false;
    }
    return MB_OK;
}

int mb_set_frame(MbCtx* h, const float* xyz, size_t n_atoms, const float* box9) {
    if (!h || !xyz || n_atoms == 0) return fail(MB_ERR_ARG, "mb_set_frame: null/empty frame");
    Ctx& c = h->c;
    MB_CUDA(cudaSetDevice(c.device));
    MB_TRY(set_box(c, box9));
    size_t bytes = n_atoms * 3 * sizeof(float);
    MB_TRY(c.xyz_own.reserve(bytes));
    // host buffer may be pageable: cudaMemcpyAsync stages it; ordering is by stream
    MB_CUDA(cudaMemcpyAsync(c.xyz_own.p, xyz, bytes, cudaMemcpyHostToDevice, c.stream));
    MB_CUDA(cudaStreamSynchronize(c.stream));  // caller may free xyz right after return
    c.d_xyz = c.xyz_own.as<float>();
    c.n_atoms = n_atoms;
    return MB_OK;
}

int mb_set_frame_device(MbCtx* h, const float* xyz_dev, size_t n_atoms, const float* box9) {
    if (!h || !xyz_dev || n_atoms == 0) return fail(MB_ERR_ARG, "mb_set_frame_device: null/empty frame");
    Ctx& c = h->c;
    MB_TRY(set_box(c, box9));
    c.d_xyz = xyz_dev;
    c.n_atoms = n_atoms;
    return MB_OK;
}

int mb_get_frame(MbCtx* h, float* out, size_t n_atoms) {
    if (!h || !out) return fail(MB_ERR_ARG, "null argument");
    Ctx& c = h->c;
    if (!c.d_xyz || n_atoms != c.n_atoms) return fail(MB_ERR_ARG, "mb_get_frame: no frame or size mismatch");
    MB_CUDA(cudaSetDevice(c.device));
    MB_CUDA(cudaMemcpyAsync(out, c.d_xyz, n_atoms * 3 * sizeof(float), cudaMemcpyDeviceToHost, c.stream));
    MB_CUDA(cudaStreamSynchronize(c.stream));
    return MB_OK;
}

int mb_set_masses(MbCtx* h, const float* masses, size_t n_atoms) {
    if (!h || !masses || n_atoms == 0) return fail(MB_ERR_ARG, "mb_set_masses: null/empty");
    Ctx& c = h->c;
    MB_CUDA(cudaSetDevice(c.device));
    MB_TRY(c.masses.reserve(n_atoms * sizeof(float)));
    MB_CUDA(cudaMemcpyAsync(c.masses.p, masses, n_atoms * sizeof(float), cudaMemcpyHostToDevice, c.stream));
    MB_CUDA(cudaStreamSynchronize(c.stream));
    c.n_masses = n_atoms;
    return MB_OK;
}

int mb_set_frame2(MbCtx* h, const float* xyz, size_t n_atoms) {
    if (!h || !xyz || n_atoms == 0) return fail(MB_ERR_ARG, "mb_set_frame2: null/empty");
    Ctx& c = h->c;
    MB_CUDA(cudaSetDevice(c.device));
    MB_TRY(c.xyz2.reserve(n_atoms * 3 * sizeof(float)));
    MB_CUDA(cudaMemcpyAsync(c.xyz2.p, xyz, n_atoms * 3 * sizeof(float), cudaMemcpyHostToDevice, c.stream));
    MB_CUDA(cudaStreamSynchronize(c.stream));
    c.n_atoms2 = n_atoms;
    return MB_OK;
}

int mb_get_masses(MbCtx* h, float* out, size_t n_atoms) {
    if (!h || !out) return fail(MB_ERR_ARG, "null argument");
    Ctx& c = h->c;
    if (!c.masses.p || n_atoms != c.n_masses) return fail(MB_ERR_ARG, "mb_get_masses: no masses or size mismatch");
    MB_CUDA(cudaSetDevice(c.device));
    MB_CUDA(cudaMemcpyAsync(out, c.masses.p, n_atoms * sizeof(float), cudaMemcpyDeviceToHost, c.stream));
    MB_CUDA(cudaStreamSynchronize(c.stream));
    return MB_OK;
}

uint64_t mb_launch_count(MbCtx* h) { return h ? h->c.launches : 0; }

int mb_get_stat(MbCtx* h, const char* key, double* out) {
    if (!h || !key || !out) return fail(MB_ERR_ARG, "null argument");
    Ctx& c = h->c;
    if (!strcmp(key, "search_kernel_ms")) *out = c.prof_search_ms;
    else if (!strcmp(key, "search_kernel_launches")) *out = (double)c.prof_search_launches;
    else if (!strcmp(key, "pair_capacity")) *out = (double)c.pair_cap;
    else if (!strcmp(key, "sm_count")) *out = (double)c.sm_count;
    else if (!strcmp(key, "search_tests_per_frame")) *out = c.last_tests_per_frame;  // last count-only batch search
    else if (!strcmp(key, "traj_h2d_ms")) *out = c.traj_h2d_ms;        // last mb_batch_load_traj: raw bytes to the device
    else if (!strcmp(key, "traj_decode_ms")) *out = c.traj_decode_ms;  // ... all decode kernels (CUDA events)
    else if (!strcmp(key, "traj_scan_ms")) *out = c.traj_scan_ms;      // ... of which xtc_scan_kernel
    else if (!strcmp(key, "traj_raw_bytes")) *out = c.traj_raw_bytes;
    else return fail(MB_ERR_ARG, "unknown stat '%s'", key);
    return MB_OK;
}

}  // extern "C"
