// mb_measure_pbc.cu — periodic variants of the selection reductions, inertia tensor, principal axes.
//
// Replaces (molar/src/measure.rs):
//   center_of_geometry                       :37-45
//   center_of_geometry_pbc[_dims]            :142-168
//   center_of_mass_pbc[_dims]                :172-214
//   gyration_pbc                             :216-226
//   inertia / inertia_pbc + do_inertia       :88-98, 228-238, 573-610
//   principal_transform[_pbc]                :100-108, 240-252, 645-649
//
// The reference walks the selection serially, but nothing in these loops is a chain: every atom's
// image is taken relative to the FIRST atom (closest_image(c, p0), periodic_box.rs:321-330) or to
// the centre (shortest_vector(pos - c)), so each is one data-parallel pass:
//   pass A  image of every atom next to atom 0, evaluated in f32 with the reference's operation
//           order (so the choice of image is the reference's), accumulated in f64;
//   pass B  second moments of d = shortest_vector(pos - c) (or pos - c) in f64.
// The 3x3 symmetric eigenproblem is solved on the host (cyclic Jacobi, f64).
// Reference quirk kept: in center_of_mass_pbc the first atom enters the numerator with weight 1, not
// with its mass (`let mut cm = p0.coords`, :180,200), while the denominator holds every mass.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "mb_common.cuh"
#include "mb_reduce.cuh"

namespace mb {

constexpr int PRED_THREADS = 256;

struct PbcRedParams {
    const float* xyz;
    const unsigned long long* ids;
    int n;
    const float* masses;  // NULL: unit weights (centre of geometry)
    int use_box;          // pass A: images next to atom 0; pass B: d = shortest_vector(pos - c)
    unsigned w;           // PbcDims of pass A
    DevBox box;
    double c[3];          // pass B: centre
    double* partials;
    unsigned* ticket;
    double* results;
};

__device__ __forceinline__ size_t pgid(const PbcRedParams& P, int k) { return P.ids ? (size_t)P.ids[k] : (size_t)k; }

// pass A.  v: [0] sum of weights (every atom), [1..3] sum over k >= 1 of w_k * (im_k - p0), im_k the f32
// image the reference forms: p0 + shortest_vector_dims(p_k - p0).  results: centre[3], W, status.
__global__ void __launch_bounds__(PRED_THREADS) center_pbc_kernel(const __grid_constant__ PbcRedParams P) {
    const size_t g0 = pgid(P, 0);
    const float p0x = P.xyz[3 * g0], p0y = P.xyz[3 * g0 + 1], p0z = P.xyz[3 * g0 + 2];
    double v[4] = {0, 0, 0, 0};
    // four atoms per thread and round, all loads first: the kernel is a stream of 16 B per atom and what bounds it is
    // the number of loads in flight
    constexpr int U = 4;
    const int stride = gridDim.x * PRED_THREADS;
    for (int k0 = blockIdx.x * PRED_THREADS + threadIdx.x; k0 < P.n; k0 += U * stride) {
        float x[U], y[U], z[U], mf[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int k = k0 + u * stride;
            const size_t g = pgid(P, k < P.n ? k : k0);
            x[u] = P.xyz[3 * g];
            y[u] = P.xyz[3 * g + 1];
            z[u] = P.xyz[3 * g + 2];
            mf[u] = P.masses ? P.masses[g] : 1.0f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int k = k0 + u * stride;
            if (k >= P.n) continue;
            const double m = (double)mf[u];
            v[0] += m;
            if (k == 0) continue;
            float ix = x[u], iy = y[u], iz = z[u];
            if (P.use_box) {
                float s0, s1, s2;
                shortest_vector_dev(P.box, xsub(x[u], p0x), xsub(y[u], p0y), xsub(z[u], p0z), P.w, s0, s1, s2);
                ix = xadd(p0x, s0);
                iy = xadd(p0y, s1);
                iz = xadd(p0z, s2);
            }
            v[1] += m * ((double)ix - (double)p0x);
            v[2] += m * ((double)iy - (double)p0y);
            v[3] += m * ((double)iz - (double)p0z);
        }
    }
    __shared__ double res[4];
    if (!grid_reduce<4, PRED_THREADS>(v, P.partials, P.ticket, blockIdx.x, gridDim.x, res)) return;
    if (threadIdx.x == 0) {
        const double W = res[0];
        if (W == 0.0) {
            P.results[0] = P.results[1] = P.results[2] = nan("");
            P.results[3] = 0.0;
            P.results[4] = 1.0;
            return;
        }
        // reference: cm = p0 (weight ONE) + sum_{k>=1} w_k im_k ; centre = cm / W.  Without a box the same
        // formula with the first atom weighted properly (plain centre of geometry / mass).
        const double w0 = P.masses ? (double)P.masses[g0] : 1.0;
        const double first = P.use_box ? 1.0 : w0;
        const double f = (first + (W - w0)) / W;
        P.results[0] = (double)p0x * f + res[1] / W;
        P.results[1] = (double)p0y * f + res[2] / W;
        P.results[2] = (double)p0z * f + res[3] / W;
        P.results[3] = W;
        P.results[4] = 0.0;
    }
}

// pass B.  v: [0] sum m, [1..3] sum m d_a^2 (xx,yy,zz), [4..6] sum m d_a d_b (xy,xz,yz)
__global__ void __launch_bounds__(PRED_THREADS) tensor_kernel(const __grid_constant__ PbcRedParams P) {
    const float cx = (float)P.c[0], cy = (float)P.c[1], cz = (float)P.c[2];  // the reference's centre is a Pos (f32)
    double v[7] = {0, 0, 0, 0, 0, 0, 0};
    constexpr int U = 4;  // atoms per thread and round, loads first (see center_pbc_kernel)
    const int stride = gridDim.x * PRED_THREADS;
    for (int k0 = blockIdx.x * PRED_THREADS + threadIdx.x; k0 < P.n; k0 += U * stride) {
        float x[U], y[U], z[U], mf[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int k = k0 + u * stride;
            const size_t g = pgid(P, k < P.n ? k : k0);
            x[u] = P.xyz[3 * g];
            y[u] = P.xyz[3 * g + 1];
            z[u] = P.xyz[3 * g + 2];
            mf[u] = P.masses ? P.masses[g] : 1.0f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (k0 + u * stride >= P.n) continue;
            const double m = (double)mf[u];
            double dx, dy, dz;
            if (P.use_box) {
                float s0, s1, s2;
                shortest_vector_dev(P.box, xsub(x[u], cx), xsub(y[u], cy), xsub(z[u], cz), 7u, s0, s1, s2);
                dx = s0;
                dy = s1;
                dz = s2;
            } else {
                dx = (double)x[u] - P.c[0];
                dy = (double)y[u] - P.c[1];
                dz = (double)z[u] - P.c[2];
            }
            v[0] += m;
            v[1] += m * dx * dx;
            v[2] += m * dy * dy;
            v[3] += m * dz * dz;
            v[4] += m * dx * dy;
            v[5] += m * dx * dz;
            v[6] += m * dy * dz;
        }
    }
    __shared__ double res[7];
    if (!grid_reduce<7, PRED_THREADS>(v, P.partials, P.ticket, blockIdx.x, gridDim.x, res)) return;
    if (threadIdx.x < 7) P.results[threadIdx.x] = res[threadIdx.x];
}

// ---- host side ---------------------------------------------------------------------------------
constexpr size_t P_TICKET_BYTES = 256;

static int pbc_scratch(Ctx* c, int nb, int K, PbcRedParams& P) {
    // own scratch block: [ticket 256 B, zeroed once][results 256 B][partials]
    const size_t need = P_TICKET_BYTES + 256 + (size_t)nb * K * sizeof(double);
    const bool fresh = need > c->pbc_tmp.cap;
    MB_TRY(c->pbc_tmp.reserve(need));
    if (fresh) MB_CUDA(cudaMemsetAsync(c->pbc_tmp.p, 0, P_TICKET_BYTES, c->stream));
    char* base = static_cast<char*>(c->pbc_tmp.p);
    P.ticket = reinterpret_cast<unsigned*>(base);
    MB_TRY(c->host_results());
    P.results = c->d_res;  // mapped host memory: the finishing thread's stores ARE the transfer
    P.partials = reinterpret_cast<double*>(base + P_TICKET_BYTES + 256);
    return MB_OK;
}

static int pbc_prepare(Ctx* c, const uint64_t* ids, size_t n, bool need_mass, bool need_box, const char* what,
                       PbcRedParams& P, int* nb_out) {
    if (!c->d_xyz) return fail(MB_ERR_STATE, "no frame set");
    MB_CUDA(cudaSetDevice(c->device));
    MB_TRY(validate_sel(ids, n, c->n_atoms, what));
    if (need_mass && (!c->masses.p || c->n_masses < c->n_atoms))
        return fail(MB_ERR_STATE, "masses not set (mb_set_masses) for %zu atoms", c->n_atoms);
    if (need_box && !c->has_box) return fail(MB_ERR_NO_PBC, "%s: the frame has no periodic box", what);
    memset(&P, 0, sizeof(P));
    P.xyz = c->d_xyz;
    P.ids = nullptr;
    if (ids) {
        MB_TRY(c->ids1.reserve(n * sizeof(uint64_t)));
        MB_CUDA(cudaMemcpyAsync(c->ids1.p, ids, n * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
        P.ids = c->ids1.as<unsigned long long>();
    }
    P.n = (int)n;
    P.masses = need_mass ? c->masses.as<float>() : nullptr;
    if (c->has_box) P.box = to_dev_box(c->box);
    const size_t want = (n + (size_t)PRED_THREADS * 4 - 1) / ((size_t)PRED_THREADS * 4);
    *nb_out = (int)std::max<size_t>(1, std::min<size_t>(want, (size_t)c->sm_count * 4));
    return MB_OK;
}

// centre with (use_box) or without periodic images; out5: centre[3], W, status
static int run_center(Ctx* c, PbcRedParams& P, int nb, int use_box, unsigned w, double out5[5]) {
    P.use_box = use_box;
    P.w = w;
    MB_TRY(pbc_scratch(c, nb, 4, P));
    center_pbc_kernel<<<nb, PRED_THREADS, 0, c->stream>>>(P);
    c->launches++;
    MB_CUDA(cudaGetLastError());
    MB_CUDA(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < 5; ++k) out5[k] = c->h_res[k];
    if (out5[4] != 0.0) return fail(MB_ERR_ZERO_MASS, "zero mass");
    return MB_OK;
}

static int run_tensor(Ctx* c, PbcRedParams& P, int nb, int use_box, const double centre[3], double out7[7]) {
    P.use_box = use_box;
    for (int d = 0; d < 3; ++d) P.c[d] = centre[d];
    MB_TRY(pbc_scratch(c, nb, 7, P));
    tensor_kernel<<<nb, PRED_THREADS, 0, c->stream>>>(P);
    c->launches++;
    MB_CUDA(cudaGetLastError());
    MB_CUDA(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < 7; ++k) out7[k] = c->h_res[k];
    return MB_OK;
}

// Symmetric 3x3 eigenproblem (cyclic Jacobi, f64): eigenvalues ascending, as do_inertia sorts them
// (measure.rs:590-597); col0, col1 normalised, col2 = col0 x col1 (right-handed, :601-603).  An
// eigenvector is defined up to its sign: col0 and col1 are returned with their largest component positive.
static void inertia_axes(const double S[7], double moments[3], double axes[3][3]) {
    double A[3][3] = {{S[2] + S[3], -S[4], -S[5]}, {-S[4], S[1] + S[3], -S[6]}, {-S[5], -S[6], S[1] + S[2]}};
    double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 64; ++sweep) {
        double off = std::fabs(A[0][1]) + std::fabs(A[0][2]) + std::fabs(A[1][2]);
        double diag = std::fabs(A[0][0]) + std::fabs(A[1][1]) + std::fabs(A[2][2]);
        if (off <= 1e-300 || off <= 1e-17 * diag) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (A[p][q] == 0.0) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double cs = 1.0 / std::sqrt(t * t + 1.0), sn = t * cs;
                for (int k = 0; k < 3; ++k) {
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = cs * akp - sn * akq;
                    A[k][q] = sn * akp + cs * akq;
                }
                for (int k = 0; k < 3; ++k) {
                    const double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = cs * apk - sn * aqk;
                    A[q][k] = sn * apk + cs * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = cs * vkp - sn * vkq;
                    V[k][q] = sn * vkp + cs * vkq;
                }
            }
    }
    int ord[3] = {0, 1, 2};
    std::sort(ord, ord + 3, [&](int a, int b) { return A[a][a] < A[b][b]; });
    for (int j = 0; j < 3; ++j) moments[j] = A[ord[j]][ord[j]];
    double col[2][3];
    for (int j = 0; j < 2; ++j) {
        double nn = 0;
        for (int k = 0; k < 3; ++k) nn += V[k][ord[j]] * V[k][ord[j]];
        nn = std::sqrt(nn);
        int big = 0;
        for (int k = 0; k < 3; ++k) {
            col[j][k] = V[k][ord[j]] / nn;
            if (std::fabs(col[j][k]) > std::fabs(col[j][big])) big = k;
        }
        if (col[j][big] < 0)
            for (int k = 0; k < 3; ++k) col[j][k] = -col[j][k];
    }
    const double c2[3] = {col[0][1] * col[1][2] - col[0][2] * col[1][1], col[0][2] * col[1][0] - col[0][0] * col[1][2],
                          col[0][0] * col[1][1] - col[0][1] * col[1][0]};
    for (int k = 0; k < 3; ++k) {
        axes[k][0] = col[0][k];
        axes[k][1] = col[1][k];
        axes[k][2] = c2[k];
    }
}

static int inertia_impl(Ctx* c, const uint64_t* ids, size_t n, int pbc, double centre[3], double moments[3],
                        double axes[3][3]) {
    PbcRedParams P;
    int nb = 1;
    MB_TRY(pbc_prepare(c, ids, n, true, pbc != 0, pbc ? "inertia_pbc" : "inertia", P, &nb));
    double c5[5], S[7];
    MB_TRY(run_center(c, P, nb, pbc ? 1 : 0, 7u, c5));
    MB_TRY(run_tensor(c, P, nb, pbc ? 1 : 0, c5, S));
    for (int d = 0; d < 3; ++d) centre[d] = c5[d];
    inertia_axes(S, moments, axes);
    return MB_OK;
}

}  // namespace mb

using namespace mb;

extern "C" {

int mb_center_of_geometry(MbCtx* h, const uint64_t* ids, size_t n, double out3[3]) {
    if (!h || !out3) return fail(MB_ERR_ARG, "null argument");
    PbcRedParams P;
    int nb = 1;
    MB_TRY(pbc_prepare(&h->c, ids, n, false, false, "center_of_geometry", P, &nb));
    double c5[5];
    MB_TRY(run_center(&h->c, P, nb, 0, 0u, c5));
    for (int d = 0; d < 3; ++d) out3[d] = c5[d];
    return MB_OK;
}

int mb_center_pbc(MbCtx* h, const uint64_t* ids, size_t n, int mass_weighted, uint8_t pbc_dims, double out3[3]) {
    if (!h || !out3) return fail(MB_ERR_ARG, "null argument");
    PbcRedParams P;
    int nb = 1;
    MB_TRY(pbc_prepare(&h->c, ids, n, mass_weighted != 0, true, mass_weighted ? "center_of_mass_pbc" : "center_of_geometry_pbc",
                       P, &nb));
    double c5[5];
    MB_TRY(run_center(&h->c, P, nb, 1, pbc_dims & 7u, c5));
    for (int d = 0; d < 3; ++d) out3[d] = c5[d];
    return MB_OK;
}

int mb_gyration_pbc(MbCtx* h, const uint64_t* ids, size_t n, double* out) {
    if (!h || !out) return fail(MB_ERR_ARG, "null argument");
    PbcRedParams P;
    int nb = 1;
    MB_TRY(pbc_prepare(&h->c, ids, n, true, true, "gyration_pbc", P, &nb));
    double c5[5], S[7];
    MB_TRY(run_center(&h->c, P, nb, 1, 7u, c5));
    MB_TRY(run_tensor(&h->c, P, nb, 1, c5, S));
    *out = std::sqrt((S[1] + S[2] + S[3]) / S[0]);
    return MB_OK;
}

int mb_inertia(MbCtx* h, const uint64_t* ids, size_t n, int pbc, double moments3[3], double axes9_colmajor[9]) {
    if (!h || !moments3 || !axes9_colmajor) return fail(MB_ERR_ARG, "null argument");
    double centre[3], axes[3][3];
    MB_TRY(inertia_impl(&h->c, ids, n, pbc, centre, moments3, axes));
    for (int col = 0; col < 3; ++col)
        for (int row = 0; row < 3; ++row) axes9_colmajor[col * 3 + row] = axes[row][col];
    return MB_OK;
}

int mb_principal_transform(MbCtx* h, const uint64_t* ids, size_t n, int pbc, double R9_colmajor[9], double t3[3]) {
    if (!h || !R9_colmajor || !t3) return fail(MB_ERR_ARG, "null argument");
    double centre[3], moments[3], axes[3][3];
    MB_TRY(inertia_impl(&h->c, ids, n, pbc, centre, moments, axes));
    // do_principal_transform: T = trans(cm) * Rot(axes^-1) * trans(-cm); axes is orthonormal: inverse = transpose
    double R[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[i][j] = axes[j][i];
    for (int i = 0; i < 3; ++i) {
        t3[i] = centre[i] - (R[i][0] * centre[0] + R[i][1] * centre[1] + R[i][2] * centre[2]);
        for (int j = 0; j < 3; ++j) R9_colmajor[j * 3 + i] = R[i][j];
    }
    return MB_OK;
}

}  // extern "C"
