// mb_reduce.cuh — deterministic f64 block/grid reduction and the 3x3 Kabsch solve.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace mb {

// One-sided (Hestenes) Jacobi SVD of a 3x3 matrix A (row-major) in f64, singular values sorted
// descending (what nalgebra::SVD::new returns, measure.rs:626), followed by the reflection fix
// and R = U * diag(1,1,d) * V^T   (measure.rs:631-642).  Returns false on non-finite input
// (MeasureError::Svd).
__host__ __device__ inline bool kabsch_rotation(const double A[9], double R[9]) {
    double a[3][3], v[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            a[i][j] = A[i * 3 + j];
            v[i][j] = i == j ? 1.0 : 0.0;
            if (!isfinite(a[i][j])) return false;
        }
    const double eps = 2.220446049250313e-16;
    for (int sweep = 0; sweep < 60; ++sweep) {
        bool rotated = false;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int i = 0; i < 3; ++i) {
                    alpha += a[i][p] * a[i][p];
                    beta += a[i][q] * a[i][q];
                    gamma += a[i][p] * a[i][q];
                }
                // |gamma| <= eps sqrt(alpha beta), without the square root
                if (gamma == 0.0 || gamma * gamma <= (eps * eps) * (alpha * beta)) continue;
                rotated = true;
                // t = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)) with zeta = (beta - alpha) / (2 gamma), multiplied through
                // by |2 gamma|: one square root, one division and one reciprocal square root per rotation instead of
                // six such operations — on the device this chain of serial f64 div / sqrt sequences IS the latency of
                // a frame's fit (three rotations per sweep, six to eight sweeps)
                const double da = beta - alpha, dg = 2.0 * gamma;
                const double hyp = sqrt(da * da + dg * dg);
                double t = (da >= 0 ? dg : -dg) / (fabs(da) + hyp);
#ifdef __CUDA_ARCH__
                double c = rsqrt(1.0 + t * t), s = c * t;
#else
                double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#endif
                for (int i = 0; i < 3; ++i) {
                    double x = a[i][p], y = a[i][q];
                    a[i][p] = c * x - s * y;
                    a[i][q] = s * x + c * y;
                    x = v[i][p];
                    y = v[i][q];
                    v[i][p] = c * x - s * y;
                    v[i][q] = s * x + c * y;
                }
            }
        if (!rotated) break;
    }
    double sv[3];
    for (int j = 0; j < 3; ++j) sv[j] = sqrt(a[0][j] * a[0][j] + a[1][j] * a[1][j] + a[2][j] * a[2][j]);
    int o0 = 0, o1 = 1, o2 = 2;
    if (sv[o0] < sv[o1]) { int t = o0; o0 = o1; o1 = t; }
    if (sv[o1] < sv[o2]) { int t = o1; o1 = o2; o2 = t; }
    if (sv[o0] < sv[o1]) { int t = o0; o0 = o1; o1 = t; }
    const int ord[3] = {o0, o1, o2};
    double U[3][3], V[3][3], s[3];
    for (int jj = 0; jj < 3; ++jj) {
        int j = ord[jj];
        s[jj] = sv[j];
        for (int i = 0; i < 3; ++i) {
            V[i][jj] = v[i][j];
            U[i][jj] = sv[j] > 0 ? a[i][j] / sv[j] : 0.0;
        }
    }
    const double tiny = s[0] * eps * 8.0;
    auto det = [](const double M[3][3]) {
        return M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
               M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
    };
    if (!(s[0] > 0)) {
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) U[i][j] = i == j ? 1.0 : 0.0;
    } else {
        if (s[1] <= tiny) {
            int k = 0;
            for (int i = 1; i < 3; ++i)
                if (fabs(U[i][0]) < fabs(U[k][0])) k = i;
            double w[3], nn = 0, dotp = U[k][0];
            for (int i = 0; i < 3; ++i) {
                w[i] = (i == k ? 1.0 : 0.0) - dotp * U[i][0];
                nn += w[i] * w[i];
            }
            nn = sqrt(nn);
            for (int i = 0; i < 3; ++i) U[i][1] = w[i] / nn;
        }
        if (s[2] <= tiny) {
            double cx = U[1][0] * U[2][1] - U[2][0] * U[1][1];
            double cy = U[2][0] * U[0][1] - U[0][0] * U[2][1];
            double cz = U[0][0] * U[1][1] - U[1][0] * U[0][1];
            double sg = det(V) < 0 ? -1.0 : 1.0;
            U[0][2] = sg * cx;
            U[1][2] = sg * cy;
            U[2][2] = sg * cz;
        }
    }
    // d = sign(det(U * V^T)) = sign(det U * det V)
    double UVt[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) UVt[i][j] = U[i][0] * V[j][0] + U[i][1] * V[j][1] + U[i][2] * V[j][2];
    double d = det(UVt) < 0.0 ? -1.0 : 1.0;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[i * 3 + j] = U[i][0] * V[j][0] + U[i][1] * V[j][1] + d * U[i][2] * V[j][2];
    return true;
}

#ifdef __CUDACC__
// Reduce K per-thread doubles over the block, store the block partial, and let the LAST block to
// finish (ticket) fold all partials in block order — bitwise deterministic for a fixed launch shape.
// Returns true in every thread of the last block after `result[0..K)` is complete (in shared memory).
template <int K, int THREADS>
__device__ __forceinline__ bool grid_reduce(double (&v)[K], double* __restrict__ partials /*[nblk][K]*/,
                                            unsigned* __restrict__ ticket, int blk, int nblk,
                                            double* sh_result /*[K] shared*/) {
    __shared__ double sh[THREADS / 32][K];
    __shared__ bool is_last;
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) sh[wid][k] = x;
    }
    __syncthreads();
    if (threadIdx.x < K) {
        double x = 0;
#pragma unroll
        for (int w = 0; w < THREADS / 32; ++w) x += sh[w][threadIdx.x];
        partials[(size_t)blk * K + threadIdx.x] = x;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(ticket, 1u) == (unsigned)(nblk - 1);
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
    // Fold the per-block partials with the whole block: warp w takes components w, w+nwarps, ...;
    // lane l sums blocks l, l+32, ... in order, then a fixed shuffle tree.  (A first version let K
    // threads walk all blocks serially: ~300 dependent L2 latencies on the critical path of every frame.)
    for (int k = (int)wid; k < K; k += THREADS / 32) {
        double x = 0;
        for (int b = (int)lane; b < nblk; b += 32) x += __ldcg(&partials[(size_t)b * K + k]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) sh_result[k] = x;
    }
    if (threadIdx.x == 0) *ticket = 0;  // re-arm for the next launch
    __syncthreads();
    return true;
}
#endif

}  // namespace mb
