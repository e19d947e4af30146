// mb_measure.cu — selection reductions and Kabsch alignment on sm_100a.
//
// Replaces Measure::center_of_mass / gyration (molar/src/measure.rs:60-87,561-570), rmsd / rmsd_mw
// (:485-504,538-558), fit_transform[_at_origin] + rot_transform (:507-535,613-643) and
// Modify::apply_transform (molar/src/modify.rs:32-36).
//
// The reference makes 2 (gyration) to 5 (fit + superpose + rmsd) serial passes over the selection.
// Here every quantity is ONE pass that accumulates raw moments about a pivot atom in f64
// (inputs stay f32; products of two f32 are exact in f64), reduced deterministically, and the
// 3x3 covariance / SVD / reflection fix are finished by one thread of the last block.  Results are
// f64 and agree with the reference's f64 build to ~1e-12 relative (tolerance asked: 1e-6).
//
// HBM traffic per frame: 12 B/atom coordinates (+4 B/atom masses, L2-resident across frames);
// the batched superposition re-reads the frame from L2 and writes 12 B/atom.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "mb_common.cuh"
#include "mb_reduce.cuh"

namespace mb {

constexpr int RED_THREADS = 256;

struct SelView {
    const float* xyz;
    const unsigned long long* ids;
    int n;
};

__device__ __forceinline__ size_t sel_gid(const SelView& s, int k) { return s.ids ? (size_t)s.ids[k] : (size_t)k; }

// 4 consecutive atoms (identity selection, 16-byte aligned frame, n % 4 == 0): three 128-bit loads
__device__ __forceinline__ void load4(const float* __restrict__ xyz, int a0, float (&x)[4], float (&y)[4], float (&z)[4]) {
    const float4* p = reinterpret_cast<const float4*>(xyz + 3 * (size_t)a0);
    float4 u = __ldg(p), v = __ldg(p + 1), w = __ldg(p + 2);
    x[0] = u.x; y[0] = u.y; z[0] = u.z;
    x[1] = u.w; y[1] = v.x; z[1] = v.y;
    x[2] = v.z; y[2] = v.w; z[2] = w.x;
    x[3] = w.y; y[3] = w.z; z[3] = w.w;
}
__device__ __forceinline__ bool vec_ok(const SelView& s) {
    return s.ids == nullptr && (s.n & 3) == 0 && ((reinterpret_cast<uintptr_t>(s.xyz) & 15u) == 0);
}

// ---- COM / gyration moments: {M, S=sum m q, Q=sum m |q|^2}, q = p - pivot --------------------
// results row (8 doubles): com[3], rg, M, status(0 ok / 1 zero mass), unused[2]
__global__ void __launch_bounds__(RED_THREADS) moments1_kernel(const float* __restrict__ xyz_base, size_t frame_stride,
                                                               const unsigned long long* __restrict__ ids, int n,
                                                               const float* __restrict__ masses,
                                                               double* __restrict__ partials,
                                                               unsigned* __restrict__ tickets,
                                                               double* __restrict__ results) {
    const int frame = blockIdx.y;
    SelView s{xyz_base + (size_t)frame * frame_stride, ids, n};
    const size_t g0 = sel_gid(s, 0);
    const double ox = s.xyz[3 * g0], oy = s.xyz[3 * g0 + 1], oz = s.xyz[3 * g0 + 2];
    double v[5] = {0, 0, 0, 0, 0};
    auto acc = [&](float x, float y, float z, float mf) {
        double m = mf, qx = (double)x - ox, qy = (double)y - oy, qz = (double)z - oz;
        v[0] += m;
        v[1] += m * qx;
        v[2] += m * qy;
        v[3] += m * qz;
        v[4] += m * (qx * qx + qy * qy + qz * qz);
    };
    if (vec_ok(s) && ((reinterpret_cast<uintptr_t>(masses) & 15u) == 0)) {
        for (int a0 = 4 * (blockIdx.x * RED_THREADS + threadIdx.x); a0 < n; a0 += 4 * gridDim.x * RED_THREADS) {
            float x[4], y[4], z[4];
            load4(s.xyz, a0, x, y, z);
            float4 m4 = __ldg(reinterpret_cast<const float4*>(masses + a0));
            acc(x[0], y[0], z[0], m4.x);
            acc(x[1], y[1], z[1], m4.y);
            acc(x[2], y[2], z[2], m4.z);
            acc(x[3], y[3], z[3], m4.w);
        }
    } else {
        for (int k = blockIdx.x * RED_THREADS + threadIdx.x; k < n; k += gridDim.x * RED_THREADS) {
            size_t g = sel_gid(s, k);
            acc(s.xyz[3 * g], s.xyz[3 * g + 1], s.xyz[3 * g + 2], masses[g]);
        }
    }
    __shared__ double res[5];
    if (!grid_reduce<5, RED_THREADS>(v, partials + (size_t)frame * gridDim.x * 5, tickets + frame, blockIdx.x,
                                     gridDim.x, res))
        return;
    if (threadIdx.x == 0) {
        double* out = results + (size_t)frame * 8;
        double M = res[0];
        if (M == 0.0) {
            out[0] = out[1] = out[2] = out[3] = nan("");
            out[4] = 0.0;
            out[5] = 1.0;
        } else {
            double ax = res[1] / M, ay = res[2] / M, az = res[3] / M;
            out[0] = ox + ax;
            out[1] = oy + ay;
            out[2] = oz + az;
            double r2 = res[4] / M - (ax * ax + ay * ay + az * az);
            out[3] = sqrt(r2 > 0.0 ? r2 : 0.0);
            out[4] = M;
            out[5] = 0.0;
        }
    }
}

// ---- rmsd / rmsd_mw: {sum w |p2-p1|^2, sum w} --------------------------------------------------
__global__ void __launch_bounds__(RED_THREADS) rmsd_kernel(SelView s1, SelView s2, const float* __restrict__ masses,
                                                           double* __restrict__ partials, unsigned* __restrict__ ticket,
                                                           double* __restrict__ results) {
    double v[2] = {0, 0};
    for (int k = blockIdx.x * RED_THREADS + threadIdx.x; k < s1.n; k += gridDim.x * RED_THREADS) {
        size_t g1 = sel_gid(s1, k), g2 = sel_gid(s2, k);
        double dx = (double)s2.xyz[3 * g2] - (double)s1.xyz[3 * g1];
        double dy = (double)s2.xyz[3 * g2 + 1] - (double)s1.xyz[3 * g1 + 1];
        double dz = (double)s2.xyz[3 * g2 + 2] - (double)s1.xyz[3 * g1 + 2];
        double w = masses ? (double)masses[g1] : 1.0;
        v[0] += w * (dx * dx + dy * dy + dz * dz);
        v[1] += w;
    }
    __shared__ double res[2];
    if (!grid_reduce<2, RED_THREADS>(v, partials, ticket, blockIdx.x, gridDim.x, res)) return;
    if (threadIdx.x == 0) {
        results[0] = res[0];
        results[1] = res[1];
    }
}

// ---- Kabsch moments --------------------------------------------------------------------------
// v: [0] M1, [1..3] S1 = sum m1 q1, [4..6] T2 = sum m1 q2, [7..15] C = sum m1 q2 q1^T (row-major),
//    [16] M2, [17..19] S2 = sum m2 q2 (only when SEP2: sel2 has its own masses)
// results row (16 doubles): R[9] row-major, t[3], status (0 ok, 1 zero mass, 3 svd), pad
template <bool SEP2>
__device__ __forceinline__ void fit_finalize(const double* res, const double o1[3], const double o2[3], int at_origin,
                                             double* out) {
    const double M1 = res[0];
    const double M2 = SEP2 ? res[16] : res[0];
    if (M1 == 0.0 || (!at_origin && M2 == 0.0)) {
        for (int i = 0; i < 12; ++i) out[i] = nan("");
        out[12] = 1.0;
        return;
    }
    double cov[9], R[9];
    double a1[3], a2[3];
    for (int i = 0; i < 3; ++i) {
        a1[i] = res[1 + i] / M1;
        a2[i] = SEP2 ? res[17 + i] / M2 : res[4 + i] / M1;
    }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double C = res[7 + 3 * i + j], T2 = res[4 + i], S1 = res[1 + j];
            cov[3 * i + j] = at_origin ? C + T2 * o1[j] + o2[i] * S1 + M1 * o2[i] * o1[j]
                                       : C - T2 * a1[j] - a2[i] * S1 + M1 * a2[i] * a1[j];
        }
    if (!kabsch_rotation(cov, R)) {
        for (int i = 0; i < 12; ++i) out[i] = nan("");
        out[12] = 3.0;
        return;
    }
    for (int i = 0; i < 9; ++i) out[i] = R[i];
    for (int i = 0; i < 3; ++i) {
        if (at_origin) {
            out[9 + i] = 0.0;
        } else {
            double cm1[3] = {o1[0] + a1[0], o1[1] + a1[1], o1[2] + a1[2]};
            double cm2i = o2[i] + a2[i];
            out[9 + i] = cm2i - (R[3 * i] * cm1[0] + R[3 * i + 1] * cm1[1] + R[3 * i + 2] * cm1[2]);
        }
    }
    out[12] = 0.0;
}

template <bool SEP2>
__global__ void __launch_bounds__(RED_THREADS, 3) fit_moments_kernel(const float* __restrict__ xyz1_base, size_t stride1,
                                                                  const unsigned long long* __restrict__ ids1,
                                                                  const float* __restrict__ xyz2,
                                                                  const unsigned long long* __restrict__ ids2, int n,
                                                                  const float* __restrict__ masses, int at_origin,
                                                                  double* __restrict__ partials,
                                                                  unsigned* __restrict__ tickets,
                                                                  double* __restrict__ results) {
    constexpr int K = SEP2 ? 20 : 16;
    const int frame = blockIdx.y;
    SelView s1{xyz1_base + (size_t)frame * stride1, ids1, n};
    SelView s2{xyz2, ids2, n};
    const size_t g10 = sel_gid(s1, 0), g20 = sel_gid(s2, 0);
    const double o1[3] = {s1.xyz[3 * g10], s1.xyz[3 * g10 + 1], s1.xyz[3 * g10 + 2]};
    const double o2[3] = {s2.xyz[3 * g20], s2.xyz[3 * g20 + 1], s2.xyz[3 * g20 + 2]};
    double v[K];
#pragma unroll
    for (int i = 0; i < K; ++i) v[i] = 0.0;
    auto acc = [&](float x1, float y1, float z1, float x2, float y2, float z2, float m1f, float m2f) {
        double m = m1f;
        double q1x = (double)x1 - o1[0], q1y = (double)y1 - o1[1], q1z = (double)z1 - o1[2];
        double q2x = (double)x2 - o2[0], q2y = (double)y2 - o2[1], q2z = (double)z2 - o2[2];
        v[0] += m;
        v[1] += m * q1x;
        v[2] += m * q1y;
        v[3] += m * q1z;
        double wx = m * q2x, wy = m * q2y, wz = m * q2z;
        v[4] += wx;
        v[5] += wy;
        v[6] += wz;
        v[7] += wx * q1x;  v[8] += wx * q1y;  v[9] += wx * q1z;
        v[10] += wy * q1x; v[11] += wy * q1y; v[12] += wy * q1z;
        v[13] += wz * q1x; v[14] += wz * q1y; v[15] += wz * q1z;
        if (SEP2) {
            double m2 = m2f;
            v[16] += m2;
            v[17] += m2 * q2x;
            v[18] += m2 * q2y;
            v[19] += m2 * q2z;
        }
    };
    if (!SEP2 && vec_ok(s1) && vec_ok(s2) && ((reinterpret_cast<uintptr_t>(masses) & 15u) == 0)) {
        for (int a0 = 4 * (blockIdx.x * RED_THREADS + threadIdx.x); a0 < n; a0 += 4 * gridDim.x * RED_THREADS) {
            float x1[4], y1[4], z1[4], x2[4], y2[4], z2[4];
            load4(s1.xyz, a0, x1, y1, z1);
            load4(s2.xyz, a0, x2, y2, z2);
            float4 m4 = __ldg(reinterpret_cast<const float4*>(masses + a0));
            acc(x1[0], y1[0], z1[0], x2[0], y2[0], z2[0], m4.x, 0.f);
            acc(x1[1], y1[1], z1[1], x2[1], y2[1], z2[1], m4.y, 0.f);
            acc(x1[2], y1[2], z1[2], x2[2], y2[2], z2[2], m4.z, 0.f);
            acc(x1[3], y1[3], z1[3], x2[3], y2[3], z2[3], m4.w, 0.f);
        }
    } else {
        for (int k = blockIdx.x * RED_THREADS + threadIdx.x; k < n; k += gridDim.x * RED_THREADS) {
            size_t g1 = sel_gid(s1, k), g2 = sel_gid(s2, k);
            acc(s1.xyz[3 * g1], s1.xyz[3 * g1 + 1], s1.xyz[3 * g1 + 2], s2.xyz[3 * g2], s2.xyz[3 * g2 + 1],
                s2.xyz[3 * g2 + 2], masses[g1], SEP2 ? masses[g2] : 0.f);
        }
    }
    __shared__ double res[K];
    if (!grid_reduce<K, RED_THREADS>(v, partials + (size_t)frame * gridDim.x * K, tickets + frame, blockIdx.x,
                                     gridDim.x, res))
        return;
    if (threadIdx.x == 0) fit_finalize<SEP2>(res, o1, o2, at_origin, results + (size_t)frame * 16);
}

// ---- apply_transform (modify.rs:32-36): p <- R p + t, f64 math, f32 store ----------------------
struct Xform {
    double R[9];
    double t[3];
};
__global__ void __launch_bounds__(256) apply_transform_kernel(float* __restrict__ xyz,
                                                              const unsigned long long* __restrict__ ids, int n,
                                                              Xform X) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    size_t g = ids ? (size_t)ids[k] : (size_t)k;
    double x = xyz[3 * g], y = xyz[3 * g + 1], z = xyz[3 * g + 2];
    xyz[3 * g] = (float)(X.R[0] * x + X.R[1] * y + X.R[2] * z + X.t[0]);
    xyz[3 * g + 1] = (float)(X.R[3] * x + X.R[4] * y + X.R[5] * z + X.t[1]);
    xyz[3 * g + 2] = (float)(X.R[6] * x + X.R[7] * y + X.R[8] * z + X.t[2]);
}

// ---- batched superposition + RMSD after the fit (config 4, second pass; the frame is L2-hot) ---
// fitres: rows of 16 doubles from fit_moments_kernel.  rmsd_out[frame] = sqrt(sum |R p + t - ref|^2 / n)
__global__ void __launch_bounds__(RED_THREADS) superpose_rmsd_kernel(float* __restrict__ xyz_base, size_t stride,
                                                                     const float* __restrict__ ref, int n,
                                                                     const double* __restrict__ fitres, int superpose,
                                                                     double* __restrict__ partials,
                                                                     unsigned* __restrict__ tickets,
                                                                     double* __restrict__ rmsd_out) {
    // frames are visited in REVERSE launch order: the moments pass that ran just before this kernel
    // left the last frames of the group in L2, so they are re-read from there instead of HBM
    const int frame = gridDim.y - 1 - blockIdx.y;
    float* xyz = xyz_base + (size_t)frame * stride;
    const double* fr = fitres + (size_t)frame * 16;
    double R[9], t[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = fr[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) t[i] = fr[9 + i];
    double v[1] = {0.0};
    auto one = [&](float x, float y, float z, float rx, float ry, float rz, float& ox, float& oy, float& oz) {
        double px = R[0] * x + R[1] * y + R[2] * z + t[0];
        double py = R[3] * x + R[4] * y + R[5] * z + t[1];
        double pz = R[6] * x + R[7] * y + R[8] * z + t[2];
        double dx = px - rx, dy = py - ry, dz = pz - rz;
        v[0] += dx * dx + dy * dy + dz * dz;
        ox = (float)px;
        oy = (float)py;
        oz = (float)pz;
    };
    const bool vec = (n & 3) == 0 && ((reinterpret_cast<uintptr_t>(xyz) & 15u) == 0) &&
                     ((reinterpret_cast<uintptr_t>(ref) & 15u) == 0);
    if (vec) {
        for (int a0 = 4 * (blockIdx.x * RED_THREADS + threadIdx.x); a0 < n; a0 += 4 * gridDim.x * RED_THREADS) {
            float x[4], y[4], z[4], rx[4], ry[4], rz[4], ox[4], oy[4], oz[4];
            load4(xyz, a0, x, y, z);
            load4(ref, a0, rx, ry, rz);
#pragma unroll
            for (int i = 0; i < 4; ++i) one(x[i], y[i], z[i], rx[i], ry[i], rz[i], ox[i], oy[i], oz[i]);
            if (superpose) {
                float4* p = reinterpret_cast<float4*>(xyz + 3 * (size_t)a0);
                p[0] = make_float4(ox[0], oy[0], oz[0], ox[1]);
                p[1] = make_float4(oy[1], oz[1], ox[2], oy[2]);
                p[2] = make_float4(oz[2], ox[3], oy[3], oz[3]);
            }
        }
    } else {
        for (int k = blockIdx.x * RED_THREADS + threadIdx.x; k < n; k += gridDim.x * RED_THREADS) {
            float ox, oy, oz;
            one(xyz[3 * (size_t)k], xyz[3 * (size_t)k + 1], xyz[3 * (size_t)k + 2], ref[3 * (size_t)k],
                ref[3 * (size_t)k + 1], ref[3 * (size_t)k + 2], ox, oy, oz);
            if (superpose) {
                xyz[3 * (size_t)k] = ox;
                xyz[3 * (size_t)k + 1] = oy;
                xyz[3 * (size_t)k + 2] = oz;
            }
        }
    }
    __shared__ double res[1];
    if (!grid_reduce<1, RED_THREADS>(v, partials + (size_t)frame * gridDim.x, tickets + frame, blockIdx.x, gridDim.x, res))
        return;
    if (threadIdx.x == 0) rmsd_out[frame] = sqrt(res[0] / (double)n);
}

// ---------------------------------------------------------------------------------------------
// Fused, persistent Kabsch fit + superposition + RMSD over a block of device-resident frames
// (config 4).  One cooperative launch; the grid is exactly the number of co-resident CTAs.  CTA b
// owns slice b (a contiguous atom range) of EVERY frame:
//   * the slice of the reference frame and of the mass column is staged in shared memory once;
//   * the slice of frame f is brought in by ONE 1-D TMA bulk copy (cp.async.bulk + mbarrier),
//     three buffers deep, so HBM latency is hidden behind the f64 arithmetic of the previous frame;
//   * pass 1 accumulates the Kabsch moments from shared memory, the per-frame partials are folded
//     by the last CTA to arrive, which also runs the 3x3 SVD and publishes (R,t) with a release flag;
//   * pass 2 of frame f is delayed by one frame (so the SVD latency is hidden), transforms the
//     slice in place in shared memory, accumulates sum |Rp+t-ref|^2 and writes the slice back with one
//     TMA bulk store.
// HBM traffic per frame: 12 B/atom read + 12 B/atom written — the algorithmic minimum.
// ---------------------------------------------------------------------------------------------
constexpr int FUSED_THREADS = 256;
constexpr int FUSED_NBUF = 4;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)  // suspend-time hint (ns): a waiting warp should not spin on issue slots
            : "memory");
    } while (!ok);
}
// the same for a role that has nothing else to do (producers, storer, publisher): back off between polls so that the
// spinning does not take issue slots from the computing warps of the same scheduler
__device__ __forceinline__ void mbar_wait_idle(unsigned long long* bar, unsigned parity) {
    for (;;) {
        unsigned ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) break;
        __nanosleep(40);
    }
}
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst_gmem, const void* src_smem, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
                 "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct FusedParams {
    float* frames;        // [nf][n][3], superposed in place
    const float* ref;     // [n][3]
    const float* masses;  // [n]
    int n, nf, slice;     // atoms per slice (multiple of 4)
    int superpose;
    double* part_fit;     // [nf][grid][16]
    double* part_sup;     // [nf][grid]
    unsigned* tick_fit;   // [nf]
    unsigned* tick_sup;   // [nf]
    unsigned* flag;       // [nf]  1 when fitres[f] is valid
    double* fitres;       // [nf][16]
    double* rmsd;         // [nf]
};

// block-level sum of K doubles per thread -> sh_out[K] (valid in all threads after return)
template <int K>
__device__ __forceinline__ void block_sum(double (&v)[K], double* sh_out) {
    __shared__ double shw[FUSED_THREADS / 32][K];
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) shw[wid][k] = x;
    }
    __syncthreads();
    if (threadIdx.x < K) {
        double x = 0;
#pragma unroll
        for (int w = 0; w < FUSED_THREADS / 32; ++w) x += shw[w][threadIdx.x];
        sh_out[threadIdx.x] = x;
    }
    __syncthreads();
}

constexpr int FUSED_DELAY = 2;  // pass 2 of frame f runs during iteration f + FUSED_DELAY

__global__ void __launch_bounds__(FUSED_THREADS, 2) fit_fused_kernel(const FusedParams P) {
    extern __shared__ __align__(128) unsigned char fsm[];
    const int slice = P.slice;
    float* buf[FUSED_NBUF];
    for (int i = 0; i < FUSED_NBUF; ++i) buf[i] = reinterpret_cast<float*>(fsm) + (size_t)i * slice * 3;
    float* sref = reinterpret_cast<float*>(fsm) + (size_t)FUSED_NBUF * slice * 3;
    float* smass = sref + (size_t)slice * 3;
    __shared__ __align__(8) unsigned long long bar[FUSED_NBUF];
    __shared__ double res[16];
    __shared__ double sRt[12];
    __shared__ double r2[1];

    const int b = blockIdx.x, nblk = gridDim.x;
    const int a0 = b * slice;
    const int cnt = max(0, min(slice, P.n - a0));  // atoms of this CTA's slice
    const unsigned bytes = (unsigned)cnt * 12u;
    const int tid = threadIdx.x;
    const unsigned lane = tid & 31u, wid = tid >> 5;

    if (tid == 0) {
        for (int i = 0; i < FUSED_NBUF; ++i) mbar_init(&bar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // reference slice + masses: resident for the whole kernel
    for (int i = tid; i < cnt * 3; i += FUSED_THREADS) sref[i] = P.ref[(size_t)a0 * 3 + i];
    for (int i = tid; i < cnt; i += FUSED_THREADS) smass[i] = P.masses[a0 + i];
    __syncthreads();
    auto prefetch = [&](int f) {
        if (tid == 0 && cnt > 0 && f < P.nf) {
            unsigned long long* br = &bar[f % FUSED_NBUF];
            mbar_expect_tx(br, bytes);
            bulk_load(buf[f % FUSED_NBUF], P.frames + ((size_t)f * P.n + a0) * 3, bytes, br);
        }
    };
    for (int f = 0; f < FUSED_NBUF - FUSED_DELAY; ++f) prefetch(f);
    const double o2x = P.ref[0], o2y = P.ref[1], o2z = P.ref[2];  // pivot of the reference frame: atom 0

    auto pass2 = [&](int f) {
        // wait for (R,t) of frame f (published FUSED_DELAY frames ago by that frame's finisher CTA)
        if (tid == 0) {
            const volatile unsigned* fl = P.flag + f;
            while (*fl == 0u) { }
            __threadfence();
        }
        __syncthreads();
        if (tid < 12) sRt[tid] = __ldcg(&P.fitres[(size_t)f * 16 + tid]);
        __syncthreads();
        double R[9], t[3];
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = sRt[i];
#pragma unroll
        for (int i = 0; i < 3; ++i) t[i] = sRt[9 + i];
        float* x = buf[f % FUSED_NBUF];
        double v[1] = {0.0};
        for (int i = tid; i < cnt; i += FUSED_THREADS) {
            const double px0 = x[3 * i], py0 = x[3 * i + 1], pz0 = x[3 * i + 2];
            const double px = R[0] * px0 + R[1] * py0 + R[2] * pz0 + t[0];
            const double py = R[3] * px0 + R[4] * py0 + R[5] * pz0 + t[1];
            const double pz = R[6] * px0 + R[7] * py0 + R[8] * pz0 + t[2];
            const double dx = px - (double)sref[3 * i], dy = py - (double)sref[3 * i + 1], dz = pz - (double)sref[3 * i + 2];
            v[0] += dx * dx + dy * dy + dz * dz;
            if (P.superpose) {
                x[3 * i] = (float)px;
                x[3 * i + 1] = (float)py;
                x[3 * i + 2] = (float)pz;
            }
        }
        if (P.superpose) fence_async_smem();  // generic-proxy writes -> visible to the bulk store
        block_sum<1>(v, r2);                  // (contains the __syncthreads the store needs)
        if (tid == 0) {
            P.part_sup[(size_t)f * nblk + b] = r2[0];  // folded in block order by finish_rmsd_kernel
            if (P.superpose && cnt > 0) {
                bulk_store(P.frames + ((size_t)f * P.n + a0) * 3, x, bytes);
                bulk_store_wait_read();  // the buffer may be refilled only after the store has read it
            }
        }
        fence_async_smem();
        __syncthreads();
    };

    for (int f = 0; f < P.nf; ++f) {
        float* x = buf[f % FUSED_NBUF];
        if (cnt > 0) mbar_wait(&bar[f % FUSED_NBUF], (unsigned)((f / FUSED_NBUF) & 1));
        // ---- pass 1: Kabsch moments of this slice (pivots: atom 0 of the frame / of the reference)
        const float* fr0 = P.frames + (size_t)f * P.n * 3;
        const double o1x = __ldcg(fr0), o1y = __ldcg(fr0 + 1), o1z = __ldcg(fr0 + 2);
        double v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0.0;
        for (int i = tid; i < cnt; i += FUSED_THREADS) {
            const double m = smass[i];
            const double q1x = (double)x[3 * i] - o1x, q1y = (double)x[3 * i + 1] - o1y, q1z = (double)x[3 * i + 2] - o1z;
            const double q2x = (double)sref[3 * i] - o2x, q2y = (double)sref[3 * i + 1] - o2y, q2z = (double)sref[3 * i + 2] - o2z;
            v[0] += m;
            v[1] += m * q1x; v[2] += m * q1y; v[3] += m * q1z;
            const double wx = m * q2x, wy = m * q2y, wz = m * q2z;
            v[4] += wx; v[5] += wy; v[6] += wz;
            v[7] += wx * q1x;  v[8] += wx * q1y;  v[9] += wx * q1z;
            v[10] += wy * q1x; v[11] += wy * q1y; v[12] += wy * q1z;
            v[13] += wz * q1x; v[14] += wz * q1y; v[15] += wz * q1z;
        }
        block_sum<16>(v, res);
        // publish this CTA's partial and arrive WITHOUT waiting for the counter's old value: only the
        // frame's finisher CTA (role rotates: f % grid) ever waits on the ticket
        if (tid < 16) P.part_fit[((size_t)f * nblk + b) * 16 + tid] = res[tid];
        __threadfence();
        __syncthreads();
        if (tid == 0) atomicAdd(P.tick_fit + f, 1u);
        if (b == f % nblk) {
            if (tid == 0) {
                const volatile unsigned* tk = P.tick_fit + f;
                while (*tk < (unsigned)nblk) { }
                __threadfence();
            }
            __syncthreads();
            const double* part = P.part_fit + (size_t)f * nblk * 16;
            for (int k = (int)wid; k < 16; k += FUSED_THREADS / 32) {
                double xsum = 0;
                for (int bb = (int)lane; bb < nblk; bb += 32) xsum += __ldcg(&part[(size_t)bb * 16 + k]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) xsum += __shfl_xor_sync(0xffffffffu, xsum, o);
                if (lane == 0) res[k] = xsum;
            }
            __syncthreads();
            if (tid == 0) {
                const double o1[3] = {o1x, o1y, o1z}, o2[3] = {o2x, o2y, o2z};
                fit_finalize<false>(res, o1, o2, 0, P.fitres + (size_t)f * 16);
                __threadfence();
                atomicExch(P.flag + f, 1u);
                P.tick_fit[f] = 0;  // re-arm
            }
            __syncthreads();
        }
        // ---- pass 2 of an earlier frame (its fold + SVD ran while the grid moved on)
        if (f >= FUSED_DELAY) pass2(f - FUSED_DELAY);
        prefetch(f + FUSED_NBUF - FUSED_DELAY);  // into the buffer released just now (or still unused)
    }
    for (int f = max(0, P.nf - FUSED_DELAY); f < P.nf; ++f) pass2(f);
}

// ---------------------------------------------------------------------------------------------------------------
// Single-pass Kabsch batch, second design (config 4; measure.rs:507-522,613-643 + modify.rs:32-36 + measure.rs:485-504
// per frame): every byte of a frame crosses HBM once in each direction.
//
// The first fused kernel (above) is bound by its grid-wide dependency: the CTA that folds the partials and runs the
// SVD of frame f also owns a slice, so it falls behind by the whole finish latency, frame f+1 cannot complete without
// its partial, and the frames serialise on (finish latency + pass 1).  Here the roles are split:
//   * WORKER CTAs (blockIdx < W) own a slice of every frame.  Slices arrive by 1-D TMA bulk copies into a ring of
//     NBUF shared-memory buffers; pass 1 (16 f64 moments) publishes a partial and a ticket and never waits; pass 2
//     (superpose in shared memory, sum |Rp+t-ref|^2, TMA bulk store) of frame f runs DELAY frames later, when the
//     rotation has normally long been published.
//   * FINISHER CTAs (blockIdx >= W) own no atoms: finisher k waits for the W tickets of frames f = k (mod NFIN), folds
//     the partials with the whole CTA in a fixed order, runs the 3x3 Kabsch SVD and publishes (R, t) + a flag.
// Workers therefore run at the speed of the TMA ring as long as the finish latency stays below DELAY frame times.
constexpr int FS_NFIN = 4;

template <int NBUF, int DELAY>
__global__ void __launch_bounds__(FUSED_THREADS, 1) fit_stream_kernel(const FusedParams P, int W) {
    static_assert(NBUF >= DELAY + 2, "one buffer must be free for the load in flight");
    extern __shared__ __align__(128) unsigned char fsm[];
    const int b = blockIdx.x, tid = threadIdx.x;
    const unsigned lane = tid & 31u, wid = tid >> 5;
    __shared__ double res[16];
    const double o2x = P.ref[0], o2y = P.ref[1], o2z = P.ref[2];  // pivot of the reference frame: atom 0

    if (b >= W) {
        // ------------------------------ finisher ------------------------------
        for (int f = b - W; f < P.nf; f += FS_NFIN) {
            if (tid == 0) {
                const volatile unsigned* tk = P.tick_fit + f;
                while (*tk < (unsigned)W) { }
                __threadfence();
            }
            __syncthreads();
            // fold: output k by warp (k % 8), partials strided over the lanes, shuffle tree — a fixed order
            const double* part = P.part_fit + (size_t)f * W * 16;
            for (int k = (int)wid; k < 16; k += FUSED_THREADS / 32) {
                double xsum = 0;
                for (int bb = (int)lane; bb < W; bb += 32) xsum += __ldcg(&part[(size_t)bb * 16 + k]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) xsum += __shfl_xor_sync(0xffffffffu, xsum, o);
                if (lane == 0) res[k] = xsum;
            }
            __syncthreads();
            if (tid == 0) {
                const float* fr0 = P.frames + (size_t)f * P.n * 3;
                const double o1[3] = {__ldcg(fr0), __ldcg(fr0 + 1), __ldcg(fr0 + 2)};
                const double o2[3] = {o2x, o2y, o2z};
                fit_finalize<false>(res, o1, o2, 0, P.fitres + (size_t)f * 16);
                P.tick_fit[f] = 0;  // re-arm for the next launch
                __threadfence();
                atomicExch(P.flag + f, 1u);
            }
            __syncthreads();
        }
        return;
    }

    // ------------------------------ worker ------------------------------
    const int slice = P.slice;
    float* buf[NBUF];
#pragma unroll
    for (int i = 0; i < NBUF; ++i) buf[i] = reinterpret_cast<float*>(fsm) + (size_t)i * slice * 3;
    float* sref = reinterpret_cast<float*>(fsm) + (size_t)NBUF * slice * 3;
    __shared__ __align__(8) unsigned long long bar[NBUF];
    __shared__ double sRt[12];
    __shared__ double r2[1];
    const int a0 = b * slice;
    const int cnt = max(0, min(slice, P.n - a0));  // atoms of this CTA's slice
    const unsigned bytes = (unsigned)cnt * 12u;
    if (tid == 0) {
        for (int i = 0; i < NBUF; ++i) mbar_init(&bar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < cnt * 3; i += FUSED_THREADS) sref[i] = P.ref[(size_t)a0 * 3 + i];  // resident for the kernel
    __syncthreads();
    const float* __restrict__ mass = P.masses + a0;  // 4 B/atom per frame from L2
    auto prefetch = [&](int f) {
        if (tid == 0 && cnt > 0 && f < P.nf) {
            unsigned long long* br = &bar[f % NBUF];
            mbar_expect_tx(br, bytes);
            bulk_load(buf[f % NBUF], P.frames + ((size_t)f * P.n + a0) * 3, bytes, br);
        }
    };
    for (int f = 0; f < NBUF - DELAY; ++f) prefetch(f);

    auto pass2 = [&](int f) {
        if (tid == 0) {
            const volatile unsigned* fl = P.flag + f;
            while (*fl == 0u) { }
            __threadfence();
        }
        __syncthreads();
        if (tid < 12) sRt[tid] = __ldcg(&P.fitres[(size_t)f * 16 + tid]);
        __syncthreads();
        double R[9], t[3];
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = sRt[i];
#pragma unroll
        for (int i = 0; i < 3; ++i) t[i] = sRt[9 + i];
        float* x = buf[f % NBUF];
        double v[1] = {0.0};
        for (int i = tid; i < cnt; i += FUSED_THREADS) {
            const double px0 = x[3 * i], py0 = x[3 * i + 1], pz0 = x[3 * i + 2];
            const double px = R[0] * px0 + R[1] * py0 + R[2] * pz0 + t[0];
            const double py = R[3] * px0 + R[4] * py0 + R[5] * pz0 + t[1];
            const double pz = R[6] * px0 + R[7] * py0 + R[8] * pz0 + t[2];
            const double dx = px - (double)sref[3 * i], dy = py - (double)sref[3 * i + 1], dz = pz - (double)sref[3 * i + 2];
            v[0] += dx * dx + dy * dy + dz * dz;
            if (P.superpose) {
                x[3 * i] = (float)px;
                x[3 * i + 1] = (float)py;
                x[3 * i + 2] = (float)pz;
            }
        }
        if (P.superpose) fence_async_smem();  // generic-proxy writes -> visible to the bulk store
        block_sum<1>(v, r2);                  // (contains the __syncthreads the store needs)
        if (tid == 0) {
            P.part_sup[(size_t)f * W + b] = r2[0];  // folded in block order by finish_rmsd_kernel
            if (P.superpose && cnt > 0) {
                bulk_store(P.frames + ((size_t)f * P.n + a0) * 3, x, bytes);
                bulk_store_wait_read();  // the buffer may be refilled only after the store has read it
            }
        }
        fence_async_smem();
        __syncthreads();
    };

    for (int f = 0; f < P.nf; ++f) {
        float* x = buf[f % NBUF];
        if (cnt > 0) mbar_wait(&bar[f % NBUF], (unsigned)((f / NBUF) & 1));
        const float* fr0 = P.frames + (size_t)f * P.n * 3;
        // the frame's pivot (its atom 0) must be read BEFORE any worker superposes the frame in place: every CTA
        // reads it here, and pass 2 of frame f starts only after all W tickets of frame f are in
        const double o1x = __ldcg(fr0), o1y = __ldcg(fr0 + 1), o1z = __ldcg(fr0 + 2);
        double v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0.0;
        for (int i = tid; i < cnt; i += FUSED_THREADS) {
            const double m = __ldg(mass + i);
            const double q1x = (double)x[3 * i] - o1x, q1y = (double)x[3 * i + 1] - o1y, q1z = (double)x[3 * i + 2] - o1z;
            const double q2x = (double)sref[3 * i] - o2x, q2y = (double)sref[3 * i + 1] - o2y, q2z = (double)sref[3 * i + 2] - o2z;
            v[0] += m;
            v[1] += m * q1x; v[2] += m * q1y; v[3] += m * q1z;
            const double wx = m * q2x, wy = m * q2y, wz = m * q2z;
            v[4] += wx; v[5] += wy; v[6] += wz;
            v[7] += wx * q1x;  v[8] += wx * q1y;  v[9] += wx * q1z;
            v[10] += wy * q1x; v[11] += wy * q1y; v[12] += wy * q1z;
            v[13] += wz * q1x; v[14] += wz * q1y; v[15] += wz * q1z;
        }
        block_sum<16>(v, res);
        if (tid < 16) P.part_fit[((size_t)f * W + b) * 16 + tid] = res[tid];
        __threadfence();
        __syncthreads();
        if (tid == 0) atomicAdd(P.tick_fit + f, 1u);  // never waits: the finisher CTAs do
        if (f >= DELAY) pass2(f - DELAY);
        prefetch(f + NBUF - DELAY);  // into the buffer released just now (or still unused)
    }
    for (int f = max(0, P.nf - DELAY); f < P.nf; ++f) pass2(f);
}

// ---------------------------------------------------------------------------------------------------------------
// Single-pass Kabsch batch, third design (fused_fit = 3): ONE persistent kernel, second pass served by L2.
//
// Every CTA owns the same slice of every frame and alternates between pass 1 of frame f (16 f64 moments -> partial +
// ticket; the CTA that takes the last ticket folds the partials and runs the SVD) and pass 2 of frame f - LAG
// (superpose + sum |Rp+t-ref|^2).  LAG frames is long enough (tens of microseconds) for (R, t) of that frame to have
// been published, so nobody ever waits, and short enough for the frame to be still resident in the 126 MB L2: HBM
// sees each frame once in each direction (12 B/atom read by pass 1, 12 B/atom written by pass 2), the re-read is an
// L2 hit.  No shared-memory staging, so any atom count works and several CTAs per SM hide the f64 latency.
struct LagParams {
    float* frames;        // [nf][n][3], superposed in place
    const float* ref;     // [n][3]
    const float* masses;  // [n]
    int n, nf, superpose, lag;
    int per;              // atoms per CTA slice (multiple of 4)
    int teams;            // T: the grid is T teams of gridDim.x / T CTAs; team t owns frames t, t + T, ...
    int use_smem;         // the CTA's slice of the reference frame and of the masses lives in shared memory (16 B/atom)
    double* part_fit;     // [nf][team size][16]
    double* part_sup;     // [nf][team size]
    unsigned* tick_fit;   // [nf]
    unsigned* flag;       // [nf]
    double* fitres;       // [nf][16]
};

constexpr int LAG_THREADS = 224;            // streaming threads of a CTA (7 warps: with the solver warp a CTA is 8 warps — warps are
                                            // allocated in groups of four, a ninth would cost the registers of twelve)
constexpr int LAG_BLOCK = LAG_THREADS + 32;  // + one solver warp (fold + 3x3 SVD of the frames this CTA is responsible for)

// barrier of the streaming warps only (the solver warp never joins it)
__device__ __forceinline__ void bar_stream() { asm volatile("bar.sync 1, %0;" ::"n"(LAG_THREADS) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Warp total of 16 doubles per lane by a butterfly that halves the values kept per lane at every step (32 shuffles
// instead of 160): afterwards lane l holds the warp total of value (l >> 1).  Fixed order => deterministic.
__device__ __forceinline__ double warp_sum16(const double (&v)[16], unsigned lane) {
    double a[8], b4[4], c2[2];
    const bool h16 = lane & 16u, h8 = lane & 8u, h4 = lane & 4u, h2 = lane & 2u;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const double mine = h16 ? v[8 + k] : v[k], other = h16 ? v[k] : v[8 + k];
        a[k] = mine + __shfl_xor_sync(0xffffffffu, other, 16);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double mine = h8 ? a[4 + k] : a[k], other = h8 ? a[k] : a[4 + k];
        b4[k] = mine + __shfl_xor_sync(0xffffffffu, other, 8);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const double mine = h4 ? b4[2 + k] : b4[k], other = h4 ? b4[k] : b4[2 + k];
        c2[k] = mine + __shfl_xor_sync(0xffffffffu, other, 4);
    }
    const double mine = h2 ? c2[1] : c2[0], other = h2 ? c2[0] : c2[1];
    double d = mine + __shfl_xor_sync(0xffffffffu, other, 2);
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    return d;
}

// One persistent kernel for the whole batch.  Every CTA owns one slice of atoms of EVERY frame: its 256 streaming
// threads run pass 1 (moments) of frame f and pass 2 (superposition + RMSD sum) of frame f - lag in the same iteration,
// so pass 2 re-reads its slice from L2 and a frame costs 12 B/atom of DRAM reads + 12 B/atom of writes.  The fold of
// the per-CTA moments and the 3x3 SVD of frame f belong to the SOLVER WARP of CTA f mod G — a ninth warp that never
// joins the streaming barriers.  (When the last CTA to arrive did the solve, the ~10 us of serial f64 Jacobi made that
// CTA the last one of the next frame as well: every solve landed on the same CTA, one after the other, 19 us per frame.)
__global__ void __launch_bounds__(LAG_BLOCK, 2) fit_lag_kernel(const LagParams P) {
    __shared__ double wsum[LAG_THREADS / 32][16];
    __shared__ double sRt[12];
    // teams: CTA blockIdx.x is member b of team `team`; a team works through its frames one after the other, the T
    // teams run T frames concurrently (a frame's chain of dependent latencies — loads, block reduction, ticket, flag,
    // L2 re-read — is several microseconds whatever the slice size, so one team alone cannot reach the HBM rate)
    const int T = P.teams, team = blockIdx.x % T, b = blockIdx.x / T, G = gridDim.x / T, tid = threadIdx.x;
    const unsigned lane = tid & 31u, wid = tid >> 5;
    const double o2x = P.ref[0], o2y = P.ref[1], o2z = P.ref[2];  // pivot of the reference frame: atom 0
    const int nft = P.nf > team ? (P.nf - team + T - 1) / T : 0;  // frames of this team
    if (wid == LAG_THREADS / 32) {
        // ---------------- solver warp: every G-th frame of the team ----------------
        for (int i = b; i < nft; i += G) {
            const int f = team + i * T;
            if (lane == 0)
                while (ld_acquire_u32(P.tick_fit + f) != (unsigned)G) __nanosleep(200);
            __syncwarp();
            const double* part = P.part_fit + (size_t)f * G * 16;
            double res[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                double xsum = 0;
                for (int bb = (int)lane; bb < G; bb += 32) xsum += __ldcg(&part[(size_t)bb * 16 + k]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) xsum += __shfl_xor_sync(0xffffffffu, xsum, o);
                res[k] = xsum;
            }
            if (lane == 0) {
                const float* fr = P.frames + (size_t)f * P.n * 3;
                const double o1[3] = {(double)__ldcg(fr), (double)__ldcg(fr + 1), (double)__ldcg(fr + 2)};
                const double o2[3] = {o2x, o2y, o2z};
                fit_finalize<false>(res, o1, o2, 0, P.fitres + (size_t)f * 16);
                P.tick_fit[f] = 0;  // re-arm for the next launch
                __threadfence();
                st_release_u32(P.flag + f, 1u);
            }
            __syncwarp();
        }
        return;
    }
    const int a0 = min(P.n, b * P.per), a1 = min(P.n, a0 + P.per);
    const bool vec = (P.n & 3) == 0 && ((reinterpret_cast<uintptr_t>(P.frames) | reinterpret_cast<uintptr_t>(P.ref) |
                                         reinterpret_cast<uintptr_t>(P.masses)) & 15u) == 0;
    // The slice of the reference frame and of the masses is the same for every frame: staged once in shared memory, so
    // that the only global loads of the frame loop are the frame's own bytes (loads in flight are what bounds a
    // latency-exposed streaming loop; L2 reads of the reference would take 4/7 of them)
    extern __shared__ __align__(16) float lag_dyn[];
    float* sref = lag_dyn;
    float* smass = lag_dyn + 3 * (size_t)P.per;
    if (P.use_smem) {
        for (int k = tid; k < 3 * (a1 - a0); k += LAG_THREADS) sref[k] = P.ref[3 * (size_t)a0 + k];
        for (int k = tid; k < a1 - a0; k += LAG_THREADS) smass[k] = P.masses[a0 + k];
        bar_stream();
    }
    // reference coordinates + masses of the four atoms starting at atom a (a - a0 is a multiple of 4)
    auto ref4 = [&](int a, float (&x)[4], float (&y)[4], float (&z)[4], float4& m4) {
        if (P.use_smem) {
            const float4* p = reinterpret_cast<const float4*>(sref + 3 * (a - a0));
            const float4 u = p[0], v = p[1], w = p[2];
            x[0] = u.x; y[0] = u.y; z[0] = u.z; x[1] = u.w; y[1] = v.x; z[1] = v.y;
            x[2] = v.z; y[2] = v.w; z[2] = w.x; x[3] = w.y; y[3] = w.z; z[3] = w.w;
            m4 = *reinterpret_cast<const float4*>(smass + (a - a0));
        } else {
            load4(P.ref, a, x, y, z);
            m4 = __ldg(reinterpret_cast<const float4*>(P.masses + a));
        }
    };
    constexpr int LAG_U = 3;  // four-atom groups whose loads a thread issues before it uses the first

    for (int i = 0; i < nft + P.lag; ++i) {
        const int f = team + i * T;
        if (i < nft) {
            // ---------------- pass 1 of frame f ----------------
            const float* fr = P.frames + (size_t)f * P.n * 3;
            const double o1x = __ldcg(fr), o1y = __ldcg(fr + 1), o1z = __ldcg(fr + 2);
            double v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = 0.0;
            auto acc = [&](float x1, float y1, float z1, float x2, float y2, float z2, float mf) {
                const double m = mf;
                const double q1x = (double)x1 - o1x, q1y = (double)y1 - o1y, q1z = (double)z1 - o1z;
                const double q2x = (double)x2 - o2x, q2y = (double)y2 - o2y, q2z = (double)z2 - o2z;
                v[0] += m;
                v[1] += m * q1x; v[2] += m * q1y; v[3] += m * q1z;
                const double wx = m * q2x, wy = m * q2y, wz = m * q2z;
                v[4] += wx; v[5] += wy; v[6] += wz;
                v[7] += wx * q1x;  v[8] += wx * q1y;  v[9] += wx * q1z;
                v[10] += wy * q1x; v[11] += wy * q1y; v[12] += wy * q1z;
                v[13] += wz * q1x; v[14] += wz * q1y; v[15] += wz * q1z;
            };
            if (vec) {
                for (int a = a0 + 4 * tid; a < a1; a += 4 * LAG_THREADS * LAG_U) {
                    float4 q[LAG_U][3];
#pragma unroll
                    for (int u = 0; u < LAG_U; ++u) {
                        const int au = a + u * 4 * LAG_THREADS;
                        if (au < a1) {
                            const float4* p = reinterpret_cast<const float4*>(fr + 3 * (size_t)au);
                            q[u][0] = __ldcg(p);
                            q[u][1] = __ldcg(p + 1);
                            q[u][2] = __ldcg(p + 2);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < LAG_U; ++u) {
                        const int au = a + u * 4 * LAG_THREADS;
                        if (au < a1) {
                            float x2[4], y2[4], z2[4];
                            float4 m4;
                            ref4(au, x2, y2, z2, m4);
                            acc(q[u][0].x, q[u][0].y, q[u][0].z, x2[0], y2[0], z2[0], m4.x);
                            acc(q[u][0].w, q[u][1].x, q[u][1].y, x2[1], y2[1], z2[1], m4.y);
                            acc(q[u][1].z, q[u][1].w, q[u][2].x, x2[2], y2[2], z2[2], m4.z);
                            acc(q[u][2].y, q[u][2].z, q[u][2].w, x2[3], y2[3], z2[3], m4.w);
                        }
                    }
                }
            } else {
                for (int a = a0 + tid; a < a1; a += LAG_THREADS)
                    acc(fr[3 * (size_t)a], fr[3 * (size_t)a + 1], fr[3 * (size_t)a + 2], P.ref[3 * (size_t)a],
                        P.ref[3 * (size_t)a + 1], P.ref[3 * (size_t)a + 2], P.masses[a]);
            }
            const double w = warp_sum16(v, lane);
            if ((lane & 1u) == 0u) wsum[wid][lane >> 1] = w;
            bar_stream();
            if (tid < 16) {
                double x = 0.0;
#pragma unroll
                for (int k = 0; k < LAG_THREADS / 32; ++k) x += wsum[k][tid];
                P.part_fit[((size_t)f * G + b) * 16 + tid] = x;
                __threadfence();
            }
            bar_stream();
            if (tid == 0) atomicAdd(P.tick_fit + f, 1u);  // the G-th ticket releases the frame to its solver warp
        }
        const int g = f - P.lag * T;
        if (i >= P.lag) {
            // ---------------- pass 2 of frame g: (R, t) was published LAG frames ago ----------------
            if (tid == 0)
                while (ld_acquire_u32(P.flag + g) == 0u) __nanosleep(100);
            bar_stream();
            if (tid < 12) sRt[tid] = __ldcg(&P.fitres[(size_t)g * 16 + tid]);
            bar_stream();
            double R[9], t[3];
#pragma unroll
            for (int i = 0; i < 9; ++i) R[i] = sRt[i];
#pragma unroll
            for (int i = 0; i < 3; ++i) t[i] = sRt[9 + i];
            float* fr = P.frames + (size_t)g * P.n * 3;
            double r2 = 0.0;
            auto sup = [&](float x0, float y0, float z0, float rx, float ry, float rz, float& ox, float& oy, float& oz) {
                const double px0 = x0, py0 = y0, pz0 = z0;
                const double px = R[0] * px0 + R[1] * py0 + R[2] * pz0 + t[0];
                const double py = R[3] * px0 + R[4] * py0 + R[5] * pz0 + t[1];
                const double pz = R[6] * px0 + R[7] * py0 + R[8] * pz0 + t[2];
                const double dx = px - (double)rx, dy = py - (double)ry, dz = pz - (double)rz;
                r2 += dx * dx + dy * dy + dz * dz;
                ox = (float)px;
                oy = (float)py;
                oz = (float)pz;
            };
            if (vec) {
                for (int a = a0 + 4 * tid; a < a1; a += 4 * LAG_THREADS * LAG_U) {
                    float4 q[LAG_U][3];
#pragma unroll
                    for (int u = 0; u < LAG_U; ++u) {
                        const int au = a + u * 4 * LAG_THREADS;
                        if (au < a1) {
                            const float4* p = reinterpret_cast<const float4*>(fr + 3 * (size_t)au);
                            q[u][0] = __ldcg(p);  // L2 hits (the slice was read `lag` frames of this team ago)
                            q[u][1] = __ldcg(p + 1);
                            q[u][2] = __ldcg(p + 2);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < LAG_U; ++u) {
                        const int au = a + u * 4 * LAG_THREADS;
                        if (au < a1) {
                            float x2[4], y2[4], z2[4], ox[4], oy[4], oz[4];
                            float4 m4;
                            ref4(au, x2, y2, z2, m4);
                            sup(q[u][0].x, q[u][0].y, q[u][0].z, x2[0], y2[0], z2[0], ox[0], oy[0], oz[0]);
                            sup(q[u][0].w, q[u][1].x, q[u][1].y, x2[1], y2[1], z2[1], ox[1], oy[1], oz[1]);
                            sup(q[u][1].z, q[u][1].w, q[u][2].x, x2[2], y2[2], z2[2], ox[2], oy[2], oz[2]);
                            sup(q[u][2].y, q[u][2].z, q[u][2].w, x2[3], y2[3], z2[3], ox[3], oy[3], oz[3]);
                            if (P.superpose) {
                                float4* qo = reinterpret_cast<float4*>(fr + 3 * (size_t)au);
                                __stcs(qo, make_float4(ox[0], oy[0], oz[0], ox[1]));
                                __stcs(qo + 1, make_float4(oy[1], oz[1], ox[2], oy[2]));
                                __stcs(qo + 2, make_float4(oz[2], ox[3], oy[3], oz[3]));
                            }
                        }
                    }
                }
            } else {
                for (int a = a0 + tid; a < a1; a += LAG_THREADS) {
                    float ox, oy, oz;
                    sup(__ldcg(fr + 3 * (size_t)a), __ldcg(fr + 3 * (size_t)a + 1), __ldcg(fr + 3 * (size_t)a + 2),
                        P.ref[3 * (size_t)a], P.ref[3 * (size_t)a + 1], P.ref[3 * (size_t)a + 2], ox, oy, oz);
                    if (P.superpose) {
                        fr[3 * (size_t)a] = ox;
                        fr[3 * (size_t)a + 1] = oy;
                        fr[3 * (size_t)a + 2] = oz;
                    }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) r2 += __shfl_xor_sync(0xffffffffu, r2, o);
            bar_stream();  // wsum is free again (pass 1 of this iteration has consumed it)
            if (lane == 0) wsum[wid][0] = r2;
            bar_stream();
            if (tid == 0) {
                double x = 0.0;
#pragma unroll
                for (int k = 0; k < LAG_THREADS / 32; ++k) x += wsum[k][0];
                P.part_sup[(size_t)g * G + b] = x;  // folded in block order by finish_rmsd_kernel
            }
            bar_stream();
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Single-pass Kabsch batch, fourth design (fused_fit = 4): the persistent kernel of design 3 with WARP-SPECIALISED
// roles, so that no computing warp ever waits on a global load, on a fence or on another warp.  One CTA per SM owns a
// slice of every frame; the slice moves in chunks through two shared-memory rings filled by cp.async.bulk (TMA, SASS
// UBLKCP); every hand-over is an mbarrier full / empty pair — there is no CTA or group barrier in the frame loop:
//   warp 0  (one lane)  producer of ring 1: the frame's chunks from DRAM, never more than `max_lead` frames ahead of
//                       pass 2 (keeps the frames in flight inside L2)
//   warp 1  (one lane)  producer of ring 2: waits for (R, t) of frame g to be published, copies it to shared memory and
//                       re-loads the frame's chunks — L2 hits, pass 1 read them a few frames ago
//   warp 2              publisher of pass 1: adds the warps' moment sums of a frame, stores the CTA's partial, fence, ticket
//   warp 3  (one lane)  storer of pass 2: one bulk store per finished chunk, frees the slot once the store has read it,
//                       adds the warps' RMSD sums of a frame
//   warps 4-9           pass 1: 16 f64 moments of the slice (reference slice + masses resident in shared memory)
//   warps 10-15         pass 2: p <- R p + t in place in the ring slot, sum |p - ref|^2
//   warp 16             solver: fold of the per-CTA partials + 3x3 SVD of the frames f = blockIdx.x (mod grid)
// HBM sees each frame once in each direction (12 B/atom read + 12 B/atom written).
struct WsParams {
    float* frames;        // [nf][n][3], superposed in place (n % 4 == 0, 16-byte aligned)
    const float* ref;     // [n][3]
    const float* masses;  // [n]
    int n, nf, superpose;
    int per;              // atoms per CTA slice (multiple of 4); every CTA of the grid owns at least one atom
    int chunk;            // atoms per ring slot (multiple of 4)
    int use_smem;         // the CTA's slice of the reference frame and of the masses lives in shared memory (16 B/atom)
    int max_lead;         // frames pass 1 may run ahead of pass 2
    double* part_fit;     // [nf][grid][16]
    double* part_sup;     // [nf][grid]
    unsigned* tick_fit;   // [nf]
    unsigned* flag;       // [nf]
    double* fitres;       // [nf][16]
};
#ifndef MB_WS_S1
#define MB_WS_S1 6
#endif
#ifndef MB_WS_S2
#define MB_WS_S2 4
#endif
constexpr int WS_S1 = MB_WS_S1;   // slots of ring 1 (DRAM latency)
constexpr int WS_S = MB_WS_S2;    // slots of ring 2 (L2 latency + store read-out)
#ifndef MB_WS_NW
#define MB_WS_NW 6
#endif
constexpr int WS_NW = MB_WS_NW;   // warps of a computing group
constexpr int WS_NT = WS_NW * 32;
#ifndef MB_WS_UB
#define MB_WS_UB 3
#endif
constexpr int WS_UB = MB_WS_UB;   // atoms a thread has in flight
constexpr int WS_BLOCK = 32 * (4 + 2 * WS_NW + 1);
constexpr int WS_RT = 4;   // (R, t) of the frames pass 2 has in flight

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <bool REF_SMEM>
__global__ void __launch_bounds__(WS_BLOCK, 1) fit_ws_kernel(const WsParams P) {
    extern __shared__ __align__(128) unsigned char ws_dyn[];
    __shared__ __align__(8) unsigned long long full1[WS_S1], empty1[WS_S1], full2[WS_S], empty2[WS_S], done2[WS_S];
    __shared__ __align__(8) unsigned long long red1_full[2], red1_empty[2], red2_full[2], red2_empty[2];
    __shared__ __align__(16) float piv[WS_S1][4];  // first 16 bytes of the frame whose chunk 0 sits in the slot (pivot = atom 0)
    __shared__ double sRt[WS_RT][12];
    __shared__ double wsum1[2][WS_NW][16];
    __shared__ double wsum2[2][WS_NW];
    __shared__ volatile int p2_done;  // frames pass 2 of this CTA has finished

    const int tid = threadIdx.x, wid = tid >> 5;
    const unsigned lane = tid & 31u;
    const int b = blockIdx.x, G = gridDim.x;
    const int a0 = b * P.per, cnt = min(P.per, P.n - a0);
    const int K = (cnt + P.chunk - 1) / P.chunk;  // chunks of this CTA's slice
    float* ring1 = reinterpret_cast<float*>(ws_dyn);
    float* ring2 = ring1 + (size_t)WS_S1 * P.chunk * 3;
    float* sref = ring2 + (size_t)WS_S * P.chunk * 3;
    float* smass = sref + 3 * (size_t)P.per;
    if (tid == 0) {
        for (int i = 0; i < WS_S1; ++i) {
            mbar_init(&full1[i], 1);
            mbar_init(&empty1[i], WS_NT);
        }
        for (int i = 0; i < WS_S; ++i) {
            mbar_init(&full2[i], 1);
            mbar_init(&empty2[i], 1);
            mbar_init(&done2[i], WS_NT);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&red1_full[i], WS_NW * 16);
            mbar_init(&red1_empty[i], 16);
            mbar_init(&red2_full[i], WS_NW);
            mbar_init(&red2_empty[i], 1);
        }
        p2_done = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (REF_SMEM) {
        for (int k = tid; k < 3 * cnt; k += WS_BLOCK) sref[k] = P.ref[3 * (size_t)a0 + k];
        for (int k = tid; k < cnt; k += WS_BLOCK) smass[k] = P.masses[a0 + k];
    }
    __syncthreads();
    const size_t fstride = (size_t)P.n * 3;
    auto chunk_atoms = [&](int c) { return min(P.chunk, cnt - c * P.chunk); };

    if (wid == 0) {
        // ---------------- producer of ring 1 ----------------
        if (lane != 0) return;
        int i = 0;
        for (int f = 0; f < P.nf; ++f) {
            while (f - p2_done > P.max_lead) __nanosleep(200);
            const float* fr = P.frames + (size_t)f * fstride;
            for (int c = 0; c < K; ++c, ++i) {
                const int s = i % WS_S1, u = i / WS_S1;
                if (u >= 1) mbar_wait_idle(&empty1[s], (unsigned)(u - 1) & 1u);
                const unsigned bytes = (unsigned)chunk_atoms(c) * 12u;
                mbar_expect_tx(&full1[s], bytes + (c == 0 ? 16u : 0u));
                if (c == 0) bulk_load(piv[s], fr, 16u, &full1[s]);
                bulk_load(ring1 + (size_t)s * P.chunk * 3, fr + 3 * (size_t)(a0 + c * P.chunk), bytes, &full1[s]);
            }
        }
        return;
    }
    if (wid == 1) {
        // ---------------- producer of ring 2 ----------------
        if (lane != 0) return;
        int j = 0;
        for (int g = 0; g < P.nf; ++g) {
            while (ld_acquire_u32(P.flag + g) == 0u) __nanosleep(100);
            const float* fr = P.frames + (size_t)g * fstride;
            for (int c = 0; c < K; ++c, ++j) {
                const int s = j % WS_S, u = j / WS_S;
                if (u >= 1) mbar_wait_idle(&empty2[s], (unsigned)(u - 1) & 1u);
                if (c == 0) {
                    double rt[12];
#pragma unroll
                    for (int k = 0; k < 12; ++k) rt[k] = __ldcg(&P.fitres[(size_t)g * 16 + k]);
#pragma unroll
                    for (int k = 0; k < 12; ++k) sRt[g % WS_RT][k] = rt[k];
                }
                const unsigned bytes = (unsigned)chunk_atoms(c) * 12u;
                mbar_expect_tx(&full2[s], bytes);
                bulk_load(ring2 + (size_t)s * P.chunk * 3, fr + 3 * (size_t)(a0 + c * P.chunk), bytes, &full2[s]);
            }
        }
        return;
    }
    if (wid == 2) {
        // ---------------- publisher of pass 1 ----------------
        for (int f = 0; f < P.nf; ++f) {
            mbar_wait_idle(&red1_full[f & 1], (unsigned)(f >> 1) & 1u);  // every reading lane acquires the phase itself
            if (lane < 16) {
                double x = 0.0;
#pragma unroll
                for (int k = 0; k < WS_NW; ++k) x += wsum1[f & 1][k][lane];
                mbar_arrive(&red1_empty[f & 1]);
                P.part_fit[((size_t)f * G + b) * 16 + lane] = x;
                __threadfence();
            }
            __syncwarp();
            if (lane == 0) atomicAdd(P.tick_fit + f, 1u);  // the G-th ticket releases the frame to its solver warp
        }
        return;
    }
    if (wid == 3) {
        // ---------------- storer of pass 2 ----------------
        if (lane != 0) return;
        int j = 0;
        for (int g = 0; g < P.nf; ++g) {
            float* fr = P.frames + (size_t)g * fstride;
            for (int c = 0; c < K; ++c, ++j) {
                const int s = j % WS_S, u = j / WS_S;
                mbar_wait_idle(&done2[s], (unsigned)u & 1u);
                if (P.superpose) {
                    bulk_store(fr + 3 * (size_t)(a0 + c * P.chunk), ring2 + (size_t)s * P.chunk * 3, (unsigned)chunk_atoms(c) * 12u);
                    // the slot of the PREVIOUS chunk is free once its store has read it out
                    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    if (j >= 1) mbar_arrive(&empty2[(j - 1) % WS_S]);
                } else {
                    mbar_arrive(&empty2[s]);
                }
            }
            mbar_wait_idle(&red2_full[g & 1], (unsigned)(g >> 1) & 1u);
            double x = 0.0;
#pragma unroll
            for (int k = 0; k < WS_NW; ++k) x += wsum2[g & 1][k];
            mbar_arrive(&red2_empty[g & 1]);
            P.part_sup[(size_t)g * G + b] = x;  // folded in block order by finish_rmsd_kernel
            p2_done = g + 1;
        }
        if (P.superpose) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        return;
    }
    if (wid == 4 + 2 * WS_NW) {
        // ---------------- solver warp: frames b, b + G, ... ----------------
        const double o2[3] = {(double)P.ref[0], (double)P.ref[1], (double)P.ref[2]};
        for (int f = b; f < P.nf; f += G) {
            // the pivot (atom 0 of the frame) does not change before pass 2 of this frame: fetched ahead of the wait
            const float* fr = P.frames + (size_t)f * fstride;
            const double o1[3] = {(double)__ldcg(fr), (double)__ldcg(fr + 1), (double)__ldcg(fr + 2)};
            if (lane == 0)
                while (ld_acquire_u32(P.tick_fit + f) != (unsigned)G) __nanosleep(100);
            __syncwarp();
            // fold: lane l sums the partials of CTAs l, l + 32, ... (all loads of a lane independent, 16 bytes each),
            // then one butterfly over the 16 values
            const double2* part = reinterpret_cast<const double2*>(P.part_fit + (size_t)f * G * 16);
            double acc[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) acc[k] = 0.0;
            for (int bb = (int)lane; bb < G; bb += 32) {
                double2 q[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) q[k] = __ldcg(&part[(size_t)bb * 8 + k]);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    acc[2 * k] += q[k].x;
                    acc[2 * k + 1] += q[k].y;
                }
            }
            const double tot = warp_sum16(acc, lane);  // lane l holds the total of value l >> 1
            double res[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) res[k] = __shfl_sync(0xffffffffu, tot, 2 * k);
            if (lane == 0) {
                fit_finalize<false>(res, o1, o2, 0, P.fitres + (size_t)f * 16);
                P.tick_fit[f] = 0;  // re-arm for the next launch
                __threadfence();
                st_release_u32(P.flag + f, 1u);
            }
            __syncwarp();
        }
        return;
    }
    if (wid < 4 + WS_NW) {
        // ---------------- pass 1: moments ----------------
        const int t = tid - 4 * 32, w = wid - 4;
        const double o2x = P.ref[0], o2y = P.ref[1], o2z = P.ref[2];
        int i = 0;
        for (int f = 0; f < P.nf; ++f) {
            double v[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] = 0.0;
            double o1x = 0, o1y = 0, o1z = 0;
            for (int c = 0; c < K; ++c, ++i) {
                const int s = i % WS_S1, u = i / WS_S1;
                mbar_wait(&full1[s], (unsigned)u & 1u);
                if (c == 0) {
                    o1x = piv[s][0];
                    o1y = piv[s][1];
                    o1z = piv[s][2];
                }
                const float* sb = ring1 + (size_t)s * P.chunk * 3;
                const int cc = chunk_atoms(c), roff = c * P.chunk;
                for (int ab = t; ab < cc; ab += WS_NT * WS_UB) {
                    float x1[WS_UB], y1[WS_UB], z1[WS_UB], x2[WS_UB], y2[WS_UB], z2[WS_UB], mf[WS_UB];
#pragma unroll
                    for (int q = 0; q < WS_UB; ++q) {
                        const int a = ab + q * WS_NT;
                        const bool ok = a < cc;
                        const int as = ok ? a : t;  // in range: t < cc here
                        x1[q] = sb[3 * as]; y1[q] = sb[3 * as + 1]; z1[q] = sb[3 * as + 2];
                        if (REF_SMEM) {
                            x2[q] = sref[3 * (roff + as)];
                            y2[q] = sref[3 * (roff + as) + 1];
                            z2[q] = sref[3 * (roff + as) + 2];
                            mf[q] = smass[roff + as];
                        } else {
                            const size_t ga = (size_t)a0 + roff + as;
                            x2[q] = __ldg(P.ref + 3 * ga);
                            y2[q] = __ldg(P.ref + 3 * ga + 1);
                            z2[q] = __ldg(P.ref + 3 * ga + 2);
                            mf[q] = __ldg(P.masses + ga);
                        }
                        if (!ok) mf[q] = 0.0f;  // a zero mass adds exact zeros to every moment
                    }
#pragma unroll
                    for (int q = 0; q < WS_UB; ++q) {
                        const double m = mf[q];
                        const double q1x = (double)x1[q] - o1x, q1y = (double)y1[q] - o1y, q1z = (double)z1[q] - o1z;
                        const double q2x = (double)x2[q] - o2x, q2y = (double)y2[q] - o2y, q2z = (double)z2[q] - o2z;
                        v[0] += m;
                        v[1] += m * q1x; v[2] += m * q1y; v[3] += m * q1z;
                        const double wx = m * q2x, wy = m * q2y, wz = m * q2z;
                        v[4] += wx; v[5] += wy; v[6] += wz;
                        v[7] += wx * q1x;  v[8] += wx * q1y;  v[9] += wx * q1z;
                        v[10] += wy * q1x; v[11] += wy * q1y; v[12] += wy * q1z;
                        v[13] += wz * q1x; v[14] += wz * q1y; v[15] += wz * q1z;
                    }
                }
                mbar_arrive(&empty1[s]);
            }
            const double ws = warp_sum16(v, lane);
            if (f >= 2) mbar_wait(&red1_empty[f & 1], (unsigned)((f >> 1) - 1) & 1u);
            if ((lane & 1u) == 0u) {
                wsum1[f & 1][w][lane >> 1] = ws;
                mbar_arrive(&red1_full[f & 1]);
            }
        }
        return;
    }
    // ---------------- pass 2: superposition + RMSD sum ----------------
    {
        const int t = tid - (4 + WS_NW) * 32, w = wid - 4 - WS_NW;
        int j = 0;
        for (int g = 0; g < P.nf; ++g) {
            double R[9], tr[3];
            double r2 = 0.0;
            for (int c = 0; c < K; ++c, ++j) {
                const int s = j % WS_S, u = j / WS_S;
                mbar_wait(&full2[s], (unsigned)u & 1u);
                if (c == 0) {
#pragma unroll
                    for (int k = 0; k < 9; ++k) R[k] = sRt[g % WS_RT][k];
#pragma unroll
                    for (int k = 0; k < 3; ++k) tr[k] = sRt[g % WS_RT][9 + k];
                }
                float* sb = ring2 + (size_t)s * P.chunk * 3;
                const int cc = chunk_atoms(c), roff = c * P.chunk;
                for (int ab = t; ab < cc; ab += WS_NT * WS_UB) {
                    float x0[WS_UB], y0[WS_UB], z0[WS_UB], rx[WS_UB], ry[WS_UB], rz[WS_UB];
#pragma unroll
                    for (int q = 0; q < WS_UB; ++q) {
                        const int a = ab + q * WS_NT;
                        const int as = a < cc ? a : t;
                        x0[q] = sb[3 * as]; y0[q] = sb[3 * as + 1]; z0[q] = sb[3 * as + 2];
                        if (REF_SMEM) {
                            rx[q] = sref[3 * (roff + as)];
                            ry[q] = sref[3 * (roff + as) + 1];
                            rz[q] = sref[3 * (roff + as) + 2];
                        } else {
                            const size_t ga = (size_t)a0 + roff + as;
                            rx[q] = __ldg(P.ref + 3 * ga);
                            ry[q] = __ldg(P.ref + 3 * ga + 1);
                            rz[q] = __ldg(P.ref + 3 * ga + 2);
                        }
                    }
#pragma unroll
                    for (int q = 0; q < WS_UB; ++q) {
                        const int a = ab + q * WS_NT;
                        if (a < cc) {
                            const double px0 = x0[q], py0 = y0[q], pz0 = z0[q];
                            const double px = R[0] * px0 + R[1] * py0 + R[2] * pz0 + tr[0];
                            const double py = R[3] * px0 + R[4] * py0 + R[5] * pz0 + tr[1];
                            const double pz = R[6] * px0 + R[7] * py0 + R[8] * pz0 + tr[2];
                            const double dx = px - (double)rx[q], dy = py - (double)ry[q], dz = pz - (double)rz[q];
                            r2 += dx * dx + dy * dy + dz * dz;
                            if (P.superpose) {
                                sb[3 * a] = (float)px;
                                sb[3 * a + 1] = (float)py;
                                sb[3 * a + 2] = (float)pz;
                            }
                        }
                    }
                }
                if (P.superpose) fence_async_smem();  // the slot's new contents become visible to the bulk store
                mbar_arrive(&done2[s]);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) r2 += __shfl_xor_sync(0xffffffffu, r2, o);
            if (lane == 0) {
                if (g >= 2) mbar_wait(&red2_empty[g & 1], (unsigned)((g >> 1) - 1) & 1u);
                wsum2[g & 1][w] = r2;
                mbar_arrive(&red2_full[g & 1]);
            }
            __syncwarp();
        }
    }
}

// rmsd[f] = sqrt(sum_b part[f][b] / n), partials folded in block order (deterministic)
__global__ void __launch_bounds__(32) finish_rmsd_kernel(const double* __restrict__ part, int nblk, int n,
                                                         double* __restrict__ rmsd) {
    const int f = blockIdx.x;
    double x = 0;
    for (int b = threadIdx.x; b < nblk; b += 32) x += part[(size_t)f * nblk + b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (threadIdx.x == 0) rmsd[f] = sqrt(x / (double)n);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int upload_ids(Ctx* c, DevBuf& buf, const uint64_t* ids, size_t n, const unsigned long long** out) {
    *out = nullptr;
    if (!ids) return MB_OK;
    MB_TRY(buf.reserve(n * sizeof(uint64_t)));
    MB_CUDA(cudaMemcpyAsync(buf.p, ids, n * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
    *out = buf.as<unsigned long long>();
    return MB_OK;
}

static int check_sel(const uint64_t* ids, size_t n, size_t n_atoms, const char* what) {
    return validate_sel(ids, n, n_atoms, what);
}

static int red_blocks(const Ctx* c, size_t n, int per_thread) {
    size_t want = (n + (size_t)RED_THREADS * per_thread - 1) / ((size_t)RED_THREADS * per_thread);
    return (int)std::max<size_t>(1, std::min<size_t>(want, (size_t)c->sm_count * 4));
}

// scratch layout in reduce_tmp: [tickets: 64 KB, zeroed at allocation and re-armed by the kernels]
// [results][partials]
struct RedScratch {
    unsigned* tickets;
    double* results;
    double* partials;
};
constexpr size_t TICKET_BYTES = 64 * 1024;
static int red_scratch(Ctx* c, size_t frames, size_t partial_doubles, size_t result_doubles, RedScratch* s) {
    if (frames * sizeof(unsigned) > TICKET_BYTES) return fail(MB_ERR_ARG, "too many frames per reduction group");
    size_t res_bytes = ((result_doubles * sizeof(double) + 255) / 256) * 256;
    size_t need = TICKET_BYTES + res_bytes + partial_doubles * sizeof(double);
    bool fresh = need > c->reduce_tmp.cap;
    MB_TRY(c->reduce_tmp.reserve(need));
    if (fresh) MB_CUDA(cudaMemsetAsync(c->reduce_tmp.p, 0, TICKET_BYTES, c->stream));
    s->tickets = c->reduce_tmp.as<unsigned>();
    s->results = reinterpret_cast<double*>(static_cast<char*>(c->reduce_tmp.p) + TICKET_BYTES);
    s->partials = reinterpret_cast<double*>(static_cast<char*>(c->reduce_tmp.p) + TICKET_BYTES + res_bytes);
    return MB_OK;
}

static int need_masses(const Ctx* c) {
    if (!c->masses.p || c->n_masses < c->n_atoms)
        return fail(MB_ERR_STATE, "masses not set (mb_set_masses) for %zu atoms", c->n_atoms);
    return MB_OK;
}

static int com_gyr(Ctx* c, const uint64_t* ids, size_t n, double out8[8]) {
    if (!c->d_xyz) return fail(MB_ERR_STATE, "no frame set");
    MB_CUDA(cudaSetDevice(c->device));
    MB_TRY(need_masses(c));
    MB_TRY(check_sel(ids, n, c->n_atoms, "center_of_mass"));
    const unsigned long long* d_ids;
    MB_TRY(upload_ids(c, c->ids1, ids, n, &d_ids));
    int nb = red_blocks(c, n, 8);
    RedScratch s;
    MB_TRY(red_scratch(c, 1, (size_t)nb * 5, 8, &s));
    MB_TRY(c->host_results());  // the finishing thread writes the row into mapped host memory: no D2H copy
    moments1_kernel<<<dim3(nb, 1), RED_THREADS, 0, c->stream>>>(c->d_xyz, 0, d_ids, (int)n, c->masses.as<float>(),
                                                               s.partials, s.tickets, c->d_res);
    c->launches++;
    MB_CUDA(cudaGetLastError());
    MB_CUDA(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < 8; ++k) out8[k] = c->h_res[k];
    if (out8[5] != 0.0) return fail(MB_ERR_ZERO_MASS, "zero mass");
    return MB_OK;
}

// ---- batch (device-resident trajectory) ------------------------------------------------------
// ref_frame == SIZE_MAX: keep the reference frame already staged in batch_ref (streaming: the ring no longer holds it)
int batch_fit_impl(Ctx* c, size_t ref_frame, size_t f0, size_t f1, int superpose, double* rmsd_out) {
    const bool keep_ref = ref_frame == (size_t)-1;
    if (!c->batch.p || f1 > c->batch_frames || f0 >= f1 || (!keep_ref && ref_frame >= c->batch_frames))
        return fail(MB_ERR_ARG, "batch_fit: bad frame range");
    if (keep_ref && c->batch_ref.cap < c->batch_atoms * 3 * sizeof(float))
        return fail(MB_ERR_STATE, "batch_fit: no reference frame staged");
    MB_CUDA(cudaSetDevice(c->device));
    if (!c->masses.p || c->n_masses < c->batch_atoms) return fail(MB_ERR_STATE, "masses not set for the batch");
    const size_t n = c->batch_atoms, nf = f1 - f0;
    // private copy of the reference frame: it may itself be superposed in place below
    if (!keep_ref) {
        MB_TRY(c->batch_ref.reserve(n * 3 * sizeof(float)));
        MB_CUDA(cudaMemcpyAsync(c->batch_ref.p, c->batch.as<float>() + ref_frame * n * 3, n * 3 * sizeof(float),
                                cudaMemcpyDeviceToDevice, c->stream));
    }
    const float* ref = c->batch_ref.as<float>();
    // ---- single-pass path, fourth design (fused_fit = 4): warp-specialised persistent kernel, TMA rings ----
    if (c->opt_fused_fit == 4 && (n % 4) == 0 && !(reinterpret_cast<uintptr_t>(c->batch.p) & 15u)) {
        int per = (int)((((n + c->sm_count - 1) / c->sm_count) + 3) / 4 * 4);
        const int grid = (int)((n + per - 1) / per);  // every CTA owns at least one atom
        // ring slot: a multiple of one batched round of a group (WS_NT x WS_UB atoms), as large as the shared memory
        // left next to the resident reference slice allows, at most half a slice (>= 2 chunks in flight per frame)
        const size_t smem_budget = (size_t)200 * 1024;
        const int round = WS_NT * WS_UB;
        int use_smem = (size_t)per * 16 + (size_t)(WS_S1 + WS_S) * round * 12 <= smem_budget ? 1 : 0;
        const size_t ring_budget = smem_budget - (use_smem ? (size_t)per * 16 : 0);
        int mult = (int)std::max<size_t>(1, std::min<size_t>(ring_budget / ((size_t)(WS_S1 + WS_S) * round * 12),
                                                              std::max<size_t>(1, ((size_t)per / 2 + round - 1) / round)));
        int chunk = c->opt_fit_group > 0 ? (c->opt_fit_group + 3) / 4 * 4 : mult * round;
        chunk = std::min(chunk, per);
        const size_t ring_bytes = (size_t)(WS_S1 + WS_S) * chunk * 12;
        if (ring_bytes + (use_smem ? (size_t)per * 16 : 0) > smem_budget) use_smem = 0;
        if (ring_bytes > smem_budget) return fail(MB_ERR_ARG, "batch_fit: fit_group too large for the shared-memory rings");
        const size_t dyn_smem = ring_bytes + (use_smem ? (size_t)per * 16 : 0);
        auto ws_kern = use_smem ? fit_ws_kernel<true> : fit_ws_kernel<false>;
        MB_CUDA(cudaFuncSetAttribute(ws_kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));
        int occ_real = 0;
        MB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_real, ws_kern, WS_BLOCK, dyn_smem));
        if ((long long)occ_real * c->sm_count < grid)
            return fail(MB_ERR_STATE, "batch_fit: persistent grid of %d CTAs (%zu B shared) does not fit the device", grid, dyn_smem);
        const size_t fb = n * 12;
        int lead = (int)std::max<size_t>(2, std::min<size_t>(64, ((size_t)40 << 20) / std::max<size_t>(fb, 1)));
        if (c->opt_fit_lag > 0) lead = c->opt_fit_lag;
        const size_t fchunk = std::min<size_t>(nf, 4096);
        RedScratch s{};
        MB_TRY(red_scratch(c, 2 * fchunk, fchunk * (size_t)grid * 17, nf * 17, &s));
        double* fitres = s.results;
        double* d_rmsd = s.results + nf * 16;
        for (size_t g0 = 0; g0 < nf; g0 += fchunk) {
            const size_t gn = std::min(fchunk, nf - g0);
            WsParams P;
            P.frames = c->batch.as<float>() + (f0 + g0) * n * 3;
            P.ref = ref;
            P.masses = c->masses.as<float>();
            P.n = (int)n;
            P.nf = (int)gn;
            P.superpose = superpose;
            P.per = per;
            P.chunk = chunk;
            P.use_smem = use_smem;
            P.max_lead = lead;
            P.part_fit = s.partials;
            P.part_sup = s.partials + fchunk * (size_t)grid * 16;
            P.tick_fit = s.tickets;
            P.flag = s.tickets + fchunk;
            P.fitres = fitres + g0 * 16;
            MB_CUDA(cudaMemsetAsync(P.flag, 0, gn * sizeof(unsigned), c->stream));
            void* args[] = {&P};
            MB_CUDA(cudaLaunchCooperativeKernel((const void*)ws_kern, dim3(grid), dim3(WS_BLOCK), args, dyn_smem, c->stream));
            finish_rmsd_kernel<<<(unsigned)gn, 32, 0, c->stream>>>(P.part_sup, grid, (int)n, d_rmsd + g0);
            c->launches += 2;
        }
        MB_CUDA(cudaGetLastError());
        if (rmsd_out)
            MB_CUDA(cudaMemcpyAsync(rmsd_out, d_rmsd, nf * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        MB_CUDA(cudaStreamSynchronize(c->stream));
        return MB_OK;
    }
    // ---- single-pass path, third design (fused_fit = 3): one persistent kernel, pass 2 lags by LAG frames and reads L2
    if (c->opt_fused_fit == 3) {
        const int occ = 2;  // CTAs per SM the kernel is sized for (<= 128 registers x 256 threads)
        // grid: all resident CTAs, but never more CTAs than slices of 4 * LAG_THREADS atoms (one vector round)
        const size_t tile = 4 * (size_t)LAG_THREADS;
        const int grid_max = (int)std::min<size_t>((size_t)c->sm_count * occ, std::max<size_t>(1, (n + tile - 1) / tile));
        const size_t ntile = (n + tile - 1) / tile;
        // T teams of grid / T CTAs: each team processes its own frames (t, t + T, ...), T frames in flight.  Automatic:
        // the largest T (<= 4) whose slices of the reference frame + masses (16 B/atom) still fit two CTAs per SM.
        const size_t smem_budget = 100 * 1024;
        auto shape = [&](int T, int& gt, int& per) {
            gt = std::max(1, grid_max / T);
            const size_t tiles_per = (ntile + gt - 1) / gt;
            gt = (int)((ntile + tiles_per - 1) / tiles_per);  // slices of whole tiles, as even as the tile count allows
            per = (int)(tiles_per * tile);
        };
        int T = c->opt_fit_teams, gt = 1, per = 0;
        const int t_cap = std::max(1, std::min(grid_max, (int)std::min<size_t>(nf, 64)));
        if (T > 0) {
            T = std::min(T, t_cap);
            shape(T, gt, per);
        } else {
            for (T = std::min(4, t_cap); T >= 1; --T) {
                shape(T, gt, per);
                if ((size_t)per * 16 <= smem_budget || T == 1) break;
            }
        }
        const int grid = gt * T;
        const int use_smem = (size_t)per * 16 <= smem_budget ? 1 : 0;
        const size_t dyn_smem = use_smem ? (size_t)per * 16 : 0;
        MB_CUDA(cudaFuncSetAttribute(fit_lag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_budget));
        MB_CUDA(cudaFuncSetAttribute(fit_lag_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        int occ_real = 0;
        MB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_real, fit_lag_kernel, LAG_BLOCK, dyn_smem));
        if ((long long)occ_real * c->sm_count < grid)
            return fail(MB_ERR_STATE, "batch_fit: persistent grid of %d CTAs (%zu B shared) does not fit the device (%d per SM)", grid,
                        dyn_smem, occ_real);
        // LAG: long enough for the finish of a frame (fold + SVD, ~10 us) to be over, short enough for the frames
        // in flight (read + written) to stay inside L2
        const size_t fb = n * 12;
        int lag = (int)std::max<size_t>(1, std::min<size_t>(64, ((size_t)48 << 20) / std::max<size_t>(fb * T, 1)));
        if (c->opt_fit_lag > 0) lag = c->opt_fit_lag;
        const size_t chunk = std::min<size_t>(nf, 4096);
        RedScratch s{};
        MB_TRY(red_scratch(c, 2 * chunk, chunk * (size_t)gt * 17, nf * 17, &s));
        double* fitres = s.results;
        double* d_rmsd = s.results + nf * 16;
        for (size_t g0 = 0; g0 < nf; g0 += chunk) {
            const size_t gn = std::min(chunk, nf - g0);
            LagParams P;
            P.teams = T;
            P.use_smem = use_smem;
            P.frames = c->batch.as<float>() + (f0 + g0) * n * 3;
            P.ref = ref;
            P.masses = c->masses.as<float>();
            P.n = (int)n;
            P.nf = (int)gn;
            P.superpose = superpose;
            P.lag = (int)std::min<size_t>((size_t)lag, std::max<size_t>(1, gn / T));
            P.per = per;
            P.part_fit = s.partials;
            P.part_sup = s.partials + chunk * (size_t)gt * 16;
            P.tick_fit = s.tickets;
            P.flag = s.tickets + chunk;
            P.fitres = fitres + g0 * 16;
            MB_CUDA(cudaMemsetAsync(P.flag, 0, gn * sizeof(unsigned), c->stream));
            void* args[] = {&P};
            MB_CUDA(cudaLaunchCooperativeKernel((const void*)fit_lag_kernel, dim3(grid), dim3(LAG_BLOCK), args, dyn_smem,
                                                c->stream));
            finish_rmsd_kernel<<<(unsigned)gn, 32, 0, c->stream>>>(P.part_sup, gt, (int)n, d_rmsd + g0);
            c->launches += 2;
        }
        MB_CUDA(cudaGetLastError());
        if (rmsd_out)
            MB_CUDA(cudaMemcpyAsync(rmsd_out, d_rmsd, nf * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        MB_CUDA(cudaStreamSynchronize(c->stream));
        return MB_OK;
    }
    // ---- single-pass path, second design (fused_fit = 2): worker CTAs stream slices through a TMA ring, dedicated
    //      finisher CTAs fold + SVD; 12 B/atom read + 12 B/atom written per frame ----
    if (c->opt_fused_fit == 2 && (n % 4) == 0 && !(reinterpret_cast<uintptr_t>(c->batch.p) & 15u) &&
        c->sm_count > 2 * FS_NFIN) {
        constexpr int NBUF = 4, DELAY = 2;
        const int grid = c->sm_count, W = grid - FS_NFIN;
        const int slice = (int)(((n + W - 1) / W + 3) / 4 * 4);
        const size_t smem = (size_t)slice * (NBUF + 1) * 12;
        int occ = 0;
        if (smem <= (size_t)225 * 1024) {
            auto kern = fit_stream_kernel<NBUF, DELAY>;
            MB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            MB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, FUSED_THREADS, smem));
        }
        if (occ >= 1) {
            const size_t chunk = std::min<size_t>(nf, 1024);
            RedScratch s{};
            MB_TRY(red_scratch(c, 3 * chunk, chunk * (size_t)W * 17, nf * 17, &s));
            double* fitres = s.results;
            double* d_rmsd = s.results + nf * 16;
            for (size_t g0 = 0; g0 < nf; g0 += chunk) {
                const size_t gn = std::min(chunk, nf - g0);
                FusedParams P;
                P.frames = c->batch.as<float>() + (f0 + g0) * n * 3;
                P.ref = ref;
                P.masses = c->masses.as<float>();
                P.n = (int)n;
                P.nf = (int)gn;
                P.slice = slice;
                P.superpose = superpose;
                P.part_fit = s.partials;
                P.part_sup = s.partials + chunk * (size_t)W * 16;
                P.tick_fit = s.tickets;
                P.tick_sup = s.tickets + chunk;
                P.flag = s.tickets + 2 * chunk;
                P.fitres = fitres + g0 * 16;
                P.rmsd = d_rmsd + g0;
                MB_CUDA(cudaMemsetAsync(P.flag, 0, gn * sizeof(unsigned), c->stream));
                int Wv = W;
                void* args[] = {&P, &Wv};
                MB_CUDA(cudaLaunchCooperativeKernel((const void*)fit_stream_kernel<NBUF, DELAY>, dim3(grid),
                                                    dim3(FUSED_THREADS), args, smem, c->stream));
                finish_rmsd_kernel<<<(unsigned)gn, 32, 0, c->stream>>>(P.part_sup, W, (int)n, P.rmsd);
                c->launches += 2;
            }
            MB_CUDA(cudaGetLastError());
            if (rmsd_out)
                MB_CUDA(cudaMemcpyAsync(rmsd_out, d_rmsd, nf * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            MB_CUDA(cudaStreamSynchronize(c->stream));
            return MB_OK;
        }
    }
    // ---- fused persistent path, first design (fused_fit = 1) ----
    {
        int occ = 0;
        const int sm = c->sm_count;
        // try 2 CTAs/SM first, then 1
        for (int want = 2; want >= 1 && c->opt_fused_fit == 1; --want) {
            int grid = sm * want;
            int slice = (int)(((n + grid - 1) / grid + 3) / 4 * 4);
            size_t smem = (size_t)slice * (FUSED_NBUF + 1) * 12 + (size_t)slice * 4;
            if (smem > (size_t)(want == 2 ? 110 : 220) * 1024) continue;
            if ((n % 4) != 0 || (reinterpret_cast<uintptr_t>(c->batch.p) & 15u)) break;
            MB_CUDA(cudaFuncSetAttribute(fit_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            MB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fit_fused_kernel, FUSED_THREADS, smem));
            if (occ < want) continue;
            const size_t chunk = std::min<size_t>(nf, 256);
            RedScratch s;
            MB_TRY(red_scratch(c, 3 * chunk, chunk * (size_t)grid * 17, nf * 17, &s));
            double* fitres = s.results;
            double* d_rmsd = s.results + nf * 16;
            for (size_t g0 = 0; g0 < nf; g0 += chunk) {
                const size_t gn = std::min(chunk, nf - g0);
                FusedParams P;
                P.frames = c->batch.as<float>() + (f0 + g0) * n * 3;
                P.ref = ref;
                P.masses = c->masses.as<float>();
                P.n = (int)n;
                P.nf = (int)gn;
                P.slice = slice;
                P.superpose = superpose;
                P.part_fit = s.partials;
                P.part_sup = s.partials + chunk * (size_t)grid * 16;
                P.tick_fit = s.tickets;
                P.tick_sup = s.tickets + chunk;
                P.flag = s.tickets + 2 * chunk;
                P.fitres = fitres + g0 * 16;
                P.rmsd = d_rmsd + g0;
                MB_CUDA(cudaMemsetAsync(P.flag, 0, gn * sizeof(unsigned), c->stream));
                void* args[] = {&P};
                MB_CUDA(cudaLaunchCooperativeKernel((const void*)fit_fused_kernel, dim3(grid), dim3(FUSED_THREADS), args,
                                                    smem, c->stream));
                finish_rmsd_kernel<<<(unsigned)gn, 32, 0, c->stream>>>(P.part_sup, grid, (int)n, P.rmsd);
                c->launches += 2;
            }
            if (rmsd_out)
                MB_CUDA(cudaMemcpyAsync(rmsd_out, d_rmsd, nf * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            MB_CUDA(cudaStreamSynchronize(c->stream));
            return MB_OK;
        }
    }
    // ---- two-kernel path (any n, any alignment) ----
    // Group frames so that a group (moments pass + superposition pass) stays L2-resident, and
    // alternate groups over two streams so that the serial tail of one group's kernels (last-block
    // reduction + 3x3 SVD on one thread) overlaps the streaming part of the other group's.
    const size_t frame_bytes = n * 12;
    // Large groups: each launch must carry enough bytes per thread to reach HBM speed (a 6-frame group
    // sized for L2 residency ran at 1.3 TB/s: launch ramp + reduction tail dominated).  The
    // superposition pass re-reads the group in reverse order, so its first ~90 MB still hit L2.
    size_t group = std::max<size_t>(1, std::min<size_t>(nf, (size_t)(768u << 20) / std::max<size_t>(frame_bytes, 1)));
    group = std::min<size_t>(group, 2048);
    if (c->opt_fit_group > 0) group = std::min<size_t>(nf, (size_t)c->opt_fit_group);
    int nb = (int)std::max<size_t>(1, std::min<size_t>((n + RED_THREADS * 8 - 1) / (RED_THREADS * 8),
                                                      std::max<size_t>(4, (size_t)c->sm_count * 8 / group)));
    const int NS = c->opt_fit_streams > 0 ? std::min(c->opt_fit_streams, 4) : 2;
    for (int i = 0; i < NS; ++i)
        if (!c->aux_stream[i]) MB_CUDA(cudaStreamCreateWithFlags(&c->aux_stream[i], cudaStreamNonBlocking));
    if (!c->aux_event) MB_CUDA(cudaEventCreateWithFlags(&c->aux_event, cudaEventDisableTiming));
    RedScratch s;
    // per stream: tickets for 2 kernels x group frames, partials for both kernels;
    // results: 16 doubles per frame for the whole range + rmsd
    const size_t part_per_stream = (size_t)group * nb * 17;
    MB_TRY(red_scratch(c, (size_t)NS * 2 * group, (size_t)NS * part_per_stream, nf * 17, &s));
    double* fitres = s.results;
    double* d_rmsd = s.results + nf * 16;
    // the aux streams start after everything already queued on the context stream (the ref copy)
    MB_CUDA(cudaEventRecord(c->aux_event, c->stream));
    for (int i = 0; i < NS; ++i) MB_CUDA(cudaStreamWaitEvent(c->aux_stream[i], c->aux_event, 0));
    size_t gi = 0;
    for (size_t g0 = 0; g0 < nf; g0 += group, ++gi) {
        const int si = (int)(gi % NS);
        cudaStream_t st = c->aux_stream[si];
        size_t gn = std::min(group, nf - g0);
        float* base = c->batch.as<float>() + (f0 + g0) * n * 3;
        double* part_fit = s.partials + (size_t)si * part_per_stream;
        double* part_sup = part_fit + (size_t)group * nb * 16;
        unsigned* tick = s.tickets + (size_t)si * 2 * group;
        fit_moments_kernel<false><<<dim3(nb, (unsigned)gn), RED_THREADS, 0, st>>>(
            base, n * 3, nullptr, ref, nullptr, (int)n, c->masses.as<float>(), 0, part_fit, tick, fitres + g0 * 16);
        superpose_rmsd_kernel<<<dim3(nb, (unsigned)gn), RED_THREADS, 0, st>>>(
            base, n * 3, ref, (int)n, fitres + g0 * 16, superpose, part_sup, tick + group, d_rmsd + g0);
        c->launches += 2;
    }
    MB_CUDA(cudaGetLastError());
    // join: the context stream continues after both aux streams
    for (int i = 0; i < NS; ++i) {
        MB_CUDA(cudaEventRecord(c->aux_event, c->aux_stream[i]));
        MB_CUDA(cudaStreamWaitEvent(c->stream, c->aux_event, 0));
    }
    if (rmsd_out) {
        MB_CUDA(cudaMemcpyAsync(rmsd_out, d_rmsd, nf * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    }
    MB_CUDA(cudaStreamSynchronize(c->stream));
    return MB_OK;
}

// COM + Rg of frames [f0,f1) of the batch into rows of 8 doubles at d_rows (device)
int enqueue_batch_moments(Ctx* c, size_t f0, size_t f1, double* d_rows8, double* partials, unsigned* tickets, int nb) {
    const size_t n = c->batch_atoms;
    const float* base = c->batch.as<float>() + f0 * n * 3;
    moments1_kernel<<<dim3(nb, (unsigned)(f1 - f0)), RED_THREADS, 0, c->stream>>>(
        base, n * 3, nullptr, (int)n, c->masses.as<float>(), partials, tickets, d_rows8);
    c->launches++;
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

}  // namespace mb

using namespace mb;

extern "C" {

int mb_center_of_mass(MbCtx* h, const uint64_t* ids, size_t n, double out3[3]) {
    if (!h || !out3) return fail(MB_ERR_ARG, "null argument");
    double r[8];
    MB_TRY(com_gyr(&h->c, ids, n, r));
    out3[0] = r[0];
    out3[1] = r[1];
    out3[2] = r[2];
    return MB_OK;
}

int mb_gyration(MbCtx* h, const uint64_t* ids, size_t n, double* out) {
    if (!h || !out) return fail(MB_ERR_ARG, "null argument");
    double r[8];
    MB_TRY(com_gyr(&h->c, ids, n, r));
    *out = r[3];
    return MB_OK;
}

int mb_rmsd(MbCtx* h, const uint64_t* ids1, size_t n1, const uint64_t* ids2, size_t n2, int use_frame2,
            int mass_weighted, double* out) {
    if (!h || !out) return fail(MB_ERR_ARG, "null argument");
    Ctx* c = &h->c;
    if (!c->d_xyz) return fail(MB_ERR_STATE, "no frame set");
    if (n1 != n2) return fail(MB_ERR_SIZES, "incompatible sizes: %zu and %zu", n1, n2);  // measure.rs:494-496
    MB_CUDA(cudaSetDevice(c->device));
    if (use_frame2 && !c->xyz2.p) return fail(MB_ERR_STATE, "frame2 not set");
    const float* xyz2 = use_frame2 ? c->xyz2.as<float>() : c->d_xyz;
    size_t natoms2 = use_frame2 ? c->n_atoms2 : c->n_atoms;
    MB_TRY(check_sel(ids1, n1, c->n_atoms, "rmsd sel1"));
    MB_TRY(check_sel(ids2, n2, natoms2, "rmsd sel2"));
    if (mass_weighted) MB_TRY(need_masses(c));
    const unsigned long long *d1, *d2;
    MB_TRY(upload_ids(c, c->ids1, ids1, n1, &d1));
    MB_TRY(upload_ids(c, c->ids2, ids2, n2, &d2));
    int nb = red_blocks(c, n1, 4);
    RedScratch s;
    MB_TRY(red_scratch(c, 1, (size_t)nb * 2, 8, &s));
    MB_TRY(c->host_results());
    rmsd_kernel<<<nb, RED_THREADS, 0, c->stream>>>(SelView{c->d_xyz, d1, (int)n1}, SelView{xyz2, d2, (int)n2},
                                                  mass_weighted ? c->masses.as<float>() : nullptr, s.partials,
                                                  s.tickets, c->d_res);
    c->launches++;
    MB_CUDA(cudaGetLastError());
    MB_CUDA(cudaStreamSynchronize(c->stream));
    const double r[2] = {c->h_res[0], c->h_res[1]};
    if (mass_weighted) {
        if (r[1] == 0.0) return fail(MB_ERR_ZERO_MASS, "zero mass");
        *out = std::sqrt(r[0] / r[1]);
    } else {
        *out = std::sqrt(r[0] / (double)n1);
    }
    return MB_OK;
}

int mb_fit_transform(MbCtx* h, const uint64_t* ids1, size_t n1, const uint64_t* ids2, size_t n2, int use_frame2,
                     int at_origin, double R9[9], double t3[3]) {
    if (!h || !R9 || !t3) return fail(MB_ERR_ARG, "null argument");
    Ctx* c = &h->c;
    if (!c->d_xyz) return fail(MB_ERR_STATE, "no frame set");
    // izip! stops at the shortest (measure.rs:621); the bindings only ever pass equal sizes
    if (n1 != n2) return fail(MB_ERR_SIZES, "incompatible sizes: %zu and %zu", n1, n2);
    MB_CUDA(cudaSetDevice(c->device));
    if (use_frame2 && !c->xyz2.p) return fail(MB_ERR_STATE, "frame2 not set");
    const float* xyz2 = use_frame2 ? c->xyz2.as<float>() : c->d_xyz;
    size_t natoms2 = use_frame2 ? c->n_atoms2 : c->n_atoms;
    MB_TRY(need_masses(c));
    if (c->n_masses < natoms2) return fail(MB_ERR_STATE, "masses cover %zu atoms, frame2 has %zu", c->n_masses, natoms2);
    MB_TRY(check_sel(ids1, n1, c->n_atoms, "fit sel1"));
    MB_TRY(check_sel(ids2, n2, natoms2, "fit sel2"));
    const unsigned long long *d1, *d2;
    MB_TRY(upload_ids(c, c->ids1, ids1, n1, &d1));
    MB_TRY(upload_ids(c, c->ids2, ids2, n2, &d2));
    int nb = red_blocks(c, n1, 4);
    RedScratch s;
    MB_TRY(red_scratch(c, 1, (size_t)nb * 20, 16, &s));
    MB_TRY(c->host_results());
    fit_moments_kernel<true><<<dim3(nb, 1), RED_THREADS, 0, c->stream>>>(c->d_xyz, 0, d1, xyz2, d2, (int)n1,
                                                                        c->masses.as<float>(), at_origin, s.partials,
                                                                        s.tickets, c->d_res);
    c->launches++;
    MB_CUDA(cudaGetLastError());
    MB_CUDA(cudaStreamSynchronize(c->stream));
    double r[13];
    for (int k = 0; k < 13; ++k) r[k] = c->h_res[k];
    if (r[12] == 1.0) return fail(MB_ERR_ZERO_MASS, "zero mass");
    if (r[12] == 3.0) return fail(MB_ERR_SVD, "SVD failed");
    for (int col = 0; col < 3; ++col)
        for (int row = 0; row < 3; ++row) R9[col * 3 + row] = r[row * 3 + col];
    for (int i = 0; i < 3; ++i) t3[i] = r[9 + i];
    return MB_OK;
}

int mb_apply_transform(MbCtx* h, const uint64_t* ids, size_t n, const double R9[9], const double t3[3]) {
    if (!h || !R9 || !t3) return fail(MB_ERR_ARG, "null argument");
    Ctx* c = &h->c;
    if (!c->d_xyz) return fail(MB_ERR_STATE, "no frame set");
    MB_CUDA(cudaSetDevice(c->device));
    MB_TRY(check_sel(ids, n, c->n_atoms, "apply_transform"));
    const unsigned long long* d_ids;
    MB_TRY(upload_ids(c, c->ids1, ids, n, &d_ids));
    Xform X;
    for (int col = 0; col < 3; ++col)
        for (int row = 0; row < 3; ++row) X.R[row * 3 + col] = R9[col * 3 + row];
    for (int i = 0; i < 3; ++i) X.t[i] = t3[i];
    apply_transform_kernel<<<(int)((n + 255) / 256), 256, 0, c->stream>>>(const_cast<float*>(c->d_xyz), d_ids, (int)n, X);
    c->launches++;
    MB_CUDA(cudaGetLastError());
    MB_CUDA(cudaStreamSynchronize(c->stream));
    return MB_OK;
}

}  // extern "C"
