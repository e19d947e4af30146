// mb_reduce_many.cu — segmented selection reductions: thousands of small selections in ONE launch.
//
// The reference's idiom for per-residue / per-molecule analysis is a rayon loop over many small selections, each
// calling Measure::center_of_mass / center_of_geometry / gyration (molar/src/selection.rs:318-322,
// selection/par_split.rs:100-125; the per-selection bodies are measure.rs:37-45,60-87,561-570).  One
// mb_center_of_mass call per selection costs a launch and a synchronisation (~50 us) — a hundred times the CPU cost at
// 20 atoms — so this entry point takes all selections at once in CSR form (ids + offsets) and gives every selection
// to one warp (or one CTA when the selections are large): 12 B/atom + 4 B/atom of masses gathered through the ids,
// f64 accumulation of raw moments about the selection's first atom (the arithmetic of moments1_kernel), reduced in a
// fixed order, so results are deterministic and agree with the reference's f64 build to ~1e-12.
#include <algorithm>
#include <cmath>

#include "mb_common.cuh"

namespace mb {

struct ManyParams {
    const float* xyz;
    const float* masses;                 // NULL for the centre of geometry
    const unsigned long long* ids;       // concatenated selections (global atom ids)
    const unsigned long long* offsets;   // n_sel + 1
    unsigned n_sel;
    int what;                            // MB_REDUCE_*
    int width;                           // doubles per output row
    double* out;                         // n_sel x width
    int* status;                         // n_sel: 0 ok, 1 zero mass, 2 empty selection
};

template <int GROUP>
__device__ __forceinline__ double group_sum(double v, double* scratch) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (GROUP == 32) return v;
    // CTA-wide: warp totals through shared memory, summed in warp order by every thread (deterministic)
    const int w = threadIdx.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[w] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < GROUP / 32; ++k) t += scratch[k];
    return t;
}

// GROUP threads per selection: 32 (one warp, selections of up to a few hundred atoms) or 256 (one CTA)
template <int GROUP>
__global__ void __launch_bounds__(256) reduce_many_kernel(const __grid_constant__ ManyParams P) {
    __shared__ double scratch[8];
    const unsigned per_block = 256 / GROUP;
    const unsigned t = threadIdx.x % GROUP;
    // (with GROUP == 256 every thread of the CTA sees the same s, so the __syncthreads in group_sum are uniform)
    for (unsigned s = blockIdx.x * per_block + threadIdx.x / GROUP; s < P.n_sel; s += gridDim.x * per_block) {
        const unsigned long long b = P.offsets[s], e = P.offsets[s + 1];
        double* out = P.out + (size_t)s * P.width;
        if (e <= b) {
            if (t == 0) {
                for (int k = 0; k < P.width; ++k) out[k] = nan("");
                P.status[s] = 2;
            }
            continue;
        }
        const size_t g0 = (size_t)P.ids[b];
        const double ox = P.xyz[3 * g0], oy = P.xyz[3 * g0 + 1], oz = P.xyz[3 * g0 + 2];
        double m0 = 0, sx = 0, sy = 0, sz = 0, q = 0;
        for (unsigned long long k = b + t; k < e; k += GROUP) {
            const size_t g = (size_t)P.ids[k];
            const double m = P.masses ? (double)P.masses[g] : 1.0;
            const double qx = (double)P.xyz[3 * g] - ox, qy = (double)P.xyz[3 * g + 1] - oy,
                         qz = (double)P.xyz[3 * g + 2] - oz;
            m0 += m;
            sx += m * qx;
            sy += m * qy;
            sz += m * qz;
            q += m * (qx * qx + qy * qy + qz * qz);
        }
        m0 = group_sum<GROUP>(m0, scratch);
        sx = group_sum<GROUP>(sx, scratch);
        sy = group_sum<GROUP>(sy, scratch);
        sz = group_sum<GROUP>(sz, scratch);
        if (P.what != MB_REDUCE_COM && P.what != MB_REDUCE_COG) q = group_sum<GROUP>(q, scratch);
        if (t == 0) {
            if (m0 == 0.0) {  // MeasureError::ZeroMass (measure.rs:70-72)
                for (int k = 0; k < P.width; ++k) out[k] = nan("");
                P.status[s] = 1;
            } else {
                const double ax = sx / m0, ay = sy / m0, az = sz / m0;
                double r2 = q / m0 - (ax * ax + ay * ay + az * az);
                const double rg = sqrt(r2 > 0.0 ? r2 : 0.0);
                if (P.what == MB_REDUCE_GYRATION) {
                    out[0] = rg;
                } else {
                    out[0] = ox + ax;
                    out[1] = oy + ay;
                    out[2] = oz + az;
                    if (P.what == MB_REDUCE_COM_GYRATION) out[3] = rg;
                }
                P.status[s] = 0;
            }
        }
    }
}

}  // namespace mb

using namespace mb;

extern "C" {

int mb_reduce_many(MbCtx* h, const uint64_t* ids, const uint64_t* offsets, size_t n_sel, int what, double* out,
                   int* status_out) {
    if (!h || !ids || !offsets || !out) return fail(MB_ERR_ARG, "mb_reduce_many: null argument");
    Ctx& c = h->c;
    if (!c.d_xyz) return fail(MB_ERR_STATE, "no frame set");
    if (n_sel == 0) return fail(MB_ERR_ARG, "mb_reduce_many: no selections");
    if (n_sel > 0x7fffffffull) return fail(MB_ERR_ARG, "mb_reduce_many: too many selections");
    int width;
    switch (what) {
        case MB_REDUCE_COM: width = 3; break;
        case MB_REDUCE_COG: width = 3; break;
        case MB_REDUCE_GYRATION: width = 1; break;
        case MB_REDUCE_COM_GYRATION: width = 4; break;
        default: return fail(MB_ERR_ARG, "mb_reduce_many: unknown reduction %d", what);
    }
    const bool need_mass = what != MB_REDUCE_COG;
    if (need_mass && (!c.masses.p || c.n_masses < c.n_atoms))
        return fail(MB_ERR_STATE, "masses not set (mb_set_masses) for %zu atoms", c.n_atoms);
    // offsets must be non-decreasing; every selection strictly increasing and in range (one host pass over the ids)
    if (offsets[0] != 0) return fail(MB_ERR_ARG, "mb_reduce_many: offsets[0] must be 0");
    const size_t total = (size_t)offsets[n_sel];
    for (size_t s = 0; s < n_sel; ++s) {
        if (offsets[s + 1] < offsets[s]) return fail(MB_ERR_ARG, "mb_reduce_many: offsets must be non-decreasing");
        const size_t b = (size_t)offsets[s], e = (size_t)offsets[s + 1];
        uint64_t bad = 0;
        for (size_t k = b + 1; k < e; ++k) bad |= (uint64_t)(ids[k] <= ids[k - 1]);
        if (bad) return fail(MB_ERR_ARG, "mb_reduce_many: selection %zu is not strictly increasing", s);
        if (e > b && ids[e - 1] >= c.n_atoms)
            return fail(MB_ERR_ARG, "mb_reduce_many: selection %zu: index %llu out of range (%zu atoms)", s,
                        (unsigned long long)ids[e - 1], c.n_atoms);
    }
    MB_CUDA(cudaSetDevice(c.device));
    // device staging: ids | offsets | out | status
    const size_t ids_b = (total * 8 + 255) / 256 * 256, off_b = ((n_sel + 1) * 8 + 255) / 256 * 256;
    const size_t out_b = (n_sel * (size_t)width * 8 + 255) / 256 * 256, st_b = (n_sel * 4 + 255) / 256 * 256;
    MB_TRY(c.many_tmp.reserve(ids_b + off_b + out_b + st_b + 256));
    char* base = static_cast<char*>(c.many_tmp.p);
    ManyParams P;
    P.xyz = c.d_xyz;
    P.masses = need_mass ? c.masses.as<float>() : nullptr;
    P.ids = reinterpret_cast<unsigned long long*>(base);
    P.offsets = reinterpret_cast<unsigned long long*>(base + ids_b);
    P.out = reinterpret_cast<double*>(base + ids_b + off_b);
    P.status = reinterpret_cast<int*>(base + ids_b + off_b + out_b);
    P.n_sel = (unsigned)n_sel;
    P.what = what;
    P.width = width;
    if (total) MB_CUDA(cudaMemcpyAsync(base, ids, total * 8, cudaMemcpyHostToDevice, c.stream));
    MB_CUDA(cudaMemcpyAsync(base + ids_b, offsets, (n_sel + 1) * 8, cudaMemcpyHostToDevice, c.stream));
    const double mean = (double)total / (double)n_sel;
    if (mean > 1024.0) {
        const unsigned blocks = (unsigned)std::min<size_t>(n_sel, (size_t)c.sm_count * 16);
        reduce_many_kernel<256><<<blocks, 256, 0, c.stream>>>(P);
    } else {
        const unsigned blocks = (unsigned)std::min<size_t>((n_sel + 7) / 8, (size_t)c.sm_count * 16);
        reduce_many_kernel<32><<<blocks, 256, 0, c.stream>>>(P);
    }
    c.launches++;
    MB_CUDA(cudaGetLastError());
    MB_CUDA(cudaMemcpyAsync(out, P.out, n_sel * (size_t)width * 8, cudaMemcpyDeviceToHost, c.stream));
    std::vector<int> st_local;
    int* st = status_out;
    if (!st) {
        st_local.resize(n_sel);
        st = st_local.data();
    }
    MB_CUDA(cudaMemcpyAsync(st, P.status, n_sel * sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    MB_CUDA(cudaStreamSynchronize(c.stream));
    // the first failing selection decides the return code (the caller has the per-selection status as well)
    for (size_t s = 0; s < n_sel; ++s) {
        if (st[s] == 1) return fail(MB_ERR_ZERO_MASS, "mb_reduce_many: selection %zu has zero mass", s);
        if (st[s] == 2) return fail(MB_ERR_ARG, "mb_reduce_many: selection %zu is empty", s);
    }
    return MB_OK;
}

}  // extern "C"
