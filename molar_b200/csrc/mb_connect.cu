// mb_connect.cu — consumers of the pair list that stay on the device.
//
// Replaces
//   SearchConnectivity::from_iter            molar/src/connectivity.rs:8-38   (atom -> neighbours, both directions)
//   Modify::unwrap_connectivity[_dim]        molar/src/modify.rs:64-131       (contact graph, walk, closest images)
// so that 10^8 pairs do not have to travel to the host to be turned into an adjacency map.
//
// Adjacency is CSR over the atom index space: degree count (atomics) -> exclusive scan -> fill (atomics).  The
// order of the neighbours inside a row is unspecified — in the reference it is the order rayon happened to
// deliver the pairs in.
//
// unwrap_connectivity.  The reference walks the graph depth-first from the lowest unused atom and moves every
// newly reached atom to its closest image next to the atom it was reached from.  Which spanning tree that is
// depends on the pair order, i.e. it is not a defined quantity of the reference; what is defined are the
// components, their start atoms (lowest index of each component) and — whenever the molecule is smaller than
// half the box — the unwrapped positions up to f32 rounding.  Here:
//   components   lock-free union-find over the pair list, larger root hooked under the smaller one, so the
//                representative of a component IS its lowest index = the reference's start atom;
//   unwrapping   level-synchronous breadth-first walk from all start atoms at once in ONE persistent kernel
//                (grid barrier between levels); a reached atom takes the LOWEST-index frontier neighbour as
//                its parent, which makes the result deterministic.
#include <algorithm>
#include <cstring>
#include <vector>

#include "mb_common.cuh"

namespace mb {

// ---- CSR adjacency ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) csr_count_kernel(const uint2* __restrict__ pairs, unsigned long long np,
                                                        unsigned* __restrict__ deg) {
    for (unsigned long long k = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; k < np;
         k += (unsigned long long)gridDim.x * blockDim.x) {
        const uint2 p = pairs[k];
        atomicAdd(&deg[p.x], 1u);
        atomicAdd(&deg[p.y], 1u);
    }
}
__global__ void __launch_bounds__(256) csr_fill_kernel(const uint2* __restrict__ pairs, unsigned long long np,
                                                       unsigned* __restrict__ cursor, unsigned* __restrict__ cols) {
    for (unsigned long long k = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; k < np;
         k += (unsigned long long)gridDim.x * blockDim.x) {
        const uint2 p = pairs[k];
        cols[atomicAdd(&cursor[p.x], 1u)] = p.y;
        cols[atomicAdd(&cursor[p.y], 1u)] = p.x;
    }
}
// order-independent checksum of the adjacency: the hash of mb_pairs_checksum (mix64((min << 32) | max), summed and
// xor-ed) taken once over the entries with row < column and once over the entries with row > column — a symmetric
// neighbour list gives the pair list's checksum twice.  One warp per row.
__device__ __forceinline__ unsigned long long mix64c(unsigned long long x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}
__global__ void __launch_bounds__(256) csr_checksum_kernel(const unsigned* __restrict__ row_ptr, const unsigned* __restrict__ cols,
                                                           unsigned n, unsigned long long* __restrict__ out4) {
    const unsigned lane = threadIdx.x & 31u, nw = (gridDim.x * blockDim.x) >> 5;
    unsigned long long s0 = 0, x0 = 0, s1 = 0, x1 = 0;
    for (unsigned i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += nw) {
        const unsigned e1 = row_ptr[i + 1];
        for (unsigned e = row_ptr[i] + lane; e < e1; e += 32) {
            const unsigned j = cols[e];
            if (i < j) {
                const unsigned long long h = mix64c(((unsigned long long)i << 32) | j);
                s0 += h;
                x0 ^= h;
            } else {
                const unsigned long long h = mix64c(((unsigned long long)j << 32) | i);
                s1 += h;
                x1 ^= h;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        x0 ^= __shfl_xor_sync(0xffffffffu, x0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        x1 ^= __shfl_xor_sync(0xffffffffu, x1, o);
    }
    if (lane == 0) {
        atomicAdd(&out4[0], s0);
        atomicXor(&out4[1], x0);
        atomicAdd(&out4[2], s1);
        atomicXor(&out4[3], x1);
    }
}
__global__ void widen_u32_kernel(const unsigned* __restrict__ in, size_t n, unsigned long long* __restrict__ out) {
    size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (k < n) out[k] = in[k];
}

// ---- union-find (larger root under smaller root) ---------------------------------------------
__device__ __forceinline__ unsigned uf_find(unsigned* parent, unsigned x) {
    unsigned p = parent[x];
    while (p != x) {
        const unsigned g = parent[p];
        if (g != p) parent[x] = g;  // path halving (benign race: only ever replaces a parent by an ancestor)
        x = p;
        p = g;
    }
    return x;
}
__global__ void uf_init_kernel(unsigned* __restrict__ parent, unsigned n) {
    unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) parent[k] = k;
}
__global__ void __launch_bounds__(256) uf_unite_kernel(const uint2* __restrict__ pairs, unsigned long long np,
                                                       unsigned* parent) {
    for (unsigned long long k = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; k < np;
         k += (unsigned long long)gridDim.x * blockDim.x) {
        const uint2 p = pairs[k];
        unsigned a = uf_find(parent, p.x), b = uf_find(parent, p.y);
        while (a != b) {
            if (a < b) {
                const unsigned t = a;
                a = b;
                b = t;
            }
            // a > b: hook a under b if a is still a root
            const unsigned old = atomicCAS(&parent[a], a, b);
            if (old == a) break;
            a = uf_find(parent, old);
            b = uf_find(parent, b);
        }
    }
}
// same over CSR rows (one warp per row; every edge is stored in both rows: the higher end unites)
__global__ void __launch_bounds__(256) uf_unite_rows_kernel(const unsigned* __restrict__ row_ptr,
                                                            const unsigned* __restrict__ cols, unsigned n, unsigned* parent) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned nw = (gridDim.x * blockDim.x) >> 5;
    for (unsigned i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += nw) {
        const unsigned e1 = row_ptr[i + 1];
        for (unsigned e = row_ptr[i] + lane; e < e1; e += 32) {
            const unsigned j = cols[e];
            if (j >= i) continue;
            unsigned a = uf_find(parent, i), b = uf_find(parent, j);
            while (a != b) {
                if (a < b) {
                    const unsigned t = a;
                    a = b;
                    b = t;
                }
                const unsigned old = atomicCAS(&parent[a], a, b);
                if (old == a) break;
                a = uf_find(parent, old);
                b = uf_find(parent, b);
            }
        }
    }
}
__global__ void uf_flatten_kernel(unsigned* parent, unsigned n) {
    unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) {
        unsigned r = k;
        while (parent[r] != r) r = parent[r];
        parent[k] = r;
    }
}

// ---- breadth-first unwrapping -----------------------------------------------------------------
struct BfsParams {
    float* xyz;                 // frame, modified in place
    const unsigned* row_ptr;
    const unsigned* cols;
    const unsigned* label;      // flattened union-find: label[i] == i <=> start atom
    const unsigned char* member;  // 1 for atoms of the selection
    unsigned char* visited;
    unsigned* cand;             // lowest-index frontier neighbour proposing to adopt the atom (init ~0)
    unsigned* frontier[2];
    unsigned* count;            // [0],[1] sizes of the two frontiers, [2] barrier arrivals, [3] barrier generation, [4] levels
    unsigned n;
    unsigned dims;
    DevBox box;
};

__device__ __forceinline__ void grid_barrier(unsigned* count, unsigned nblocks) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        volatile unsigned* gen = count + 3;
        const unsigned g = *gen;
        if (atomicAdd(count + 2, 1u) == nblocks - 1) {
            count[2] = 0;
            __threadfence();
            atomicAdd(count + 3, 1u);
        } else {
            while (*gen == g) __nanosleep(64);
        }
        __threadfence();
    }
    __syncthreads();
}

constexpr int BFS_THREADS = 256;
__global__ void __launch_bounds__(BFS_THREADS) unwrap_bfs_kernel(const __grid_constant__ BfsParams P) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned warp = (blockIdx.x * BFS_THREADS + threadIdx.x) >> 5, nwarps = (gridDim.x * BFS_THREADS) >> 5;
    // level 0: every start atom of the selection
    for (unsigned i = blockIdx.x * BFS_THREADS + threadIdx.x; i < P.n; i += gridDim.x * BFS_THREADS)
        if (P.member[i] && P.label[i] == i) {
            P.visited[i] = 1;
            P.frontier[0][atomicAdd(P.count, 1u)] = i;
        }
    grid_barrier(P.count, gridDim.x);
    unsigned cur = 0;
    for (;;) {
        const unsigned nf = *((volatile unsigned*)(P.count + cur));
        if (nf == 0) break;
        const unsigned* fr = P.frontier[cur];
        // Everything other CTAs wrote in an earlier phase (frontier, visited, cand, positions) is read with
        // __ldcg: L1 is not coherent across SMs and a grid barrier does not invalidate it.
        // phase 1: every frontier atom proposes itself to its unvisited neighbours; the lowest index wins
        for (unsigned w = warp; w < nf; w += nwarps) {
            const unsigned u = __ldcg(fr + w);
            const unsigned e0 = P.row_ptr[u], e1 = P.row_ptr[u + 1];
            for (unsigned e = e0 + lane; e < e1; e += 32) {
                const unsigned v = P.cols[e];
                if (!__ldcg(P.visited + v)) atomicMin(&P.cand[v], u);
            }
        }
        grid_barrier(P.count, gridDim.x);
        // phase 2: the winner moves the atom next to itself and puts it on the next frontier
        unsigned* nx = P.frontier[cur ^ 1u];
        for (unsigned w = warp; w < nf; w += nwarps) {
            const unsigned u = __ldcg(fr + w);
            const float ux = __ldcg(P.xyz + 3 * (size_t)u), uy = __ldcg(P.xyz + 3 * (size_t)u + 1),
                        uz = __ldcg(P.xyz + 3 * (size_t)u + 2);
            const unsigned e0 = P.row_ptr[u], e1 = P.row_ptr[u + 1];
            for (unsigned e = e0 + lane; e < e1; e += 32) {
                const unsigned v = P.cols[e];
                if (!__ldcg(P.visited + v) && __ldcg(P.cand + v) == u) {
                    float* pv = P.xyz + 3 * (size_t)v;
                    float s0, s1, s2;
                    // closest_image_dims(p, p0, dims) = p0 + shortest_vector_dims(p - p0, dims)  (periodic_box.rs:327-330)
                    shortest_vector_dev(P.box, xsub(__ldcg(pv), ux), xsub(__ldcg(pv + 1), uy), xsub(__ldcg(pv + 2), uz), P.dims,
                                        s0, s1, s2);
                    pv[0] = xadd(ux, s0);
                    pv[1] = xadd(uy, s1);
                    pv[2] = xadd(uz, s2);
                    P.visited[v] = 1;
                    nx[atomicAdd(P.count + (cur ^ 1u), 1u)] = v;
                }
            }
        }
        grid_barrier(P.count, gridDim.x);
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            P.count[cur] = 0;  // this frontier is consumed; it becomes the target of the level after next
            P.count[4] += 1;
        }
        cur ^= 1u;
        grid_barrier(P.count, gridDim.x);
    }
}

__global__ void mark_members_kernel(const unsigned long long* __restrict__ ids, unsigned n, unsigned char* __restrict__ member) {
    unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) member[ids ? ids[k] : k] = 1;
}

// CSR of the context's current pair list over [0, n_index)
static int build_csr(Ctx* c, size_t n_index, unsigned** row_ptr, unsigned** cols, size_t* nnz) {
    if (c->last.kind != 1 && c->last.kind != 2) return fail(MB_ERR_STATE, "no pair list on this context");
    if (n_index == 0 || n_index > 0x7fffffffull) return fail(MB_ERR_ARG, "connectivity: bad index space");
    const unsigned long long np = (unsigned long long)c->last.count;
    if (2 * np > 0xfffffff0ull) return fail(MB_ERR_ARG, "connectivity: more than 2^31 pairs");
    const size_t words = 2 * (n_index + 2);
    MB_TRY(c->conn_tmp.reserve(words * sizeof(unsigned)));
    MB_TRY(c->conn_cols.reserve((2 * np + 1) * sizeof(unsigned)));
    unsigned* deg = c->conn_tmp.as<unsigned>();
    unsigned* rp = deg + (n_index + 2);
    MB_CUDA(cudaMemsetAsync(deg, 0, (n_index + 2) * sizeof(unsigned), c->stream));
    const int blocks = (int)std::max<unsigned long long>(1, std::min<unsigned long long>((np + 255) / 256, (unsigned long long)c->sm_count * 16));
    const uint2* pairs = c->pairs.as<uint2>();
    if (np) csr_count_kernel<<<blocks, 256, 0, c->stream>>>(pairs, np, deg);
    MB_TRY(exclusive_scan_u32(c, deg, (int)n_index, rp));
    // the fill cursors start at the row starts: reuse `deg`
    MB_CUDA(cudaMemcpyAsync(deg, rp, n_index * sizeof(unsigned), cudaMemcpyDeviceToDevice, c->stream));
    if (np) csr_fill_kernel<<<blocks, 256, 0, c->stream>>>(pairs, np, deg, c->conn_cols.as<unsigned>());
    c->launches += np ? 2 : 0;
    MB_CUDA(cudaGetLastError());
    *row_ptr = rp;
    *cols = c->conn_cols.as<unsigned>();
    *nnz = (size_t)(2 * np);
    c->conn_n = n_index;
    c->conn_nnz = *nnz;
    return MB_OK;
}

}  // namespace mb

using namespace mb;

extern "C" {

int64_t mb_connectivity(MbCtx* h, size_t n_index, uint64_t* row_ptr_out) {
    if (!h) return fail(MB_ERR_ARG, "null context");
    Ctx* c = &h->c;
    MB_CUDA(cudaSetDevice(c->device));
    unsigned *rp = nullptr, *cols = nullptr;
    size_t nnz = 0;
    MB_TRY(build_csr(c, n_index, &rp, &cols, &nnz));
    if (row_ptr_out) {
        MB_TRY(c->out_ids.reserve((n_index + 1) * sizeof(unsigned long long)));
        widen_u32_kernel<<<(unsigned)((n_index + 1 + 255) / 256), 256, 0, c->stream>>>(rp, n_index + 1,
                                                                                       c->out_ids.as<unsigned long long>());
        c->launches++;
        MB_CUDA(cudaMemcpyAsync(row_ptr_out, c->out_ids.p, (n_index + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
    }
    MB_CUDA(cudaStreamSynchronize(c->stream));
    return (int64_t)nnz;
}

int64_t mb_search_connectivity(MbCtx* h, float cutoff, const uint64_t* ids, size_t n, uint8_t pbc_dims,
                               uint64_t* row_ptr_out) {
    if (!h) return fail(MB_ERR_ARG, "null context");
    Ctx* c = &h->c;
    const size_t N = c->n_atoms;
    bool rows_done = false;
    MB_TRY(neighbor_rows_cells(c, cutoff, ids, n, pbc_dims & 7u, N, 0, &rows_done));
    unsigned *rp = nullptr, *cols = nullptr;
    size_t nnz = 0;
    if (rows_done) {
        rp = c->conn_tmp.as<unsigned>() + (N + 2);
        nnz = c->conn_nnz;
    } else {
        // small or degenerate grids: pair list of the general path -> CSR
        const int keep_dist = c->opt_with_dist;
        c->opt_with_dist = 0;
        int64_t np = 0;
        int rc = search_single_impl(c, cutoff, ids, n, pbc_dims & 7u, 0, &np);
        c->opt_with_dist = keep_dist;
        if (rc < 0) return rc;
        MB_TRY(build_csr(c, N, &rp, &cols, &nnz));
    }
    if (row_ptr_out) {
        MB_TRY(c->out_ids.reserve((N + 1) * sizeof(unsigned long long)));
        widen_u32_kernel<<<(unsigned)((N + 1 + 255) / 256), 256, 0, c->stream>>>(rp, N + 1, c->out_ids.as<unsigned long long>());
        c->launches++;
        MB_CUDA(cudaMemcpyAsync(row_ptr_out, c->out_ids.p, (N + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
    }
    MB_CUDA(cudaStreamSynchronize(c->stream));
    return (int64_t)nnz;
}

int mb_connectivity_checksum(MbCtx* h, uint64_t out4[4]) {
    if (!h || !out4) return fail(MB_ERR_ARG, "null argument");
    Ctx* c = &h->c;
    if (!c->conn_cols.p || c->conn_n == 0) return fail(MB_ERR_STATE, "no connectivity on this context");
    MB_CUDA(cudaSetDevice(c->device));
    MB_TRY(c->counters.reserve(256));
    unsigned long long* d4 = c->counters.as<unsigned long long>() + 8;
    MB_CUDA(cudaMemsetAsync(d4, 0, 4 * sizeof(unsigned long long), c->stream));
    const unsigned* rp = c->conn_tmp.as<unsigned>() + (c->conn_n + 2);
    csr_checksum_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(rp, c->conn_cols.as<unsigned>(), (unsigned)c->conn_n, d4);
    c->launches++;
    MB_CUDA(cudaMemcpyAsync(out4, d4, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
    MB_CUDA(cudaStreamSynchronize(c->stream));
    return MB_OK;
}

int mb_fill_connectivity(MbCtx* h, uint64_t* cols_out) {
    if (!h || !cols_out) return fail(MB_ERR_ARG, "null argument");
    Ctx* c = &h->c;
    if (!c->conn_cols.p || c->conn_n == 0) return fail(MB_ERR_STATE, "no connectivity on this context");
    MB_CUDA(cudaSetDevice(c->device));
    if (c->conn_nnz == 0) return MB_OK;
    MB_TRY(c->out_ids.reserve(c->conn_nnz * sizeof(unsigned long long)));
    widen_u32_kernel<<<(unsigned)((c->conn_nnz + 255) / 256), 256, 0, c->stream>>>(c->conn_cols.as<unsigned>(), c->conn_nnz,
                                                                                  c->out_ids.as<unsigned long long>());
    c->launches++;
    MB_CUDA(cudaMemcpyAsync(cols_out, c->out_ids.p, c->conn_nnz * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
    MB_CUDA(cudaStreamSynchronize(c->stream));
    return MB_OK;
}

int64_t mb_unwrap_connectivity(MbCtx* h, float cutoff, const uint64_t* ids, size_t n, uint8_t image_dims,
                               int64_t* roots_out) {
    if (!h) return fail(MB_ERR_ARG, "null context");
    Ctx* c = &h->c;
    if (!c->d_xyz) return fail(MB_ERR_STATE, "no frame set");
    if (!c->has_box) return fail(MB_ERR_NO_PBC, "unwrap_connectivity: the frame has no periodic box");
    MB_CUDA(cudaSetDevice(c->device));
    // contact graph: distance_search_single_pbc(cutoff, ..., PBC_FULL)   (modify.rs:79-80), as neighbour rows written
    // by the search itself when the grid allows it, else pair list -> CSR
    const size_t N = c->n_atoms;
    // scratch after the CSR words: label[N] cand[N] frontier0[N] frontier1[N] count[8] visited[N] member[N]
    const size_t csr_words = 2 * (N + 2);
    const size_t extra_bytes = (4 * N + 8) * sizeof(unsigned) + 2 * N + 64;
    bool rows_done = false;
    MB_TRY(neighbor_rows_cells(c, cutoff, ids, n, 7, N, extra_bytes, &rows_done));
    unsigned *rp = nullptr, *cols = nullptr;
    size_t nnz = 0;
    int64_t np = 0;
    if (rows_done) {
        rp = c->conn_tmp.as<unsigned>() + (N + 2);
        cols = c->conn_cols.as<unsigned>();
        nnz = c->conn_nnz;
    } else {
        const int keep_dist = c->opt_with_dist;
        c->opt_with_dist = 0;
        int rc = search_single_impl(c, cutoff, ids, n, 7, 0, &np);
        c->opt_with_dist = keep_dist;
        if (rc < 0) return rc;
        MB_TRY(c->conn_tmp.reserve(csr_words * sizeof(unsigned) + extra_bytes));
        MB_TRY(build_csr(c, N, &rp, &cols, &nnz));
    }
    unsigned* base = c->conn_tmp.as<unsigned>() + csr_words;
    unsigned* label = base;
    unsigned* cand = base + N;
    unsigned* fr0 = base + 2 * N;
    unsigned* fr1 = base + 3 * N;
    unsigned* count = base + 4 * N;
    unsigned char* visited = reinterpret_cast<unsigned char*>(count + 8);
    unsigned char* member = visited + N;
    const unsigned nb = (unsigned)((N + 255) / 256);
    uf_init_kernel<<<nb, 256, 0, c->stream>>>(label, (unsigned)N);
    if (rows_done) {
        if (nnz) uf_unite_rows_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(rp, cols, (unsigned)N, label);
    } else {
        const unsigned long long P = (unsigned long long)np;
        const int pblocks = (int)std::max<unsigned long long>(1, std::min<unsigned long long>((P + 255) / 256, (unsigned long long)c->sm_count * 16));
        if (P) uf_unite_kernel<<<pblocks, 256, 0, c->stream>>>(c->pairs.as<uint2>(), P, label);
    }
    uf_flatten_kernel<<<nb, 256, 0, c->stream>>>(label, (unsigned)N);
    MB_CUDA(cudaMemsetAsync(cand, 0xff, N * sizeof(unsigned), c->stream));
    MB_CUDA(cudaMemsetAsync(count, 0, 8 * sizeof(unsigned), c->stream));
    MB_CUDA(cudaMemsetAsync(visited, 0, 2 * N, c->stream));
    const unsigned long long* d_ids = nullptr;
    if (ids) d_ids = c->ids1.as<unsigned long long>();  // uploaded by search_single_impl
    mark_members_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(d_ids, (unsigned)n, member);
    BfsParams B;
    memset(&B, 0, sizeof(B));
    B.xyz = const_cast<float*>(c->d_xyz);
    B.row_ptr = rp;
    B.cols = cols;
    B.label = label;
    B.member = member;
    B.visited = visited;
    B.cand = cand;
    B.frontier[0] = fr0;
    B.frontier[1] = fr1;
    B.count = count;
    B.n = (unsigned)N;
    B.dims = image_dims & 7u;
    B.box = to_dev_box(c->box);
    int per_sm = 1;
    MB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, unwrap_bfs_kernel, BFS_THREADS, 0));
    const int grid = c->sm_count * std::max(1, std::min(per_sm, 2));  // all CTAs must be co-resident (grid barrier)
    void* args[] = {&B};
    MB_CUDA(cudaLaunchCooperativeKernel((const void*)unwrap_bfs_kernel, dim3(grid), dim3(BFS_THREADS), args, 0, c->stream));
    c->launches += 5;
    MB_CUDA(cudaGetLastError());
    // labels -> host; start atoms of the selection
    std::vector<unsigned> hl(N);
    MB_CUDA(cudaMemcpyAsync(hl.data(), label, N * sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
    MB_CUDA(cudaStreamSynchronize(c->stream));
    int64_t nstart = 0;
    if (ids) {
        // ids are sorted: position of a start atom within the selection by binary search
        for (size_t k = 0; k < n; ++k) {
            const unsigned r = hl[ids[k]];
            if (r == ids[k]) ++nstart;
            if (roots_out) roots_out[k] = (int64_t)(std::lower_bound(ids, ids + n, (uint64_t)r) - ids);
        }
    } else {
        for (size_t k = 0; k < n; ++k) {
            if (hl[k] == k) ++nstart;
            if (roots_out) roots_out[k] = (int64_t)hl[k];
        }
    }
    return nstart;
}

}  // extern "C"
