// mb_batch.cu — device-resident trajectory: synthetic frame generator and the batched drivers
// that replace the per-frame loop of AnalysisTask::run (molar/src/analysis_task.rs:113-280): many
// frames stay resident in HBM, the per-frame kernels are enqueued back to back on one stream and
// only per-frame scalars come back to the host.
#include <algorithm>
#include <cstring>
#include <vector>

#include "mb_common.cuh"

namespace mb {

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    unsigned long long z = x + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__device__ __forceinline__ float unit_float(unsigned long long u) {
    return (float)(u >> 40) * 5.9604644775390625e-08f;  // exact: 24-bit integer * 2^-24
}

// SURVEY.md §8(d) generator; bit-identical to the oracle's orc_synth_frame.
__global__ void __launch_bounds__(256) synth_kernel(float* __restrict__ out, unsigned long long seed,
                                                    unsigned long long first_frame, size_t n_frames, size_t n_atoms,
                                                    DevBox box, int stray_permille) {
    size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (idx >= n_frames * n_atoms) return;
    size_t f = idx / n_atoms, a = idx % n_atoms;
    unsigned long long frame = first_frame + f;
    float s[3];
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) s[ax] = unit_float(splitmix64(seed ^ (frame << 32) ^ (unsigned long long)(a * 3 + ax)));
    if (stray_permille > 0) {
        unsigned long long h = splitmix64(seed ^ (frame << 32) ^ 0x5BD1E995C0FFEEULL ^ ((unsigned long long)a << 2));
        if ((int)(h % 1000ull) < stray_permille) {
            int dim = (int)((h >> 20) % 3ull);
            float sh = ((h >> 40) & 1ull) ? 1.0f : -1.0f;
            if (dim == 0) s[0] = xadd(s[0], sh);
            else if (dim == 1) s[1] = xadd(s[1], sh);
            else s[2] = xadd(s[2], sh);
        }
    }
    float x, y, z;
    xmatvec(box.m, s[0], s[1], s[2], x, y, z);
    out[3 * idx] = x;
    out[3 * idx + 1] = y;
    out[3 * idx + 2] = z;
}

__global__ void __launch_bounds__(256) synth_masses_kernel(float* __restrict__ out, unsigned long long seed, size_t n) {
    size_t a = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (a >= n) return;
    float s = unit_float(splitmix64(seed ^ 0xA5A5A5A5DEADBEEFULL ^ (unsigned long long)a));
    out[a] = xadd(1.0f, xmul(15.0f, s));
}

// rows8 (com xyz, rg, M, status, -, -) + u64 counters (stride 4: pairs, work, tests, -) -> rows5 {com, rg, count}
__global__ void assemble_rows_kernel(const double* __restrict__ rows8, const unsigned long long* __restrict__ counters,
                                     int nf, double* __restrict__ rows5) {
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf) return;
    rows5[5 * f] = rows8[8 * f];
    rows5[5 * f + 1] = rows8[8 * f + 1];
    rows5[5 * f + 2] = rows8[8 * f + 2];
    rows5[5 * f + 3] = rows8[8 * f + 3];
    rows5[5 * f + 4] = (double)counters[4 * f];
}

// Streaming over HOST frames: what the reference's trajectory loop does with an IO thread feeding a bounded
// channel while the consumer analyses (io.rs:209-233, analysis_task.rs:113-280), done with a copy stream of its own
// and a ring of three chunks: uploads are issued two chunks ahead of the chunk being processed, so the copy engine
// always has work queued while work() (which is synchronous) runs, and never waits behind a compute stream.
// work(b0, nf, f0): process ring frames [b0, b0+nf) = stream frames [f0, f0+nf); synchronous.
template <class Work>
static int stream_chunks(Ctx& c, const float* frames, size_t n_frames, size_t n_atoms, size_t chunk, Work&& work) {
    MB_CUDA(cudaSetDevice(c.device));
    constexpr size_t RING = 3;
    chunk = std::max<size_t>(1, std::min(chunk, n_frames));
    const size_t fbytes = n_atoms * 3 * sizeof(float);
    MB_TRY(c.batch.reserve(RING * chunk * fbytes));
    c.batch_frames = RING * chunk;
    c.batch_atoms = n_atoms;
    c.d_xyz = c.batch.as<float>();
    c.n_atoms = n_atoms;
    if (!c.copy_stream) MB_CUDA(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
    cudaStream_t copy = c.copy_stream;
    cudaEvent_t up[RING] = {nullptr, nullptr, nullptr};
    for (size_t k = 0; k < RING; ++k) MB_CUDA(cudaEventCreateWithFlags(&up[k], cudaEventDisableTiming));
    // Chunk schedule: the upload of the FIRST chunk is the only one nothing overlaps, so the stream starts with small
    // chunks (chunk/8, chunk/4, chunk/2 frames) and then runs at full size.  With eight ranks sharing the host's PCIe
    // root (20 GB/s per rank instead of 46) a full first chunk of 8 x 12 MB cost 4.8 ms of every 32-frame call.
    std::vector<std::pair<size_t, size_t>> sched;  // (first frame, frames)
    {
        size_t f = 0;
        for (size_t div = 8; f < n_frames; div = div > 1 ? div / 2 : 1) {
            const size_t nf = std::min(std::max<size_t>(1, chunk / div), n_frames - f);
            sched.emplace_back(f, nf);
            f += nf;
        }
    }
    const size_t nchunks = sched.size();
    auto upload = [&](size_t k) -> int {
        const size_t f0 = sched[k].first, nf = sched[k].second;
        MB_CUDA(cudaMemcpyAsync(c.batch.as<char>() + (k % RING) * chunk * fbytes, frames + f0 * n_atoms * 3, nf * fbytes,
                                cudaMemcpyHostToDevice, copy));
        MB_CUDA(cudaEventRecord(up[k % RING], copy));
        return MB_OK;
    };
    int rc = upload(0);
    if (rc == MB_OK && nchunks > 1) rc = upload(1);
    for (size_t k = 0; k < nchunks && rc == MB_OK; ++k) {
        // ring slot (k+2) % 3 held chunk k-1, which the previous (synchronous) work() call has finished with
        if (k + 2 < nchunks) rc = upload(k + 2);
        if (rc != MB_OK) break;
        cudaStreamWaitEvent(c.stream, up[k % RING], 0);
        rc = work((k % RING) * chunk, sched[k].second, sched[k].first);
    }
    cudaStreamSynchronize(copy);
    for (size_t k = 0; k < RING; ++k) cudaEventDestroy(up[k]);
    return rc;
}

// box9 == NULL means "no box", exactly as in mb_set_frame / mb_batch_upload: a box set by an earlier call is not
// inherited (the periodic entry points then fail with MB_ERR_NO_PBC instead of searching with a stale box)
static int stream_box(Ctx& c, const float* box9) {
    if (box9) {
        MB_TRY(host_box_from_colmajor(box9, &c.box));
        c.has_box = true;
    } else {
        c.has_box = false;
    }
    return MB_OK;
}

}  // namespace mb

using namespace mb;

extern "C" {

int mb_batch_synth(MbCtx* h, uint64_t seed, uint64_t first_frame, size_t n_frames, size_t n_atoms, const float* box9,
                   int stray_permille) {
    if (!h || !box9 || n_frames == 0 || n_atoms == 0) return fail(MB_ERR_ARG, "mb_batch_synth: bad argument");
    Ctx& c = h->c;
    MB_CUDA(cudaSetDevice(c.device));
    MB_TRY(host_box_from_colmajor(box9, &c.box));
    c.has_box = true;
    size_t total = n_frames * n_atoms;
    MB_TRY(c.batch.reserve(total * 3 * sizeof(float)));
    size_t blocks = (total + 255) / 256;
    if (blocks > 0x7fffffffull) return fail(MB_ERR_ARG, "batch too large");
    synth_kernel<<<(unsigned)blocks, 256, 0, c.stream>>>(c.batch.as<float>(), seed, first_frame, n_frames, n_atoms,
                                                        to_dev_box(c.box), stray_permille);
    c.launches++;
    MB_CUDA(cudaGetLastError());
    MB_CUDA(cudaStreamSynchronize(c.stream));
    c.batch_frames = n_frames;
    c.batch_atoms = n_atoms;
    c.d_xyz = c.batch.as<float>();
    c.n_atoms = n_atoms;
    return MB_OK;
}

int mb_batch_upload(MbCtx* h, const float* xyz, size_t n_frames, size_t n_atoms, const float* box9) {
    if (!h || !xyz || n_frames == 0 || n_atoms == 0) return fail(MB_ERR_ARG, "mb_batch_upload: bad argument");
    Ctx& c = h->c;
    MB_CUDA(cudaSetDevice(c.device));
    if (box9) {
        MB_TRY(host_box_from_colmajor(box9, &c.box));
        c.has_box = true;
    } else {
        c.has_box = false;
    }
    size_t bytes = n_frames * n_atoms * 3 * sizeof(float);
    MB_TRY(c.batch.reserve(bytes));
    MB_CUDA(cudaMemcpyAsync(c.batch.p, xyz, bytes, cudaMemcpyHostToDevice, c.stream));
    MB_CUDA(cudaStreamSynchronize(c.stream));
    c.batch_frames = n_frames;
    c.batch_atoms = n_atoms;
    c.d_xyz = c.batch.as<float>();
    c.n_atoms = n_atoms;
    return MB_OK;
}

int mb_batch_synth_masses(MbCtx* h, uint64_t seed, size_t n_atoms) {
    if (!h || n_atoms == 0) return fail(MB_ERR_ARG, "mb_batch_synth_masses: bad argument");
    Ctx& c = h->c;
    MB_CUDA(cudaSetDevice(c.device));
    MB_TRY(c.masses.reserve(n_atoms * sizeof(float)));
    synth_masses_kernel<<<(unsigned)((n_atoms + 255) / 256), 256, 0, c.stream>>>(c.masses.as<float>(), seed, n_atoms);
    c.launches++;
    MB_CUDA(cudaGetLastError());
    MB_CUDA(cudaStreamSynchronize(c.stream));
    c.n_masses = n_atoms;
    return MB_OK;
}

int mb_batch_select(MbCtx* h, size_t frame) {
    if (!h) return fail(MB_ERR_ARG, "null context");
    Ctx& c = h->c;
    if (!c.batch.p || frame >= c.batch_frames) return fail(MB_ERR_ARG, "mb_batch_select: frame out of range");
    c.d_xyz = c.batch.as<float>() + frame * c.batch_atoms * 3;
    c.n_atoms = c.batch_atoms;
    return MB_OK;
}

int mb_batch_search(MbCtx* h, float cutoff, uint8_t pbc_dims, size_t f0, size_t f1, int mode, int64_t* counts,
                    uint64_t* checksums2) {
    if (!h) return fail(MB_ERR_ARG, "null context");
    return batch_search_impl(&h->c, cutoff, pbc_dims & 7, f0, f1, mode, counts, checksums2);
}

int mb_stream_search(MbCtx* h, float cutoff, uint8_t pbc_dims, const float* frames, size_t n_frames, size_t n_atoms,
                     const float* box9, int mode, int64_t* counts) {
    if (!h || !frames || n_frames == 0 || n_atoms == 0) return fail(MB_ERR_ARG, "mb_stream_search: bad argument");
    Ctx& c = h->c;
    MB_TRY(stream_box(c, box9));
    return stream_chunks(c, frames, n_frames, n_atoms, n_atoms >= 500000 ? 8 : 32, [&](size_t b0, size_t nf, size_t f0) {
        return batch_search_impl(&c, cutoff, pbc_dims & 7, b0, b0 + nf, mode, counts ? counts + f0 : nullptr, nullptr);
    });
}

// Kabsch fit of every host frame onto the FIRST one + unweighted RMSD after the fit (config 4); the superposed
// coordinates stay on the device (last chunk) — only the RMSD values travel back.
int mb_stream_fit(MbCtx* h, const float* frames, size_t n_frames, size_t n_atoms, double* rmsd_out) {
    if (!h || !frames || n_frames == 0 || n_atoms == 0) return fail(MB_ERR_ARG, "mb_stream_fit: bad argument");
    Ctx& c = h->c;
    const size_t chunk = std::max<size_t>(8, std::min<size_t>(128, ((size_t)48 << 20) / (n_atoms * 12)));
    return stream_chunks(c, frames, n_frames, n_atoms, chunk, [&](size_t b0, size_t nf, size_t f0) {
        // the first chunk stages frame 0 as the reference; later chunks keep it
        return batch_fit_impl(&c, f0 == 0 ? b0 : (size_t)-1, b0, b0 + nf, 1, rmsd_out ? rmsd_out + f0 : nullptr);
    });
}

// Per-frame COM + gyration + contact count over host frames (config 5): out is n_frames x 5 doubles.
int mb_stream_pipeline(MbCtx* h, float cutoff, uint8_t pbc_dims, const float* frames, size_t n_frames, size_t n_atoms,
                       const float* box9, double* out) {
    if (!h || !frames || !out || n_frames == 0 || n_atoms == 0) return fail(MB_ERR_ARG, "mb_stream_pipeline: bad argument");
    Ctx& c = h->c;
    MB_TRY(stream_box(c, box9));
    return stream_chunks(c, frames, n_frames, n_atoms, n_atoms >= 500000 ? 4 : 32, [&](size_t b0, size_t nf, size_t f0) {
        return mb_batch_pipeline(h, cutoff, pbc_dims, b0, b0 + nf, out + 5 * f0);
    });
}

int mb_batch_fit(MbCtx* h, size_t ref_frame, size_t f0, size_t f1, int superpose, double* rmsd_out) {
    if (!h) return fail(MB_ERR_ARG, "null context");
    return batch_fit_impl(&h->c, ref_frame, f0, f1, superpose, rmsd_out);
}

int mb_batch_pipeline(MbCtx* h, float cutoff, uint8_t pbc_dims, size_t f0, size_t f1, double* out) {
    if (!h) return fail(MB_ERR_ARG, "null context");
    Ctx& c = h->c;
    if (!c.batch.p || f1 > c.batch_frames || f0 >= f1) return fail(MB_ERR_ARG, "batch_pipeline: bad frame range");
    if (!c.masses.p || c.n_masses < c.batch_atoms) return fail(MB_ERR_STATE, "masses not set for the batch");
    MB_CUDA(cudaSetDevice(c.device));
    const size_t n = c.batch_atoms, nf = f1 - f0;
    // moments scratch lives in pipe_tmp: [tickets 64 KB][rows8 nf*8][partials nf*nb*5];
    // the contact counts come from batch_search (count-only mode), whose per-frame device counters
    // stay in batch_tmp at stride 4
    int nb = (int)std::max<size_t>(1, std::min<size_t>((n + 256 * 8 - 1) / (256 * 8), 64));
    size_t tick_bytes = 64 * 1024;
    if (nf * sizeof(unsigned) > tick_bytes) return fail(MB_ERR_ARG, "batch_pipeline: at most 16384 frames per call");
    size_t rows8_bytes = nf * 8 * sizeof(double);
    size_t part_bytes = nf * (size_t)nb * 5 * sizeof(double);
    size_t need = tick_bytes + rows8_bytes + part_bytes;
    bool fresh = need > c.pipe_tmp.cap;
    MB_TRY(c.pipe_tmp.reserve(need));
    if (fresh) MB_CUDA(cudaMemsetAsync(c.pipe_tmp.p, 0, tick_bytes, c.stream));  // tickets re-arm themselves
    char* base = static_cast<char*>(c.pipe_tmp.p);
    unsigned* tickets = reinterpret_cast<unsigned*>(base);
    double* rows8 = reinterpret_cast<double*>(base + tick_bytes);
    double* partials = reinterpret_cast<double*>(base + tick_bytes + rows8_bytes);
    MB_TRY(c.batch_scalars.reserve(nf * 5 * sizeof(double)));
    MB_TRY(enqueue_batch_moments(&c, f0, f1, rows8, partials, tickets, nb));
    // count-only neighbour search of every frame, frames alternating over the stream slots
    MB_TRY(batch_search_impl(&c, cutoff, pbc_dims & 7, f0, f1, 1, nullptr, nullptr));
    const unsigned long long* counters = c.batch_tmp.as<unsigned long long>();
    assemble_rows_kernel<<<(unsigned)((nf + 127) / 128), 128, 0, c.stream>>>(rows8, counters, (int)nf,
                                                                            c.batch_scalars.as<double>());
    c.launches++;
    MB_CUDA(cudaGetLastError());
    c.batch_rows = nf;
    c.batch_row_doubles = 5;
    if (out) MB_CUDA(cudaMemcpyAsync(out, c.batch_scalars.p, nf * 5 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    MB_CUDA(cudaStreamSynchronize(c.stream));
    c.harvest_profile();
    return MB_OK;
}

const void* mb_batch_scalars_device(MbCtx* h, size_t* n_rows, size_t* row_doubles) {
    if (!h) return nullptr;
    if (n_rows) *n_rows = h->c.batch_rows;
    if (row_doubles) *row_doubles = h->c.batch_row_doubles;
    return h->c.batch_scalars.p;
}

}  // extern "C"
