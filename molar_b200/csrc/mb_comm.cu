// mb_comm.cu — the one collective of the path: gathering the per-frame scalars of frame-sharded ranks.
//
// Frames are independent, so ranks (one GPU each) never exchange coordinates or pair lists; what the per-frame loop
// of the reference produces per frame — AnalysisTask::process_frame results, analysis_task.rs:113-280 — are a few
// scalars, and those are all-gathered over NCCL (NVLink / NVSwitch) at the end of a pass.  libnccl.so.2 is opened
// with dlopen at mb_comm_init time, so a single-GPU host without NCCL can still load libmolar_b200.so.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

#include "mb_common.cuh"

namespace mb {

struct NcclApi {
    void* lib = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
};

static NcclApi* nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        const char* names[] = {getenv("MOLAR_B200_NCCL"), "libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            if (!nm || !*nm) continue;
            api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (api.lib) {
#define MB_SYM(field, name) api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, name))
            MB_SYM(GetUniqueId, "ncclGetUniqueId");
            MB_SYM(CommInitRank, "ncclCommInitRank");
            MB_SYM(CommInitAll, "ncclCommInitAll");
            MB_SYM(CommDestroy, "ncclCommDestroy");
            MB_SYM(AllGather, "ncclAllGather");
            MB_SYM(AllReduce, "ncclAllReduce");
            MB_SYM(GetErrorString, "ncclGetErrorString");
            MB_SYM(GetVersion, "ncclGetVersion");
#undef MB_SYM
            if (!api.GetUniqueId || !api.CommInitRank || !api.CommInitAll || !api.CommDestroy || !api.AllGather ||
                !api.AllReduce || !api.GetErrorString) {
                dlclose(api.lib);
                api.lib = nullptr;
            }
        }
    }
    return api.lib ? &api : nullptr;
}

#define MB_NCCL(api, call)                                                                                   \
    do {                                                                                                     \
        ncclResult_t _r = (call);                                                                            \
        if (_r != ncclSuccess)                                                                               \
            return ::mb::fail(MB_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, (api)->GetErrorString(_r)); \
    } while (0)

void comm_destroy(Ctx* c) {
    if (c->nccl_comm) {
        NcclApi* a = nccl_api();
        if (a) a->CommDestroy(static_cast<ncclComm_t>(c->nccl_comm));
        c->nccl_comm = nullptr;
    }
    c->comm_rank = 0;
    c->comm_world = 1;
}

}  // namespace mb

using namespace mb;

extern "C" {

int mb_comm_unique_id(unsigned char* id128) {
    if (!id128) return fail(MB_ERR_ARG, "null argument");
    static_assert(sizeof(ncclUniqueId) == MB_COMM_ID_BYTES, "NCCL unique id size");
    NcclApi* a = nccl_api();
    if (!a) return fail(MB_ERR_STATE, "libnccl.so.2 not found (%s)", dlerror() ? dlerror() : "dlopen failed");
    ncclUniqueId id;
    MB_NCCL(a, a->GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return MB_OK;
}

int mb_comm_init(MbCtx* h, int rank, int world, const unsigned char* id128) {
    if (!h || !id128) return fail(MB_ERR_ARG, "null argument");
    if (world < 1 || rank < 0 || rank >= world) return fail(MB_ERR_ARG, "mb_comm_init: rank %d of %d", rank, world);
    Ctx& c = h->c;
    NcclApi* a = nccl_api();
    if (!a) return fail(MB_ERR_STATE, "libnccl.so.2 not found");
    MB_CUDA(cudaSetDevice(c.device));
    comm_destroy(&c);
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t comm = nullptr;
    MB_NCCL(a, a->CommInitRank(&comm, world, id, rank));
    c.nccl_comm = comm;
    c.comm_rank = rank;
    c.comm_world = world;
    return MB_OK;
}

int mb_comm_init_all(int n, MbCtx* const* ctxs) {
    if (n < 1 || !ctxs) return fail(MB_ERR_ARG, "mb_comm_init_all: bad arguments");
    NcclApi* a = nccl_api();
    if (!a) return fail(MB_ERR_STATE, "libnccl.so.2 not found");
    std::vector<int> devs(n);
    for (int i = 0; i < n; ++i) {
        if (!ctxs[i]) return fail(MB_ERR_ARG, "mb_comm_init_all: null context %d", i);
        devs[i] = ctxs[i]->c.device;
        comm_destroy(&ctxs[i]->c);
    }
    std::vector<ncclComm_t> comms(n, nullptr);
    MB_NCCL(a, a->CommInitAll(comms.data(), n, devs.data()));
    for (int i = 0; i < n; ++i) {
        ctxs[i]->c.nccl_comm = comms[i];
        ctxs[i]->c.comm_rank = i;
        ctxs[i]->c.comm_world = n;
    }
    return MB_OK;
}

void mb_comm_destroy(MbCtx* h) {
    if (h) comm_destroy(&h->c);
}

int mb_comm_info(MbCtx* h, int* rank, int* world, int* nccl_version) {
    if (!h) return fail(MB_ERR_ARG, "null context");
    if (rank) *rank = h->c.comm_rank;
    if (world) *world = h->c.nccl_comm ? h->c.comm_world : 1;
    if (nccl_version) {
        *nccl_version = 0;
        NcclApi* a = nccl_api();
        if (a && a->GetVersion) a->GetVersion(nccl_version);
    }
    return MB_OK;
}

// rows == NULL: contribute the rows the last mb_batch_pipeline / mb_batch_fit left on the device
int mb_gather_scalars(MbCtx* h, const double* rows, size_t n_rows, size_t n_cols, double* out_all) {
    if (!h || !out_all) return fail(MB_ERR_ARG, "null argument");
    Ctx& c = h->c;
    const size_t cnt = n_rows * n_cols;
    if (cnt == 0) return fail(MB_ERR_ARG, "mb_gather_scalars: empty rows");
    MB_CUDA(cudaSetDevice(c.device));
    const int world = c.nccl_comm ? c.comm_world : 1;
    const double* d_send = nullptr;
    if (rows) {
        MB_TRY(c.comm_send.reserve(cnt * sizeof(double)));
        MB_CUDA(cudaMemcpyAsync(c.comm_send.p, rows, cnt * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        d_send = c.comm_send.as<double>();
    } else {
        if (!c.batch_scalars.p || c.batch_rows * c.batch_row_doubles < cnt)
            return fail(MB_ERR_STATE, "mb_gather_scalars: no device rows of that size on this context");
        d_send = c.batch_scalars.as<double>();
    }
    if (world == 1) {
        MB_CUDA(cudaMemcpyAsync(out_all, d_send, cnt * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        MB_CUDA(cudaStreamSynchronize(c.stream));
        return MB_OK;
    }
    NcclApi* a = nccl_api();
    MB_TRY(c.comm_recv.reserve(cnt * world * sizeof(double)));
    MB_NCCL(a, a->AllGather(d_send, c.comm_recv.p, cnt, ncclDouble, static_cast<ncclComm_t>(c.nccl_comm), c.stream));
    MB_CUDA(cudaMemcpyAsync(out_all, c.comm_recv.p, cnt * world * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    MB_CUDA(cudaStreamSynchronize(c.stream));
    return MB_OK;
}

int mb_comm_max(MbCtx* h, double* inout, size_t n) {
    if (!h || !inout || n == 0) return fail(MB_ERR_ARG, "null argument");
    Ctx& c = h->c;
    if (!c.nccl_comm || c.comm_world == 1) return MB_OK;
    MB_CUDA(cudaSetDevice(c.device));
    NcclApi* a = nccl_api();
    MB_TRY(c.comm_send.reserve(n * sizeof(double)));
    MB_CUDA(cudaMemcpyAsync(c.comm_send.p, inout, n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    MB_NCCL(a, a->AllReduce(c.comm_send.p, c.comm_send.p, n, ncclDouble, ncclMax, static_cast<ncclComm_t>(c.nccl_comm),
                            c.stream));
    MB_CUDA(cudaMemcpyAsync(inout, c.comm_send.p, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    MB_CUDA(cudaStreamSynchronize(c.stream));
    return MB_OK;
}

int mb_comm_barrier(MbCtx* h) {
    double one = 1.0;
    return mb_comm_max(h, &one, 1);
}

// ---- device-side timing for a host without its own CUDA binding (bench drivers, the Rust crate) ----------------
// Events live on the context's stream; slots 0..7.
int mb_timer_record(MbCtx* h, int slot) {
    if (!h || slot < 0 || slot >= 8) return fail(MB_ERR_ARG, "mb_timer_record: bad slot");
    Ctx& c = h->c;
    MB_CUDA(cudaSetDevice(c.device));
    if (!c.timer_ev[slot]) MB_CUDA(cudaEventCreate(&c.timer_ev[slot]));
    MB_CUDA(cudaEventRecord(c.timer_ev[slot], c.stream));
    return MB_OK;
}

int mb_timer_elapsed_ms(MbCtx* h, int slot_begin, int slot_end, double* ms) {
    if (!h || !ms || slot_begin < 0 || slot_begin >= 8 || slot_end < 0 || slot_end >= 8)
        return fail(MB_ERR_ARG, "mb_timer_elapsed_ms: bad arguments");
    Ctx& c = h->c;
    if (!c.timer_ev[slot_begin] || !c.timer_ev[slot_end]) return fail(MB_ERR_STATE, "timer slot never recorded");
    MB_CUDA(cudaSetDevice(c.device));
    MB_CUDA(cudaEventSynchronize(c.timer_ev[slot_end]));
    float f = 0.f;
    MB_CUDA(cudaEventElapsedTime(&f, c.timer_ev[slot_begin], c.timer_ev[slot_end]));
    *ms = (double)f;
    return MB_OK;
}

// pinned host memory for a host without a CUDA binding (frames that mb_stream_* upload should be pinned)
void* mb_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) {
        set_error("cudaMallocHost(%zu) failed", bytes);
        return nullptr;
    }
    return p;
}
void mb_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

}  // extern "C"
