// facenet_b200 -- NCCL behind the C ABI: one process per GPU, the communicator lives in the handle.
//
// The reference shards nothing (pure NumPy); north_star's multi-GPU form is "each GPU takes row blocks of the pair matrix, the
// small embedding matrix is all-gathered over NVLink with NCCL, and the per-threshold count histograms are all-reduced"
// (BASELINE.json).  A host without torch.distributed binds these entry points directly: rank 0 makes the id
// (fnb_comm_unique_id), ships the 128 bytes to the other processes by whatever means it has, every process calls
// fnb_comm_init, and fnb_pair_histogram_sharded (fnb_api.cu) does the exchange and the reduction itself.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2): inside a torch process that is the copy torch has already loaded,
// so both share one NCCL; elsewhere the system library, or the path in FNB_NCCL_LIB.  Only the types come from <nccl.h>.
#include "fnb_host.h"

#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>

namespace fnb {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitRankConfig)(ncclComm_t*, int, ncclUniqueId, int, ncclConfig_t*) = nullptr;      // optional
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    std::string error;
};

static NcclApi g_nccl;

static const NcclApi* nccl_api() {
    if (g_nccl.lib) return &g_nccl;
    const char* env = getenv("FNB_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        if (!nm || !*nm) continue;
        // RTLD_NOLOAD first: the copy a framework in this process has loaded already
        void* l = dlopen(nm, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        if (!l) l = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (l) { g_nccl.lib = l; break; }
    }
    if (!g_nccl.lib) { g_nccl.error = std::string("libnccl.so.2 not found (set FNB_NCCL_LIB): ") + (dlerror() ? dlerror() : ""); return nullptr; }
    bool ok = true;
    auto sym = [&](const char* n) { void* p = dlsym(g_nccl.lib, n); if (!p) { ok = false; g_nccl.error = std::string("NCCL symbol missing: ") + n; } return p; };
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))sym("ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))sym("ncclCommInitRank");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))sym("ncclCommDestroy");
    g_nccl.AllGather = (decltype(g_nccl.AllGather))sym("ncclAllGather");
    g_nccl.Broadcast = (decltype(g_nccl.Broadcast))sym("ncclBroadcast");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))sym("ncclAllReduce");
    g_nccl.GroupStart = (decltype(g_nccl.GroupStart))sym("ncclGroupStart");
    g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))sym("ncclGroupEnd");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))sym("ncclGetErrorString");
    g_nccl.GetVersion = (decltype(g_nccl.GetVersion))sym("ncclGetVersion");
    if (!ok) { dlclose(g_nccl.lib); g_nccl.lib = nullptr; return nullptr; }
    g_nccl.CommInitRankConfig = (decltype(g_nccl.CommInitRankConfig))dlsym(g_nccl.lib, "ncclCommInitRankConfig");
    return &g_nccl;
}

constexpr int kSharedCounters = 256;
constexpr int kMaxStealWorld = 8;

struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    // Tile-queue counters of every rank, mapped into every rank through CUDA IPC (peer access over NVLink): ctr[r] is rank r's
    // array (ctr[rank] is this rank's own allocation).  A rank drains its OWN queue with local atomics and, when that is empty,
    // takes tiles from the other ranks' queues (GramParams::steal_counter) -- a GPU that runs slower under the power limit is
    // helped out by the others instead of holding up the all-reduce.
    unsigned long long* ctr[kMaxStealWorld] = {};
    bool shared = false;
};

#define NCK(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) \
    return h->fail(FNB_ERR_CUDA, "%s: %s", #call, g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "NCCL error"); } while (0)

int comm_world(const fnb_context* h) { return h->comm ? h->comm->world : 1; }
int comm_rank(const fnb_context* h) { return h->comm ? h->comm->rank : 0; }

int comm_all_gather(fnb_context* h, const void* send, void* recv, size_t bytes_per_rank, cudaStream_t s) {
    if (!h->comm) return h->fail(FNB_ERR_INVALID, "no communicator (fnb_comm_init)");
    NCK(g_nccl.AllGather(send, recv, bytes_per_rank, ncclUint8, h->comm->comm, s));
    return FNB_OK;
}

int comm_broadcast(fnb_context* h, const void* send, void* recv, size_t bytes, int root, cudaStream_t s) {
    if (!h->comm) return h->fail(FNB_ERR_INVALID, "no communicator (fnb_comm_init)");
    NCK(g_nccl.Broadcast(send, recv, bytes, ncclUint8, root, h->comm->comm, s));
    return FNB_OK;
}

int comm_group_start(fnb_context* h) { NCK(g_nccl.GroupStart()); return FNB_OK; }
int comm_group_end(fnb_context* h) { NCK(g_nccl.GroupEnd()); return FNB_OK; }

int comm_all_reduce_u64(fnb_context* h, void* buf, size_t count, bool max_op, cudaStream_t s) {
    if (!h->comm) return h->fail(FNB_ERR_INVALID, "no communicator (fnb_comm_init)");
    NCK(g_nccl.AllReduce(buf, buf, count, ncclUint64, max_op ? ncclMax : ncclSum, h->comm->comm, s));
    return FNB_OK;
}

// rank r's counters as mapped into this process (NULL: the ranks do not share their queues)
unsigned long long* comm_shared_counters(const fnb_context* h, int r, int* count) {
    if (count) *count = kSharedCounters;
    return (h->comm && h->comm->shared && r >= 0 && r < h->comm->world) ? h->comm->ctr[r] : nullptr;
}

static void drop_shared_counters(Comm* c) {
    for (int r = 0; r < kMaxStealWorld; ++r) {
        if (!c->ctr[r]) continue;
        if (r == c->rank) cudaFree(c->ctr[r]); else cudaIpcCloseMemHandle(c->ctr[r]);
        c->ctr[r] = nullptr;
    }
    c->shared = false;
    cudaGetLastError();
}

// every rank allocates its counters, the IPC handles travel by ncclAllGather, every rank maps the others'; the ranks then agree
// (all-reduce of a flag) on whether everybody succeeded -- if not, nobody shares
static void setup_shared_counters(fnb_context* h) {
    Comm* c = h->comm;
    if (c->world < 2 || c->world > kMaxStealWorld || getenv("FNB_NO_SHARED_QUEUE")) return;
    const size_t hb = sizeof(cudaIpcMemHandle_t);
    unsigned long long ok = 1;
    unsigned char* tmp = nullptr;
    if (cudaMalloc(&tmp, hb * (c->world + 1) + 8) != cudaSuccess) { cudaGetLastError(); return; }
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    if (cudaMalloc(&c->ctr[c->rank], kSharedCounters * 8) != cudaSuccess || cudaMemset(c->ctr[c->rank], 0, kSharedCounters * 8) != cudaSuccess ||
        cudaIpcGetMemHandle(&mine, c->ctr[c->rank]) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    std::vector<cudaIpcMemHandle_t> all((size_t)c->world);
    cudaMemcpyAsync(tmp + hb * c->world, &mine, hb, cudaMemcpyHostToDevice, h->stream);
    const bool nccl_ok = g_nccl.AllGather(tmp + hb * c->world, tmp, hb, ncclUint8, c->comm, h->stream) == ncclSuccess;
    cudaMemcpyAsync(all.data(), tmp, hb * c->world, cudaMemcpyDeviceToHost, h->stream);
    cudaStreamSynchronize(h->stream);
    unsigned long long* flag = (unsigned long long*)(tmp + hb * (c->world + 1));
    if (!nccl_ok) ok = 0;
    // every rank must have allocated before anybody maps: first round of agreement
    cudaMemcpyAsync(flag, &ok, 8, cudaMemcpyHostToDevice, h->stream);
    if (g_nccl.AllReduce(flag, flag, 1, ncclUint64, ncclMin, c->comm, h->stream) != ncclSuccess) ok = 0;
    unsigned long long all_ok = 0;
    cudaMemcpyAsync(&all_ok, flag, 8, cudaMemcpyDeviceToHost, h->stream);
    cudaStreamSynchronize(h->stream);
    if (ok && all_ok) {
        for (int r = 0; r < c->world; ++r) {
            if (r == c->rank) continue;
            void* p = nullptr;
            if (cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
            c->ctr[r] = (unsigned long long*)p;
        }
    } else {
        ok = 0;
    }
    cudaMemcpyAsync(flag, &ok, 8, cudaMemcpyHostToDevice, h->stream);
    if (g_nccl.AllReduce(flag, flag, 1, ncclUint64, ncclMin, c->comm, h->stream) != ncclSuccess) ok = 0;
    cudaMemcpyAsync(&all_ok, flag, 8, cudaMemcpyDeviceToHost, h->stream);
    cudaStreamSynchronize(h->stream);
    cudaFree(tmp);
    c->shared = ok && all_ok;
    if (!c->shared) drop_shared_counters(c);
}

void comm_release(fnb_context* h) {
    if (h->comm) drop_shared_counters(h->comm);
    if (h->comm) {
        if (h->comm->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm->comm);
        delete h->comm;
        h->comm = nullptr;
    }
}

}  // namespace fnb

using namespace fnb;

static thread_local std::string g_comm_error;

extern "C" const char* fnb_comm_last_error(void) { return g_comm_error.c_str(); }

extern "C" int fnb_comm_unique_id(void* id128) {
    if (!id128) { g_comm_error = "fnb_comm_unique_id: NULL"; return FNB_ERR_INVALID; }
    const NcclApi* api = nccl_api();
    if (!api) { g_comm_error = g_nccl.error; return FNB_ERR_UNSUPPORTED; }
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    ncclResult_t r = api->GetUniqueId(&id);
    if (r != ncclSuccess) { g_comm_error = std::string("ncclGetUniqueId: ") + api->GetErrorString(r); return FNB_ERR_CUDA; }
    memcpy(id128, &id, sizeof(id));
    return FNB_OK;
}

extern "C" int fnb_comm_init(fnb_handle h, const void* id128, int rank, int world) {
    if (!h) return FNB_ERR_INVALID;
    if (!id128 || world < 1 || rank < 0 || rank >= world) return h->fail(FNB_ERR_INVALID, "fnb_comm_init: bad id / rank %d / world %d", rank, world);
    const NcclApi* api = nccl_api();
    if (!api) return h->fail(FNB_ERR_UNSUPPORTED, "%s", g_nccl.error.c_str());
    cudaError_t e = cudaSetDevice(h->device);
    if (e != cudaSuccess) return h->fail(FNB_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    comm_release(h);
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    Comm* c = new Comm();
    c->rank = rank; c->world = world;
    // The row exchange runs UNDER the persistent Gram kernels, which hold one CTA per SM on all but kAuxReserveSms of the SMs
    // (132 in two-pair clusters + plain pairs on the rest): NCCL's CTAs have those SMs.  More CTAs than that would sit in the
    // launch queue until a Gram launch ends -- with the ranks' collectives waiting on each other meanwhile -- so the
    // communicator is capped (FNB_NCCL_MAX_CTAS).
    ncclResult_t r;
    if (api->CommInitRankConfig) {
        ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
        const char* env = getenv("FNB_NCCL_MAX_CTAS");
        cfg.maxCTAs = env ? atoi(env) : kAuxReserveSms;
        if (cfg.maxCTAs <= 0) cfg.maxCTAs = NCCL_CONFIG_UNDEF_INT;
        // ... and those 16 SMs are the leftovers of the GPCs (two per GPC: the Gram clusters take four SMs each), so NCCL's own
        // default of 4-CTA clusters (CGA) could not be placed on them at all: measured at 2 GPUs, every broadcast then finished
        // right AFTER the Gram launch it was meant to run under.  Cluster size 1 lets NCCL's CTAs take any free SM.
        const char* cga = getenv("FNB_NCCL_CGA");
        cfg.cgaClusterSize = cga ? atoi(cga) : 1;
        r = api->CommInitRankConfig(&c->comm, world, id, rank, &cfg);
    } else {
        r = api->CommInitRank(&c->comm, world, id, rank);
    }
    if (r != ncclSuccess) { delete c; return h->fail(FNB_ERR_CUDA, "ncclCommInitRank: %s", api->GetErrorString(r)); }
    h->comm = c;
    setup_shared_counters(h);
    return FNB_OK;
}

extern "C" int fnb_comm_destroy(fnb_handle h) {
    if (!h) return FNB_ERR_INVALID;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
    comm_release(h);
    return FNB_OK;
}

extern "C" int fnb_comm_shared_queue(fnb_handle h) { return (h && h->comm && h->comm->shared) ? 1 : 0; }

extern "C" int fnb_comm_info(fnb_handle h, int* rank, int* world, int* nccl_version) {
    if (!h) return FNB_ERR_INVALID;
    if (rank) *rank = comm_rank(h);
    if (world) *world = comm_world(h);
    if (nccl_version) { *nccl_version = 0; if (g_nccl.lib && g_nccl.GetVersion) g_nccl.GetVersion(nccl_version); }
    return FNB_OK;
}
