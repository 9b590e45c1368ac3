// facenet_b200 -- host -> device staging of kDLCPU tensors (NumPy arrays: pageable memory).
//
// The reference hands over what np.concatenate returned (facenet/facenet.py:184-201, callbacks.py:21-28): pageable host
// memory.  A cudaMemcpyAsync from pageable memory makes the driver stage the bytes itself, single-threaded and
// synchronously; here the copy is pipelined instead:
//   * a ring of pinned slots (kRingSlots x kSlotBytes);
//   * a small pool of host threads copies the caller's bytes into a slot (parallel memcpy: one core does ~10 GB/s,
//     PCIe Gen5 x16 wants ~50);
//   * each filled slot goes to the device with cudaMemcpyAsync on a dedicated copy stream while the threads fill the next;
//   * the handle's stream waits for the last chunk with an event -- work queued on it BEFORE the copy (label sort) overlaps it.
// Memory that is already pinned (cudaHostAlloc / cudaHostRegister, e.g. torch pin_memory) is copied in place, in one call.
#include "fnb_host.h"

#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>

namespace fnb {

constexpr size_t kStageSlotBytes = (size_t)16 << 20;
constexpr int kRingSlots = 4;

struct HostCopier {
    // a contiguous range, or (perm != NULL) a gather of `rows` rows of row_bytes each: dst row i <- src + perm[i] * row_bytes
    struct Part { char* dst; const char* src; size_t bytes; const long long* perm = nullptr; size_t row_bytes = 0; size_t rows = 0; };
    static void run_part(const Part& p) {
        if (!p.perm) { memcpy(p.dst, p.src, p.bytes); return; }
        for (size_t i = 0; i < p.rows; ++i) memcpy(p.dst + i * p.row_bytes, p.src + (size_t)p.perm[i] * p.row_bytes, p.row_bytes);
    }
    std::vector<std::thread> workers;
    std::mutex m;
    std::condition_variable cv_work, cv_done;
    std::vector<Part> parts;
    size_t next = 0;
    int busy = 0;
    bool stop = false;

    explicit HostCopier(int nthreads) {
        for (int i = 0; i < nthreads; ++i) workers.emplace_back([this] { run(); });
    }
    ~HostCopier() {
        { std::lock_guard<std::mutex> g(m); stop = true; }
        cv_work.notify_all();
        for (auto& t : workers) t.join();
    }
    void run() {
        std::unique_lock<std::mutex> g(m);
        for (;;) {
            cv_work.wait(g, [this] { return stop || next < parts.size(); });
            if (stop) return;
            const Part p = parts[next++];
            ++busy;
            g.unlock();
            run_part(p);
            g.lock();
            if (--busy == 0 && next >= parts.size()) cv_done.notify_all();
        }
    }
    // dst <- src with every thread of the pool and the caller
    void copy(void* dst, const void* src, size_t bytes) {
        const size_t nparts = workers.size() + 1;
        const size_t chunk = ((bytes / nparts + 4095) / 4096) * 4096;
        if (workers.empty() || bytes < ((size_t)1 << 20)) { memcpy(dst, src, bytes); return; }
        {
            std::lock_guard<std::mutex> g(m);
            parts.clear(); next = 0;
            for (size_t off = 0; off < bytes; off += chunk)
                parts.push_back(Part{(char*)dst + off, (const char*)src + off, std::min(chunk, bytes - off)});
        }
        drain();
    }
    // dst row i <- src_base + perm[i] * row_bytes, i < rows (the class-order gather of a streamed upload)
    void gather(void* dst, const void* src_base, const long long* perm, size_t rows, size_t row_bytes) {
        const size_t nparts = 4 * (workers.size() + 1);
        const size_t per = std::max<size_t>(1, (rows + nparts - 1) / nparts);
        {
            std::lock_guard<std::mutex> g(m);
            parts.clear(); next = 0;
            for (size_t r = 0; r < rows; r += per) {
                Part p{(char*)dst + r * row_bytes, (const char*)src_base, 0};
                p.perm = perm + r; p.row_bytes = row_bytes; p.rows = std::min(per, rows - r);
                parts.push_back(p);
            }
        }
        drain();
    }
    // the caller works too, then waits for the pool
    void drain() {
        cv_work.notify_all();
        for (;;) {
            Part p;
            {
                std::lock_guard<std::mutex> g(m);
                if (next >= parts.size()) break;
                p = parts[next++];
                ++busy;
            }
            run_part(p);
            std::lock_guard<std::mutex> g(m);
            --busy;
        }
        std::unique_lock<std::mutex> g(m);
        cv_done.wait(g, [this] { return busy == 0 && next >= parts.size(); });
    }
};

void destroy_copier(HostCopier* c) { delete c; }

#define CKS(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return h->fail(FNB_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } while (0)

static int ensure_copy_stream(fnb_context* h) {
    if (h->copy_stream) return FNB_OK;
    CKS(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    return FNB_OK;
}

static int ensure_staging(fnb_context* h) {
    int rc = ensure_copy_stream(h);
    if (rc) return rc;
    if (h->copier) return FNB_OK;
    CKS(h->ring.ensure(kRingSlots * kStageSlotBytes));
    for (int i = 0; i < kRingSlots; ++i) CKS(cudaEventCreateWithFlags(&h->ring_ev[i], cudaEventDisableTiming));
    unsigned hw = std::thread::hardware_concurrency();
    int n = (int)std::min(7u, hw > 2 ? hw / 2 - 1 : 0u);          // up to 7 workers + the caller; fewer on small hosts
    // a sharded job runs one process per GPU on the same host: the ranks share its cores
    const int world = comm_world(h);
    if (world > 1) n = std::max(1, std::min(n, (int)(hw / (2 * (unsigned)world))));
    const char* env = getenv("FNB_COPY_THREADS");
    if (env) n = std::max(0, atoi(env) - 1);
    h->copier = new HostCopier(n);
    return FNB_OK;
}

// dst (device) <- src (host, `bytes`), ordered after everything already queued on the handle's stream that touches dst and
// before everything queued on it afterwards.  Returns with the source fully consumed (the caller may reuse it).
int stage_to_device(fnb_context* h, void* dst, const void* src, size_t bytes) {
    if (bytes == 0) return FNB_OK;
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, src);
    const bool pinned = (e == cudaSuccess) && (attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged);
    if (e != cudaSuccess) cudaGetLastError();
    if (pinned || bytes < ((size_t)256 << 10)) {
        // pinned: DMA straight from the caller's buffer; tiny: one synchronous driver-side staging is cheaper than the ring
        const bool timed = pinned && bytes >= ((size_t)1 << 20) && bytes > h->h2d_timed_bytes;     // the events bracket the largest copy of the call
        if (timed) CKS(cudaEventRecord(h->copy_ev[1], h->stream));
        CKS(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream));
        if (timed) { CKS(cudaEventRecord(h->copy_ev[2], h->stream)); h->h2d_timed = true; h->h2d_timed_bytes = bytes; }
        h->last_h2d_bytes += bytes;
        return FNB_OK;
    }
    int rc = ensure_staging(h);
    if (rc) return rc;
    const bool timed = bytes > h->h2d_timed_bytes;
    // the copy stream starts after the work queued so far (an earlier launch may still read dst)
    CKS(cudaEventRecord(h->copy_ev[0], h->stream));
    CKS(cudaStreamWaitEvent(h->copy_stream, h->copy_ev[0], 0));
    if (timed) CKS(cudaEventRecord(h->copy_ev[1], h->copy_stream));
    char* ring = h->ring.as<char>();
    int slot = 0;
    for (size_t off = 0; off < bytes; off += kStageSlotBytes, slot = (slot + 1) % kRingSlots) {
        const size_t len = std::min(kStageSlotBytes, bytes - off);
        CKS(cudaEventSynchronize(h->ring_ev[slot]));                 // the DMA that last used this slot has finished
        h->copier->copy(ring + (size_t)slot * kStageSlotBytes, (const char*)src + off, len);
        CKS(cudaMemcpyAsync((char*)dst + off, ring + (size_t)slot * kStageSlotBytes, len, cudaMemcpyHostToDevice, h->copy_stream));
        CKS(cudaEventRecord(h->ring_ev[slot], h->copy_stream));
    }
    CKS(cudaEventRecord(timed ? h->copy_ev[2] : h->copy_ev[0], h->copy_stream));
    CKS(cudaStreamWaitEvent(h->stream, timed ? h->copy_ev[2] : h->copy_ev[0], 0));
    h->last_h2d_bytes += bytes;
    if (timed) { h->h2d_timed = true; h->h2d_timed_bytes = bytes; }
    return FNB_OK;
}

// One chunk of a streamed upload (fnb_pair_histogram_bins over host rows): the bytes travel on the copy stream -- from the
// caller's buffer when it is pinned, through the ring otherwise -- and the handle's stream waits for them, so the split and the
// Gram launch queued next start when the chunk has landed while the launches queued EARLIER keep the GPU busy.
int stage_chunk(fnb_context* h, void* dst, const void* src, size_t bytes, bool first, const long long* perm, size_t row_bytes) {
    if (bytes == 0) return FNB_OK;
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, src);
    // a gather (perm != NULL: dst row i <- src row perm[i]) always goes through the ring: the host threads collect the rows
    const bool pinned = !perm && (e == cudaSuccess) && (attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged);
    if (e != cudaSuccess) cudaGetLastError();
    int rc = pinned ? ensure_copy_stream(h) : ensure_staging(h);
    if (rc) return rc;
    if (first) {
        // dst may still be read by work queued earlier on the handle's stream
        CKS(cudaEventRecord(h->copy_ev[0], h->stream));
        CKS(cudaStreamWaitEvent(h->copy_stream, h->copy_ev[0], 0));
        CKS(cudaEventRecord(h->copy_ev[1], h->copy_stream));
        h->h2d_timed = true; h->h2d_timed_bytes = 0;
    }
    if (pinned) {
        CKS(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->copy_stream));
    } else {
        char* ring = h->ring.as<char>();
        int slot = h->ring_next;
        const size_t step = perm ? (kStageSlotBytes / row_bytes) * row_bytes : kStageSlotBytes;
        for (size_t off = 0; off < bytes; off += step, slot = (slot + 1) % kRingSlots) {
            const size_t len = std::min(step, bytes - off);
            CKS(cudaEventSynchronize(h->ring_ev[slot]));                 // the DMA that last used this slot has finished
            if (perm) h->copier->gather(ring + (size_t)slot * kStageSlotBytes, src, perm + off / row_bytes, len / row_bytes, row_bytes);
            else h->copier->copy(ring + (size_t)slot * kStageSlotBytes, (const char*)src + off, len);
            CKS(cudaMemcpyAsync((char*)dst + off, ring + (size_t)slot * kStageSlotBytes, len, cudaMemcpyHostToDevice, h->copy_stream));
            CKS(cudaEventRecord(h->ring_ev[slot], h->copy_stream));
        }
        h->ring_next = slot;
    }
    CKS(cudaEventRecord(h->copy_ev[2], h->copy_stream));
    CKS(cudaStreamWaitEvent(h->stream, h->copy_ev[2], 0));
    h->last_h2d_bytes += bytes;
    h->h2d_timed_bytes += bytes;
    return FNB_OK;
}

}  // namespace fnb
