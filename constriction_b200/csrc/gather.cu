// gather.cu -- the one exchange step of the multi-GPU path: every rank ends up holding every rank's compressed
// container (BASELINE.json: "a single all-gather over NVLink only to concatenate the per-shard compressed
// buffers"; SURVEY section 8b: ctr_gather_compressed).  Two implementations behind the C ABI:
//
//   ctr_gather_*  (peer memory)  Every rank owns receive buffers that are mapped into every rank's address space
//       (symmetric memory over NVLink; the caller allocates and exchanges them, e.g. with
//       torch.distributed._symmetric_memory or cuMemCreate/cuMemExportToShareableHandle).  A buffer is divided
//       into fixed SLOTS, one per source rank, so no destination address depends on another rank's size: the
//       encoder writes its container straight into its own slot, and as soon as the encode kernel has finished a
//       worker thread of this library pushes the used part of the slot into the same slot of every peer with
//       copy-engine transfers (no SM: the coder kernels keep every SM), ordered by stream memory operations on
//       flags in peer memory (cuStreamWriteValue32 / cuStreamWaitValue32).  The caller's threads never block:
//       the only host wait is the worker's wait for the rank's OWN encode (it needs the size to issue the copies),
//       there is no cross-rank host synchronisation and no size exchange.
//
//   ctr_gather_compressed_nccl   The same result with NCCL collectives on a caller-supplied communicator
//       (ncclAllGather of the sizes, one host wait, grouped ncclBroadcast of the words and offset tables).  NCCL is
//       resolved at run time (dlsym), so the library has no link-time dependency on it.
//
// Container layout (both): words  u32[n_buffers][world][slot_words], offsets u64[n_buffers][world][slot_streams + 1];
// slot (b, r) holds rank r's container of turn q (b = q mod n_buffers) exactly as its encoder wrote it (offsets are
// relative to the slot), so `ctr_*_decode(model, slot words, slot offsets, ...)` decodes rank r's streams.
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/constriction_b200.h"
#include "host_common.h"

namespace {

using StreamValue32Fn = CUresult (*)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
StreamValue32Fn driver_fn(const char *name) {
    void *f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return (StreamValue32Fn)f;
}
StreamValue32Fn write_value_fn() {
    static const StreamValue32Fn fn = driver_fn("cuStreamWriteValue32");
    return fn;
}
StreamValue32Fn wait_value_fn() {
    static const StreamValue32Fn fn = driver_fn("cuStreamWaitValue32");
    return fn;
}

constexpr int kMaxPushStreams = 8;
constexpr int kJobRing = 16;  // pinned size slots / events: more turns than can be in flight

struct Job {
    uint32_t turn;
    uint64_t n_streams;
    int slot;  // index into totals / events
};

}  // namespace

struct ctr_gather_s {
    uint32_t world = 0, rank = 0, nb = 0;
    uint64_t slot_words = 0, slot_streams = 0;
    std::vector<char *> words, offsets;  // peer-mapped bases, index = rank
    std::vector<char *> flags;
    int device = 0;
    int n_push = 1;
    bool lockstep = true;
    cudaStream_t push[kMaxPushStreams] = {};
    cudaStream_t aux = nullptr;       // size read-back: off the caller's stream, which must not wait for a copy engine
    cudaEvent_t enc_done[kJobRing] = {};
    // The flag waits / writes of ctr_gather_wait / ctr_gather_release run on two streams of their own and are tied to
    // the caller's stream by one event each: a stream memory operation costs the stream it is in several
    // microseconds between two kernels, and a step would have 2 (world - 1) of them.  (Two streams: a release must
    // never queue behind a wait that depends on a peer which in turn waits for that release.)
    cudaStream_t wait_stream = nullptr, release_stream = nullptr;
    cudaEvent_t sync_ev[kJobRing] = {};
    uint32_t sync_posted = 0;
    uint64_t *totals = nullptr;  // pinned, kJobRing entries
    cudaEvent_t encoded[kJobRing] = {};
    uint32_t seq = 0;
    uint32_t posted = 0;
    // worker
    std::thread worker;
    std::mutex mu;
    std::condition_variable cv, idle_cv;
    std::deque<Job> jobs;
    bool stop = false, busy = false;
    std::atomic<int> error{0};
    std::string error_text;

    char *words_slot(uint32_t r, uint32_t b, uint32_t src) const { return words[r] + ((uint64_t)b * world + src) * slot_words * 4; }
    char *offsets_slot(uint32_t r, uint32_t b, uint32_t src) const {
        return offsets[r] + ((uint64_t)b * world + src) * (slot_streams + 1) * 8;
    }
    // flags of rank r: u32[3][nb][world] -- [0] arrived[b][src], [1] released[b][consumer], [2] encoded[b][src]
    char *flag(uint32_t r, uint32_t kind, uint32_t b, uint32_t who) const {
        return flags[r] + (((uint64_t)kind * nb + b) * world + who) * 4;
    }
};

namespace {

void set_error(ctr_gather_s *g, int code, const std::string &text) {
    int expected = 0;
    if (g->error.compare_exchange_strong(expected, code)) g->error_text = text;
}

// the worker: waits for the rank's own encode, then issues the pushes of that turn
void run_push(ctr_gather_s *g, const Job &job) {
    const uint32_t world = g->world, me = g->rank, b = job.turn % g->nb;
    cudaError_t e = cudaEventSynchronize(g->encoded[job.slot]);
    if (e != cudaSuccess) set_error(g, CTR_ERR_CUDA, std::string("cudaEventSynchronize: ") + cudaGetErrorString(e));
    uint64_t total = g->totals[job.slot];
    if (total > g->slot_words) {  // the peers must still see the turn arrive (with an empty container)
        set_error(g, CTR_ERR_OUT_OF_SPACE, "ctr_gather_push: the container exceeds slot_words");
        total = 0;
    }
    const char *src_words = g->words_slot(me, b, me);
    const char *src_off = g->offsets_slot(me, b, me);
    if (g->lockstep && world > 2) {
        // All ranks start the pushes of a turn together and visit their destinations in the same rotation (round i:
        // rank r -> rank r + i), so that in every round the transfers form a permutation: each NVLink port carries one
        // outgoing and one incoming transfer at full rate.  Unsynchronised pushes collide (two sources into one port
        // while another port idles) and measured 1.9x slower on 8 GPUs.  The start is a flag barrier on the push
        // stream: "my turn is encoded" to every peer, then wait for every peer's.
        cudaStream_t s0 = g->push[0];
        for (uint32_t i = 1; i < world; ++i) {
            const uint32_t d = (me + i) % world;
            const CUresult r = write_value_fn()((CUstream)s0, (CUdeviceptr)(uintptr_t)g->flag(d, 2, b, me), job.turn,
                                                CU_STREAM_WRITE_VALUE_DEFAULT);
            if (r != CUDA_SUCCESS) set_error(g, CTR_ERR_CUDA, "cuStreamWriteValue32 failed (" + std::to_string((int)r) + ")");
        }
        for (uint32_t i = 1; i < world; ++i) {
            const uint32_t d = (me + i) % world;
            const CUresult r = wait_value_fn()((CUstream)s0, (CUdeviceptr)(uintptr_t)g->flag(me, 2, b, d), job.turn,
                                               CU_STREAM_WAIT_VALUE_GEQ);
            if (r != CUDA_SUCCESS) set_error(g, CTR_ERR_CUDA, "cuStreamWaitValue32 failed (" + std::to_string((int)r) + ")");
        }
    }
    for (uint32_t i = 1; i < world; ++i) {
        const uint32_t d = (me + i) % world;  // start with my right neighbour: destinations form a permutation
        cudaStream_t s = g->push[(i - 1) % g->n_push];
        // peer d has released the previous use of buffer b (written into MY flags by d's consumer stream)
        if (job.turn > g->nb) {
            const CUresult r = wait_value_fn()((CUstream)s, (CUdeviceptr)(uintptr_t)g->flag(me, 1, b, d), job.turn - g->nb,
                                               CU_STREAM_WAIT_VALUE_GEQ);
            if (r != CUDA_SUCCESS) set_error(g, CTR_ERR_CUDA, "cuStreamWaitValue32 failed (" + std::to_string((int)r) + ")");
        }
        if (total) {
            e = cudaMemcpyAsync(g->words_slot(d, b, me), src_words, total * 4, cudaMemcpyDeviceToDevice, s);
            if (e != cudaSuccess) set_error(g, CTR_ERR_CUDA, std::string("push words: ") + cudaGetErrorString(e));
        }
        e = cudaMemcpyAsync(g->offsets_slot(d, b, me), src_off, (job.n_streams + 1) * 8, cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) set_error(g, CTR_ERR_CUDA, std::string("push offsets: ") + cudaGetErrorString(e));
        const CUresult r = write_value_fn()((CUstream)s, (CUdeviceptr)(uintptr_t)g->flag(d, 0, b, me), job.turn,
                                            CU_STREAM_WRITE_VALUE_DEFAULT);
        if (r != CUDA_SUCCESS) set_error(g, CTR_ERR_CUDA, "cuStreamWriteValue32 failed (" + std::to_string((int)r) + ")");
    }
}

void worker_main(ctr_gather_s *g) {
    cudaSetDevice(g->device);
    for (;;) {
        Job job;
        {
            std::unique_lock<std::mutex> lock(g->mu);
            g->busy = false;
            g->idle_cv.notify_all();
            g->cv.wait(lock, [&] { return g->stop || !g->jobs.empty(); });
            if (g->jobs.empty()) return;  // stop requested and nothing left
            job = g->jobs.front();
            g->jobs.pop_front();
            g->busy = true;
        }
        run_push(g, job);
    }
}

}  // namespace

extern "C" int ctr_gather_create(uint32_t world, uint32_t rank, uint32_t n_buffers, uint64_t slot_words,
                                 uint64_t slot_streams, void *const *words_bases, void *const *offsets_bases,
                                 void *const *flags_bases, ctr_gather_t *out) {
    if (!out || world == 0 || rank >= world || n_buffers == 0 || !words_bases || !offsets_bases || !flags_bases)
        return CTR_ERR_BAD_ARGUMENT;
    if (slot_words % 4 != 0) return CTR_ERR_BAD_ARGUMENT;  // slots stay 16-byte aligned (decoders)
    for (uint32_t r = 0; r < world; ++r)
        if (!words_bases[r] || !offsets_bases[r] || !flags_bases[r] || (uintptr_t)words_bases[r] % 16 != 0) return CTR_ERR_BAD_ARGUMENT;
    if (ctr_device_count() == 0) return ctr::host_cuda_fail(cudaErrorNoDevice, "no CUDA device");
    if (!write_value_fn() || !wait_value_fn()) return ctr::host_fail("stream memory operations are not available");
    ctr_gather_s *g = new ctr_gather_s();
    g->world = world;
    g->rank = rank;
    g->nb = n_buffers;
    g->slot_words = slot_words;
    g->slot_streams = slot_streams;
    for (uint32_t r = 0; r < world; ++r) {
        g->words.push_back(static_cast<char *>(words_bases[r]));
        g->offsets.push_back(static_cast<char *>(offsets_bases[r]));
        g->flags.push_back(static_cast<char *>(flags_bases[r]));
    }
    if (const char *e = getenv("CTR_PUSH_STREAMS")) g->n_push = atoi(e);
    if (const char *e = getenv("CTR_PUSH_LOCKSTEP")) g->lockstep = atoi(e) != 0;
    if (g->lockstep) g->n_push = 1;  // one rotation, in order
    if (g->n_push < 1) g->n_push = 1;
    if (g->n_push > kMaxPushStreams) g->n_push = kMaxPushStreams;
    cudaError_t e = cudaGetDevice(&g->device);
    for (int i = 0; i < g->n_push && e == cudaSuccess; ++i) e = cudaStreamCreateWithFlags(&g->push[i], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&g->aux, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&g->wait_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&g->release_stream, cudaStreamNonBlocking);
    for (int i = 0; i < kJobRing && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&g->sync_ev[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaHostAlloc((void **)&g->totals, kJobRing * sizeof(uint64_t), cudaHostAllocDefault);
    for (int i = 0; i < kJobRing && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&g->encoded[i], cudaEventDisableTiming);
    for (int i = 0; i < kJobRing && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&g->enc_done[i], cudaEventDisableTiming);
    if (e != cudaSuccess) {
        const int rc = ctr::host_cuda_fail(e, "ctr_gather_create");
        ctr_gather_destroy(g);
        return rc;
    }
    g->worker = std::thread(worker_main, g);
    *out = g;
    return CTR_OK;
}

extern "C" int ctr_gather_destroy(ctr_gather_t g) {
    if (!g) return CTR_OK;
    if (g->worker.joinable()) {
        {
            std::lock_guard<std::mutex> lock(g->mu);
            g->stop = true;
        }
        g->cv.notify_all();
        g->worker.join();
    }
    for (int i = 0; i < kMaxPushStreams; ++i)
        if (g->push[i]) {
            cudaStreamSynchronize(g->push[i]);
            cudaStreamDestroy(g->push[i]);
        }
    for (cudaStream_t st : {g->aux, g->wait_stream, g->release_stream})
        if (st) {
            cudaStreamSynchronize(st);
            cudaStreamDestroy(st);
        }
    for (int i = 0; i < kJobRing; ++i)
        if (g->sync_ev[i]) cudaEventDestroy(g->sync_ev[i]);
    for (int i = 0; i < kJobRing; ++i) {
        if (g->encoded[i]) cudaEventDestroy(g->encoded[i]);
        if (g->enc_done[i]) cudaEventDestroy(g->enc_done[i]);
    }
    if (g->totals) cudaFreeHost(g->totals);
    delete g;
    return CTR_OK;
}

extern "C" int ctr_gather_begin_turn(ctr_gather_t g, uint32_t *turn, uint32_t **words_slot_dev,
                                     uint64_t *words_capacity, uint64_t **offsets_slot_dev) {
    if (!g || !turn) return CTR_ERR_BAD_ARGUMENT;
    const uint32_t q = ++g->seq;
    *turn = q;
    const uint32_t b = q % g->nb;
    if (words_slot_dev) *words_slot_dev = reinterpret_cast<uint32_t *>(g->words_slot(g->rank, b, g->rank));
    if (words_capacity) *words_capacity = g->slot_words;
    if (offsets_slot_dev) *offsets_slot_dev = reinterpret_cast<uint64_t *>(g->offsets_slot(g->rank, b, g->rank));
    return CTR_OK;
}

extern "C" int ctr_gather_slot(ctr_gather_t g, uint32_t turn, uint32_t src_rank, const uint32_t **words_slot_dev,
                               const uint64_t **offsets_slot_dev) {
    if (!g || src_rank >= g->world) return CTR_ERR_BAD_ARGUMENT;
    const uint32_t b = turn % g->nb;
    if (words_slot_dev) *words_slot_dev = reinterpret_cast<const uint32_t *>(g->words_slot(g->rank, b, src_rank));
    if (offsets_slot_dev) *offsets_slot_dev = reinterpret_cast<const uint64_t *>(g->offsets_slot(g->rank, b, src_rank));
    return CTR_OK;
}

extern "C" int ctr_gather_push(ctr_gather_t g, uint32_t turn, uint64_t n_streams, void *encode_stream) {
    if (!g || n_streams > g->slot_streams || turn == 0) return CTR_ERR_BAD_ARGUMENT;
    if (g->error.load()) return ctr::host_fail(g->error_text), g->error.load();
    cudaStream_t s = (cudaStream_t)encode_stream;
    const int slot = (int)(g->posted++ % kJobRing);
    {  // a ring slot is reused only after the worker is done with the job that used it
        std::unique_lock<std::mutex> lock(g->mu);
        g->idle_cv.wait(lock, [&] { return g->jobs.size() + (g->busy ? 1 : 0) < (size_t)kJobRing - 1; });
    }
    const uint32_t b = turn % g->nb;
    const char *off = g->offsets_slot(g->rank, b, g->rank);
    CTR_HOST_TRY(cudaEventRecord(g->enc_done[slot], s));
    CTR_HOST_TRY(cudaStreamWaitEvent(g->aux, g->enc_done[slot], 0));
    CTR_HOST_TRY(cudaMemcpyAsync(&g->totals[slot], off + n_streams * 8, 8, cudaMemcpyDeviceToHost, g->aux));
    CTR_HOST_TRY(cudaEventRecord(g->encoded[slot], g->aux));
    {
        std::lock_guard<std::mutex> lock(g->mu);
        g->jobs.push_back(Job{turn, n_streams, slot});
    }
    g->cv.notify_one();
    return CTR_OK;
}

extern "C" int ctr_gather_wait(ctr_gather_t g, uint32_t turn, void *consumer_stream) {
    if (!g || turn == 0) return CTR_ERR_BAD_ARGUMENT;
    const uint32_t b = turn % g->nb;
    if (g->world == 1) return CTR_OK;
    for (uint32_t r = 0; r < g->world; ++r) {
        if (r == g->rank) continue;
        const CUresult rc = wait_value_fn()((CUstream)g->wait_stream, (CUdeviceptr)(uintptr_t)g->flag(g->rank, 0, b, r), turn,
                                            CU_STREAM_WAIT_VALUE_GEQ);
        if (rc != CUDA_SUCCESS) return ctr::host_fail("cuStreamWaitValue32 failed (" + std::to_string((int)rc) + ")");
    }
    cudaEvent_t ev = g->sync_ev[g->sync_posted++ % kJobRing];
    CTR_HOST_TRY(cudaEventRecord(ev, g->wait_stream));
    CTR_HOST_TRY(cudaStreamWaitEvent((cudaStream_t)consumer_stream, ev, 0));
    return CTR_OK;
}

extern "C" int ctr_gather_release(ctr_gather_t g, uint32_t turn, void *consumer_stream) {
    if (!g || turn == 0) return CTR_ERR_BAD_ARGUMENT;
    const uint32_t b = turn % g->nb;
    if (g->world == 1) return CTR_OK;
    cudaEvent_t ev = g->sync_ev[g->sync_posted++ % kJobRing];
    CTR_HOST_TRY(cudaEventRecord(ev, (cudaStream_t)consumer_stream));
    CTR_HOST_TRY(cudaStreamWaitEvent(g->release_stream, ev, 0));
    for (uint32_t i = 1; i < g->world; ++i) {
        const uint32_t r = (g->rank + i) % g->world;
        const CUresult rc = write_value_fn()((CUstream)g->release_stream, (CUdeviceptr)(uintptr_t)g->flag(r, 1, b, g->rank), turn,
                                             CU_STREAM_WRITE_VALUE_DEFAULT);
        if (rc != CUDA_SUCCESS) return ctr::host_fail("cuStreamWriteValue32 failed (" + std::to_string((int)rc) + ")");
    }
    return CTR_OK;
}

extern "C" int ctr_gather_sync(ctr_gather_t g) {
    if (!g) return CTR_ERR_BAD_ARGUMENT;
    {
        std::unique_lock<std::mutex> lock(g->mu);
        g->idle_cv.wait(lock, [&] { return g->jobs.empty() && !g->busy; });
    }
    for (int i = 0; i < g->n_push; ++i) CTR_HOST_TRY(cudaStreamSynchronize(g->push[i]));
    CTR_HOST_TRY(cudaStreamSynchronize(g->release_stream));
    if (g->error.load()) return ctr::host_fail(g->error_text), g->error.load();
    return CTR_OK;
}

// =====================================================================================================
// NCCL variant (communicator supplied by the caller; NCCL resolved at run time)
// =====================================================================================================
namespace {

struct Nccl {
    int (*all_gather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    int (*broadcast)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*group_start)() = nullptr;
    int (*group_end)() = nullptr;
    const char *(*error_string)(int) = nullptr;
    bool ok = false;
};

const Nccl &nccl() {
    static const Nccl n = [] {
        Nccl r;
        void *h = nullptr;
        if (const char *path = getenv("CTR_NCCL_LIB")) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
        // already loaded by the host program (PyTorch, a Rust / C host linked against NCCL)?
        if (!h && dlsym(RTLD_DEFAULT, "ncclAllGather")) h = RTLD_DEFAULT;
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return r;
        r.all_gather = (decltype(r.all_gather))dlsym(h, "ncclAllGather");
        r.broadcast = (decltype(r.broadcast))dlsym(h, "ncclBroadcast");
        r.group_start = (decltype(r.group_start))dlsym(h, "ncclGroupStart");
        r.group_end = (decltype(r.group_end))dlsym(h, "ncclGroupEnd");
        r.error_string = (decltype(r.error_string))dlsym(h, "ncclGetErrorString");
        r.ok = r.all_gather && r.broadcast && r.group_start && r.group_end;
        return r;
    }();
    return n;
}
constexpr int kNcclUint32 = 3, kNcclUint64 = 5;  // ncclDataType_t

int nccl_fail(int code, const char *what) {
    const Nccl &n = nccl();
    return ctr::host_fail(std::string(what) + ": " + (n.error_string ? n.error_string(code) : "NCCL error") + " (" +
                          std::to_string(code) + ")");
}

}  // namespace

extern "C" int ctr_gather_compressed_nccl(void *nccl_comm, uint32_t world, uint32_t rank, const uint32_t *words_dev,
                                          const uint64_t *offsets_dev, uint64_t n_streams, uint64_t slot_words,
                                          uint64_t slot_streams, uint32_t *words_out_dev, uint64_t *offsets_out_dev,
                                          uint64_t *meta_dev, uint64_t *meta_host, void *stream) {
    if (!nccl_comm || world == 0 || rank >= world || !offsets_dev || !words_out_dev || !offsets_out_dev || !meta_dev || !meta_host)
        return CTR_ERR_BAD_ARGUMENT;
    if (n_streams > slot_streams || slot_words % 4 != 0) return CTR_ERR_BAD_ARGUMENT;
    if (ctr_device_count() == 0) return ctr::host_cuda_fail(cudaErrorNoDevice, "no CUDA device");
    const Nccl &n = nccl();
    if (!n.ok) return ctr::host_fail("NCCL is not loaded in this process (set CTR_NCCL_LIB)");
    cudaStream_t s = (cudaStream_t)stream;
    // 1. sizes: {total words, streams} of every rank; the one host wait of the exchange
    uint64_t *mine = meta_dev + 2 * world;  // scratch behind the gathered table: meta_dev is u64[2 * world + 2]
    CTR_HOST_TRY(cudaMemcpyAsync(mine, offsets_dev + n_streams, 8, cudaMemcpyDeviceToDevice, s));
    CTR_HOST_TRY(cudaMemcpyAsync(mine + 1, &n_streams, 8, cudaMemcpyHostToDevice, s));
    int rc = n.all_gather(mine, meta_dev, 2, kNcclUint64, nccl_comm, s);
    if (rc) return nccl_fail(rc, "ncclAllGather(sizes)");
    CTR_HOST_TRY(cudaMemcpyAsync(meta_host, meta_dev, 16 * (size_t)world, cudaMemcpyDeviceToHost, s));
    CTR_HOST_TRY(cudaStreamSynchronize(s));
    for (uint32_t r = 0; r < world; ++r)
        if (meta_host[2 * r] > slot_words || meta_host[2 * r + 1] > slot_streams) return CTR_ERR_OUT_OF_SPACE;
    // 2. every rank's words and offset table straight into its slot on every rank
    if ((rc = n.group_start())) return nccl_fail(rc, "ncclGroupStart");
    for (uint32_t r = 0; r < world; ++r) {
        uint32_t *w_dst = words_out_dev + (uint64_t)r * slot_words;
        uint64_t *o_dst = offsets_out_dev + (uint64_t)r * (slot_streams + 1);
        if (meta_host[2 * r]) {
            rc = n.broadcast(r == rank ? (const void *)words_dev : (const void *)w_dst, w_dst, meta_host[2 * r], kNcclUint32, (int)r,
                             nccl_comm, s);
            if (rc) break;
        }
        rc = n.broadcast(r == rank ? (const void *)offsets_dev : (const void *)o_dst, o_dst, meta_host[2 * r + 1] + 1, kNcclUint64,
                         (int)r, nccl_comm, s);
        if (rc) break;
    }
    const int rc_end = n.group_end();
    if (rc) return nccl_fail(rc, "ncclBroadcast");
    if (rc_end) return nccl_fail(rc_end, "ncclGroupEnd");
    return CTR_OK;
}
