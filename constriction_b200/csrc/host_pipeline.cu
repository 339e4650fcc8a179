// host_pipeline.cu -- the host-buffer entry points (ctr_*_host, ctr_*_host_async): what a pyo3 / Rust binding for
// `encode_iid_symbols_reverse` / `decode_iid_symbols` over many coders calls with host memory on both sides.
//
// A call is PCIe-bound (4 bytes per symbol cross the bus, the coder kernels need 2 ns per 1000 symbols), so the
// batch is cut into chunks of consecutive streams that flow through a three-stage pipeline on kSlots CUDA streams:
//     H2D copy of chunk c+1   |   coder kernel of chunk c   |   D2H copy of chunk c-1
// Streams are independent coders, so a chunk is a complete batch of its own:
//   * interleaved deal  -- the chunk's symbols are a column strip of the [rows][K] symbol matrix: one strided
//     (2-D) DMA turns it into a dense [rows][K_chunk] matrix on the device, i.e. an interleaved batch of K_chunk
//     streams (+ a 1-D copy for the ragged last row);
//   * contiguous layout -- the chunk's symbols are one contiguous range; its offsets are rebased on the device.
// The only host waits inside an encode call are for the size of a finished chunk (its words are copied back to
// the place where the previous chunk's words end), taken while later chunks are already in flight; a decode call
// has none.  Device buffers come from the stream-ordered memory pool, once per call and pipeline slot.
//
// `_async` variants run the same pipeline on a thread of the library and return a job handle, so that one host
// thread can keep both directions of the bus busy (upload of the batch being encoded, download of the batch
// being decoded).
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/constriction_b200.h"
#include "host_common.h"

namespace {

constexpr int kSlots = 3;
constexpr uint64_t kMinChunkSymbols = 1ull << 21;  // do not cut batches into pieces smaller than 8 MB of symbols
constexpr uint64_t kMinStripStreams = 1024;        // interleaved deal: strips of >= 4 KB per row keep the 2-D DMA efficient
constexpr int kMaxChunks = 16;

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// out[i] = in[i] - in[0]
__global__ void rebase_offsets_kernel(const uint64_t *in, uint64_t *out, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] - in[0];
}

struct SlotRes {
    cudaStream_t s = nullptr;
    cudaEvent_t ev_size = nullptr, ev_done = nullptr;
    uint64_t *h_meta = nullptr;  // pinned: {total words, status[0..3] as 2 x u64}
};
struct PipeRes {
    int device = -1;
    SlotRes slot[kSlots];
};

std::mutex g_pool_mutex;
std::vector<PipeRes *> g_pool;

PipeRes *acquire_pipe() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    {
        std::lock_guard<std::mutex> lock(g_pool_mutex);
        for (size_t i = 0; i < g_pool.size(); ++i)
            if (g_pool[i]->device == dev) {
                PipeRes *r = g_pool[i];
                g_pool.erase(g_pool.begin() + i);
                return r;
            }
    }
    PipeRes *r = new PipeRes();
    r->device = dev;
    ctr::host_keep_pool_memory();
    bool ok = true;
    for (int i = 0; i < kSlots && ok; ++i) {
        ok = cudaStreamCreateWithFlags(&r->slot[i].s, cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&r->slot[i].ev_size, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&r->slot[i].ev_done, cudaEventDisableTiming) == cudaSuccess &&
             cudaHostAlloc((void **)&r->slot[i].h_meta, 64, cudaHostAllocDefault) == cudaSuccess;
    }
    if (!ok) {
        cudaGetLastError();
        delete r;  // (leaks what was created; this only happens when the device is unusable)
        return nullptr;
    }
    return r;
}
void release_pipe(PipeRes *r) {
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    g_pool.push_back(r);
}

struct DevBuf {
    void *p = nullptr;
    cudaStream_t s = nullptr;
    int alloc(size_t bytes, cudaStream_t stream) {
        s = stream;
        CTR_HOST_TRY(cudaMallocAsync(&p, bytes ? align_up(bytes, 16) : 16, stream));
        return CTR_OK;
    }
    template <typename T>
    T *as() const {
        return static_cast<T *>(p);
    }
    void release() {  // stream-ordered: the pool hands the memory to the next allocation on this stream
        if (p) cudaFreeAsync(p, s);
        p = nullptr;
    }
    ~DevBuf() { release(); }
};

// geometry of the whole batch and of one chunk of streams [k0, k1)
struct Batch {
    uint64_t N, K;
    const uint64_t *sym_off;  // host, or null (interleaved)
    uint64_t T, last;         // interleaved: rows, streams owning a symbol in the last row
};
struct Chunk {
    uint64_t k0, k1, kc;
    uint64_t n;          // symbols
    uint64_t s0;         // contiguous: first symbol
    uint64_t rows_full;  // interleaved: rows in which every stream of the chunk owns a symbol
    uint64_t tail;       // interleaved: streams of the chunk that own one more symbol
};

Batch make_batch(uint64_t N, uint64_t K, const uint64_t *sym_off) {
    Batch b{N, K, sym_off, 0, 0};
    if (!sym_off && K) {
        b.T = (N + K - 1) / K;
        b.last = b.T ? N - (b.T - 1) * K : 0;
    }
    return b;
}

std::vector<Chunk> plan_chunks(const Batch &b) {
    uint64_t want = b.N / kMinChunkSymbols;
    want = std::max<uint64_t>(1, std::min<uint64_t>(want, kMaxChunks));
    if (!b.sym_off) want = std::min<uint64_t>(want, std::max<uint64_t>(1, b.K / kMinStripStreams));
    want = std::min<uint64_t>(want, std::max<uint64_t>(1, b.K));
    std::vector<Chunk> out;
    uint64_t k0 = 0;
    for (uint64_t c = 0; c < want; ++c) {
        uint64_t k1 = c + 1 == want ? b.K : (b.K * (c + 1) / want) / 32 * 32;  // strips start on a multiple of 32 streams
        if (b.sym_off && c + 1 != want) {
            // contiguous: cut by symbols, not by streams (streams may be ragged)
            const uint64_t target = b.N * (c + 1) / want;
            k1 = std::lower_bound(b.sym_off + k0, b.sym_off + b.K, target) - b.sym_off;
        }
        if (k1 <= k0) continue;
        Chunk ch{};
        ch.k0 = k0;
        ch.k1 = k1;
        ch.kc = k1 - k0;
        if (b.sym_off) {
            ch.s0 = b.sym_off[k0];
            ch.n = b.sym_off[k1] - ch.s0;
        } else {
            const bool ragged = b.last != b.K;
            ch.rows_full = ragged ? b.T - 1 : b.T;
            ch.tail = ragged ? std::min(ch.kc, b.last > k0 ? b.last - k0 : 0) : 0;
            ch.n = ch.rows_full * ch.kc + ch.tail;
        }
        out.push_back(ch);
        k0 = k1;
    }
    return out;
}

// sym_offsets must describe slices of the symbol array (checked here because chunks are cut from them)
bool offsets_valid(const uint64_t *sym_off, uint64_t K, uint64_t N) {
    if (sym_off[0] > N) return false;
    for (uint64_t k = 0; k < K; ++k)
        if (sym_off[k + 1] < sym_off[k] || sym_off[k + 1] > N) return false;
    return true;
}

// H2D (to_device) or D2H of one per-symbol array (4-byte elements) of a chunk
int copy_symbol_array(void *dev, const void *host_base, const Batch &b, const Chunk &c, bool to_device, cudaStream_t s) {
    const cudaMemcpyKind kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
    char *h = const_cast<char *>(static_cast<const char *>(host_base));
    char *d = static_cast<char *>(dev);
    if (b.sym_off) {
        if (!c.n) return CTR_OK;
        if (to_device)
            CTR_HOST_TRY(cudaMemcpyAsync(d, h + c.s0 * 4, c.n * 4, kind, s));
        else
            CTR_HOST_TRY(cudaMemcpyAsync(h + c.s0 * 4, d, c.n * 4, kind, s));
        return CTR_OK;
    }
    if (c.rows_full) {
        if (to_device)
            CTR_HOST_TRY(cudaMemcpy2DAsync(d, c.kc * 4, h + c.k0 * 4, b.K * 4, c.kc * 4, c.rows_full, kind, s));
        else
            CTR_HOST_TRY(cudaMemcpy2DAsync(h + c.k0 * 4, b.K * 4, d, c.kc * 4, c.kc * 4, c.rows_full, kind, s));
    }
    if (c.tail) {
        char *hp = h + (c.rows_full * b.K + c.k0) * 4, *dp = d + c.rows_full * c.kc * 4;
        if (to_device)
            CTR_HOST_TRY(cudaMemcpyAsync(dp, hp, c.tail * 4, kind, s));
        else
            CTR_HOST_TRY(cudaMemcpyAsync(hp, dp, c.tail * 4, kind, s));
    }
    return CTR_OK;
}

struct SlotBufs {
    DevBuf sym, idx, off_raw, off, ws, words, offsets, status;
    void release() {
        for (DevBuf *b : {&sym, &idx, &off_raw, &off, &ws, &words, &offsets, &status}) b->release();
    }
};

struct Call {  // arguments of one host call (both directions)
    bool range, decode;
    ctr_model_t model;
    const int32_t *symbols_in;
    int32_t *symbols_out;
    uint64_t N, K;
    const uint64_t *sym_off;
    const uint32_t *model_index;
    int32_t index_mode;
    uint32_t *words_out;
    uint64_t words_capacity;
    uint64_t *offsets_out;
    const uint32_t *words_in;
    const uint64_t *offsets_in;
    int *data_status;
    uint64_t *failing_stream;
};

void note_status(const uint64_t *h_meta, uint64_t k0, int *status, uint64_t *bad) {
    const uint32_t *st = reinterpret_cast<const uint32_t *>(h_meta + 1);
    if ((int)st[0] > *status) {
        *status = (int)st[0];
        *bad = k0 + (((uint64_t)st[3] << 32) | st[2]);
    }
}

int fill_layout(ctr_layout *L, const Call &a, const Batch &b, const Chunk &c, SlotBufs &B, cudaStream_t s) {
    memset(L, 0, sizeof *L);
    L->n_streams = c.kc;
    L->n_symbols = c.n;
    L->model_index_mode = a.index_mode;
    int rc;
    if (b.sym_off) {
        if ((rc = B.off_raw.alloc((c.kc + 1) * 8, s)) || (rc = B.off.alloc((c.kc + 1) * 8, s))) return rc;
        CTR_HOST_TRY(cudaMemcpyAsync(B.off_raw.p, b.sym_off + c.k0, (c.kc + 1) * 8, cudaMemcpyHostToDevice, s));
        rebase_offsets_kernel<<<(unsigned)((c.kc + 1 + 255) / 256), 256, 0, s>>>(B.off_raw.as<uint64_t>(), B.off.as<uint64_t>(), c.kc + 1);
        ctr::host_count_launch();
        L->sym_offsets_dev = B.off.as<uint64_t>();
    }
    if (a.index_mode == CTR_INDEX_PER_SYMBOL) {
        if ((rc = B.idx.alloc(c.n * 4, s))) return rc;
        if ((rc = copy_symbol_array(B.idx.p, a.model_index, b, c, true, s))) return rc;
        L->model_index_dev = B.idx.as<uint32_t>();
    } else if (a.index_mode == CTR_INDEX_PER_STREAM) {
        if ((rc = B.idx.alloc(c.kc * 4, s))) return rc;
        CTR_HOST_TRY(cudaMemcpyAsync(B.idx.p, a.model_index + c.k0, c.kc * 4, cudaMemcpyHostToDevice, s));
        L->model_index_dev = B.idx.as<uint32_t>();
    }
    return CTR_OK;
}

int run_encode(const Call &a, PipeRes *res) {
    const Batch b = make_batch(a.N, a.K, a.sym_off);
    const std::vector<Chunk> chunks = plan_chunks(b);
    const size_t n = chunks.size();
    std::vector<SlotBufs> bufs(n);  // buffers are released (stream-ordered) when the call returns
    std::vector<uint64_t> cap(n);
    uint64_t base = 0;  // words of the chunks finished so far
    int status = 0, rc = CTR_OK;
    uint64_t bad = 0;
    bool out_of_space = false;

    auto issue = [&](size_t i) -> int {
        const Chunk &c = chunks[i];
        SlotRes &S = res->slot[i % kSlots];
        SlotBufs &B = bufs[i];
        int r;
        ctr_layout L;
        if ((r = B.sym.alloc(c.n * 4, S.s))) return r;
        if ((r = copy_symbol_array(B.sym.p, a.symbols_in, b, c, true, S.s))) return r;
        if ((r = fill_layout(&L, a, b, c, B, S.s))) return r;
        const size_t ws_bytes = ctr_ans_encode_workspace_bytes(&L);
        cap[i] = ctr_ans_max_compressed_words(&L);
        if ((r = B.ws.alloc(ws_bytes, S.s)) || (r = B.words.alloc(cap[i] * 4, S.s)) || (r = B.offsets.alloc((c.kc + 1) * 8, S.s)) ||
            (r = B.status.alloc(16, S.s)))
            return r;
        CTR_HOST_TRY(cudaMemsetAsync(B.status.p, 0, 16, S.s));
        r = a.range ? ctr_range_encode(a.model, B.sym.as<int32_t>(), &L, nullptr, B.ws.p, ws_bytes, B.words.as<uint32_t>(), cap[i],
                                       B.offsets.as<uint64_t>(), nullptr, B.status.as<uint32_t>(), S.s)
                    : ctr_ans_encode_reverse(a.model, B.sym.as<int32_t>(), &L, nullptr, B.ws.p, ws_bytes, B.words.as<uint32_t>(),
                                             cap[i], B.offsets.as<uint64_t>(), nullptr, B.status.as<uint32_t>(), S.s);
        if (r) return r;
        // chunk-relative offsets go straight to their place; the host rebases them once the chunk's base is known
        CTR_HOST_TRY(cudaMemcpyAsync(a.offsets_out + c.k0, B.offsets.p, c.kc * 8, cudaMemcpyDeviceToHost, S.s));
        CTR_HOST_TRY(cudaMemcpyAsync(S.h_meta, B.offsets.as<uint64_t>() + c.kc, 8, cudaMemcpyDeviceToHost, S.s));
        CTR_HOST_TRY(cudaMemcpyAsync(S.h_meta + 1, B.status.p, 16, cudaMemcpyDeviceToHost, S.s));
        CTR_HOST_TRY(cudaEventRecord(S.ev_size, S.s));
        return CTR_OK;
    };
    auto finish = [&](size_t i) -> int {
        const Chunk &c = chunks[i];
        SlotRes &S = res->slot[i % kSlots];
        CTR_HOST_TRY(cudaEventSynchronize(S.ev_size));
        const uint64_t total = S.h_meta[0];
        note_status(S.h_meta, c.k0, &status, &bad);
        if (base + total > a.words_capacity || total > cap[i]) {
            out_of_space = true;
        } else if (total) {
            CTR_HOST_TRY(cudaMemcpyAsync(a.words_out + base, bufs[i].words.p, total * 4, cudaMemcpyDeviceToHost, S.s));
        }
        CTR_HOST_TRY(cudaEventRecord(S.ev_done, S.s));
        if (base)
            for (uint64_t k = c.k0; k < c.k1; ++k) a.offsets_out[k] += base;
        base += total;
        return CTR_OK;
    };

    for (size_t i = 0; i < n && !rc; ++i) {
        if (i >= (size_t)kSlots) {  // the chunk that used this slot has left the device: its buffers go back to the pool
            rc = cudaEventSynchronize(res->slot[i % kSlots].ev_done) == cudaSuccess ? CTR_OK : CTR_ERR_CUDA;
            bufs[i - kSlots].release();
        }
        if (!rc) rc = issue(i);
        if (!rc && i >= 1) rc = finish(i - 1);
    }
    if (!rc && n) rc = finish(n - 1);
    for (int i = 0; i < kSlots; ++i) {
        const cudaError_t e = cudaStreamSynchronize(res->slot[i].s);
        if (e != cudaSuccess && !rc) rc = ctr::host_cuda_fail(e, "cudaStreamSynchronize");
    }
    if (rc) return rc;
    a.offsets_out[a.K] = base;
    if (a.data_status) *a.data_status = status;
    if (a.failing_stream) *a.failing_stream = bad;
    return out_of_space ? CTR_ERR_OUT_OF_SPACE : CTR_OK;
}

int run_decode(const Call &a, PipeRes *res) {
    const Batch b = make_batch(a.N, a.K, a.sym_off);
    const std::vector<Chunk> chunks = plan_chunks(b);
    const size_t n = chunks.size();
    std::vector<SlotBufs> bufs(n);
    int status = 0, rc = CTR_OK;
    uint64_t bad = 0;

    auto issue = [&](size_t i) -> int {
        const Chunk &c = chunks[i];
        SlotRes &S = res->slot[i % kSlots];
        SlotBufs &B = bufs[i];
        int r;
        ctr_layout L;
        const uint64_t w0 = a.offsets_in[c.k0], w1 = a.offsets_in[c.k1];
        if (w1 < w0) return CTR_ERR_BAD_ARGUMENT;
        if ((r = B.words.alloc((w1 - w0) * 4, S.s))) return r;
        if (w1 > w0) CTR_HOST_TRY(cudaMemcpyAsync(B.words.p, a.words_in + w0, (w1 - w0) * 4, cudaMemcpyHostToDevice, S.s));
        if ((r = B.ws.alloc((c.kc + 1) * 8, S.s)) || (r = B.offsets.alloc((c.kc + 1) * 8, S.s))) return r;
        CTR_HOST_TRY(cudaMemcpyAsync(B.ws.p, a.offsets_in + c.k0, (c.kc + 1) * 8, cudaMemcpyHostToDevice, S.s));
        rebase_offsets_kernel<<<(unsigned)((c.kc + 1 + 255) / 256), 256, 0, S.s>>>(B.ws.as<uint64_t>(), B.offsets.as<uint64_t>(), c.kc + 1);
        ctr::host_count_launch();
        if ((r = fill_layout(&L, a, b, c, B, S.s))) return r;
        if ((r = B.sym.alloc(c.n * 4, S.s)) || (r = B.status.alloc(16, S.s))) return r;
        CTR_HOST_TRY(cudaMemsetAsync(B.status.p, 0, 16, S.s));
        r = a.range ? ctr_range_decode(a.model, B.words.as<uint32_t>(), B.offsets.as<uint64_t>(), &L, nullptr, B.sym.as<int32_t>(),
                                       nullptr, nullptr, B.status.as<uint32_t>(), S.s)
                    : ctr_ans_decode(a.model, B.words.as<uint32_t>(), B.offsets.as<uint64_t>(), &L, nullptr, B.sym.as<int32_t>(),
                                     nullptr, nullptr, B.status.as<uint32_t>(), S.s);
        if (r) return r;
        if ((r = copy_symbol_array(B.sym.p, a.symbols_out, b, c, false, S.s))) return r;
        CTR_HOST_TRY(cudaMemcpyAsync(S.h_meta + 1, B.status.p, 16, cudaMemcpyDeviceToHost, S.s));
        CTR_HOST_TRY(cudaEventRecord(S.ev_done, S.s));
        return CTR_OK;
    };
    auto finish = [&](size_t i) -> int {  // slot reuse / end of call: the chunk has left the device
        SlotRes &S = res->slot[i % kSlots];
        CTR_HOST_TRY(cudaEventSynchronize(S.ev_done));
        note_status(S.h_meta, chunks[i].k0, &status, &bad);
        return CTR_OK;
    };
    for (size_t i = 0; i < n && !rc; ++i) {
        if (i >= (size_t)kSlots) {
            rc = finish(i - kSlots);
            bufs[i - kSlots].release();
        }
        if (!rc) rc = issue(i);
    }
    for (size_t i = n >= (size_t)kSlots ? n - kSlots : 0; i < n && !rc; ++i) rc = finish(i);
    for (int i = 0; i < kSlots; ++i) {
        const cudaError_t e = cudaStreamSynchronize(res->slot[i].s);
        if (e != cudaSuccess && !rc) rc = ctr::host_cuda_fail(e, "cudaStreamSynchronize");
    }
    if (rc) return rc;
    if (a.data_status) *a.data_status = status;
    if (a.failing_stream) *a.failing_stream = bad;
    return CTR_OK;
}

int run_call(const Call &a) {
    if (!a.model) return CTR_ERR_BAD_ARGUMENT;
    if (a.decode ? (!a.offsets_in || (!a.symbols_out && a.N) || (!a.words_in && a.K && a.offsets_in[a.K]))
                 : (!a.words_out || !a.offsets_out || (!a.symbols_in && a.N)))
        return CTR_ERR_BAD_ARGUMENT;
    if (a.index_mode < 0 || a.index_mode > 2 || (a.index_mode != CTR_INDEX_NONE && !a.model_index)) return CTR_ERR_BAD_ARGUMENT;
    if (a.K == 0 && a.N != 0) return CTR_ERR_BAD_ARGUMENT;
    if (ctr_device_count() == 0) return ctr::host_cuda_fail(cudaErrorNoDevice, "no CUDA device");
    if (a.data_status) *a.data_status = 0;
    if (a.failing_stream) *a.failing_stream = 0;
    if (a.K == 0) {
        if (!a.decode) a.offsets_out[0] = 0;
        return CTR_OK;
    }
    if (a.sym_off && !offsets_valid(a.sym_off, a.K, a.N)) return CTR_ERR_BAD_ARGUMENT;
    PipeRes *res = acquire_pipe();
    if (!res) return ctr::host_fail("host pipeline: cannot create streams / pinned staging");
    const int rc = a.decode ? run_decode(a, res) : run_encode(a, res);
    release_pipe(res);
    return rc;
}

}  // namespace

struct ctr_host_job_s {
    std::thread thread;
    int rc = CTR_OK;
    int device = 0;
    std::string error;
};

namespace {
int start_job(const Call &a, ctr_host_job_t *job) {
    if (!job) return CTR_ERR_BAD_ARGUMENT;
    ctr_host_job_s *j = new ctr_host_job_s();
    if (cudaGetDevice(&j->device) != cudaSuccess) {
        cudaGetLastError();
        j->device = 0;
    }
    j->thread = std::thread([j, a] {
        cudaSetDevice(j->device);
        j->rc = run_call(a);
        if (j->rc == CTR_ERR_CUDA) j->error = ctr_last_cuda_error();  // the error text is thread-local
    });
    *job = j;
    return CTR_OK;
}
Call encode_call(bool range, ctr_model_t model, const int32_t *symbols, uint64_t N, uint64_t K, const uint64_t *sym_off,
                 const uint32_t *model_index, int32_t index_mode, uint32_t *words_out, uint64_t words_capacity,
                 uint64_t *offsets_out, int *data_status, uint64_t *failing_stream) {
    Call a{};
    a.range = range;
    a.decode = false;
    a.model = model;
    a.symbols_in = symbols;
    a.N = N;
    a.K = K;
    a.sym_off = sym_off;
    a.model_index = model_index;
    a.index_mode = index_mode;
    a.words_out = words_out;
    a.words_capacity = words_capacity;
    a.offsets_out = offsets_out;
    a.data_status = data_status;
    a.failing_stream = failing_stream;
    return a;
}
Call decode_call(bool range, ctr_model_t model, const uint32_t *words, const uint64_t *offsets, uint64_t N, uint64_t K,
                 const uint64_t *sym_off, const uint32_t *model_index, int32_t index_mode, int32_t *symbols_out,
                 int *data_status, uint64_t *failing_stream) {
    Call a{};
    a.range = range;
    a.decode = true;
    a.model = model;
    a.words_in = words;
    a.offsets_in = offsets;
    a.N = N;
    a.K = K;
    a.sym_off = sym_off;
    a.model_index = model_index;
    a.index_mode = index_mode;
    a.symbols_out = symbols_out;
    a.data_status = data_status;
    a.failing_stream = failing_stream;
    return a;
}
}  // namespace

#define CTR_ENCODE_ARGS                                                                                                   \
    ctr_model_t model, const int32_t *symbols_host, uint64_t n_symbols, uint64_t n_streams, const uint64_t *sym_offsets_host, \
        const uint32_t *model_index_host, int32_t model_index_mode, uint32_t *words_out_host, uint64_t words_capacity,    \
        uint64_t *offsets_out_host, int *data_status, uint64_t *failing_stream
#define CTR_ENCODE_PASS                                                                                                   \
    model, symbols_host, n_symbols, n_streams, sym_offsets_host, model_index_host, model_index_mode, words_out_host,       \
        words_capacity, offsets_out_host, data_status, failing_stream
#define CTR_DECODE_ARGS                                                                                                   \
    ctr_model_t model, const uint32_t *words_host, const uint64_t *offsets_host, uint64_t n_symbols, uint64_t n_streams,  \
        const uint64_t *sym_offsets_host, const uint32_t *model_index_host, int32_t model_index_mode,                     \
        int32_t *symbols_out_host, int *data_status, uint64_t *failing_stream
#define CTR_DECODE_PASS                                                                                                   \
    model, words_host, offsets_host, n_symbols, n_streams, sym_offsets_host, model_index_host, model_index_mode,          \
        symbols_out_host, data_status, failing_stream

extern "C" int ctr_ans_encode_reverse_host(CTR_ENCODE_ARGS) { return run_call(encode_call(false, CTR_ENCODE_PASS)); }
extern "C" int ctr_range_encode_host(CTR_ENCODE_ARGS) { return run_call(encode_call(true, CTR_ENCODE_PASS)); }
extern "C" int ctr_ans_decode_host(CTR_DECODE_ARGS) { return run_call(decode_call(false, CTR_DECODE_PASS)); }
extern "C" int ctr_range_decode_host(CTR_DECODE_ARGS) { return run_call(decode_call(true, CTR_DECODE_PASS)); }

extern "C" int ctr_ans_encode_reverse_host_async(CTR_ENCODE_ARGS, ctr_host_job_t *job) {
    return start_job(encode_call(false, CTR_ENCODE_PASS), job);
}
extern "C" int ctr_range_encode_host_async(CTR_ENCODE_ARGS, ctr_host_job_t *job) {
    return start_job(encode_call(true, CTR_ENCODE_PASS), job);
}
extern "C" int ctr_ans_decode_host_async(CTR_DECODE_ARGS, ctr_host_job_t *job) {
    return start_job(decode_call(false, CTR_DECODE_PASS), job);
}
extern "C" int ctr_range_decode_host_async(CTR_DECODE_ARGS, ctr_host_job_t *job) {
    return start_job(decode_call(true, CTR_DECODE_PASS), job);
}

extern "C" int ctr_host_job_wait(ctr_host_job_t job) {
    if (!job) return CTR_ERR_BAD_ARGUMENT;
    if (job->thread.joinable()) job->thread.join();
    const int rc = job->rc;
    if (rc == CTR_ERR_CUDA) ctr::host_fail(job->error);
    delete job;
    return rc;
}
