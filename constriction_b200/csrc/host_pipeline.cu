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
// has none.  Every pipeline slot keeps one grow-only block of device memory between calls.
//
// `_async` variants run the same pipeline on a thread of the library and return a job handle, so that one host
// thread can keep both directions of the bus busy (upload of the batch being encoded, download of the batch
// being decoded).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <stdio.h>

#include <algorithm>
#include <chrono>
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

// CTR_HOST_TRACE=1: host-side timeline of the pipeline on stderr (diagnostics)
bool trace_on() {
    static const bool on = getenv("CTR_HOST_TRACE") != nullptr;
    return on;
}
double now_ms() {
    static const auto t0 = std::chrono::steady_clock::now();
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}
#define CTR_TRACE(...)                                  \
    do {                                                \
        if (trace_on()) {                               \
            fprintf(stderr, "[%9.3f] ", now_ms());      \
            fprintf(stderr, __VA_ARGS__);               \
            fputc('\n', stderr);                        \
        }                                               \
    } while (0)

// out[i] = in[i] - in[0]
__global__ void rebase_offsets_kernel(const uint64_t *in, uint64_t *out, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] - in[0];
}

// ---- encode results straight into pinned host memory ---------------------------------------------------------------
// When the caller's output buffers are pinned (device-accessible under unified addressing), a chunk's results leave
// through SM stores instead of copy-engine transfers: one small kernel per chunk writes the chunk's words to the place
// where the previous chunk's words end, its offsets rebased by that amount, and hands the running total (and the
// merged data status) to the next chunk through device memory.  The host then never needs a chunk's size -- an
// encode call has no host wait besides slot reuse -- and the small downloads do not queue behind a concurrent decode
// job's symbol downloads on the copy engine.
struct EmitChain {  // device memory, one entry per chunk + 1
    uint64_t base;       // words of all earlier chunks
    uint32_t status[4];  // merged data status so far {code, overflow flag, failing stream lo, hi}
};

__global__ void emit_chunk_kernel(const uint32_t *__restrict__ words, const uint64_t *__restrict__ offsets, uint64_t kc, uint64_t k0,
                                  uint64_t chunk_capacity, const uint32_t *__restrict__ chunk_status, const EmitChain *in,
                                  EmitChain *out, uint32_t *host_words, uint64_t capacity, uint64_t *host_offsets,
                                  uint64_t *host_final /* {total, status lo/hi words} or null unless last chunk */) {
    const uint64_t base = in->base;
    const uint64_t total = offsets[kc];
    const bool fits = total <= chunk_capacity && base + total <= capacity;
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t k = tid; k < kc; k += nthreads) host_offsets[k0 + k] = offsets[k] + base;
    if (fits) {
        // 16-byte stores once the destination is 16-byte aligned (the source then is not, in general: 4 scalar loads)
        uint32_t *dst = host_words + base;
        const uint64_t head = min(total, (uint64_t)((16 - ((uintptr_t)dst & 15)) & 15) / 4);
        if (tid < head) dst[tid] = words[tid];
        const uint64_t vecs = (total - head) / 4;
        uint4 *dst4 = reinterpret_cast<uint4 *>(dst + head);
        const uint32_t *src = words + head;
        for (uint64_t v = tid; v < vecs; v += nthreads) {
            const uint32_t *p = src + 4 * v;
            dst4[v] = make_uint4(p[0], p[1], p[2], p[3]);
        }
        for (uint64_t j = head + 4 * vecs + tid; j < total; j += nthreads) dst[j] = words[j];
    }
    if (tid == 0) {
        EmitChain o = *in;
        o.base = base + (fits ? total : 0);
        if (!fits) o.status[1] = 1u;
        if (chunk_status[0] > o.status[0]) {
            const uint64_t bad = k0 + (((uint64_t)chunk_status[3] << 32) | chunk_status[2]);
            o.status[0] = chunk_status[0];
            o.status[2] = (uint32_t)bad;
            o.status[3] = (uint32_t)(bad >> 32);
        }
        *out = o;
        if (host_final) {
            host_offsets[k0 + kc] = o.base;
            host_final[0] = o.base;
            host_final[1] = ((uint64_t)o.status[1] << 32) | o.status[0];
            host_final[2] = ((uint64_t)o.status[3] << 32) | o.status[2];
        }
    }
}

// device-accessible host memory (cudaHostAlloc / cudaHostRegister)?
bool is_pinned(const void *p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeHost && attr.devicePointer != nullptr;
}

struct SlotRes {
    cudaStream_t s = nullptr;
    cudaEvent_t ev_size = nullptr, ev_done = nullptr;
    uint64_t *h_meta = nullptr;  // pinned: {total words, status[0..3] as 2 x u64}
    // device memory of the chunk in this slot: one grow-only block per slot, kept between calls (allocating from
    // the stream-ordered pool per call made two concurrent pipelines -- an encode and a decode job -- wait for each
    // other's frees), carved up by a bump pointer
    char *arena = nullptr;
    size_t arena_bytes = 0, used = 0;
};
struct PipeRes {
    int device = -1;
    SlotRes slot[kSlots];
    // decode calls: the compressed words of ALL chunks are uploaded on their own stream as soon as the call starts
    // (they are small next to the symbols that come back), so that no kernel ever waits for an upload that queues
    // behind a concurrent encode job's symbol copies
    EmitChain *chain = nullptr;            // device, kMaxChunks + 1 entries (encode calls with pinned outputs)
    cudaEvent_t emit_ev[kMaxChunks] = {};
    cudaStream_t up = nullptr;
    cudaEvent_t up_ev[kMaxChunks] = {};
    char *up_arena = nullptr;
    size_t up_bytes = 0;
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
    {  // small, latency-critical uploads: highest priority
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        ok = ok && cudaStreamCreateWithPriority(&r->up, cudaStreamNonBlocking, hi) == cudaSuccess;
    }
    for (int i = 0; i < kMaxChunks && ok; ++i)
        ok = cudaEventCreateWithFlags(&r->up_ev[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&r->emit_ev[i], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaMalloc((void **)&r->chain, sizeof(EmitChain) * (kMaxChunks + 1)) == cudaSuccess;
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

// the slot's stream must be idle (its previous chunk has left the device) when a chunk's memory is laid out
int arena_reserve(SlotRes &S, size_t bytes) {
    S.used = 0;
    if (bytes <= S.arena_bytes) return CTR_OK;
    CTR_HOST_TRY(cudaStreamSynchronize(S.s));
    if (S.arena) cudaFree(S.arena);
    S.arena = nullptr;
    S.arena_bytes = 0;
    const size_t want = align_up(bytes + bytes / 8, 1 << 20);
    CTR_HOST_TRY(cudaMalloc((void **)&S.arena, want));
    S.arena_bytes = want;
    return CTR_OK;
}
inline size_t arena_size(size_t bytes) { return align_up(bytes ? bytes : 16, 256); }
template <typename T>
T *arena_take(SlotRes &S, size_t bytes) {
    T *p = reinterpret_cast<T *>(S.arena + S.used);
    S.used += arena_size(bytes);
    return p;
}

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
    if (const char *e = getenv("CTR_HOST_CHUNKS")) want = std::max(1, atoi(e));  // experiments
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

// Big copies are issued in pieces so that the copies of a concurrent job in the same direction (an encode job's
// symbols and a decode job's words both go up) interleave at piece granularity instead of chunk granularity.
size_t piece_bytes() {
    static const size_t v = [] {
        const char *e = getenv("CTR_HOST_PIECE_MB");
        const long mb = e ? atol(e) : 16;
        return mb > 0 ? (size_t)mb << 20 : ~(size_t)0;
    }();
    return v;
}

int copy_1d(char *dst, const char *src, size_t bytes, cudaMemcpyKind kind, cudaStream_t s) {
    const size_t piece = piece_bytes();
    for (size_t done = 0; done < bytes;) {
        const size_t m = std::min(piece, bytes - done);
        CTR_HOST_TRY(cudaMemcpyAsync(dst + done, src + done, m, kind, s));
        done += m;
    }
    return CTR_OK;
}

// H2D (to_device) or D2H of one per-symbol array (4-byte elements) of a chunk
int copy_symbol_array(void *dev, const void *host_base, const Batch &b, const Chunk &c, bool to_device, cudaStream_t s) {
    const cudaMemcpyKind kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
    char *h = const_cast<char *>(static_cast<const char *>(host_base));
    char *d = static_cast<char *>(dev);
    if (b.sym_off) {
        if (!c.n) return CTR_OK;
        return to_device ? copy_1d(d, h + c.s0 * 4, c.n * 4, kind, s) : copy_1d(h + c.s0 * 4, d, c.n * 4, kind, s);
    }
    const uint64_t rows_per_piece = std::max<uint64_t>(1, piece_bytes() / (c.kc * 4));
    for (uint64_t r0 = 0; r0 < c.rows_full; r0 += rows_per_piece) {
        const uint64_t rows = std::min(rows_per_piece, c.rows_full - r0);
        char *hp = h + (r0 * b.K + c.k0) * 4, *dp = d + r0 * c.kc * 4;
        if (to_device)
            CTR_HOST_TRY(cudaMemcpy2DAsync(dp, c.kc * 4, hp, b.K * 4, c.kc * 4, rows, kind, s));
        else
            CTR_HOST_TRY(cudaMemcpy2DAsync(hp, b.K * 4, dp, c.kc * 4, c.kc * 4, rows, kind, s));
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

struct SlotBufs {  // device buffers of one chunk (pointers into its slot's arena)
    int32_t *sym = nullptr;
    uint32_t *idx = nullptr, *words = nullptr, *status = nullptr;
    uint64_t *off_raw = nullptr, *off = nullptr, *offsets = nullptr;
    void *ws = nullptr;
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

size_t layout_bytes(const Call &a, const Batch &b, const Chunk &c) {
    size_t n = 0;
    if (b.sym_off) n += 2 * arena_size((c.kc + 1) * 8);
    if (a.index_mode == CTR_INDEX_PER_SYMBOL) n += arena_size(c.n * 4);
    if (a.index_mode == CTR_INDEX_PER_STREAM) n += arena_size(c.kc * 4);
    return n;
}

int fill_layout(ctr_layout *L, const Call &a, const Batch &b, const Chunk &c, SlotBufs &B, SlotRes &S) {
    cudaStream_t s = S.s;
    memset(L, 0, sizeof *L);
    L->n_streams = c.kc;
    L->n_symbols = c.n;
    L->model_index_mode = a.index_mode;
    int rc;
    if (b.sym_off) {
        B.off_raw = arena_take<uint64_t>(S, (c.kc + 1) * 8);
        B.off = arena_take<uint64_t>(S, (c.kc + 1) * 8);
        CTR_HOST_TRY(cudaMemcpyAsync(B.off_raw, b.sym_off + c.k0, (c.kc + 1) * 8, cudaMemcpyHostToDevice, s));
        rebase_offsets_kernel<<<(unsigned)((c.kc + 1 + 255) / 256), 256, 0, s>>>(B.off_raw, B.off, c.kc + 1);
        ctr::host_count_launch();
        L->sym_offsets_dev = B.off;
    }
    if (a.index_mode == CTR_INDEX_PER_SYMBOL) {
        B.idx = arena_take<uint32_t>(S, c.n * 4);
        if ((rc = copy_symbol_array(B.idx, a.model_index, b, c, true, s))) return rc;
        L->model_index_dev = B.idx;
    } else if (a.index_mode == CTR_INDEX_PER_STREAM) {
        B.idx = arena_take<uint32_t>(S, c.kc * 4);
        CTR_HOST_TRY(cudaMemcpyAsync(B.idx, a.model_index + c.k0, c.kc * 4, cudaMemcpyHostToDevice, s));
        L->model_index_dev = B.idx;
    }
    return CTR_OK;
}

int run_encode(const Call &a, PipeRes *res) {
    const Batch b = make_batch(a.N, a.K, a.sym_off);
    const std::vector<Chunk> chunks = plan_chunks(b);
    const size_t n = chunks.size();
    std::vector<SlotBufs> bufs(n);
    std::vector<uint64_t> cap(n);
    uint64_t base = 0;  // words of the chunks finished so far
    int status = 0, rc = CTR_OK;
    uint64_t bad = 0;
    bool out_of_space = false;
    // pinned outputs: results leave through SM stores (emit_chunk_kernel), no host wait for sizes
    const bool direct = n <= (size_t)kMaxChunks && is_pinned(a.words_out) && is_pinned(a.offsets_out) && !getenv("CTR_HOST_NO_DIRECT");
    if (direct) CTR_HOST_TRY(cudaMemsetAsync(res->chain, 0, sizeof(EmitChain), res->slot[0].s));

    auto issue = [&](size_t i) -> int {
        const Chunk &c = chunks[i];
        SlotRes &S = res->slot[i % kSlots];
        SlotBufs &B = bufs[i];
        int r;
        ctr_layout L;
        memset(&L, 0, sizeof L);
        L.n_streams = c.kc;
        L.n_symbols = c.n;
        L.sym_offsets_dev = b.sym_off ? reinterpret_cast<const uint64_t *>(16) : nullptr;  // (sizes only)
        const size_t ws_bytes = ctr_ans_encode_workspace_bytes(&L);
        cap[i] = ctr_ans_max_compressed_words(&L);
        if ((r = arena_reserve(S, arena_size(c.n * 4) + layout_bytes(a, b, c) + arena_size(ws_bytes) + arena_size(cap[i] * 4) +
                                      arena_size((c.kc + 1) * 8) + arena_size(16))))
            return r;
        B.sym = arena_take<int32_t>(S, c.n * 4);
        if ((r = copy_symbol_array(B.sym, a.symbols_in, b, c, true, S.s))) return r;
        if ((r = fill_layout(&L, a, b, c, B, S))) return r;
        B.ws = arena_take<char>(S, ws_bytes);
        B.words = arena_take<uint32_t>(S, cap[i] * 4);
        B.offsets = arena_take<uint64_t>(S, (c.kc + 1) * 8);
        B.status = arena_take<uint32_t>(S, 16);
        CTR_HOST_TRY(cudaMemsetAsync(B.status, 0, 16, S.s));
        r = a.range ? ctr_range_encode(a.model, B.sym, &L, nullptr, B.ws, ws_bytes, B.words, cap[i], B.offsets, nullptr, B.status, S.s)
                    : ctr_ans_encode_reverse(a.model, B.sym, &L, nullptr, B.ws, ws_bytes, B.words, cap[i], B.offsets, nullptr,
                                             B.status, S.s);
        if (r) return r;
        if (direct) {
            if (i > 0) CTR_HOST_TRY(cudaStreamWaitEvent(S.s, res->emit_ev[i - 1], 0));  // the chain: base and status of chunk i-1
            const bool last = i + 1 == n;
            emit_chunk_kernel<<<64, 256, 0, S.s>>>(B.words, B.offsets, c.kc, c.k0, cap[i], B.status, res->chain + i, res->chain + i + 1,
                                                   a.words_out, a.words_capacity, a.offsets_out, last ? res->slot[0].h_meta : nullptr);
            ctr::host_count_launch();
            CTR_HOST_TRY(cudaEventRecord(res->emit_ev[i], S.s));
            CTR_HOST_TRY(cudaEventRecord(S.ev_done, S.s));
            CTR_TRACE("enc issue %zu done (direct)", i);
            return CTR_OK;
        }
        // chunk-relative offsets go straight to their place; the host rebases them once the chunk's base is known
        CTR_HOST_TRY(cudaMemcpyAsync(a.offsets_out + c.k0, B.offsets, c.kc * 8, cudaMemcpyDeviceToHost, S.s));
        CTR_HOST_TRY(cudaMemcpyAsync(S.h_meta, B.offsets + c.kc, 8, cudaMemcpyDeviceToHost, S.s));
        CTR_HOST_TRY(cudaMemcpyAsync(S.h_meta + 1, B.status, 16, cudaMemcpyDeviceToHost, S.s));
        CTR_HOST_TRY(cudaEventRecord(S.ev_size, S.s));
        CTR_TRACE("enc issue %zu done (k %llu..%llu)", i, (unsigned long long)c.k0, (unsigned long long)c.k1);
        return CTR_OK;
    };
    auto finish = [&](size_t i) -> int {
        const Chunk &c = chunks[i];
        SlotRes &S = res->slot[i % kSlots];
        CTR_TRACE("enc finish %zu wait", i);
        CTR_HOST_TRY(cudaEventSynchronize(S.ev_size));
        CTR_TRACE("enc finish %zu size known", i);
        const uint64_t total = S.h_meta[0];
        note_status(S.h_meta, c.k0, &status, &bad);
        if (base + total > a.words_capacity || total > cap[i]) {
            out_of_space = true;
        } else if (total) {
            CTR_HOST_TRY(cudaMemcpyAsync(a.words_out + base, bufs[i].words, total * 4, cudaMemcpyDeviceToHost, S.s));
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
        }
        if (!rc) rc = issue(i);
        if (!rc && i >= 1 && !direct) rc = finish(i - 1);
    }
    if (!rc && n && !direct) rc = finish(n - 1);
    for (int i = 0; i < kSlots; ++i) {
        const cudaError_t e = cudaStreamSynchronize(res->slot[i].s);
        if (e != cudaSuccess && !rc) rc = ctr::host_cuda_fail(e, "cudaStreamSynchronize");
    }
    if (rc) return rc;
    if (direct) {  // the last chunk's kernel wrote offsets[K], the total and the merged status
        const uint64_t *m = res->slot[0].h_meta;
        if (a.data_status) *a.data_status = (int)(uint32_t)m[1];
        if (a.failing_stream) *a.failing_stream = m[2];
        return (m[1] >> 32) ? CTR_ERR_OUT_OF_SPACE : CTR_OK;
    }
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

    // all compressed words and offset tables go up first, chunk by chunk, on the upload stream
    {
        size_t need = 0;
        for (const Chunk &c : chunks) {
            if (a.offsets_in[c.k1] < a.offsets_in[c.k0]) return CTR_ERR_BAD_ARGUMENT;
            need += arena_size((a.offsets_in[c.k1] - a.offsets_in[c.k0]) * 4) + arena_size((c.kc + 1) * 8);
        }
        if (need > res->up_bytes) {
            if (res->up_arena) cudaFree(res->up_arena);
            res->up_arena = nullptr;
            res->up_bytes = 0;
            const size_t want = align_up(need + need / 8, 1 << 20);
            CTR_HOST_TRY(cudaMalloc((void **)&res->up_arena, want));
            res->up_bytes = want;
        }
        size_t at = 0;
        for (size_t i = 0; i < n; ++i) {
            const Chunk &c = chunks[i];
            const uint64_t w0 = a.offsets_in[c.k0], w1 = a.offsets_in[c.k1];
            bufs[i].words = reinterpret_cast<uint32_t *>(res->up_arena + at);  // (256-byte blocks: the decoders' 16-byte reads stay inside)
            at += arena_size((w1 - w0) * 4);
            bufs[i].ws = res->up_arena + at;
            at += arena_size((c.kc + 1) * 8);
            int r = copy_1d(reinterpret_cast<char *>(bufs[i].words), reinterpret_cast<const char *>(a.words_in + w0), (w1 - w0) * 4,
                            cudaMemcpyHostToDevice, res->up);
            if (r) return r;
            CTR_HOST_TRY(cudaMemcpyAsync(bufs[i].ws, a.offsets_in + c.k0, (c.kc + 1) * 8, cudaMemcpyHostToDevice, res->up));
            CTR_HOST_TRY(cudaEventRecord(res->up_ev[i], res->up));
        }
    }

    auto issue = [&](size_t i) -> int {
        const Chunk &c = chunks[i];
        SlotRes &S = res->slot[i % kSlots];
        SlotBufs &B = bufs[i];
        int r;
        ctr_layout L;
        if ((r = arena_reserve(S, arena_size((c.kc + 1) * 8) + layout_bytes(a, b, c) + arena_size(c.n * 4) + arena_size(16)))) return r;
        B.offsets = arena_take<uint64_t>(S, (c.kc + 1) * 8);
        CTR_HOST_TRY(cudaStreamWaitEvent(S.s, res->up_ev[i], 0));
        rebase_offsets_kernel<<<(unsigned)((c.kc + 1 + 255) / 256), 256, 0, S.s>>>(static_cast<const uint64_t *>(B.ws), B.offsets, c.kc + 1);
        ctr::host_count_launch();
        if ((r = fill_layout(&L, a, b, c, B, S))) return r;
        B.sym = arena_take<int32_t>(S, c.n * 4);
        B.status = arena_take<uint32_t>(S, 16);
        CTR_HOST_TRY(cudaMemsetAsync(B.status, 0, 16, S.s));
        r = a.range ? ctr_range_decode(a.model, B.words, B.offsets, &L, nullptr, B.sym, nullptr, nullptr, B.status, S.s)
                    : ctr_ans_decode(a.model, B.words, B.offsets, &L, nullptr, B.sym, nullptr, nullptr, B.status, S.s);
        if (r) return r;
        if ((r = copy_symbol_array(B.sym, a.symbols_out, b, c, false, S.s))) return r;
        CTR_HOST_TRY(cudaMemcpyAsync(S.h_meta + 1, B.status, 16, cudaMemcpyDeviceToHost, S.s));
        CTR_HOST_TRY(cudaEventRecord(S.ev_done, S.s));
        CTR_TRACE("dec issue %zu done", i);
        return CTR_OK;
    };
    auto finish = [&](size_t i) -> int {  // slot reuse / end of call: the chunk has left the device
        SlotRes &S = res->slot[i % kSlots];
        CTR_TRACE("dec finish %zu wait", i);
        CTR_HOST_TRY(cudaEventSynchronize(S.ev_done));
        CTR_TRACE("dec finish %zu left the device", i);
        note_status(S.h_meta, chunks[i].k0, &status, &bad);
        return CTR_OK;
    };
    for (size_t i = 0; i < n && !rc; ++i) {
        if (i >= (size_t)kSlots) rc = finish(i - kSlots);
        if (!rc) rc = issue(i);
    }
    for (size_t i = n >= (size_t)kSlots ? n - kSlots : 0; i < n && !rc; ++i) rc = finish(i);
    for (int i = 0; i < kSlots; ++i) {
        const cudaError_t e = cudaStreamSynchronize(res->slot[i].s);
        if (e != cudaSuccess && !rc) rc = ctr::host_cuda_fail(e, "cudaStreamSynchronize");
    }
    {
        const cudaError_t e = cudaStreamSynchronize(res->up);
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
    CTR_TRACE("%s call start", a.decode ? "dec" : "enc");
    const int rc = a.decode ? run_decode(a, res) : run_encode(a, res);
    CTR_TRACE("%s call end", a.decode ? "dec" : "enc");
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
