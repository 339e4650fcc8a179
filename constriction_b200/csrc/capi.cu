// capi.cu -- the C ABI declared in include/constriction_b200.h: argument checking, workspace layout,
// kernel launches.  No torch types, no host synchronisation on the device-pointer entry points, no
// CPU fallback (every compute entry point needs a CUDA device).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <new>
#include <string>
#include <utility>
#include <vector>

#include "../../include/constriction_b200.h"
#include "host_common.h"
#include "ans_kernels.cuh"
#include "compact.cuh"
#include "gauss_kernels.cuh"
#include "launch.cuh"
#include "model_tables.cuh"

using namespace ctr;

namespace {

thread_local std::string g_last_cuda_error;
std::atomic<uint64_t> g_launches{0};

// ---- optional kernel timing (ctr_profile_*) ---------------------------------------------------------
struct ProfileSlot {
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;  // recorded, not yet read
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> free_list;
};
std::atomic<int> g_profile_on{0};
std::mutex g_profile_mutex;
ProfileSlot g_profile[4];

struct ProfileScope {  // records start now, stop at destruction
    int which;
    cudaStream_t s;
    std::pair<cudaEvent_t, cudaEvent_t> ev{nullptr, nullptr};
    ProfileScope(int which_, cudaStream_t s_) : which(which_), s(s_) {
        if (!g_profile_on.load(std::memory_order_relaxed)) return;
        std::lock_guard<std::mutex> lock(g_profile_mutex);
        ProfileSlot &slot = g_profile[which];
        if (!slot.free_list.empty()) {
            ev = slot.free_list.back();
            slot.free_list.pop_back();
        } else {
            cudaEventCreate(&ev.first);
            cudaEventCreate(&ev.second);
        }
        cudaEventRecord(ev.first, s);
    }
    ~ProfileScope() {
        if (!ev.first) return;
        cudaEventRecord(ev.second, s);
        std::lock_guard<std::mutex> lock(g_profile_mutex);
        g_profile[which].pending.push_back(ev);
    }
};

int cuda_fail(cudaError_t e, const char *what) {
    g_last_cuda_error = std::string(what) + ": " + cudaGetErrorString(e);
    return CTR_ERR_CUDA;
}
}  // namespace

// shared with the other host translation units (host_common.h)
namespace ctr {
int host_cuda_fail(cudaError_t e, const char *what) { return cuda_fail(e, what); }
int host_fail(const std::string &what) {
    g_last_cuda_error = what;
    return CTR_ERR_CUDA;
}
void host_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace ctr
namespace {
void keep_pool_memory();
}
namespace ctr {
void host_keep_pool_memory() { keep_pool_memory(); }
}  // namespace ctr

namespace {

#define CUDA_TRY(expr)                                       \
    do {                                                     \
        cudaError_t e__ = (expr);                            \
        if (e__ != cudaSuccess) return cuda_fail(e__, #expr); \
    } while (0)

#define LAUNCH_CHECK(name)                                    \
    do {                                                      \
        g_launches.fetch_add(1, std::memory_order_relaxed);   \
        cudaError_t e__ = cudaGetLastError();                 \
        if (e__ != cudaSuccess) return cuda_fail(e__, name);  \
    } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Scratch memory of the entry points that need some (per-symbol Gaussian entries, virtual streams of a chunked
// decode, the host-buffer calls) comes from the device's stream-ordered memory pool.  By default the pool gives
// freed memory back to the driver at every synchronisation, which makes the next call pay for a fresh mapping
// (milliseconds); keep it instead.
void keep_pool_memory() {
    static thread_local int configured_device = -1;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev == configured_device) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t threshold = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
    }
    configured_device = dev;
}
inline unsigned grid_for(uint64_t n, unsigned block) { return (unsigned)((n + block - 1) / block); }

}  // namespace

// build state of one lazily built table (see ctr_model_s)
struct LazyTable {
    cudaEvent_t built = nullptr;       // recorded behind the build kernel(s)
    cudaStream_t builder = nullptr;    // the stream they were enqueued on
    std::atomic<bool> complete{false};  // the event has been observed complete: no further waits needed
};

struct ctr_model_s {
    uint32_t n_models = 0, alphabet = 0;
    int32_t min_symbol = 0;
    uint32_t *d_cdf = nullptr;  // [n_models][alphabet+1]
    uint4 *d_enc = nullptr;     // [n_models][alphabet + 1], built on first encode
    uint4 *d_enc_rep = nullptr; // model 0, every entry 8 times (small alphabets only)
    uint32_t *d_dec = nullptr;  // model 0: quantile index + cdf, for the shared-memory decoders
    uint32_t *d_dec_big = nullptr;  // the same with 2^kBigLutBits buckets (large-batch ANS decoder), built on first use
    uint8_t *d_cidx = nullptr;  // coarse quantile index of every model, for the global-table decoders
    uint32_t dec_cdf_bytes = 0;  // (alphabet + 2) * 4 rounded up to 16: the cdf part of d_dec
    bool shared_ok = false;  // small enough for the shared-memory table kernels
    bool enc_f64 = false;    // d_enc holds double-precision reciprocals (CTR_DIV=f64)
    // transient "model" of the ctr_*_gaussian entry points: no tables, (mean, std) per symbol
    const double *lazy_means = nullptr, *lazy_stds = nullptr;
    double lazy_free_weight = 0.0;
    // Derived tables are built on first use, on the stream of the call that needs them.  A pointer is published
    // only after its build kernel is enqueued (under `lazy_mutex`) together with an event recorded behind it;
    // calls on other streams / host threads wait for that event before their kernel reads the table.
    std::mutex lazy_mutex;
    LazyTable enc_ready, dec_ready, dec_big_ready, cidx_ready;
};

namespace {

// ---- model construction ---------------------------------------------------------------------------

int model_alloc(uint32_t n_models, uint32_t alphabet, int32_t min_symbol, ctr_model_s **out) {
    if (n_models == 0 || alphabet < 2 || alphabet > kTotal) return CTR_ERR_BAD_MODEL;
    ctr_model_s *m = new (std::nothrow) ctr_model_s();
    if (!m) return CTR_ERR_BAD_ARGUMENT;
    m->n_models = n_models;
    m->alphabet = alphabet;
    m->min_symbol = min_symbol;
    m->dec_cdf_bytes = (uint32_t)align_up(((size_t)alphabet + 2) * 4, 16);
    m->shared_ok = alphabet <= kMaxSharedAlphabet;
    // (rounded up to 16 bytes: the pool decoders bulk-copy the whole array into shared memory)
    cudaError_t e = cudaMalloc(&m->d_cdf, align_up((size_t)n_models * ((size_t)alphabet + 1) * 4, 16));
    if (e != cudaSuccess) {
        delete m;
        return cuda_fail(e, "cudaMalloc(cdf)");
    }
    *out = m;
    return CTR_OK;
}

// The ANS encoder's quotient estimate runs on the FP64 pipe by default (3 instructions instead of a
// 7-instruction 64x64 high multiply; measured ~8 % faster on B200).  CTR_DIV=int selects the integer
// estimate; both are followed by the same exact integer correction and give identical words.
bool use_f64_division() {
    static const bool on = [] {
        const char *e = getenv("CTR_DIV");
        return !(e && strcmp(e, "int") == 0);
    }();
    return on;
}

// Makes the calling stream wait for a table that another stream built (no-op once the build is known complete).
int join_lazy_table(LazyTable &t, cudaStream_t s) {
    if (t.complete.load(std::memory_order_acquire) || !t.built) return CTR_OK;
    if (cudaEventQuery(t.built) == cudaSuccess) {
        t.complete.store(true, std::memory_order_release);
        return CTR_OK;
    }
    cudaGetLastError();  // cudaErrorNotReady is not an error
    if (s != t.builder) CUDA_TRY(cudaStreamWaitEvent(s, t.built, 0));
    return CTR_OK;
}
int publish_lazy_table(LazyTable &t, cudaStream_t s) {
    CUDA_TRY(cudaEventCreateWithFlags(&t.built, cudaEventDisableTiming));
    CUDA_TRY(cudaEventRecord(t.built, s));
    t.builder = s;
    return CTR_OK;
}

int ensure_enc_table(ctr_model_s *m, cudaStream_t s) {
    std::lock_guard<std::mutex> lock(m->lazy_mutex);
    if (m->d_enc) return join_lazy_table(m->enc_ready, s);
    const uint64_t entries = (uint64_t)m->n_models * ((uint64_t)m->alphabet + 1);
    uint4 *d_enc = nullptr, *d_rep = nullptr;
    CUDA_TRY(cudaMalloc(&d_enc, entries * 16));
    m->enc_f64 = use_f64_division();
    build_enc_table_kernel<<<grid_for(entries, 256), 256, 0, s>>>(m->d_cdf, m->n_models, m->alphabet,
                                                                  m->enc_f64 ? 1 : 0, d_enc);
    LAUNCH_CHECK("build_enc_table_kernel");
    if (m->alphabet <= kMaxSharedEncAlphabet) {
        const uint32_t n = (m->alphabet + 1) * 8u;
        CUDA_TRY(cudaMalloc(&d_rep, (size_t)n * 16));
        replicate_enc_table_kernel<<<grid_for(n, 256), 256, 0, s>>>(d_enc, m->alphabet, d_rep);
        LAUNCH_CHECK("replicate_enc_table_kernel");
    }
    const int rc = publish_lazy_table(m->enc_ready, s);
    m->d_enc_rep = d_rep;
    m->d_enc = d_enc;
    return rc;
}

int ensure_dec_table(ctr_model_s *m, cudaStream_t s) {
    if (!m->shared_ok) return CTR_OK;
    std::lock_guard<std::mutex> lock(m->lazy_mutex);
    if (m->d_dec) return join_lazy_table(m->dec_ready, s);
    uint32_t *d_dec = nullptr;
    CUDA_TRY(cudaMalloc(&d_dec, kLutBytes + m->dec_cdf_bytes));
    const uint32_t threads = m->alphabet + 2 > (uint32_t)kLutSize ? m->alphabet + 2 : (uint32_t)kLutSize;
    build_dec_table_kernel<<<grid_for(threads, 256), 256, 0, s>>>(m->d_cdf, m->alphabet, d_dec, kLutBits);
    LAUNCH_CHECK("build_dec_table_kernel");
    const int rc = publish_lazy_table(m->dec_ready, s);
    m->d_dec = d_dec;
    return rc;
}

// the finer quantile index of the large-batch ANS decoder (ans_kernels.cuh: kDecBlockShared)
size_t dec_big_bytes(const ctr_model_s *m) { return (size_t)kBigLutBytes + m->dec_cdf_bytes; }
int ensure_dec_big_table(ctr_model_s *m, cudaStream_t s) {
    if (!m->shared_ok) return CTR_OK;
    std::lock_guard<std::mutex> lock(m->lazy_mutex);
    if (m->d_dec_big) return join_lazy_table(m->dec_big_ready, s);
    uint32_t *d_dec = nullptr;
    CUDA_TRY(cudaMalloc(&d_dec, dec_big_bytes(m)));
    const uint32_t threads = std::max<uint32_t>(m->alphabet + 2, 1u << kBigLutBits);
    build_dec_table_kernel<<<grid_for(threads, 256), 256, 0, s>>>(m->d_cdf, m->alphabet, d_dec, kBigLutBits);
    LAUNCH_CHECK("build_dec_table_kernel");
    const int rc = publish_lazy_table(m->dec_big_ready, s);
    m->d_dec_big = d_dec;
    return rc;
}

// coarse index for decoding with global tables (built on the first such decode; alphabets up to 65536)
int ensure_coarse_index(ctr_model_s *m, cudaStream_t s) {
    if (m->alphabet > 65536u) return CTR_OK;
    std::lock_guard<std::mutex> lock(m->lazy_mutex);
    if (m->d_cidx) return join_lazy_table(m->cidx_ready, s);
    const int wide = m->alphabet > 256 ? 1 : 0;
    const uint64_t entries = (uint64_t)m->n_models * 257;
    uint8_t *d_cidx = nullptr;
    CUDA_TRY(cudaMalloc(&d_cidx, align_up(entries * (wide ? 2 : 1), 16)));
    build_coarse_index_kernel<<<grid_for(entries, 256), 256, 0, s>>>(m->d_cdf, m->n_models, m->alphabet, wide, d_cidx);
    LAUNCH_CHECK("build_coarse_index_kernel");
    const int rc = publish_lazy_table(m->cidx_ready, s);
    m->d_cidx = d_cidx;
    return rc;
}

// runs the validation kernel, builds the derived tables (the encoder table only when it is small;
// huge model pools get it on first encode) and turns device error bits into a status.
// Synchronises `s`: model construction is not on the hot path.
int model_finish(ctr_model_s *m, uint32_t *d_err, int strict, cudaStream_t s) {
    const uint64_t entries = (uint64_t)m->n_models * ((uint64_t)m->alphabet + 1);
    validate_cdf_kernel<<<grid_for(entries, 256), 256, 0, s>>>(m->d_cdf, m->n_models, m->alphabet, strict, d_err);
    LAUNCH_CHECK("validate_cdf_kernel");
    int rc = ensure_dec_table(m, s);
    if (rc) return rc;
    if (entries <= (8ull << 20) && (rc = ensure_enc_table(m, s))) return rc;
    uint32_t h_err = 0;
    CUDA_TRY(cudaMemcpyAsync(&h_err, d_err, 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    if (m->d_enc) m->enc_ready.complete.store(true, std::memory_order_release);
    if (m->d_dec) m->dec_ready.complete.store(true, std::memory_order_release);
    return h_err ? CTR_ERR_BAD_MODEL : CTR_OK;
}

struct ErrWord {
    uint32_t *d = nullptr;
    int init(cudaStream_t s) {
        CUDA_TRY(cudaMalloc(&d, 4));
        CUDA_TRY(cudaMemsetAsync(d, 0, 4, s));
        return CTR_OK;
    }
    ~ErrWord() {
        if (d) cudaFree(d);
    }
};

ModelView model_view(const ctr_model_s *m) {
    ModelView v;
    v.cdf = m->d_cdf;
    v.enc = m->d_enc;
    v.enc_rep = m->d_enc_rep;
    v.cidx = m->d_cidx;
    v.dec = m->d_dec;
    v.dec_big = m->d_dec_big;
    v.n_models = m->n_models;
    v.alphabet = m->alphabet;
    v.min_symbol = m->min_symbol;
    v.dec_cdf_bytes = m->dec_cdf_bytes;
    v.pool_cdf_bytes = v.pool_cidx_bytes = 0;
    return v;
}

// ---- layout checks / workspace -----------------------------------------------------------------------

int check_layout(const ctr_layout *L) {
    if (!L) return CTR_ERR_BAD_ARGUMENT;
    if (L->model_index_mode < 0 || L->model_index_mode > 2) return CTR_ERR_BAD_ARGUMENT;
    if (L->model_index_mode != CTR_INDEX_NONE && !L->model_index_dev) return CTR_ERR_BAD_ARGUMENT;
    if (L->n_streams == 0 && L->n_symbols != 0) return CTR_ERR_BAD_ARGUMENT;
    if (L->flags & CTR_FLAG_CHECKPOINTS) {  // contiguous layout, C a multiple of 32, both arrays present
        if (!L->sym_offsets_dev || L->checkpoint_every == 0 || L->checkpoint_every % 32 != 0) return CTR_ERR_BAD_ARGUMENT;
        if (!L->ckpt_offsets_dev || !L->checkpoints_dev) return CTR_ERR_BAD_ARGUMENT;
    }
    return CTR_OK;
}

struct EncodeWorkspace {
    uint64_t scratch_words;
    size_t status_off, ticket_off, total;
    uint64_t n_tiles;
};

// [scratch regions][tile status u64[n_tiles]][ticket u32]
EncodeWorkspace encode_workspace(const ctr_layout *L) {
    EncodeWorkspace w;
    w.scratch_words = scratch_start(L->n_symbols, L->n_streams) + 32;
    w.n_tiles = (L->n_streams + 31) / 32;  // enough for every CTA size (the chain kernels code 32 streams per CTA)
    w.status_off = align_up((size_t)w.scratch_words * 4, 256);
    w.ticket_off = w.status_off + (size_t)w.n_tiles * 8;
    w.total = align_up(w.ticket_off + 8, 256);
    return w;
}

}  // namespace

// =====================================================================================================
// library
// =====================================================================================================
extern "C" int ctr_abi_version(void) { return CTR_ABI_VERSION; }

extern "C" const char *ctr_status_string(int code) {
    switch (code) {
        case CTR_OK: return "ok";
        case CTR_ERR_IMPOSSIBLE_SYMBOL: return "Tried to encode symbol that has zero probability under the used entropy model.";
        case CTR_ERR_INVALID_DATA: return "Tried to decode invalid compressed data.";
        case CTR_ERR_TRAILING_ZERO: return "Invalid compressed data: ANS compressed data never ends in a zero word.";
        case CTR_ERR_NOT_SEALED: return "Cannot unseal compressed data because it doesn't fit into integer number of words.";
        case CTR_ERR_BAD_MODEL: return "Probability distribution not normalizable or invalid model parameter.";
        case CTR_ERR_SEEK: return "Invalid coder state or tried to seek past end of stream.";
        case CTR_ERR_OUT_OF_SPACE: return "Output buffer too small for the compressed data.";
        case CTR_ERR_BAD_ARGUMENT: return "Bad argument.";
        case CTR_ERR_CUDA: return "CUDA error (see ctr_last_cuda_error).";
        default: return "unknown status";
    }
}

extern "C" const char *ctr_last_cuda_error(void) { return g_last_cuda_error.c_str(); }

extern "C" int ctr_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" uint64_t ctr_kernel_launch_count(void) { return g_launches.load(); }

extern "C" void ctr_profile_enable(int on) { g_profile_on.store(on ? 1 : 0); }

extern "C" int ctr_profile_read(int which, double *total_ms, uint64_t *launches) {
    if (which < 0 || which > 3) return CTR_ERR_BAD_ARGUMENT;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> evs;
    {
        std::lock_guard<std::mutex> lock(g_profile_mutex);
        evs.swap(g_profile[which].pending);
    }
    double total = 0.0;
    for (auto &ev : evs) {
        CUDA_TRY(cudaEventSynchronize(ev.second));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, ev.first, ev.second));
        total += ms;
    }
    if (total_ms) *total_ms = total;
    if (launches) *launches = evs.size();
    std::lock_guard<std::mutex> lock(g_profile_mutex);
    for (auto &ev : evs) g_profile[which].free_list.push_back(ev);
    return CTR_OK;
}

// =====================================================================================================
// models
// =====================================================================================================
extern "C" int ctr_model_quantized_gaussian(int32_t min_symbol, int32_t max_symbol, const double *means_host,
                                            const double *stds_host, uint32_t n_models, void *stream,
                                            ctr_model_t *out) {
    if (!out || !means_host || !stds_host || n_models == 0) return CTR_ERR_BAD_ARGUMENT;
    if (!(max_symbol > min_symbol)) return CTR_ERR_BAD_MODEL;
    const uint64_t support = (uint64_t)((int64_t)max_symbol - (int64_t)min_symbol) + 1;
    if (support > kTotal) return CTR_ERR_BAD_MODEL;
    for (uint32_t i = 0; i < n_models; ++i)
        if (!(stds_host[i] > 0.0) || !(means_host[i] == means_host[i])) return CTR_ERR_BAD_MODEL;
    cudaStream_t s = (cudaStream_t)stream;
    ctr_model_s *m = nullptr;
    int rc = model_alloc(n_models, (uint32_t)support, min_symbol, &m);
    if (rc) return rc;
    double *d_params = nullptr;
    ErrWord err;
    auto cleanup = [&](int code) {
        if (d_params) cudaFree(d_params);
        if (code) ctr_model_destroy(m);
        return code;
    };
    if ((rc = err.init(s))) return cleanup(rc);
    if (cudaMalloc(&d_params, (size_t)n_models * 16) != cudaSuccess) return cleanup(cuda_fail(cudaGetLastError(), "cudaMalloc(params)"));
    if (cudaMemcpyAsync(d_params, means_host, (size_t)n_models * 8, cudaMemcpyHostToDevice, s) != cudaSuccess ||
        cudaMemcpyAsync(d_params + n_models, stds_host, (size_t)n_models * 8, cudaMemcpyHostToDevice, s) != cudaSuccess)
        return cleanup(cuda_fail(cudaGetLastError(), "cudaMemcpyAsync(params)"));
    const uint64_t entries = (uint64_t)n_models * (support + 1);
    qgauss_cdf_kernel<<<grid_for(entries, 128), 128, 0, s>>>(min_symbol, max_symbol, d_params, d_params + n_models,
                                                             n_models, (uint32_t)support, m->d_cdf, err.d);
    g_launches.fetch_add(1);
    if (cudaGetLastError() != cudaSuccess) return cleanup(cuda_fail(cudaGetLastError(), "qgauss_cdf_kernel"));
    rc = model_finish(m, err.d, /*strict=*/1, s);
    if (rc) return cleanup(rc);
    *out = m;
    return cleanup(CTR_OK);
}

namespace {
template <typename F>
int model_categorical(const F *pmf, int is_device, uint32_t n_models, uint32_t alphabet, void *stream,
                      ctr_model_t *out) {
    if (!out || !pmf || n_models == 0) return CTR_ERR_BAD_ARGUMENT;
    if (alphabet < 2 || alphabet >= kTotal - 1) return CTR_ERR_BAD_MODEL;  // categorical.rs:30-34
    cudaStream_t s = (cudaStream_t)stream;
    ctr_model_s *m = nullptr;
    int rc = model_alloc(n_models, alphabet, 0, &m);
    if (rc) return rc;
    F *d_pmf = nullptr;
    ErrWord err;
    auto cleanup = [&](int code) {
        if (d_pmf) cudaFree(d_pmf);
        if (code) ctr_model_destroy(m);
        return code;
    };
    if ((rc = err.init(s))) return cleanup(rc);
    const F *src = pmf;
    if (!is_device) {
        const size_t bytes = (size_t)n_models * alphabet * sizeof(F);
        if (cudaMalloc(&d_pmf, bytes) != cudaSuccess) return cleanup(cuda_fail(cudaGetLastError(), "cudaMalloc(pmf)"));
        if (cudaMemcpyAsync(d_pmf, pmf, bytes, cudaMemcpyHostToDevice, s) != cudaSuccess)
            return cleanup(cuda_fail(cudaGetLastError(), "cudaMemcpyAsync(pmf)"));
        src = d_pmf;
    }
    categorical_cdf_kernel<F><<<grid_for(n_models, 64), 64, 0, s>>>(src, n_models, alphabet, m->d_cdf, err.d);
    g_launches.fetch_add(1);
    if (cudaGetLastError() != cudaSuccess) return cleanup(cuda_fail(cudaGetLastError(), "categorical_cdf_kernel"));
    rc = model_finish(m, err.d, /*strict=*/0, s);
    if (rc) return cleanup(rc);
    *out = m;
    return cleanup(CTR_OK);
}
}  // namespace

extern "C" int ctr_model_categorical_f32(const float *pmf, int is_device, uint32_t n_models, uint32_t alphabet,
                                         void *stream, ctr_model_t *out) {
    return model_categorical<float>(pmf, is_device, n_models, alphabet, stream, out);
}
extern "C" int ctr_model_categorical_f64(const double *pmf, int is_device, uint32_t n_models, uint32_t alphabet,
                                         void *stream, ctr_model_t *out) {
    return model_categorical<double>(pmf, is_device, n_models, alphabet, stream, out);
}

// two-parameter leaky quantisers: kind 0 Gaussian, 1 Laplace, 2 Cauchy
extern "C" int ctr_model_quantized(int32_t kind, int32_t min_symbol, int32_t max_symbol, const double *p0_host,
                                   const double *p1_host, uint32_t n_models, void *stream, ctr_model_t *out) {
    if (kind == 0) return ctr_model_quantized_gaussian(min_symbol, max_symbol, p0_host, p1_host, n_models, stream, out);
    if (!out || !p0_host || !p1_host || n_models == 0 || kind < 0 || kind > 2) return CTR_ERR_BAD_ARGUMENT;
    if (!(max_symbol > min_symbol)) return CTR_ERR_BAD_MODEL;
    const uint64_t support = (uint64_t)((int64_t)max_symbol - (int64_t)min_symbol) + 1;
    if (support > kTotal) return CTR_ERR_BAD_MODEL;
    for (uint32_t i = 0; i < n_models; ++i)
        if (!(p1_host[i] > 0.0) || !(p0_host[i] == p0_host[i])) return CTR_ERR_BAD_MODEL;
    cudaStream_t s = (cudaStream_t)stream;
    ctr_model_s *m = nullptr;
    int rc = model_alloc(n_models, (uint32_t)support, min_symbol, &m);
    if (rc) return rc;
    double *d_params = nullptr;
    ErrWord err;
    auto cleanup = [&](int code) {
        if (d_params) cudaFree(d_params);
        if (code) ctr_model_destroy(m);
        return code;
    };
    if ((rc = err.init(s))) return cleanup(rc);
    if (cudaMalloc(&d_params, (size_t)n_models * 16) != cudaSuccess) return cleanup(cuda_fail(cudaGetLastError(), "cudaMalloc(params)"));
    if (cudaMemcpyAsync(d_params, p0_host, (size_t)n_models * 8, cudaMemcpyHostToDevice, s) != cudaSuccess ||
        cudaMemcpyAsync(d_params + n_models, p1_host, (size_t)n_models * 8, cudaMemcpyHostToDevice, s) != cudaSuccess)
        return cleanup(cuda_fail(cudaGetLastError(), "cudaMemcpyAsync(params)"));
    const uint64_t entries = (uint64_t)n_models * (support + 1);
    qdist_cdf_kernel<<<grid_for(entries, 128), 128, 0, s>>>(kind, min_symbol, max_symbol, d_params, d_params + n_models, n_models,
                                                            (uint32_t)support, m->d_cdf, err.d);
    g_launches.fetch_add(1);
    if (cudaGetLastError() != cudaSuccess) return cleanup(cuda_fail(cudaGetLastError(), "qdist_cdf_kernel"));
    rc = model_finish(m, err.d, /*strict=*/1, s);
    if (rc) return cleanup(rc);
    *out = m;
    return cleanup(CTR_OK);
}

extern "C" int ctr_model_binomial(const int32_t *n_host, const double *p_host, uint32_t n_models, void *stream, ctr_model_t *out) {
    if (!out || !n_host || !p_host || n_models == 0) return CTR_ERR_BAD_ARGUMENT;
    int32_t n_max = 0;
    for (uint32_t i = 0; i < n_models; ++i) {
        if (n_host[i] < 1 || !(p_host[i] >= 0.0) || !(p_host[i] <= 1.0)) return CTR_ERR_BAD_MODEL;
        n_max = std::max(n_max, n_host[i]);
    }
    if ((uint64_t)n_max + 1 > kTotal) return CTR_ERR_BAD_MODEL;
    const uint32_t alphabet = (uint32_t)n_max + 1;
    cudaStream_t s = (cudaStream_t)stream;
    ctr_model_s *m = nullptr;
    int rc = model_alloc(n_models, alphabet, 0, &m);
    if (rc) return rc;
    char *d_buf = nullptr;
    ErrWord err;
    auto cleanup = [&](int code) {
        if (d_buf) cudaFree(d_buf);
        if (code) ctr_model_destroy(m);
        return code;
    };
    if ((rc = err.init(s))) return cleanup(rc);
    const size_t scratch_bytes = (size_t)n_models * alphabet * 8, p_off = align_up(scratch_bytes, 16), n_off = p_off + (size_t)n_models * 8;
    if (cudaMalloc(&d_buf, n_off + (size_t)n_models * 4) != cudaSuccess) return cleanup(cuda_fail(cudaGetLastError(), "cudaMalloc(binomial)"));
    if (cudaMemcpyAsync(d_buf + p_off, p_host, (size_t)n_models * 8, cudaMemcpyHostToDevice, s) != cudaSuccess ||
        cudaMemcpyAsync(d_buf + n_off, n_host, (size_t)n_models * 4, cudaMemcpyHostToDevice, s) != cudaSuccess)
        return cleanup(cuda_fail(cudaGetLastError(), "cudaMemcpyAsync(binomial)"));
    binomial_cdf_kernel<<<grid_for(n_models, 64), 64, 0, s>>>(reinterpret_cast<const int32_t *>(d_buf + n_off),
                                                             reinterpret_cast<const double *>(d_buf + p_off), n_models, alphabet,
                                                             reinterpret_cast<double *>(d_buf), m->d_cdf, err.d);
    g_launches.fetch_add(1);
    if (cudaGetLastError() != cudaSuccess) return cleanup(cuda_fail(cudaGetLastError(), "binomial_cdf_kernel"));
    rc = model_finish(m, err.d, /*strict=*/0, s);  // rows of narrower models end in impossible symbols
    if (rc) return cleanup(rc);
    *out = m;
    return cleanup(CTR_OK);
}

namespace {
template <typename F>
int model_categorical_perfect(const F *pmf, int is_device, uint32_t n_models, uint32_t alphabet, void *stream, ctr_model_t *out) {
    if (!out || !pmf || n_models == 0) return CTR_ERR_BAD_ARGUMENT;
    if (alphabet < 2 || alphabet > kTotal) return CTR_ERR_BAD_MODEL;  // categorical.rs:72-74, and 2^24 units for n symbols
    cudaStream_t s = (cudaStream_t)stream;
    ctr_model_s *m = nullptr;
    int rc = model_alloc(n_models, alphabet, 0, &m);
    if (rc) return rc;
    F *d_pmf = nullptr;
    char *d_scratch = nullptr;
    ErrWord err;
    auto cleanup = [&](int code) {
        if (d_pmf) cudaFree(d_pmf);
        if (d_scratch) cudaFree(d_scratch);
        if (code) ctr_model_destroy(m);
        return code;
    };
    if ((rc = err.init(s))) return cleanup(rc);
    const F *src = pmf;
    if (!is_device) {
        const size_t bytes = (size_t)n_models * alphabet * sizeof(F);
        if (cudaMalloc(&d_pmf, bytes) != cudaSuccess) return cleanup(cuda_fail(cudaGetLastError(), "cudaMalloc(pmf)"));
        if (cudaMemcpyAsync(d_pmf, pmf, bytes, cudaMemcpyHostToDevice, s) != cudaSuccess)
            return cleanup(cuda_fail(cudaGetLastError(), "cudaMemcpyAsync(pmf)"));
        src = d_pmf;
    }
    const uint64_t n8 = ((uint64_t)alphabet + 1) / 2 * 2;
    if (cudaMalloc(&d_scratch, (size_t)n_models * n8 * (3 * 8 + 3 * 4)) != cudaSuccess)
        return cleanup(cuda_fail(cudaGetLastError(), "cudaMalloc(perfect scratch)"));
    categorical_perfect_kernel<F><<<grid_for(n_models, 32), 32, 0, s>>>(src, n_models, alphabet, d_scratch, m->d_cdf, err.d);
    g_launches.fetch_add(1);
    if (cudaGetLastError() != cudaSuccess) return cleanup(cuda_fail(cudaGetLastError(), "categorical_perfect_kernel"));
    rc = model_finish(m, err.d, /*strict=*/1, s);
    if (rc) return cleanup(rc);
    *out = m;
    return cleanup(CTR_OK);
}
}  // namespace

extern "C" int ctr_model_categorical_perfect_f32(const float *pmf, int is_device, uint32_t n_models, uint32_t alphabet,
                                                 void *stream, ctr_model_t *out) {
    return model_categorical_perfect<float>(pmf, is_device, n_models, alphabet, stream, out);
}
extern "C" int ctr_model_categorical_perfect_f64(const double *pmf, int is_device, uint32_t n_models, uint32_t alphabet,
                                                 void *stream, ctr_model_t *out) {
    return model_categorical_perfect<double>(pmf, is_device, n_models, alphabet, stream, out);
}

extern "C" int ctr_model_from_cdf(const uint32_t *cdf, int is_device, uint32_t n_models, uint32_t alphabet,
                                  int32_t min_symbol, void *stream, ctr_model_t *out) {
    if (!out || !cdf || n_models == 0) return CTR_ERR_BAD_ARGUMENT;
    cudaStream_t s = (cudaStream_t)stream;
    ctr_model_s *m = nullptr;
    int rc = model_alloc(n_models, alphabet, min_symbol, &m);
    if (rc) return rc;
    ErrWord err;
    if ((rc = err.init(s))) {
        ctr_model_destroy(m);
        return rc;
    }
    const size_t bytes = (size_t)n_models * ((size_t)alphabet + 1) * 4;
    if (cudaMemcpyAsync(m->d_cdf, cdf, bytes, is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s) !=
        cudaSuccess) {
        ctr_model_destroy(m);
        return cuda_fail(cudaGetLastError(), "cudaMemcpyAsync(cdf)");
    }
    rc = model_finish(m, err.d, /*strict=*/0, s);
    if (rc) {
        ctr_model_destroy(m);
        return rc;
    }
    *out = m;
    return CTR_OK;
}

extern "C" int ctr_model_uniform(uint32_t size, void *stream, ctr_model_t *out) {
    if (!out) return CTR_ERR_BAD_ARGUMENT;
    if (size < 2 || size > kTotal) return CTR_ERR_BAD_MODEL;  // uniform.rs:44-77
    cudaStream_t s = (cudaStream_t)stream;
    ctr_model_s *m = nullptr;
    int rc = model_alloc(1, size, 0, &m);
    if (rc) return rc;
    uniform_cdf_kernel<<<grid_for((uint64_t)size + 1, 256), 256, 0, s>>>(size, m->d_cdf);
    g_launches.fetch_add(1);
    if (cudaGetLastError() != cudaSuccess) {
        ctr_model_destroy(m);
        return cuda_fail(cudaGetLastError(), "uniform_cdf_kernel");
    }
    ErrWord err;
    if ((rc = err.init(s)) || (rc = model_finish(m, err.d, /*strict=*/1, s))) {
        ctr_model_destroy(m);
        return rc;
    }
    *out = m;
    return CTR_OK;
}

extern "C" int ctr_model_destroy(ctr_model_t m) {
    if (!m) return CTR_OK;
    if (m->d_cdf) cudaFree(m->d_cdf);
    if (m->d_enc) cudaFree(m->d_enc);
    if (m->d_enc_rep) cudaFree(m->d_enc_rep);
    if (m->d_cidx) cudaFree(m->d_cidx);
    if (m->d_dec) cudaFree(m->d_dec);
    if (m->d_dec_big) cudaFree(m->d_dec_big);
    for (LazyTable *t : {&m->enc_ready, &m->dec_ready, &m->dec_big_ready, &m->cidx_ready})
        if (t->built) cudaEventDestroy(t->built);
    delete m;
    return CTR_OK;
}

extern "C" int ctr_model_info(ctr_model_t m, uint32_t *n_models, uint32_t *alphabet, int32_t *min_symbol) {
    if (!m) return CTR_ERR_BAD_ARGUMENT;
    if (n_models) *n_models = m->n_models;
    if (alphabet) *alphabet = m->alphabet;
    if (min_symbol) *min_symbol = m->min_symbol;
    return CTR_OK;
}

extern "C" const uint32_t *ctr_model_cdf_dev(ctr_model_t m) { return m ? m->d_cdf : nullptr; }

extern "C" int ctr_model_copy_cdf_host(ctr_model_t m, uint32_t *cdf_host, void *stream) {
    if (!m || !cdf_host) return CTR_ERR_BAD_ARGUMENT;
    cudaStream_t s = (cudaStream_t)stream;
    CUDA_TRY(cudaMemcpyAsync(cdf_host, m->d_cdf, (size_t)m->n_models * ((size_t)m->alphabet + 1) * 4,
                             cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return CTR_OK;
}

// =====================================================================================================
// ANS
// =====================================================================================================
extern "C" size_t ctr_ans_encode_workspace_bytes(const ctr_layout *L) {
    if (check_layout(L)) return 0;
    return encode_workspace(L).total;
}

extern "C" uint64_t ctr_ans_max_compressed_words(const ctr_layout *L) {
    if (check_layout(L)) return 0;
    // <= 24 bits per symbol plus two state words per stream
    return scratch_words_for(L->n_symbols) + 4 * L->n_streams + 32;
}

namespace {

bool use_shared_tables(const ctr_model_s *m, const ctr_layout *L) {
    return L->model_index_mode == CTR_INDEX_NONE && m->shared_ok;
}
// the encoders' replicated table needs 128 B per entry
bool use_shared_enc_table(const ctr_model_s *m, const ctr_layout *L) {
    return use_shared_tables(m, L) && m->alphabet <= kMaxSharedEncAlphabet;
}

// Few, long streams (contiguous layout, one model per stream): the chain kernels (chain_kernels.cuh).  CTR_CHAIN=0
// forces the general kernels (tests compare the two).
bool use_chain_kernels(const ctr_layout *L, bool raw_states, bool raw_ok = false) {
    static const bool enabled = [] {
        const char *e = getenv("CTR_CHAIN");
        return !(e && strcmp(e, "0") == 0);
    }();
    if (!enabled || !L->sym_offsets_dev || L->model_index_mode == CTR_INDEX_PER_SYMBOL) return false;
    if (!raw_ok && ((L->flags & CTR_FLAG_RAW) || raw_states)) return false;
    // long streams (the time is the dependent chain of one stream), or so few streams that the general kernels cannot
    // fill the GPU anyway; many medium-length streams (e.g. 12,288 x 1024) stay with the general kernels
    if (L->n_streams > 32768 || L->n_symbols < 64 * L->n_streams) return false;
    return L->n_symbols >= 4096 * L->n_streams || L->n_streams <= 2048;
}

// CTA size: big CTAs amortise the table staging, but a batch with fewer streams than one big CTA per SM is
// latency-bound and wants its warps spread over as many SMs as possible.
constexpr uint64_t kStreamsForBigCtas = 148ull * kAnsBlock;
unsigned encode_block(const ctr_layout *L) { return L->n_streams < kStreamsForBigCtas ? kSmallBlock : kAnsBlock; }
unsigned decode_block(const ctr_layout *L, bool shared, bool contig) {
    if (L->n_streams < kStreamsForBigCtas) return kSmallBlock;
    // shared model + interleaved deal: 1024-thread CTAs stage the 32 KB quantile index once per SM
    if (shared && !contig && L->n_streams >= 148ull * kDecBlockShared) return kDecBlockShared;
    return kAnsBlock;
}

// shared memory a pool decoder may use per CTA (227 KB is the limit; static shared memory needs a little)
constexpr size_t kPoolSmemBudget = 220 * 1024;

// dynamic shared memory of a coder kernel: per-warp word staging + tables + 32x32 transposition tiles
// (contiguous layout: the encoders double-buffer the symbol tile; one more tile for per-symbol model indices)
size_t coder_smem_bytes(size_t table_bytes, const ctr_layout *L, int warps, size_t stage_words_per_warp, int sym_tiles) {
    const bool contig = L->sym_offsets_dev != nullptr;
    int tiles = 0;
    if (contig) tiles += sym_tiles + (L->model_index_mode == CTR_INDEX_PER_SYMBOL ? 1 : 0);
    return table_bytes + (size_t)warps * stage_words_per_warp * 4 + (size_t)tiles * warps * kTileWords * 4;
}

struct AnsEncodeLauncher {
    static constexpr bool kTma = true;
    static constexpr int kSlot = 0;
    static constexpr const char *kName = "ans_encode_kernel";
    static cudaError_t run(const LaunchCfg &cfg, const AnsParams &p) { return launch_ans_encode(cfg, p); }
};
struct AnsDecodeLauncher {
    static constexpr bool kTma = true;
    static constexpr int kSlot = 1;
    static constexpr const char *kName = "ans_decode_kernel";
    static cudaError_t run(const LaunchCfg &cfg, const AnsParams &p) { return launch_ans_decode(cfg, p); }
};
struct RangeEncodeLauncher {
    static constexpr bool kTma = true;
    static constexpr int kSlot = 2;
    static constexpr const char *kName = "range_encode_kernel";
    static cudaError_t run(const LaunchCfg &cfg, const AnsParams &p) { return launch_range_encode(cfg, p); }
};
struct RangeDecodeLauncher {
    static constexpr bool kTma = false;
    static constexpr int kSlot = 3;
    static constexpr const char *kName = "range_decode_kernel";
    static cudaError_t run(const LaunchCfg &cfg, const AnsParams &p) { return launch_range_decode(cfg, p); }
};

template <class Launcher>
int run_coder_kernel(const LaunchCfg &cfg, const AnsParams &p) {
    ProfileScope prof(Launcher::kSlot, cfg.stream);
    const cudaError_t e = Launcher::run(cfg, p);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (e != cudaSuccess) return cuda_fail(e, Launcher::kName);
    return CTR_OK;
}

// The interleaved symbol array as a 2-D tensor [full rows][K] of int32 for the TMA paths of the coder kernels
// (boxes of kBoxRows rows x 32 streams).  Returns false when the array cannot be described (K not a multiple of
// 4: row pitch must be a multiple of 16 bytes; misaligned base; fewer than kBoxRows full rows; no driver entry
// point), in which case the kernels use their per-row load / store paths.
using EncodeTiledFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
    static const EncodeTiledFn fn = [] {
        if (const char *e = getenv("CTR_TMA"))
            if (strcmp(e, "0") == 0) return (EncodeTiledFn) nullptr;
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            return (EncodeTiledFn) nullptr;
        }
        return (EncodeTiledFn)f;
    }();
    return fn;
}
// The decoders' TMA store path is opt-in (CTR_TMA_DECODE=1): on B200 it measures 2 % slower than per-row
// streaming stores (157 -> 160 us at 1e8 symbols), because a box needs a proxy fence and two warp barriers.
bool tma_decode_enabled() {
    static const bool on = [] {
        const char *e = getenv("CTR_TMA_DECODE");
        return e && strcmp(e, "1") == 0;
    }();
    return on;
}
bool make_symbol_tensor_map(CUtensorMap *out, const void *symbols, const ctr_layout *L, uint32_t box_streams = 32) {
    const uint64_t K = L->n_streams;
    if (K == 0 || K % 4 != 0 || K > 0xffffffffull) return false;
    const uint64_t full_rows = L->n_symbols / K;
    if (full_rows < (uint64_t)kBoxRows + 1 || full_rows > 0x7fffffffull) return false;
    if (reinterpret_cast<uintptr_t>(symbols) % 16 != 0) return false;
    const EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {K, full_rows};
    const cuuint64_t strides[1] = {K * 4};  // bytes between rows
    const cuuint32_t box[2] = {box_streams, (cuuint32_t)kBoxRows};
    const cuuint32_t elem_strides[2] = {1, 1};
    return fn(out, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, const_cast<void *>(symbols), dims, strides, box, elem_strides,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

AnsParams base_params(const ctr_model_s *m, const ctr_layout *L) {
    AnsParams p;
    memset(&p, 0, sizeof p);
    p.model = model_view(m);
    p.K = L->n_streams;
    p.N = L->n_symbols;
    p.sym_off = L->sym_offsets_dev;
    p.model_index = L->model_index_dev;
    p.index_mode = L->model_index_mode;
    p.flags = L->flags;
    if (L->flags & CTR_FLAG_CHECKPOINTS) {
        p.ckpt_every = L->checkpoint_every;
        p.ckpt_off = L->ckpt_offsets_dev;
        p.ckpt_out = L->checkpoints_dev;
    }
    return p;
}

// writes offsets[0..K] = 0 for an empty batch
int empty_offsets(uint64_t *offsets, uint64_t K, cudaStream_t s) {
    CUDA_TRY(cudaMemsetAsync(offsets, 0, (size_t)(K + 1) * 8, s));
    return CTR_OK;
}

// ---- checkpoints ------------------------------------------------------------------------------------------
// records per stream: J_k = ceil(n_k / C); exclusive prefix sums by one CTA (K is the number of *real* streams)
__global__ void checkpoint_offsets_kernel(const uint64_t *sym_off, uint64_t K, uint32_t every, uint64_t *out) {
    __shared__ uint64_t s_carry, s_warp[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint64_t base = 0; base < K; base += blockDim.x) {
        const uint64_t k = base + threadIdx.x;
        const uint64_t j = k < K ? (sym_off[k + 1] - sym_off[k] + every - 1) / every : 0;
        const uint64_t inc = warp_inclusive_scan_u64(j, lane);
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        uint64_t before = s_carry;
        for (int w = 0; w < warp; ++w) before += s_warp[w];
        if (k < K) out[k] = before + inc - j;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry = before + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[K] = s_carry;
}

// Virtual streams of a chunk-parallel decode: chunk j of stream k becomes stream v = ckpt_off[k] + j with its own
// word range, symbol range and raw start state; entries beyond the last real chunk (the arrays are sized by an
// upper bound) are empty streams.  RANGE: 4 state words and the decoder's `point` is read from the words
// (RangeDecoder::seek -> read_point, queue.rs:847-868,911-928).
template <bool RANGE>
__global__ void expand_checkpoints_kernel(const uint64_t *sym_off, const uint64_t *offsets, const uint32_t *words,
                                          const uint64_t *ckpt_off, const uint64_t *ckpt, uint32_t every, uint64_t K,
                                          uint64_t V_max, uint64_t N, const uint32_t *stream_index, uint64_t *v_begin,
                                          uint64_t *v_end, uint64_t *v_sym_off, uint64_t *v_state, uint32_t *v_index) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t V = ckpt_off[K];
    if (k < K) {
        const uint64_t o = sym_off[k], n = sym_off[k + 1] - o;
        const uint64_t J = (n + every - 1) / every;
        const uint64_t base = offsets[k], stream_end = offsets[k + 1];
        const uint64_t v0 = ckpt_off[k];
        for (uint64_t j = 0; j < J; ++j) {
            const uint64_t v = v0 + j;
            if (v >= V_max) break;
            if (RANGE) {
                const uint64_t *rec = ckpt + 4 * v;
                const uint64_t pos = base + rec[0];
                uint64_t point = 0;
                if (pos < stream_end) point = (uint64_t)words[pos] << 32;
                if (pos + 1 < stream_end) point |= words[pos + 1];
                v_sym_off[v] = o + j * every;
                v_begin[v] = pos + 2 < stream_end ? pos + 2 : stream_end;
                v_end[v] = stream_end;
                v_state[4 * v] = rec[1];
                v_state[4 * v + 1] = rec[2];
                v_state[4 * v + 2] = point;
                v_state[4 * v + 3] = 0;
            } else {
                const uint64_t *rec = ckpt + 2 * v;
                v_sym_off[v] = j == 0 ? o : o + n - (J - j) * every;
                v_end[v] = base + rec[0];
                v_begin[v] = j + 1 < J ? base + ckpt[2 * (v + 1)] : base;
                v_state[v] = rec[1];
            }
            if (v_index) v_index[v] = stream_index[k];
        }
    }
    // padding entries [V, V_max] : empty streams behind the last real stream (the last chunk ends where that
    // stream ends, which need not be the end of the symbol array)
    const uint64_t sym_end = sym_off[K] <= N ? sym_off[K] : N;
    for (uint64_t v = V + k; v <= V_max; v += (uint64_t)gridDim.x * blockDim.x) {
        v_sym_off[v] = sym_end;
        if (v < V_max) {
            v_begin[v] = 0;
            v_end[v] = 0;
            for (int i = 0; i < (RANGE ? 4 : 1); ++i) v_state[(RANGE ? 4 : 1) * v + i] = RANGE && i == 1 ? ~0ull : 0;
            if (v_index) v_index[v] = 0;
        }
    }
}

template <class EncLauncher>
int encode_common(ctr_model_t model, const int32_t *symbols_dev, const ctr_layout *L, const uint64_t *states_in,
                  void *workspace, size_t workspace_bytes, uint32_t *words_out, uint64_t capacity,
                  uint64_t *offsets_out, uint64_t *states_out, uint32_t *status, void *stream) {
    int rc = check_layout(L);
    if (rc) return rc;
    if (!model || !offsets_out || (!symbols_dev && L->n_symbols)) return CTR_ERR_BAD_ARGUMENT;
    if (ctr_device_count() == 0) return cuda_fail(cudaErrorNoDevice, "no CUDA device");
    cudaStream_t s = (cudaStream_t)stream;
    if (L->n_streams == 0) return empty_offsets(offsets_out, 0, s);
    const EncodeWorkspace w = encode_workspace(L);
    if (!workspace || workspace_bytes < w.total || !words_out) return CTR_ERR_BAD_ARGUMENT;
    if ((rc = ensure_enc_table(model, s))) return rc;

    AnsParams p = base_params(model, L);
    p.symbols_in = symbols_dev;
    p.states_in = states_in;
    p.states_out = states_out;
    p.status = status;
    char *ws = static_cast<char *>(workspace);
    p.scratch = reinterpret_cast<uint32_t *>(ws);
    p.compact.tile_status = reinterpret_cast<uint64_t *>(ws + w.status_off);
    p.compact.ticket = reinterpret_cast<unsigned int *>(ws + w.ticket_off);
    p.compact.words_out = words_out;
    p.compact.words_capacity = capacity;
    p.compact.offsets_out = offsets_out;
    CUDA_TRY(cudaMemsetAsync(ws + w.status_off, 0, w.total - w.status_off, s));

    LaunchCfg cfg;
    cfg.pool = false;
    cfg.gauss = false;
    cfg.shared = use_shared_enc_table(model, L);
    cfg.contig = L->sym_offsets_dev != nullptr;
    cfg.persym = L->model_index_mode == CTR_INDEX_PER_SYMBOL;
    cfg.f64 = model->enc_f64;
    cfg.stream = s;
    if (use_chain_kernels(L, states_in != nullptr || states_out != nullptr)) {
        // few long streams: one coder warp per 32 streams, helper warps do everything that is not the state update
        constexpr bool kAns = EncLauncher::kSlot == 0;
        cfg.block = kChainCtaThreads;
        cfg.grid = grid_for(L->n_streams, 32);
        cfg.smem = 32 * (kEncRingWords + 4) * 4 + (cfg.shared ? ((size_t)model->alphabet + 1) * 16 : 0) + 128 +
                   (size_t)kChainRingSlots * 1024 * (kAns ? 16 : 8);
        ProfileScope prof(EncLauncher::kSlot, s);
        const cudaError_t e = launch_encode_chain(cfg, p, kAns);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (e != cudaSuccess) return cuda_fail(e, "encode_chain_kernel");
        return CTR_OK;
    }
    cfg.block = encode_block(L);
    cfg.grid = grid_for(L->n_streams, cfg.block);
    // rings (the range encoder: + parking slots); replicated table (128 B per entry); two symbol tiles
    cfg.smem = coder_smem_bytes(cfg.shared ? ((size_t)model->alphabet + 1) * 128 : 0, L, cfg.block / 32,
                                32 * (EncLauncher::kSlot == 0 ? kAnsEncRingWords : kEncRingWords + 4), 2);
    if (EncLauncher::kTma && !cfg.contig && !cfg.persym && make_symbol_tensor_map(&p.tmap, symbols_dev, L)) {
        p.use_tma = 1;
        cfg.smem += (size_t)(cfg.block / 32) * kEncBoxSlots * kBoxBytes;  // TMA boxes
    }
    cfg.stream = s;
    return run_coder_kernel<EncLauncher>(cfg, p);
}

template <class DecLauncher>
int decode_common(ctr_model_t model, const uint32_t *words, const uint64_t *offsets, const ctr_layout *L,
                  const uint64_t *states_in, int32_t *symbols_out, uint64_t *states_out, uint64_t *words_left,
                  uint32_t *status, void *stream, const uint64_t *ends = nullptr) {
    int rc = check_layout(L);
    if (rc) return rc;
    if (!model || !offsets || (!symbols_out && L->n_symbols)) return CTR_ERR_BAD_ARGUMENT;
    if ((L->flags & CTR_FLAG_RAW) && !states_in) return CTR_ERR_BAD_ARGUMENT;
    if (reinterpret_cast<uintptr_t>(words) % 16 != 0) return CTR_ERR_BAD_ARGUMENT;  // 16-byte vector loads
    if (ctr_device_count() == 0) return cuda_fail(cudaErrorNoDevice, "no CUDA device");
    cudaStream_t s = (cudaStream_t)stream;
    if (L->n_streams == 0) return CTR_OK;
    const bool lazy_gauss = model->lazy_means != nullptr;
    // the finer quantile index (device_utils.cuh: kBigLutBits): large batches of one shared model in the interleaved
    // layout (one 1024-thread CTA per SM), and the chain decoders
    const bool big_lut = !lazy_gauss && use_shared_tables(model, L) &&
                         !(L->flags & CTR_FLAG_CHECKPOINTS) &&
                         (L->sym_offsets_dev ? use_chain_kernels(L, false, /*raw_ok=*/true) && L->model_index_mode != CTR_INDEX_PER_SYMBOL
                                             : decode_block(L, true, false) == (unsigned)kDecBlockShared);
    if (!lazy_gauss) {
        if ((rc = ensure_dec_table(model, s))) return rc;
        if (big_lut && (rc = ensure_dec_big_table(model, s))) return rc;
        if (!use_shared_tables(model, L) && (rc = ensure_coarse_index(model, s))) return rc;
    }

    if (L->flags & CTR_FLAG_CHECKPOINTS) {
        // chunk-parallel decode: expand the records into virtual streams (one per chunk) and decode those from
        // their raw states; no per-stream outputs (a chunk is not a stream)
        if (states_in || states_out || words_left || (L->flags & CTR_FLAG_RAW)) return CTR_ERR_BAD_ARGUMENT;
        constexpr bool kRange = DecLauncher::kSlot == 3;
        const uint64_t V_max = ctr_checkpoint_max_records(L);
        struct Pool {
            void *p = nullptr;
            cudaStream_t s;
            ~Pool() {
                if (p) cudaFreeAsync(p, s);
            }
        } pool;
        pool.s = s;
        const size_t state_words = kRange ? 4 : 1;
        const size_t bytes = (V_max * (2 + state_words) + (V_max + 1)) * 8 + V_max * 4;
        keep_pool_memory();
        CUDA_TRY(cudaMallocAsync(&pool.p, bytes, s));
        uint64_t *v_begin = static_cast<uint64_t *>(pool.p);
        uint64_t *v_end = v_begin + V_max, *v_sym_off = v_end + V_max, *v_state = v_sym_off + V_max + 1;
        uint32_t *v_index = reinterpret_cast<uint32_t *>(v_state + V_max * state_words);
        const bool per_stream = L->model_index_mode == CTR_INDEX_PER_STREAM;
        expand_checkpoints_kernel<kRange><<<grid_for(L->n_streams, 128), 128, 0, s>>>(
            L->sym_offsets_dev, offsets, words, L->ckpt_offsets_dev, L->checkpoints_dev, L->checkpoint_every, L->n_streams,
            V_max, L->n_symbols, per_stream ? L->model_index_dev : nullptr, v_begin, v_end, v_sym_off, v_state,
            per_stream ? v_index : nullptr);
        LAUNCH_CHECK("expand_checkpoints_kernel");
        ctr_layout L2 = *L;
        L2.n_streams = V_max;
        L2.sym_offsets_dev = v_sym_off;
        if (per_stream) L2.model_index_dev = v_index;
        L2.flags = (L->flags & ~CTR_FLAG_CHECKPOINTS) | CTR_FLAG_RAW;
        return decode_common<DecLauncher>(model, words, v_begin, &L2, v_state, symbols_out, nullptr, nullptr, status, stream,
                                          v_end);
    }

    AnsParams p = base_params(model, L);
    p.symbols_out = symbols_out;
    p.states_in = states_in;
    p.states_out = states_out;
    p.status = status;
    p.words = words;
    p.offsets = offsets;
    p.ends = ends;
    p.words_left = words_left;

    LaunchCfg cfg;
    cfg.shared = use_shared_tables(model, L);
    cfg.contig = L->sym_offsets_dev != nullptr;
    cfg.persym = L->model_index_mode == CTR_INDEX_PER_SYMBOL;
    cfg.f64 = false;
    cfg.stream = s;
    cfg.pool = false;
    cfg.gauss = lazy_gauss;
    if (lazy_gauss) {  // gauss_kernels.cuh: the lane searches its symbol's Gaussian; no tables, no staging
        if (!cfg.persym) return CTR_ERR_BAD_ARGUMENT;
        p.gauss_means = model->lazy_means;
        p.gauss_stds = model->lazy_stds;
        p.gauss_free_weight = model->lazy_free_weight;
        cfg.shared = false;
        cfg.block = L->n_streams < kStreamsForBigCtas ? kSmallBlock : kAnsBlock;
        cfg.grid = grid_for(L->n_streams, cfg.block);
        cfg.smem = coder_smem_bytes(0, L, cfg.block / 32, 32 * kDecRingWords, 1);
        return run_coder_kernel<DecLauncher>(cfg, p);
    }
    if (use_chain_kernels(L, false, /*raw_ok=*/true) && !cfg.persym) {
        // few long streams: one coder warp per 32 streams with a straight-line loop, a second warp writes the symbols
        size_t table_bytes = 0;
        bool pool = false;
        if (cfg.shared) {
            table_bytes = dec_big_bytes(model);
        } else if (model->d_cidx) {
            const size_t cdf_bytes = align_up((size_t)model->n_models * ((size_t)model->alphabet + 1) * 4, 16);
            const size_t cidx_bytes = align_up((size_t)model->n_models * 257 * (model->alphabet > 256 ? 2 : 1), 16);
            if (cdf_bytes + cidx_bytes <= 48 * 1024) {  // (every CTA stages the set: only worth it for small sets)
                pool = true;
                table_bytes = cdf_bytes + cidx_bytes;
                p.model.pool_cdf_bytes = (uint32_t)cdf_bytes;
                p.model.pool_cidx_bytes = (uint32_t)cidx_bytes;
            }
        }
        if (cfg.shared || pool) {
            cfg.pool = pool;
            cfg.block = 64;
            cfg.grid = grid_for(L->n_streams, 32);
            cfg.smem = 32 * kDecRingWords * 4 + table_bytes + 4 * 4096;
            ProfileScope prof(DecLauncher::kSlot, s);
            const cudaError_t e = launch_decode_chain(cfg, p, DecLauncher::kSlot == 3);
            g_launches.fetch_add(1, std::memory_order_relaxed);
            if (e != cudaSuccess) return cuda_fail(e, "decode_chain_kernel");
            return CTR_OK;
        }
    }
    // interleaved deal, one model per stream: decoded symbols leave as TMA boxes (kDecBoxSlots per warp)
    size_t box_bytes_per_warp = 0;
    if (DecLauncher::kTma && tma_decode_enabled() && !big_lut && !cfg.contig && !cfg.persym && make_symbol_tensor_map(&p.tmap, symbols_out, L)) {
        p.use_tma = 1;
        box_bytes_per_warp = (size_t)kDecBoxSlots * kBoxBytes;
    }
    if (!cfg.shared && model->d_cidx) {
        // A model set that fits shared memory next to the rings and tiles is staged there once per CTA (every
        // probe of the quantile search is then a shared-memory access instead of an L1/L2 round trip).  The
        // CTA is as big as shared memory allows, but no bigger than what spreads the batch over all SMs.
        const size_t cdf_bytes = align_up((size_t)model->n_models * ((size_t)model->alphabet + 1) * 4, 16);
        const size_t cidx_bytes = align_up((size_t)model->n_models * 257 * (model->alphabet > 256 ? 2 : 1), 16);
        const size_t per_warp = coder_smem_bytes(0, L, 1, 32 * kDecRingWords, 1) + box_bytes_per_warp;
        if (cdf_bytes + cidx_bytes + per_warp + 128 <= kPoolSmemBudget) {
            const uint64_t fit = (kPoolSmemBudget - 128 - cdf_bytes - cidx_bytes) / per_warp;
            const uint64_t want = (L->n_streams + 148ull * 32 - 1) / (148ull * 32);  // warps per CTA for one wave
            const uint64_t warps = std::max<uint64_t>(1, std::min<uint64_t>(std::min<uint64_t>(fit, want), 32));
            cfg.pool = true;
            cfg.block = (unsigned)warps * 32;
            cfg.grid = grid_for(L->n_streams, cfg.block);
            cfg.smem = coder_smem_bytes(cdf_bytes + cidx_bytes, L, (int)warps, 32 * kDecRingWords, 1) + warps * box_bytes_per_warp + (box_bytes_per_warp ? 128 : 0);
            p.model.pool_cdf_bytes = (uint32_t)cdf_bytes;
            p.model.pool_cidx_bytes = (uint32_t)cidx_bytes;
            return run_coder_kernel<DecLauncher>(cfg, p);
        }
    }
    cfg.block = decode_block(L, cfg.shared, cfg.contig);
    cfg.grid = grid_for(L->n_streams, cfg.block);
    const size_t dec_table_bytes = big_lut && cfg.block == (unsigned)kDecBlockShared ? dec_big_bytes(model)
                                                                                     : (size_t)kLutBytes + model->dec_cdf_bytes;
    cfg.smem = coder_smem_bytes(cfg.shared ? dec_table_bytes : 0, L, cfg.block / 32,
                                32 * kDecRingWords, 1) + (cfg.block / 32) * box_bytes_per_warp + (box_bytes_per_warp ? 128 : 0);
    return run_coder_kernel<DecLauncher>(cfg, p);
}

}  // namespace

extern "C" int ctr_ans_encode_reverse(ctr_model_t model, const int32_t *symbols_dev, const ctr_layout *layout,
                                      const uint64_t *states_in_dev, void *workspace_dev, size_t workspace_bytes,
                                      uint32_t *words_out_dev, uint64_t words_capacity, uint64_t *offsets_out_dev,
                                      uint64_t *states_out_dev, uint32_t *status_dev, void *stream) {
    return encode_common<AnsEncodeLauncher>(model, symbols_dev, layout, states_in_dev, workspace_dev, workspace_bytes,
                                            words_out_dev, words_capacity, offsets_out_dev, states_out_dev, status_dev,
                                            stream);
}

extern "C" int ctr_ans_decode(ctr_model_t model, const uint32_t *words_dev, const uint64_t *offsets_dev,
                              const ctr_layout *layout, const uint64_t *states_in_dev, int32_t *symbols_out_dev,
                              uint64_t *states_out_dev, uint64_t *words_left_dev, uint32_t *status_dev, void *stream) {
    return decode_common<AnsDecodeLauncher>(model, words_dev, offsets_dev, layout, states_in_dev, symbols_out_dev,
                                            states_out_dev, words_left_dev, status_dev, stream);
}

// =====================================================================================================
// Range coder
// =====================================================================================================
extern "C" size_t ctr_range_encode_workspace_bytes(const ctr_layout *L) { return ctr_ans_encode_workspace_bytes(L); }
extern "C" uint64_t ctr_range_max_compressed_words(const ctr_layout *L) { return ctr_ans_max_compressed_words(L); }

extern "C" int ctr_range_encode(ctr_model_t model, const int32_t *symbols_dev, const ctr_layout *layout,
                                const uint64_t *states_in_dev, void *workspace_dev, size_t workspace_bytes,
                                uint32_t *words_out_dev, uint64_t words_capacity, uint64_t *offsets_out_dev,
                                uint64_t *states_out_dev, uint32_t *status_dev, void *stream) {
    return encode_common<RangeEncodeLauncher>(model, symbols_dev, layout, states_in_dev, workspace_dev,
                                              workspace_bytes, words_out_dev, words_capacity, offsets_out_dev,
                                              states_out_dev, status_dev, stream);
}

extern "C" int ctr_range_decode(ctr_model_t model, const uint32_t *words_dev, const uint64_t *offsets_dev,
                                const ctr_layout *layout, const uint64_t *states_in_dev, int32_t *symbols_out_dev,
                                uint64_t *states_out_dev, uint64_t *words_read_dev, uint32_t *status_dev,
                                void *stream) {
    return decode_common<RangeDecodeLauncher>(model, words_dev, offsets_dev, layout, states_in_dev, symbols_out_dev,
                                              states_out_dev, words_read_dev, status_dev, stream);
}

// =====================================================================================================
// stream memory operations (SM-free signalling between the GPUs of a node, see constriction_b200/dist.py)
// =====================================================================================================
namespace {
using StreamValue32Fn = CUresult (*)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
StreamValue32Fn stream_value_fn(const char *name) {
    void *f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return (StreamValue32Fn)f;
}
}  // namespace

extern "C" int ctr_stream_write_value32(void *addr, uint32_t value, void *stream) {
    static const StreamValue32Fn fn = stream_value_fn("cuStreamWriteValue32");
    if (!fn || !addr) return CTR_ERR_BAD_ARGUMENT;
    const CUresult r = fn((CUstream)stream, (CUdeviceptr)(uintptr_t)addr, value, CU_STREAM_WRITE_VALUE_DEFAULT);
    if (r != CUDA_SUCCESS) {
        g_last_cuda_error = "cuStreamWriteValue32 failed (" + std::to_string((int)r) + ")";
        return CTR_ERR_CUDA;
    }
    return CTR_OK;
}

extern "C" int ctr_stream_wait_value32(void *addr, uint32_t value, void *stream) {
    static const StreamValue32Fn fn = stream_value_fn("cuStreamWaitValue32");
    if (!fn || !addr) return CTR_ERR_BAD_ARGUMENT;
    const CUresult r = fn((CUstream)stream, (CUdeviceptr)(uintptr_t)addr, value, CU_STREAM_WAIT_VALUE_GEQ);
    if (r != CUDA_SUCCESS) {
        g_last_cuda_error = "cuStreamWaitValue32 failed (" + std::to_string((int)r) + ")";
        return CTR_ERR_CUDA;
    }
    return CTR_OK;
}

// =====================================================================================================
// checkpoints
// =====================================================================================================
extern "C" uint64_t ctr_checkpoint_max_records(const ctr_layout *L) {
    if (!L || L->checkpoint_every == 0) return 0;
    return L->n_symbols / L->checkpoint_every + L->n_streams;
}

extern "C" int ctr_checkpoint_offsets(const ctr_layout *L, uint64_t *out, void *stream) {
    if (!L || !out || !L->sym_offsets_dev || L->checkpoint_every == 0 || L->checkpoint_every % 32 != 0)
        return CTR_ERR_BAD_ARGUMENT;
    if (ctr_device_count() == 0) return cuda_fail(cudaErrorNoDevice, "no CUDA device");
    checkpoint_offsets_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(L->sym_offsets_dev, L->n_streams, L->checkpoint_every, out);
    LAUNCH_CHECK("checkpoint_offsets_kernel");
    return CTR_OK;
}

// =====================================================================================================
// per-symbol Gaussian parameters (no tables)
// =====================================================================================================
namespace {

struct PoolBuf {  // stream-ordered scratch allocation, released (stream-ordered) when the call returns
    void *p = nullptr;
    cudaStream_t s = nullptr;
    int alloc(size_t bytes, cudaStream_t stream) {
        s = stream;
        keep_pool_memory();
        CUDA_TRY(cudaMallocAsync(&p, bytes ? bytes : 16, stream));
        return CTR_OK;
    }
    ~PoolBuf() {
        if (p) cudaFreeAsync(p, s);
    }
};

int check_gaussian_args(int32_t min_symbol, int32_t max_symbol, const double *means, const double *stds,
                        const ctr_layout *L, double *free_weight) {
    int rc = check_layout(L);
    if (rc) return rc;
    if (L->model_index_mode != CTR_INDEX_NONE) return CTR_ERR_BAD_ARGUMENT;  // the parameters are the per-symbol model
    if ((!means || !stds) && L->n_symbols) return CTR_ERR_BAD_ARGUMENT;
    if (L->n_symbols > 0xffffffffull) return CTR_ERR_BAD_ARGUMENT;
    if (!mm::leaky_free_weight(min_symbol, max_symbol, *free_weight)) return CTR_ERR_BAD_MODEL;
    if (ctr_device_count() == 0) return cuda_fail(cudaErrorNoDevice, "no CUDA device");
    return CTR_OK;
}

// the transient model + layout the table-free paths hand to encode_common / decode_common
void lazy_model(ctr_model_s *m, ctr_layout *L2, const ctr_layout *L, int32_t min_symbol, int32_t max_symbol,
                const uint32_t *iota) {
    m->n_models = (uint32_t)(L->n_symbols ? L->n_symbols : 1);
    m->min_symbol = min_symbol;
    m->shared_ok = false;
    *L2 = *L;
    L2->model_index_dev = iota;
    L2->model_index_mode = CTR_INDEX_PER_SYMBOL;
    (void)max_symbol;
}

template <class EncLauncher>
int encode_gaussian(int32_t min_symbol, int32_t max_symbol, const double *means, const double *stds,
                    const int32_t *symbols, const ctr_layout *L, const uint64_t *states_in, void *workspace,
                    size_t workspace_bytes, uint32_t *words_out, uint64_t capacity, uint64_t *offsets_out,
                    uint64_t *states_out, uint32_t *status, void *stream) {
    double free_weight;
    int rc = check_gaussian_args(min_symbol, max_symbol, means, stds, L, &free_weight);
    if (rc) return rc;
    if (!symbols && L->n_symbols) return CTR_ERR_BAD_ARGUMENT;
    cudaStream_t s = (cudaStream_t)stream;
    const uint64_t n = L->n_symbols;
    PoolBuf entries, iota;
    if ((rc = entries.alloc(n * 16, s)) || (rc = iota.alloc(n * 4, s))) return rc;
    ctr_model_s m;
    ctr_layout L2;
    lazy_model(&m, &L2, L, min_symbol, max_symbol, static_cast<uint32_t *>(iota.p));
    m.alphabet = 0;  // every symbol maps to entry 0 of "its" model: enc[i * 1 + 0]
    m.d_enc = static_cast<uint4 *>(entries.p);
    m.enc_f64 = use_f64_division();
    if (n) {
        qgauss_entries_kernel<<<grid_for(n, 128), 128, 0, s>>>(min_symbol, max_symbol, means, stds, symbols, n,
                                                              m.enc_f64 ? 1 : 0, m.d_enc, static_cast<uint32_t *>(iota.p), status);
        LAUNCH_CHECK("qgauss_entries_kernel");
    }
    return encode_common<EncLauncher>(&m, symbols, &L2, states_in, workspace, workspace_bytes, words_out, capacity,
                                      offsets_out, states_out, status, stream);
}

template <class DecLauncher>
int decode_gaussian(int32_t min_symbol, int32_t max_symbol, const double *means, const double *stds,
                    const uint32_t *words, const uint64_t *offsets, const ctr_layout *L, const uint64_t *states_in,
                    int32_t *symbols_out, uint64_t *states_out, uint64_t *words_left, uint32_t *status, void *stream) {
    double free_weight;
    int rc = check_gaussian_args(min_symbol, max_symbol, means, stds, L, &free_weight);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const uint64_t n = L->n_symbols;
    PoolBuf iota;
    if ((rc = iota.alloc(n * 4, s))) return rc;
    if (n) {
        iota_kernel<<<grid_for(n, 256), 256, 0, s>>>(static_cast<uint32_t *>(iota.p), n);
        LAUNCH_CHECK("iota_kernel");
    }
    ctr_model_s m;
    ctr_layout L2;
    lazy_model(&m, &L2, L, min_symbol, max_symbol, static_cast<uint32_t *>(iota.p));
    m.alphabet = (uint32_t)max_symbol - (uint32_t)min_symbol + 1u;
    m.lazy_means = means ? means : reinterpret_cast<const double *>(iota.p);
    m.lazy_stds = stds;
    m.lazy_free_weight = free_weight;
    return decode_common<DecLauncher>(&m, words, offsets, &L2, states_in, symbols_out, states_out, words_left, status,
                                      stream);
}

}  // namespace

extern "C" int ctr_ans_encode_reverse_gaussian(int32_t min_symbol, int32_t max_symbol, const double *means_dev,
    const double *stds_dev, const int32_t *symbols_dev, const ctr_layout *layout, const uint64_t *states_in_dev,
    void *workspace_dev, size_t workspace_bytes, uint32_t *words_out_dev, uint64_t words_capacity,
    uint64_t *offsets_out_dev, uint64_t *states_out_dev, uint32_t *status_dev, void *stream) {
    return encode_gaussian<AnsEncodeLauncher>(min_symbol, max_symbol, means_dev, stds_dev, symbols_dev, layout,
                                              states_in_dev, workspace_dev, workspace_bytes, words_out_dev,
                                              words_capacity, offsets_out_dev, states_out_dev, status_dev, stream);
}
extern "C" int ctr_range_encode_gaussian(int32_t min_symbol, int32_t max_symbol, const double *means_dev,
    const double *stds_dev, const int32_t *symbols_dev, const ctr_layout *layout, const uint64_t *states_in_dev,
    void *workspace_dev, size_t workspace_bytes, uint32_t *words_out_dev, uint64_t words_capacity,
    uint64_t *offsets_out_dev, uint64_t *states_out_dev, uint32_t *status_dev, void *stream) {
    return encode_gaussian<RangeEncodeLauncher>(min_symbol, max_symbol, means_dev, stds_dev, symbols_dev, layout,
                                                states_in_dev, workspace_dev, workspace_bytes, words_out_dev,
                                                words_capacity, offsets_out_dev, states_out_dev, status_dev, stream);
}
extern "C" int ctr_ans_decode_gaussian(int32_t min_symbol, int32_t max_symbol, const double *means_dev,
    const double *stds_dev, const uint32_t *words_dev, const uint64_t *offsets_dev, const ctr_layout *layout,
    const uint64_t *states_in_dev, int32_t *symbols_out_dev, uint64_t *states_out_dev, uint64_t *words_left_dev,
    uint32_t *status_dev, void *stream) {
    return decode_gaussian<AnsDecodeLauncher>(min_symbol, max_symbol, means_dev, stds_dev, words_dev, offsets_dev, layout,
                                              states_in_dev, symbols_out_dev, states_out_dev, words_left_dev, status_dev,
                                              stream);
}
extern "C" int ctr_range_decode_gaussian(int32_t min_symbol, int32_t max_symbol, const double *means_dev,
    const double *stds_dev, const uint32_t *words_dev, const uint64_t *offsets_dev, const ctr_layout *layout,
    const uint64_t *states_in_dev, int32_t *symbols_out_dev, uint64_t *states_out_dev, uint64_t *words_read_dev,
    uint32_t *status_dev, void *stream) {
    return decode_gaussian<RangeDecodeLauncher>(min_symbol, max_symbol, means_dev, stds_dev, words_dev, offsets_dev,
                                                layout, states_in_dev, symbols_out_dev, states_out_dev, words_read_dev,
                                                status_dev, stream);
}

