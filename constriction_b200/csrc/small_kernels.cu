// small_kernels.cu -- the reference's "Small" preset (Word = u16, State = u32, Probability = u16, PRECISION = 12:
// SmallAnsCoder stack.rs:153, SmallRangeEncoder / SmallRangeDecoder queue.rs:156,747) with TRUE lookup decoder models
// (ContiguousLookupDecoderModel / NonContiguousLookupDecoderModel, lookup_contiguous.rs:169-333,564-607,
// lookup_noncontiguous.rs:167,602-645): batched kernels, one lane per independent coder, behind ctr_small_*.
//
// A 12-bit model has 4096 quantiles, so the decoder's inverse CDF is a table with one 8-byte entry per quantile --
// {symbol index, left | probability << 16} -- which is 32 KB: for a single shared model it is staged in shared
// memory and a decoded symbol costs ONE shared-memory load (no search at all); per-stream models read their table
// through L1/L2.  The coder state is one 32-bit register.  These kernels favour simplicity over the last percent
// (per-lane 2-byte word accesses go through the caches, no word rings): the preset exists for small alphabets /
// small messages, where the launch latency dominates.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <utility>
#include <vector>

#include "../../include/constriction_b200.h"
#include "compact.cuh"
#include "device_utils.cuh"
#include "host_common.h"
#include "model_math.cuh"

using namespace ctr;

namespace {

constexpr uint32_t kSP = 12;             // PRECISION
constexpr uint32_t kSTotal = 1u << kSP;  // 4096
constexpr uint32_t kSQMask = kSTotal - 1u;
constexpr int kSBlock = 128;

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline unsigned grid_for(uint64_t n, unsigned block) { return (unsigned)((n + block - 1) / block); }

}  // namespace

struct ctr_small_model_s {
    uint32_t n_models = 0, alphabet = 0;
    int32_t min_symbol = 0;
    uint16_t *d_cdf = nullptr;    // [n_models][alphabet + 1]
    uint2 *d_enc = nullptr;       // [n_models][alphabet + 1] {left, prob}; [alphabet] = sentinel {0, 0}
    uint2 *d_lut = nullptr;       // [n_models][4096] {symbol index, left | prob << 16}
    int32_t *d_map = nullptr;     // optional: decoded symbol = map[index] (non-contiguous lookup decoder)
    int32_t *d_sorted = nullptr;  // optional: the map's symbols sorted, and the index of each
    uint32_t *d_sorted_idx = nullptr;
};

namespace {

struct SmallParams {
    const uint2 *enc, *lut;
    const int32_t *map, *sorted;
    const uint32_t *sorted_idx;
    uint32_t n_models, alphabet;
    int32_t min_symbol;
    uint64_t K, N;
    const uint64_t *sym_off;
    const uint32_t *model_index;  // per stream, or null
    const int32_t *symbols_in;
    int32_t *symbols_out;
    uint16_t *scratch;
    CompactParams compact;
    const uint16_t *words;
    const uint64_t *offsets;
    uint32_t *status;
};

// ---- model construction --------------------------------------------------------------------------------------------
// validates the rows and builds the encoder entries and the lookup tables: thread per (model, entry) / (model, quantile)
__global__ void small_build_kernel(const uint16_t *cdf, uint32_t n_models, uint32_t alphabet, uint2 *enc, uint2 *lut, uint32_t *err) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t per = (uint64_t)alphabet + 1;
    if (tid < per * n_models) {
        const uint64_t m = tid / per;
        const uint32_t s = (uint32_t)(tid % per);
        const uint16_t *row = cdf + m * per;
        if (s == 0 && row[0] != 0) atomicOr(err, 2u);
        if (s == alphabet) {
            if (row[s] != kSTotal) atomicOr(err, 2u);
            enc[tid] = make_uint2(0u, 0u);
        } else {
            if (row[s + 1] < row[s]) atomicOr(err, 2u);
            enc[tid] = make_uint2(row[s], (uint32_t)row[s + 1] - row[s]);
        }
    }
    if (tid < (uint64_t)kSTotal * n_models) {
        const uint64_t m = tid / kSTotal;
        const uint32_t q = (uint32_t)(tid % kSTotal);
        const uint16_t *row = cdf + m * per;
        uint32_t lo = 0, hi = alphabet - 1;  // last symbol whose left cumulative is <= q (lookup_contiguous.rs:297-333)
        while (lo < hi) {
            const uint32_t mid = (lo + hi + 1) >> 1;
            if (row[mid] <= q)
                lo = mid;
            else
                hi = mid - 1;
        }
        lut[tid] = make_uint2(lo, (uint32_t)row[lo] | (((uint32_t)row[lo + 1] - row[lo]) << 16));
    }
}

// fast_quantized_cdf (categorical.rs:16-54) at PRECISION 12 with u16 probabilities, one thread per model
template <typename F>
__global__ void small_categorical_kernel(const F *pmf, uint32_t n_models, uint32_t alphabet, uint16_t *cdf, uint32_t *err) {
    const uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_models) return;
    const F *row = pmf + m * alphabet;
    uint16_t *out = cdf + m * ((uint64_t)alphabet + 1);
    F norm = (F)0;
    for (uint32_t i = 0; i < alphabet; ++i) norm = norm + row[i];
    const F min_normal = sizeof(F) == 4 ? (F)1.17549435e-38f : (F)2.2250738585072014e-308;
    const F max_finite = sizeof(F) == 4 ? (F)3.40282347e+38f : (F)1.7976931348623157e+308;
    if (!(norm >= min_normal && norm <= max_finite)) {
        atomicOr(err, 1u);
        return;
    }
    const F scale = (F)(kSTotal - alphabet) / norm;
    F cum = (F)0;
    for (uint32_t i = 0; i < alphabet; ++i) {
        const F v = cum * scale;
        const uint32_t q = !(v > (F)0) ? 0u : (v >= (F)65535 ? 65535u : (uint32_t)v);  // Rust `as u16`
        out[i] = (uint16_t)(q + i);
        cum = cum + row[i];
    }
    out[alphabet] = (uint16_t)kSTotal;
}

// perfectly_quantized_probabilities (categorical.rs:56-177) at PRECISION 12: see categorical_perfect_kernel in
// model_tables.cuh for the order-sensitive details; alphabets are at most 4096 here, so the scratch is small
template <typename F>
__global__ void small_categorical_perfect_kernel(const F *pmf, uint32_t n_models, uint32_t alphabet, char *scratch, uint16_t *cdf,
                                                 uint32_t *err) {
    const uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_models) return;
    const uint32_t n = alphabet;
    const F *row = pmf + m * n;
    uint16_t *out = cdf + m * ((uint64_t)n + 1);
    const uint64_t n8 = ((uint64_t)n + 1) / 2 * 2;
    char *base = scratch + m * (n8 * (3 * 8 + 3 * 4));
    double *prob = reinterpret_cast<double *>(base), *win = prob + n8, *loss = win + n8;
    uint32_t *weight = reinterpret_cast<uint32_t *>(loss + n8), *order = weight + n8, *tmp = order + n8;
    auto refresh = [&](uint32_t s) {
        const double w = (double)weight[s];
        win[s] = prob[s] * mm::log1p_msun(1.0 / w);
        loss[s] = -prob[s] * mm::log1p_msun(-1.0 / w);
    };
    uint32_t remaining = kSTotal - n;
    double norm = 0.0;
    for (uint32_t i = 0; i < n; ++i) norm = norm + (double)row[i];
    if (!(norm >= 2.2250738585072014e-308 && norm <= 1.7976931348623157e+308)) {
        atomicOr(err, 1u);
        return;
    }
    const double scale = (double)remaining / norm;
    const double inf = mm::from_bits(0x7ff0000000000000ull);
    for (uint32_t i = 0; i < n; ++i) {
        const double p = (double)row[i];
        if (p < 0.0) {
            atomicOr(err, 1u);
            return;
        }
        const double v = p * scale;
        const uint32_t current = !(v > 0.0) ? 0u : (v >= 65535.0 ? 65535u : (uint32_t)v);
        remaining -= current;
        prob[i] = p;
        weight[i] = current + 1u;
        order[i] = i;
        refresh(i);
        if (weight[i] == 1u) loss[i] = inf;
    }
    while (remaining != 0u) {
        for (uint32_t width = 1; width < n; width *= 2) {
            for (uint32_t lo = 0; lo < n; lo += 2 * width) {
                const uint32_t mid = min(lo + width, n), hi = min(lo + 2 * width, n);
                uint32_t i = lo, j = mid, k = lo;
                while (i < mid && j < hi) tmp[k++] = win[order[j]] > win[order[i]] ? order[j++] : order[i++];
                while (i < mid) tmp[k++] = order[i++];
                while (j < hi) tmp[k++] = order[j++];
            }
            for (uint32_t i = 0; i < n; ++i) order[i] = tmp[i];
        }
        const uint32_t batch = min(remaining, n);
        for (uint32_t i = 0; i < batch; ++i) {
            weight[order[i]] += 1u;
            refresh(order[i]);
        }
        remaining -= batch;
    }
    for (;;) {
        uint32_t buyer = order[0], seller = order[0], buyer_pos = 0, seller_pos = 0;
        for (uint32_t pos = 1; pos < n; ++pos) {
            const uint32_t s = order[pos];
            if (win[s] >= win[buyer]) {
                buyer = s;
                buyer_pos = pos;
            }
            if (loss[s] < loss[seller]) {
                seller = s;
                seller_pos = pos;
            }
        }
        if (buyer_pos == seller_pos || win[buyer] <= loss[seller]) break;
        weight[seller] -= 1u;
        refresh(seller);
        win[seller] = -inf;
        if (weight[seller] == 1u) loss[seller] = inf;
        weight[buyer] += 1u;
        refresh(buyer);
        loss[buyer] = inf;
    }
    uint32_t acc = 0;
    for (uint32_t i = 0; i < n; ++i) {
        out[i] = (uint16_t)acc;
        acc += weight[i];
    }
    out[n] = (uint16_t)kSTotal;
    if (acc != kSTotal) atomicOr(err, 2u);
}

// ---- geometry --------------------------------------------------------------------------------------------------------
// scratch region of stream k whose symbols start at stream-major offset o: n_k + 4 words (a 12-bit model emits at
// most 12 bits = 0.75 words per symbol; 2 state / seal words)
__device__ __forceinline__ uint64_t small_scratch_start(uint64_t o, uint64_t k) { return o + 4 * k; }

struct Geometry {
    uint64_t n_k, o_k, stride;  // symbols of my stream, index of the first, distance between consecutive ones
    const int32_t *in;
    int32_t *out;
};
__device__ __forceinline__ Geometry geometry_of(const SmallParams &p, uint64_t k, bool valid, uint32_t *status) {
    Geometry g;
    if (p.sym_off) {
        uint64_t o = valid ? p.sym_off[k] : 0, n = valid ? p.sym_off[k + 1] - o : 0;
        if (o > p.N || n > p.N - o) {
            if (valid) report_error(status, kErrBadArgument, k);
            o = 0;
            n = 0;
        }
        g.n_k = n;
        g.o_k = o;
        g.stride = 1;
        g.in = p.symbols_in ? p.symbols_in + o : nullptr;
        g.out = p.symbols_out ? p.symbols_out + o : nullptr;
    } else {
        g.n_k = valid ? interleaved_len(p.N, p.K, k) : 0;
        g.o_k = valid ? interleaved_start(p.N, p.K, k) : 0;
        g.stride = p.K;
        g.in = p.symbols_in ? p.symbols_in + k : nullptr;
        g.out = p.symbols_out ? p.symbols_out + k : nullptr;
    }
    return g;
}

// symbol -> table index: contiguous alphabets subtract min_symbol; non-contiguous ones search the sorted symbols
__device__ __forceinline__ uint32_t small_index_of(const SmallParams &p, int32_t sym) {
    if (!p.sorted) return min((uint32_t)sym - (uint32_t)p.min_symbol, p.alphabet);
    uint32_t lo = 0, hi = p.alphabet;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (p.sorted[mid] < sym)
            lo = mid + 1;
        else
            hi = mid;
    }
    return (lo < p.alphabet && p.sorted[lo] == sym) ? p.sorted_idx[lo] : p.alphabet;
}

// ---- ANS --------------------------------------------------------------------------------------------------------------
// SmallAnsCoder::encode_symbol under encode_iid_symbols_reverse + into_compressed (stack.rs:1014-1048,835-849,891-895)
__global__ void __launch_bounds__(kSBlock) small_ans_encode_kernel(const SmallParams p) {
    const uint32_t tile = take_tile_ticket(p.compact.ticket);
    const uint64_t k = (uint64_t)tile * kSBlock + threadIdx.x;
    const bool valid = k < p.K;
    const Geometry g = geometry_of(p, k, valid, p.status);
    uint16_t *const begin = p.scratch + small_scratch_start(g.o_k, valid ? k : 0);
    uint16_t *w = begin;
    uint32_t m = (valid && p.model_index) ? p.model_index[k] : 0u;
    if (m >= p.n_models) m = 0;
    const uint2 *enc = p.enc + (uint64_t)m * (p.alphabet + 1);
    uint32_t state = 0;
    bool impossible = false;
    for (uint64_t i = g.n_k; i-- > 0;) {
        const uint2 e = __ldg(enc + small_index_of(p, ld_stream_s32(g.in + i * g.stride)));
        if (e.y == 0u) {
            impossible = true;
            break;
        }
        if ((state >> (32 - kSP)) >= e.y) {
            *w++ = (uint16_t)state;
            state >>= 16;
        }
        state = ((state / e.y) << kSP) | (e.x + state % e.y);
    }
    if (valid && !impossible) {  // lib.rs:719-730: the state's words, least significant first, leading zero words dropped
        if (state != 0u) *w++ = (uint16_t)state;
        if ((state >> 16) != 0u) *w++ = (uint16_t)(state >> 16);
    }
    if (valid && impossible) report_error(p.status, kErrImpossibleSymbol, k);
    compact_tail<kSBlock, uint16_t>(p.compact, tile, k, p.K, valid, begin, (valid && !impossible) ? (uint32_t)(w - begin) : 0u, p.status);
}

// from_compressed (stack.rs:299-318,440-462) + decode_symbol with a lookup decoder model (stack.rs:1070-1100,
// lookup_contiguous.rs:564-607)
template <bool SHARED>
__global__ void __launch_bounds__(kSBlock) small_ans_decode_kernel(const SmallParams p) {
    extern __shared__ __align__(16) uint2 s_lut[];
    if (SHARED) {
        for (uint32_t i = threadIdx.x; i < kSTotal; i += kSBlock) s_lut[i] = p.lut[i];
        __syncthreads();
    }
    const uint64_t k = (uint64_t)blockIdx.x * kSBlock + threadIdx.x;
    if (k >= p.K) return;
    const Geometry g = geometry_of(p, k, true, p.status);
    const uint16_t *const begin = p.words + p.offsets[k];
    const uint16_t *r = p.words + p.offsets[k + 1];  // pop from the end
    uint32_t m = p.model_index ? p.model_index[k] : 0u;
    if (m >= p.n_models) m = 0;
    const uint2 *lut = p.lut + (uint64_t)m * kSTotal;
    uint32_t state = 0;
    if (r != begin) {
        const uint32_t first = *--r;
        if (first == 0u) report_error(p.status, kErrTrailingZero, k);
        state = first;
        if (r != begin && first != 0u) state = (state << 16) | *--r;
    }
    for (uint64_t i = 0; i < g.n_k; ++i) {
        const uint32_t q = state & kSQMask;
        const uint2 e = SHARED ? s_lut[q] : __ldg(lut + q);
        const uint32_t left = e.y & 0xffffu, prob = e.y >> 16;
        state = (state >> kSP) * prob + (q - left);
        if (state < (1u << 16) && r != begin) state = (state << 16) | *--r;
        st_stream_s32(g.out + i * g.stride, p.map ? p.map[e.x] : (int32_t)((uint32_t)p.min_symbol + e.x));
    }
}

// ---- range coder ---------------------------------------------------------------------------------------------------
// SmallRangeEncoder::encode_symbol + seal (queue.rs:612-705,349-376,458-523), "eager words, late carry" as in
// coder_math.cuh: a word is written at every renormalisation and a later wrap of `lower` adds one to the words
// already written (the trailing 0xffff words wrap to zero, the word before them absorbs the carry)
__global__ void __launch_bounds__(kSBlock) small_range_encode_kernel(const SmallParams p) {
    const uint32_t tile = take_tile_ticket(p.compact.ticket);
    const uint64_t k = (uint64_t)tile * kSBlock + threadIdx.x;
    const bool valid = k < p.K;
    const Geometry g = geometry_of(p, k, valid, p.status);
    uint16_t *const begin = p.scratch + small_scratch_start(g.o_k, valid ? k : 0);
    uint16_t *w = begin;
    uint32_t m = (valid && p.model_index) ? p.model_index[k] : 0u;
    if (m >= p.n_models) m = 0;
    const uint2 *enc = p.enc + (uint64_t)m * (p.alphabet + 1);
    uint32_t lower = 0, range = 0xffffffffu;
    bool impossible = false;
    auto carry = [&]() {
        for (uint16_t *c = w; c != begin;) {
            const uint16_t v = (uint16_t)(*--c + 1u);
            *c = v;
            if (v != 0u) break;
        }
    };
    for (uint64_t i = 0; i < g.n_k; ++i) {
        const uint2 e = __ldg(enc + small_index_of(p, ld_stream_s32(g.in + i * g.stride)));
        if (e.y == 0u) {
            impossible = true;
            break;
        }
        const uint32_t scale = range >> kSP;
        const uint32_t nl = lower + scale * e.x;
        range = scale * e.y;
        if (nl < lower) carry();
        lower = nl;
        if (range < (1u << 16)) {
            *w++ = (uint16_t)(lower >> 16);
            lower <<= 16;
            range <<= 16;
        }
    }
    if (valid && !impossible && range != 0xffffffffu) {  // seal
        const uint32_t point = lower + 0xffffu;
        if (point < lower) carry();
        const uint16_t point_word = (uint16_t)(point >> 16);
        *w++ = point_word;
        if ((uint16_t)((lower + range) >> 16) == point_word) *w++ = 0;
    }
    if (valid && impossible) report_error(p.status, kErrImpossibleSymbol, k);
    compact_tail<kSBlock, uint16_t>(p.compact, tile, k, p.K, valid, begin, (valid && !impossible) ? (uint32_t)(w - begin) : 0u, p.status);
}

// SmallRangeDecoder::from_compressed / read_point / decode_symbol (queue.rs:755-773,847-868,968-1035)
template <bool SHARED>
__global__ void __launch_bounds__(kSBlock) small_range_decode_kernel(const SmallParams p) {
    extern __shared__ __align__(16) uint2 s_lut[];
    if (SHARED) {
        for (uint32_t i = threadIdx.x; i < kSTotal; i += kSBlock) s_lut[i] = p.lut[i];
        __syncthreads();
    }
    const uint64_t k = (uint64_t)blockIdx.x * kSBlock + threadIdx.x;
    if (k >= p.K) return;
    const Geometry g = geometry_of(p, k, true, p.status);
    const uint16_t *r = p.words + p.offsets[k];
    const uint16_t *const end = p.words + p.offsets[k + 1];
    uint32_t m = p.model_index ? p.model_index[k] : 0u;
    if (m >= p.n_models) m = 0;
    const uint2 *lut = p.lut + (uint64_t)m * kSTotal;
    uint32_t lower = 0, range = 0xffffffffu, point = 0;
    if (r != end) point = (uint32_t)*r++ << 16;
    if (r != end) point |= *r++;
    bool invalid = false;
    for (uint64_t i = 0; i < g.n_k; ++i) {
        const uint32_t scale = range >> kSP;
        uint32_t q = (point - lower) / scale;
        if (q >= kSTotal) {
            invalid = true;
            q = kSQMask;
        }
        const uint2 e = SHARED ? s_lut[q] : __ldg(lut + q);
        lower += scale * (e.y & 0xffffu);
        range = scale * (e.y >> 16);
        if (range < (1u << 16)) {
            lower <<= 16;
            range <<= 16;
            point <<= 16;
            if (r != end) point |= *r++;
        }
        st_stream_s32(g.out + i * g.stride, p.map ? p.map[e.x] : (int32_t)((uint32_t)p.min_symbol + e.x));
    }
    if (invalid) report_error(p.status, kErrInvalidData, k);
}

// ---- host side --------------------------------------------------------------------------------------------------------
int small_finish(ctr_small_model_s *m, uint32_t *d_err, cudaStream_t s) {
    const uint64_t threads = std::max<uint64_t>((uint64_t)m->n_models * ((uint64_t)m->alphabet + 1), (uint64_t)m->n_models * kSTotal);
    CTR_HOST_TRY(cudaMalloc(&m->d_enc, (size_t)m->n_models * ((size_t)m->alphabet + 1) * 8));
    CTR_HOST_TRY(cudaMalloc(&m->d_lut, (size_t)m->n_models * kSTotal * 8));
    small_build_kernel<<<grid_for(threads, 256), 256, 0, s>>>(m->d_cdf, m->n_models, m->alphabet, m->d_enc, m->d_lut, d_err);
    ctr::host_count_launch();
    uint32_t h_err = 0;
    CTR_HOST_TRY(cudaMemcpyAsync(&h_err, d_err, 4, cudaMemcpyDeviceToHost, s));
    CTR_HOST_TRY(cudaStreamSynchronize(s));
    return h_err ? CTR_ERR_BAD_MODEL : CTR_OK;
}

int small_alloc(uint32_t n_models, uint32_t alphabet, int32_t min_symbol, ctr_small_model_s **out) {
    if (n_models == 0 || alphabet < 2 || alphabet > kSTotal) return CTR_ERR_BAD_MODEL;
    ctr_small_model_s *m = new (std::nothrow) ctr_small_model_s();
    if (!m) return CTR_ERR_BAD_ARGUMENT;
    m->n_models = n_models;
    m->alphabet = alphabet;
    m->min_symbol = min_symbol;
    const cudaError_t e = cudaMalloc(&m->d_cdf, align_up((size_t)n_models * ((size_t)alphabet + 1) * 2, 16));
    if (e != cudaSuccess) {
        delete m;
        return ctr::host_cuda_fail(e, "cudaMalloc(small cdf)");
    }
    *out = m;
    return CTR_OK;
}

struct DevWord {
    uint32_t *d = nullptr;
    int init(cudaStream_t s) {
        CTR_HOST_TRY(cudaMalloc(&d, 4));
        CTR_HOST_TRY(cudaMemsetAsync(d, 0, 4, s));
        return CTR_OK;
    }
    ~DevWord() {
        if (d) cudaFree(d);
    }
};

template <typename F>
int small_categorical(const F *pmf, int is_device, uint32_t n_models, uint32_t alphabet, int perfect, void *stream,
                      ctr_small_model_t *out) {
    if (!out || !pmf || n_models == 0) return CTR_ERR_BAD_ARGUMENT;
    if (alphabet < 2 || (perfect ? alphabet > kSTotal : alphabet >= kSTotal - 1)) return CTR_ERR_BAD_MODEL;
    if (ctr_device_count() == 0) return ctr::host_cuda_fail(cudaErrorNoDevice, "no CUDA device");
    cudaStream_t s = (cudaStream_t)stream;
    ctr_small_model_s *m = nullptr;
    int rc = small_alloc(n_models, alphabet, 0, &m);
    if (rc) return rc;
    F *d_pmf = nullptr;
    char *d_scratch = nullptr;
    DevWord err;
    auto cleanup = [&](int code) {
        if (d_pmf) cudaFree(d_pmf);
        if (d_scratch) cudaFree(d_scratch);
        if (code) ctr_small_model_destroy(m);
        return code;
    };
    if ((rc = err.init(s))) return cleanup(rc);
    const F *src = pmf;
    if (!is_device) {
        const size_t bytes = (size_t)n_models * alphabet * sizeof(F);
        if (cudaMalloc(&d_pmf, bytes) != cudaSuccess || cudaMemcpyAsync(d_pmf, pmf, bytes, cudaMemcpyHostToDevice, s) != cudaSuccess)
            return cleanup(ctr::host_cuda_fail(cudaGetLastError(), "small model: pmf upload"));
        src = d_pmf;
    }
    if (perfect) {
        const uint64_t n8 = ((uint64_t)alphabet + 1) / 2 * 2;
        if (cudaMalloc(&d_scratch, (size_t)n_models * n8 * (3 * 8 + 3 * 4)) != cudaSuccess)
            return cleanup(ctr::host_cuda_fail(cudaGetLastError(), "cudaMalloc(perfect scratch)"));
        small_categorical_perfect_kernel<F><<<grid_for(n_models, 32), 32, 0, s>>>(src, n_models, alphabet, d_scratch, m->d_cdf, err.d);
    } else {
        small_categorical_kernel<F><<<grid_for(n_models, 64), 64, 0, s>>>(src, n_models, alphabet, m->d_cdf, err.d);
    }
    ctr::host_count_launch();
    if (cudaGetLastError() != cudaSuccess) return cleanup(ctr::host_cuda_fail(cudaGetLastError(), "small categorical kernel"));
    if ((rc = small_finish(m, err.d, s))) return cleanup(rc);
    *out = m;
    return cleanup(CTR_OK);
}

SmallParams small_params(const ctr_small_model_s *m, const ctr_layout *L) {
    SmallParams p;
    memset(&p, 0, sizeof p);
    p.enc = m->d_enc;
    p.lut = m->d_lut;
    p.map = m->d_map;
    p.sorted = m->d_sorted;
    p.sorted_idx = m->d_sorted_idx;
    p.n_models = m->n_models;
    p.alphabet = m->alphabet;
    p.min_symbol = m->min_symbol;
    p.K = L->n_streams;
    p.N = L->n_symbols;
    p.sym_off = L->sym_offsets_dev;
    p.model_index = L->model_index_mode == CTR_INDEX_PER_STREAM ? L->model_index_dev : nullptr;
    return p;
}

int small_check(const ctr_small_model_s *m, const ctr_layout *L) {
    if (!m || !L) return CTR_ERR_BAD_ARGUMENT;
    if (L->flags != 0u || L->model_index_mode == CTR_INDEX_PER_SYMBOL) return CTR_ERR_BAD_ARGUMENT;  // shared or per-stream models
    if (L->model_index_mode == CTR_INDEX_PER_STREAM && !L->model_index_dev) return CTR_ERR_BAD_ARGUMENT;
    if (L->n_streams == 0 && L->n_symbols != 0) return CTR_ERR_BAD_ARGUMENT;
    if (ctr_device_count() == 0) return ctr::host_cuda_fail(cudaErrorNoDevice, "no CUDA device");
    return CTR_OK;
}

struct SmallWorkspace {
    size_t status_off, ticket_off, total;
    uint64_t n_tiles;
};
SmallWorkspace small_workspace(const ctr_layout *L) {
    SmallWorkspace w;
    w.n_tiles = (L->n_streams + kSBlock - 1) / kSBlock;
    w.status_off = align_up((size_t)(L->n_symbols + 4 * L->n_streams + 64) * 2, 256);
    w.ticket_off = w.status_off + (size_t)w.n_tiles * 8;
    w.total = align_up(w.ticket_off + 8, 256);
    return w;
}

template <bool RANGE>
int small_encode(ctr_small_model_t model, const int32_t *symbols, const ctr_layout *L, void *workspace, size_t workspace_bytes,
                 uint16_t *words_out, uint64_t capacity, uint64_t *offsets_out, uint32_t *status, void *stream) {
    int rc = small_check(model, L);
    if (rc) return rc;
    if (!offsets_out || (!symbols && L->n_symbols)) return CTR_ERR_BAD_ARGUMENT;
    cudaStream_t s = (cudaStream_t)stream;
    if (L->n_streams == 0) {
        CTR_HOST_TRY(cudaMemsetAsync(offsets_out, 0, 8, s));
        return CTR_OK;
    }
    const SmallWorkspace w = small_workspace(L);
    if (!workspace || workspace_bytes < w.total || !words_out) return CTR_ERR_BAD_ARGUMENT;
    SmallParams p = small_params(model, L);
    char *ws = static_cast<char *>(workspace);
    p.symbols_in = symbols;
    p.scratch = reinterpret_cast<uint16_t *>(ws);
    p.compact.tile_status = reinterpret_cast<uint64_t *>(ws + w.status_off);
    p.compact.ticket = reinterpret_cast<unsigned int *>(ws + w.ticket_off);
    p.compact.words_out = reinterpret_cast<uint32_t *>(words_out);
    p.compact.words_capacity = capacity;
    p.compact.offsets_out = offsets_out;
    p.status = status;
    CTR_HOST_TRY(cudaMemsetAsync(ws + w.status_off, 0, w.total - w.status_off, s));
    if (RANGE)
        small_range_encode_kernel<<<grid_for(L->n_streams, kSBlock), kSBlock, 0, s>>>(p);
    else
        small_ans_encode_kernel<<<grid_for(L->n_streams, kSBlock), kSBlock, 0, s>>>(p);
    ctr::host_count_launch();
    CTR_HOST_TRY(cudaGetLastError());
    return CTR_OK;
}

template <bool RANGE>
int small_decode(ctr_small_model_t model, const uint16_t *words, const uint64_t *offsets, const ctr_layout *L, int32_t *symbols_out,
                 uint32_t *status, void *stream) {
    int rc = small_check(model, L);
    if (rc) return rc;
    if (!offsets || (!symbols_out && L->n_symbols)) return CTR_ERR_BAD_ARGUMENT;
    cudaStream_t s = (cudaStream_t)stream;
    if (L->n_streams == 0) return CTR_OK;
    SmallParams p = small_params(model, L);
    p.symbols_out = symbols_out;
    p.words = words;
    p.offsets = offsets;
    p.status = status;
    const bool shared = model->n_models == 1;
    const unsigned grid = grid_for(L->n_streams, kSBlock);
    if (RANGE) {
        if (shared)
            small_range_decode_kernel<true><<<grid, kSBlock, kSTotal * 8, s>>>(p);
        else
            small_range_decode_kernel<false><<<grid, kSBlock, 0, s>>>(p);
    } else {
        if (shared)
            small_ans_decode_kernel<true><<<grid, kSBlock, kSTotal * 8, s>>>(p);
        else
            small_ans_decode_kernel<false><<<grid, kSBlock, 0, s>>>(p);
    }
    ctr::host_count_launch();
    CTR_HOST_TRY(cudaGetLastError());
    return CTR_OK;
}

}  // namespace

extern "C" int ctr_small_model_from_cdf(const uint16_t *cdf, int is_device, uint32_t n_models, uint32_t alphabet, int32_t min_symbol,
                                        const int32_t *symbols_host, void *stream, ctr_small_model_t *out) {
    if (!out || !cdf || n_models == 0) return CTR_ERR_BAD_ARGUMENT;
    if (ctr_device_count() == 0) return ctr::host_cuda_fail(cudaErrorNoDevice, "no CUDA device");
    cudaStream_t s = (cudaStream_t)stream;
    ctr_small_model_s *m = nullptr;
    int rc = small_alloc(n_models, alphabet, min_symbol, &m);
    if (rc) return rc;
    DevWord err;
    auto fail = [&](int code) {
        ctr_small_model_destroy(m);
        return code;
    };
    if ((rc = err.init(s))) return fail(rc);
    if (cudaMemcpyAsync(m->d_cdf, cdf, (size_t)n_models * ((size_t)alphabet + 1) * 2, is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                        s) != cudaSuccess)
        return fail(ctr::host_cuda_fail(cudaGetLastError(), "cudaMemcpyAsync(small cdf)"));
    if (symbols_host) {  // non-contiguous alphabet: index i stands for symbols_host[i] (lookup_noncontiguous.rs:167,602-645)
        std::vector<std::pair<int32_t, uint32_t>> order(alphabet);
        for (uint32_t i = 0; i < alphabet; ++i) order[i] = {symbols_host[i], i};
        std::sort(order.begin(), order.end());
        for (uint32_t i = 1; i < alphabet; ++i)
            if (order[i].first == order[i - 1].first) return fail(CTR_ERR_BAD_MODEL);  // duplicate symbol
        std::vector<int32_t> sorted(alphabet);
        std::vector<uint32_t> idx(alphabet);
        for (uint32_t i = 0; i < alphabet; ++i) {
            sorted[i] = order[i].first;
            idx[i] = order[i].second;
        }
        if (cudaMalloc(&m->d_map, alphabet * 4) != cudaSuccess || cudaMalloc(&m->d_sorted, alphabet * 4) != cudaSuccess ||
            cudaMalloc(&m->d_sorted_idx, alphabet * 4) != cudaSuccess)
            return fail(ctr::host_cuda_fail(cudaGetLastError(), "cudaMalloc(symbol map)"));
        cudaMemcpy(m->d_map, symbols_host, alphabet * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(m->d_sorted, sorted.data(), alphabet * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(m->d_sorted_idx, idx.data(), alphabet * 4, cudaMemcpyHostToDevice);
    }
    if ((rc = small_finish(m, err.d, s))) return fail(rc);
    *out = m;
    return CTR_OK;
}

extern "C" int ctr_small_model_categorical_f32(const float *pmf, int is_device, uint32_t n_models, uint32_t alphabet, int perfect,
                                               void *stream, ctr_small_model_t *out) {
    return small_categorical<float>(pmf, is_device, n_models, alphabet, perfect, stream, out);
}
extern "C" int ctr_small_model_categorical_f64(const double *pmf, int is_device, uint32_t n_models, uint32_t alphabet, int perfect,
                                               void *stream, ctr_small_model_t *out) {
    return small_categorical<double>(pmf, is_device, n_models, alphabet, perfect, stream, out);
}

extern "C" int ctr_small_model_destroy(ctr_small_model_t m) {
    if (!m) return CTR_OK;
    for (void *p : {(void *)m->d_cdf, (void *)m->d_enc, (void *)m->d_lut, (void *)m->d_map, (void *)m->d_sorted, (void *)m->d_sorted_idx})
        if (p) cudaFree(p);
    delete m;
    return CTR_OK;
}

extern "C" int ctr_small_model_copy_cdf_host(ctr_small_model_t m, uint16_t *cdf_host, void *stream) {
    if (!m || !cdf_host) return CTR_ERR_BAD_ARGUMENT;
    CTR_HOST_TRY(cudaMemcpyAsync(cdf_host, m->d_cdf, (size_t)m->n_models * ((size_t)m->alphabet + 1) * 2, cudaMemcpyDeviceToHost,
                                 (cudaStream_t)stream));
    CTR_HOST_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return CTR_OK;
}

extern "C" size_t ctr_small_encode_workspace_bytes(const ctr_layout *L) { return L ? small_workspace(L).total : 0; }
extern "C" uint64_t ctr_small_max_compressed_words(const ctr_layout *L) { return L ? L->n_symbols + 4 * L->n_streams + 64 : 0; }

extern "C" int ctr_small_ans_encode_reverse(ctr_small_model_t model, const int32_t *symbols_dev, const ctr_layout *layout,
                                            void *workspace_dev, size_t workspace_bytes, uint16_t *words_out_dev, uint64_t words_capacity,
                                            uint64_t *offsets_out_dev, uint32_t *status_dev, void *stream) {
    return small_encode<false>(model, symbols_dev, layout, workspace_dev, workspace_bytes, words_out_dev, words_capacity, offsets_out_dev,
                               status_dev, stream);
}
extern "C" int ctr_small_range_encode(ctr_small_model_t model, const int32_t *symbols_dev, const ctr_layout *layout, void *workspace_dev,
                                      size_t workspace_bytes, uint16_t *words_out_dev, uint64_t words_capacity, uint64_t *offsets_out_dev,
                                      uint32_t *status_dev, void *stream) {
    return small_encode<true>(model, symbols_dev, layout, workspace_dev, workspace_bytes, words_out_dev, words_capacity, offsets_out_dev,
                              status_dev, stream);
}
extern "C" int ctr_small_ans_decode(ctr_small_model_t model, const uint16_t *words_dev, const uint64_t *offsets_dev, const ctr_layout *layout,
                                    int32_t *symbols_out_dev, uint32_t *status_dev, void *stream) {
    return small_decode<false>(model, words_dev, offsets_dev, layout, symbols_out_dev, status_dev, stream);
}
extern "C" int ctr_small_range_decode(ctr_small_model_t model, const uint16_t *words_dev, const uint64_t *offsets_dev,
                                      const ctr_layout *layout, int32_t *symbols_out_dev, uint32_t *status_dev, void *stream) {
    return small_decode<true>(model, words_dev, offsets_dev, layout, symbols_out_dev, status_dev, stream);
}
