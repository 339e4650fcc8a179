// model_tables.cuh -- K5: tabulation of entropy models into 24-bit fixed-point CDF rows on the device,
// and the derived encoder / decoder tables the coder kernels consume.
//
// One model = one CDF row u32[alphabet+1] (cdf[0] = 0, cdf[alphabet] = 2^24).  The reference evaluates
// the same numbers lazily, per coded symbol (quantize.rs:525-568: two erf per encoded symbol;
// lazy_contiguous.rs:228-257: O(alphabet) float adds per symbol); here they are evaluated once per
// model entry and reused for every symbol.
#pragma once
#include "device_utils.cuh"
#include "model_math.cuh"

namespace ctr {

// error flag bits written by the tabulation / validation kernels
constexpr uint32_t kTabBadParameter = 1u;   // std <= 0, NaN, normalisation not normal / negative
constexpr uint32_t kTabBadCdf = 2u;         // cdf[0] != 0, cdf[n] != 2^24, decreasing
constexpr uint32_t kTabZeroProb = 4u;       // a symbol of a model that promises nonzero probabilities has none

// QuantizedGaussian: thread per (model, entry).  quantize.rs:525-568 evaluated for every symbol.
__global__ void qgauss_cdf_kernel(int32_t min_symbol, int32_t max_symbol, const double *means, const double *stds,
                                  uint32_t n_models, uint32_t alphabet, uint32_t *cdf, uint32_t *err) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t per = (uint64_t)alphabet + 1;
    if (tid >= per * n_models) return;
    const uint32_t m = (uint32_t)(tid / per);
    const uint32_t i = (uint32_t)(tid % per);
    double free_weight;
    if (!mm::leaky_free_weight(min_symbol, max_symbol, free_weight)) {
        atomicOr(err, kTabBadParameter);
        return;
    }
    const double mean = means[m], std = stds[m];
    if (!(std > 0.0) || !(mean == mean)) {  // pybindings/stream/model.rs:654-657
        atomicOr(err, kTabBadParameter);
        return;
    }
    cdf[tid] = (i == alphabet) ? kTotal : mm::leaky_gaussian_left(free_weight, min_symbol, mean, std, i);
}

// Categorical, fast_quantized_cdf (categorical.rs:16-54): the sums are sequential in the caller's
// float type, so one thread owns one row.
template <typename F>
__global__ void categorical_cdf_kernel(const F *pmf, uint32_t n_models, uint32_t alphabet, uint32_t *cdf,
                                       uint32_t *err) {
    const uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_models) return;
    const F *row = pmf + m * alphabet;
    uint32_t *out = cdf + m * ((uint64_t)alphabet + 1);
    F norm = (F)0;
    for (uint32_t i = 0; i < alphabet; ++i) norm = norm + row[i];
    // is_normal() && is_sign_positive(): not NaN, not zero / subnormal / negative, not infinite
    const F min_normal = sizeof(F) == 4 ? (F)1.17549435e-38f : (F)2.2250738585072014e-308;
    const F max_finite = sizeof(F) == 4 ? (F)3.40282347e+38f : (F)1.7976931348623157e+308;
    const bool normal = norm >= min_normal && norm <= max_finite;
    if (!normal) {
        atomicOr(err, kTabBadParameter);
        return;
    }
    const F scale = (F)(kTotal - alphabet) / norm;
    F cum = (F)0;
    for (uint32_t i = 0; i < alphabet; ++i) {
        uint32_t q;
        if (sizeof(F) == 4)
            q = mm::f32_to_u32_sat((float)(cum * scale));
        else
            q = mm::f64_to_u32_sat((double)(cum * scale));
        out[i] = q + i;
        cum = cum + row[i];
    }
    out[alphabet] = kTotal;
}

// Uniform (uniform.rs:44-146): every bin 2^24 / size, the last one takes the remainder.
__global__ void uniform_cdf_kernel(uint32_t size, uint32_t *cdf) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > size) return;
    cdf[i] = (i == size) ? kTotal : i * (kTotal / size);
}

// cdf rows must start at 0, end at 2^24 and never decrease; `strict` additionally demands nonzero
// probabilities (leaky quantisers guarantee them; a violation means the float CDF misbehaved).
__global__ void validate_cdf_kernel(const uint32_t *cdf, uint32_t n_models, uint32_t alphabet, int strict,
                                    uint32_t *err) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t per = (uint64_t)alphabet + 1;
    if (tid >= per * n_models) return;
    const uint32_t i = (uint32_t)(tid % per);
    const uint32_t v = cdf[tid];
    if (i == 0 && v != 0) atomicOr(err, kTabBadCdf);
    if (i == alphabet) {
        if (v != kTotal) atomicOr(err, kTabBadCdf);
        return;
    }
    const uint32_t next = cdf[tid + 1];
    if (next < v) atomicOr(err, kTabBadCdf);
    if (strict && next == v) atomicOr(err, kTabZeroProb);
}

// encoder entries {left, prob, reciprocal}: thread per (model, entry); entry [alphabet] of every model is the
// all-zero sentinel that the kernels clamp out-of-range symbols to (probability 0 = impossible symbol).
// The reciprocal is floor((2^64-1)/prob) for the integer estimate, or, with `f64`, the double
// (1/prob)*(1-2^-50), which keeps trunc(double(n) * rcp) in {floor(n/prob)-1, floor(n/prob)} for all
// n < prob * 2^40 (every rounding error involved is below 2^-52 relative, the bias is 2^-50).
__global__ void build_enc_table_kernel(const uint32_t *cdf, uint32_t n_models, uint32_t alphabet, int f64, uint4 *enc) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t per = (uint64_t)alphabet + 1;
    if (tid >= per * n_models) return;
    const uint64_t m = tid / per;
    const uint32_t s = (uint32_t)(tid % per);
    if (s == alphabet) {
        enc[tid] = make_uint4(0u, 0u, 0u, 0u);
        return;
    }
    const uint32_t *row = cdf + m * per;
    const uint32_t left = row[s], prob = row[s + 1] - row[s];
    const uint64_t rcp = f64 ? reciprocal_f64_bits(prob) : reciprocal_u64(prob);
    enc[tid] = make_uint4(left, prob, (uint32_t)rcp, (uint32_t)(rcp >> 32));
}

// coarse quantile index for decoding with global tables: cidx[m][b] = last symbol of model m whose left
// cumulative is <= b << 16 (b = 0..256; the entry for b = 256 is the last symbol); u8 or u16 entries
__global__ void build_coarse_index_kernel(const uint32_t *cdf, uint32_t n_models, uint32_t alphabet, int wide,
                                          uint8_t *cidx) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t per = 257;
    if (tid >= (uint64_t)n_models * per) return;
    const uint64_t m = tid / per;
    const uint32_t b = (uint32_t)(tid % per);
    const uint32_t *row = cdf + m * ((uint64_t)alphabet + 1);
    const uint32_t q = b << 16;  // 2^24 for b == 256: every cdf[s] with s < alphabet is <= it
    uint32_t lo = 0, hi = alphabet - 1;
    while (lo < hi) {
        const uint32_t mid = (lo + hi + 1) >> 1;
        if (row[mid] <= q)
            lo = mid;
        else
            hi = mid - 1;
    }
    if (wide)
        reinterpret_cast<uint16_t *>(cidx)[tid] = (uint16_t)lo;
    else
        cidx[tid] = (uint8_t)lo;
}

// model 0's encoder entries replicated 8 times each ([alphabet + 1][8] uint4): the ANS encoder stages this
// layout in shared memory so that lane l can read copy (l & 7), which makes its LDS.128 conflict free
__global__ void replicate_enc_table_kernel(const uint4 *enc, uint32_t alphabet, uint4 *rep) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid < (alphabet + 1) * 8u) rep[tid] = enc[tid >> 3];
}

// decoder table of model 0 (see lookup_shared in ans_kernels.cuh): quantile index uint2[kLutSize], then the
// CDF row with one extra 2^24 entry.
__global__ void build_dec_table_kernel(const uint32_t *cdf, uint32_t alphabet, uint32_t *dec) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint2 *lut = reinterpret_cast<uint2 *>(dec);
    uint32_t *row = dec + kLutBytes / 4;
    if (tid <= alphabet + 1) row[tid] = tid <= alphabet ? cdf[tid] : kTotal;
    if (tid < (uint32_t)kLutSize) {
        const uint32_t q = tid << kLutShift;
        uint32_t lo = 0, hi = alphabet - 1;  // last symbol whose left cumulative is <= q
        while (lo < hi) {
            const uint32_t mid = (lo + hi + 1) >> 1;
            if (cdf[mid] <= q)
                lo = mid;
            else
                hi = mid - 1;
        }
        lut[tid] = make_uint2(cdf[lo] | ((lo & 0xffu) << 24), cdf[lo + 1] | ((lo >> 8) << 25));
    }
}

}  // namespace ctr
