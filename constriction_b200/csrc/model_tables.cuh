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

// Leaky quantisation of any two-parameter distribution (quantize.rs:525-568 with D = Gaussian / Laplace / Cauchy):
// thread per (model, entry).
__global__ void qdist_cdf_kernel(int kind, int32_t min_symbol, int32_t max_symbol, const double *p0, const double *p1,
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
    const double a = p0[m], b = p1[m];
    if (!(b > 0.0) || !(a == a)) {  // pybindings/stream/model.rs:654-657,745-748,845-848
        atomicOr(err, kTabBadParameter);
        return;
    }
    uint32_t v = kTotal;
    if (i == 0) v = 0u;
    else if (i < alphabet) {
        const int32_t symbol = (int32_t)((uint32_t)min_symbol + i);
        v = mm::f64_to_u32_sat(free_weight * mm::two_parameter_cdf(kind, (double)symbol - 0.5, a, b)) + i;
    }
    cdf[tid] = v;
}

// Binomial(n_m, p_m) over {0 .. n_m}, rows padded to the widest model (symbols above n_m are impossible).  One
// thread per model: the pmf follows from the mode by the recurrence pmf(i+1) = pmf(i) (n-i)/(i+1) p/(1-p).
// `scratch`: f64[n_models][alphabet].  (Parity with the reference's incomplete-beta evaluation is UNPINNED.)
__global__ void binomial_cdf_kernel(const int32_t *ns, const double *ps, uint32_t n_models, uint32_t alphabet, double *scratch,
                                    uint32_t *cdf, uint32_t *err) {
    const uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_models) return;
    const int64_t n = ns[m];
    const double p = ps[m];
    double free_weight;
    if (n < 1 || n + 1 > (int64_t)alphabet || !(p >= 0.0) || !(p <= 1.0) || !mm::leaky_free_weight(0, (int32_t)n, free_weight)) {
        atomicOr(err, kTabBadParameter);
        return;
    }
    double *pmf = scratch + m * alphabet;
    uint32_t *out = cdf + m * ((uint64_t)alphabet + 1);
    const double q = 1.0 - p;
    int64_t mode = (int64_t)(((double)n + 1.0) * p);
    if (mode > n) mode = n;
    pmf[mode] = 1.0;
    for (int64_t i = mode; i < n; ++i) pmf[i + 1] = q > 0.0 ? pmf[i] * ((double)(n - i) / (double)(i + 1)) * (p / q) : 0.0;
    for (int64_t i = mode; i > 0; --i) pmf[i - 1] = p > 0.0 ? pmf[i] * ((double)i / (double)(n - i + 1)) * (q / p) : 0.0;
    double norm = 0.0;
    for (int64_t i = 0; i <= n; ++i) norm = norm + pmf[i];
    double cum = 0.0;
    out[0] = 0u;
    for (int64_t i = 1; i <= n; ++i) {
        cum = cum + pmf[i - 1];
        out[i] = mm::f64_to_u32_sat(free_weight * (cum / norm)) + (uint32_t)i;
    }
    for (uint64_t i = (uint64_t)n + 1; i <= alphabet; ++i) out[i] = kTotal;
}

// Categorical, perfectly_quantized_probabilities (categorical.rs:56-177) + contiguous.rs:301-312: what the reference's
// Python `Categorical(p)` / `Bernoulli(p)` build by default (perfect=True).  The optimisation is sequential (a stable
// sort, then greedy exchanges of one unit of weight at a time), so one thread owns one model.  Computed in f64
// whatever the caller's type (`F: Into<f64>`).  Details that decide ties: the sort is stable and descending in
// `win`; the slots stay in sorted order afterwards; the buyer is the LAST maximum of `win`, the seller the FIRST
// minimum of `loss`.
//   scratch per model: prob f64[n], win f64[n], loss f64[n], weight u32[n], order u32[n], tmp u32[n]
struct PerfectScratch {
    double *prob, *win, *loss;
    uint32_t *weight, *order, *tmp;
};
__device__ inline void perfect_refresh(const PerfectScratch &w, uint32_t s) {
    const double weight = (double)w.weight[s];
    w.win[s] = w.prob[s] * mm::log1p_msun(1.0 / weight);
    w.loss[s] = -w.prob[s] * mm::log1p_msun(-1.0 / weight);
}
template <typename F>
__global__ void categorical_perfect_kernel(const F *pmf, uint32_t n_models, uint32_t alphabet, char *scratch, uint32_t *cdf,
                                           uint32_t *err) {
    const uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_models) return;
    const uint32_t n = alphabet;
    const F *row = pmf + m * n;
    uint32_t *out = cdf + m * ((uint64_t)n + 1);
    const uint64_t n8 = ((uint64_t)n + 1) / 2 * 2;  // keep the u32 arrays 8-byte aligned
    char *base = scratch + m * (n8 * (3 * 8 + 3 * 4));
    PerfectScratch w;
    w.prob = reinterpret_cast<double *>(base);
    w.win = w.prob + n8;
    w.loss = w.win + n8;
    w.weight = reinterpret_cast<uint32_t *>(w.loss + n8);
    w.order = w.weight + n8;
    w.tmp = w.order + n8;

    uint32_t remaining = kTotal - n;
    double norm = 0.0;
    for (uint32_t i = 0; i < n; ++i) norm = norm + (double)row[i];
    if (!(norm >= 2.2250738585072014e-308 && norm <= 1.7976931348623157e+308)) {
        atomicOr(err, kTabBadParameter);
        return;
    }
    const double scale = (double)remaining / norm;
    const double inf = mm::from_bits(0x7ff0000000000000ull);
    for (uint32_t i = 0; i < n; ++i) {
        const double prob = (double)row[i];
        if (prob < 0.0) {
            atomicOr(err, kTabBadParameter);
            return;
        }
        const uint32_t current = mm::f64_to_u32_sat(prob * scale);
        remaining -= current;
        w.prob[i] = prob;
        w.weight[i] = current + 1u;
        w.order[i] = i;
        perfect_refresh(w, i);
        if (w.weight[i] == 1u) w.loss[i] = inf;
    }
    // distribute the remaining weight among the symbols with the highest wins
    while (remaining != 0u) {
        // stable merge sort of `order`, descending in win
        for (uint32_t width = 1; width < n; width *= 2) {
            for (uint32_t lo = 0; lo < n; lo += 2 * width) {
                const uint32_t mid = min(lo + width, n), hi = min(lo + 2 * width, n);
                uint32_t i = lo, j = mid, k = lo;
                while (i < mid && j < hi) {
                    if (w.win[w.order[j]] > w.win[w.order[i]])
                        w.tmp[k++] = w.order[j++];
                    else
                        w.tmp[k++] = w.order[i++];
                }
                while (i < mid) w.tmp[k++] = w.order[i++];
                while (j < hi) w.tmp[k++] = w.order[j++];
            }
            for (uint32_t i = 0; i < n; ++i) w.order[i] = w.tmp[i];
        }
        const uint32_t batch = min(remaining, n);
        for (uint32_t i = 0; i < batch; ++i) {
            const uint32_t s = w.order[i];
            w.weight[s] += 1u;
            perfect_refresh(w, s);
        }
        remaining -= batch;
    }
    for (;;) {
        uint32_t buyer = w.order[0], seller = w.order[0];
        uint32_t buyer_pos = 0, seller_pos = 0;
        for (uint32_t pos = 1; pos < n; ++pos) {
            const uint32_t s = w.order[pos];
            if (w.win[s] >= w.win[buyer]) {
                buyer = s;
                buyer_pos = pos;
            }
            if (w.loss[s] < w.loss[seller]) {
                seller = s;
                seller_pos = pos;
            }
        }
        if (buyer_pos == seller_pos) break;
        if (w.win[buyer] <= w.loss[seller]) break;
        w.weight[seller] -= 1u;
        perfect_refresh(w, seller);
        w.win[seller] = -inf;
        if (w.weight[seller] == 1u) w.loss[seller] = inf;
        w.weight[buyer] += 1u;
        perfect_refresh(w, buyer);
        w.loss[buyer] = inf;
    }
    uint32_t acc = 0;
    for (uint32_t i = 0; i < n; ++i) {
        out[i] = acc;
        acc += w.weight[i];
    }
    out[n] = kTotal;
    if (acc != kTotal) atomicOr(err, kTabBadCdf);
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

// decoder table of model 0 (see lookup_shared in ans_kernels.cuh): quantile index uint2[2^lut_bits], then the
// CDF row with one extra 2^24 entry.
__global__ void build_dec_table_kernel(const uint32_t *cdf, uint32_t alphabet, uint32_t *dec, int lut_bits) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint2 *lut = reinterpret_cast<uint2 *>(dec);
    uint32_t *row = dec + (2u << lut_bits);
    if (tid <= alphabet + 1) row[tid] = tid <= alphabet ? cdf[tid] : kTotal;
    if (tid < (1u << lut_bits)) {
        const uint32_t q = tid << (kPrecision - lut_bits);
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
