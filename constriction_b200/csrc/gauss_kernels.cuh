// gauss_kernels.cuh -- QuantizedGaussian models whose (mean, std) differ per symbol, evaluated on the device
// without tabulating a CDF row per symbol (reference: pybindings/stream/model/internals.rs:188-249
// builds one LeakilyQuantizedDistribution per symbol, quantize.rs:525-568 / 580-779 evaluate it).
//
// Encoding needs (left, probability) of the *known* symbol: two CDF evaluations per symbol, independent of the
// coder state, so they run in an embarrassingly parallel pre-pass (one thread per symbol) that writes the same
// 16-byte encoder entries {left, prob, reciprocal} the table kernels produce; the coder kernels then consume them
// through their per-symbol model-index path (model "i" = entry i).
//
// Decoding needs the inverse: the symbol whose interval contains the quantile taken from the coder state.  The
// lane that owns the stream searches the (monotone) left-cumulative function of its current symbol's Gaussian:
// an approximate inverse CDF gives the starting symbol, a galloping + bisecting search over exact evaluations of
// quantize.rs:539-547 finishes; the result is the unique symbol with left <= quantile < right, exactly what the
// reference's quantile_function (quantize.rs:580-779) returns.
//
// All floating-point code is model_math.cuh's (bit-identical to the host oracle; compiled with -fmad=false).
#pragma once
#include "device_utils.cuh"
#include "model_math.cuh"

namespace ctr {

// left-sided cumulative of symbol index i in [0, alphabet]: 0, the leaky formula, 2^24
__device__ __forceinline__ uint32_t gauss_left(double free_weight, int32_t min_symbol, uint32_t alphabet, double mean,
                                               double std, uint32_t i) {
    if (i >= alphabet) return kTotal;
    return mm::leaky_gaussian_left(free_weight, min_symbol, mean, std, i);
}

// pre-pass of the encoders: entries[i] = {left, prob, reciprocal} of symbols[i] under Gaussian(means[i], stds[i]);
// index[i] = i.  An out-of-range symbol (or a vanishing probability) gives the all-zero sentinel entry, which
// the coder kernels report as an impossible symbol; a std that is not > 0 is reported as a bad model.
static __global__ void qgauss_entries_kernel(int32_t min_symbol, int32_t max_symbol, const double *means, const double *stds,
                                      const int32_t *symbols, uint64_t n, int f64, uint4 *entries, uint32_t *index,
                                      uint32_t *status) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    index[i] = (uint32_t)i;
    double free_weight = 0.0;
    mm::leaky_free_weight(min_symbol, max_symbol, free_weight);
    const uint32_t alphabet = (uint32_t)max_symbol - (uint32_t)min_symbol + 1u;
    const double mean = means[i], std = stds[i];
    uint4 e = make_uint4(0u, 0u, 0u, 0u);
    if (!(std > 0.0)) {  // pybindings/stream/model.rs:654-657
        report_error(status, kErrBadModel, i);
    } else {
        const uint32_t s = (uint32_t)symbols[i] - (uint32_t)min_symbol;
        if (s < alphabet) {
            const uint32_t left = gauss_left(free_weight, min_symbol, alphabet, mean, std, s);
            const uint32_t right = gauss_left(free_weight, min_symbol, alphabet, mean, std, s + 1u);
            const uint32_t prob = right - left;  // quantize.rs:562-565 (wrapping)
            if (prob != 0u && prob <= kTotal) {
                const uint64_t rcp = f64 ? reciprocal_f64_bits(prob) : reciprocal_u64(prob);
                e = make_uint4(left, prob, (uint32_t)rcp, (uint32_t)(rcp >> 32));
            }
        }
    }
    entries[i] = e;
}

static __global__ void iota_kernel(uint32_t *index, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) index[i] = (uint32_t)i;
}

// The symbol index s in [0, alphabet) with left(s) <= q < left(s + 1), and those two cumulatives.
static __device__ __noinline__ uint32_t gauss_quantile(uint32_t q, double mean, double std, double free_weight, int32_t min_symbol,
                                                uint32_t alphabet, uint32_t &left, uint32_t &right) {
    // starting point: the symbol nearest to the float quantile of (q + 1/2) / 2^24 (any value would be correct)
    const float pr = ((float)q + 0.5f) * (1.0f / 16777216.0f);
    const double x = mean + std * (double)normcdfinvf(pr);
    const double rel = rint(x) - (double)min_symbol;
    uint32_t lo, hi, l_lo, l_hi;
    uint32_t start = 0;
    if (rel > 0.0) start = rel >= (double)(alphabet - 1u) ? alphabet - 1u : (uint32_t)rel;
    const uint32_t l_start = gauss_left(free_weight, min_symbol, alphabet, mean, std, start);
    uint32_t step = 1;
    if (l_start <= q) {  // gallop right until left(hi) > q   (left(alphabet) = 2^24 > q ends it)
        lo = start;
        l_lo = l_start;
        for (;;) {
            hi = alphabet - lo <= step ? alphabet : lo + step;
            l_hi = gauss_left(free_weight, min_symbol, alphabet, mean, std, hi);
            if (l_hi > q) break;
            lo = hi;
            l_lo = l_hi;
            step <<= 1;
        }
    } else {  // gallop left until left(lo) <= q   (left(0) = 0 <= q ends it)
        hi = start;
        l_hi = l_start;
        for (;;) {
            lo = hi <= step ? 0u : hi - step;
            l_lo = gauss_left(free_weight, min_symbol, alphabet, mean, std, lo);
            if (l_lo <= q) break;
            hi = lo;
            l_hi = l_lo;
            step <<= 1;
        }
    }
    while (hi - lo > 1u) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        const uint32_t l_mid = gauss_left(free_weight, min_symbol, alphabet, mean, std, mid);
        if (l_mid <= q) {
            lo = mid;
            l_lo = l_mid;
        } else {
            hi = mid;
            l_hi = l_mid;
        }
    }
    left = l_lo;
    right = l_hi;
    return lo;
}

}  // namespace ctr
