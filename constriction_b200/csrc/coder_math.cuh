// coder_math.cuh -- per-symbol state updates of the ANS and range coders, "Default" preset
// (Word=u32, State=u64, Probability=u32, PRECISION=24), as `__host__ __device__` inline functions.
//
// The arithmetic is written so that the results equal the reference's (bamler-lab/constriction
// v0.5.0) for every input, but none of it is a translation: the u64 division of the ANS encoder
// and of the range decoder is replaced by reciprocal multiplication plus one exact fix-up, and all
// state lives in registers.  The same functions are compiled for the host by
// tests/host_math_harness.cpp and fuzzed there against the oracle.
//
// Reference behaviour being matched:
//   ANS   encode  src/stream/stack.rs:1014-1048      decode  src/stream/stack.rs:1070-1100
//   Range encode  src/stream/queue.rs:612-705        decode  src/stream/queue.rs:968-1035
//   Range seal    src/stream/queue.rs:349-376,458-523
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define CTR_HD __host__ __device__ __forceinline__
#else
#define CTR_HD inline
#endif

namespace ctr {

constexpr uint32_t kPrecision = 24;
constexpr uint32_t kTotal = 1u << kPrecision;
constexpr uint32_t kQuantileMask = kTotal - 1u;

// status codes (shared with include/constriction_b200.h)
constexpr uint32_t kOk = 0;
constexpr uint32_t kErrImpossibleSymbol = 1;  // lib.rs:376  -> KeyError
constexpr uint32_t kErrInvalidData = 2;       // queue.rs:1401 -> AssertionError
constexpr uint32_t kErrTrailingZero = 3;      // stack.rs:1555 -> ValueError
constexpr uint32_t kErrOutOfSpace = 7;        // backends.rs:1512 BoundedWriteError::OutOfSpace

CTR_HD uint64_t mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * (unsigned __int128)b) >> 64);
#endif
}

// floor((2^64-1)/d).  For every n < 2^64:  mulhi64(n, rcp) is floor(n/d) or floor(n/d)-1, because
// n*rcp/2^64 = n/d - (n/2^64)*((1+e)/d) with e = (2^64-1) mod d < d, i.e. the deficit is < 1.
CTR_HD uint64_t reciprocal_u64(uint32_t d) { return d ? ~0ull / (uint64_t)d : 0ull; }

// Exact n / d and n % d for any n (d > 0), without a divider.
CTR_HD void divmod_by_reciprocal(uint64_t n, uint32_t d, uint64_t rcp, uint64_t &quot, uint32_t &rem) {
    uint64_t q = mulhi64(n, rcp);
    // true remainder of the estimate lies in [0, 2d) and 2d < 2^33; d < 2^25 for every caller, so
    // the low 32 bits carry it exactly.
    uint32_t r = (uint32_t)n - (uint32_t)q * d;
    if (r >= d) {
        r -= d;
        q += 1;
    }
    quot = q;
    rem = r;
}

// ---------------------------------------------------------------- ANS (stack) ----------------

// One encoder table entry: what the encoder needs for one symbol of one model.
struct EncEntry {
    uint32_t left;    // left-sided cumulative, < 2^24
    uint32_t prob;    // probability, in [0, 2^24]; 0 marks an impossible symbol
    uint32_t rcp_lo;  // reciprocal_u64(prob)
    uint32_t rcp_hi;
};

// stack.rs:1035-1040: returns true (and the word to push) if the state must shed a word first.
CTR_HD bool ans_encode_needs_flush(uint64_t state, uint32_t prob) {
    return (uint32_t)(state >> (64 - kPrecision)) >= prob;  // state>>40 < 2^24, prob <= 2^24
}

// ---- the encoder's division: state / prob and state % prob without a divider --------------------------
// Two interchangeable quotient estimates, both in {q - 1, q} for every n < prob * 2^40 (the renormalised
// state), followed by the same exact correction:
//   integer : mulhi64(n, floor((2^64-1)/prob))                       (see reciprocal_u64)
//   FP64    : trunc(rz(double(n)) * ((1/prob) * (1 - 2^-50)))        3 instructions on the GPU
//             every rounding involved is below 2^-52 relative and the bias is 2^-50, so the product never
//             exceeds n/prob and falls short of it by less than 2^40 * 1.7 * 2^-50 < 1.
CTR_HD uint64_t reciprocal_f64_bits(uint32_t d) {
    if (d == 0) return 0ull;
    const double r = (1.0 / (double)d) * (1.0 - 8.8817841970012523e-16);  // 2^-50
    uint64_t b;
#if defined(__CUDA_ARCH__)
    b = (uint64_t)__double_as_longlong(r);
#else
    memcpy(&b, &r, 8);
#endif
    return b;
}

template <bool F64>
CTR_HD uint64_t ans_quotient_estimate(uint64_t n, uint32_t rcp_lo, uint32_t rcp_hi) {
    if (F64) {
#if defined(__CUDA_ARCH__)
        return __double2ull_rz(__ull2double_rz(n) * __hiloint2double((int)rcp_hi, (int)rcp_lo));
#else
        double nd = (double)n;  // round to nearest; step down if that rounded up (round toward zero)
        if (nd >= 18446744073709551616.0 || (uint64_t)nd > n) nd = nextafter(nd, 0.0);
        const uint64_t bits = ((uint64_t)rcp_hi << 32) | rcp_lo;
        double r;
        memcpy(&r, &bits, 8);
        return (uint64_t)(nd * r);
#endif
    }
    return mulhi64(n, ((uint64_t)rcp_hi << 32) | rcp_lo);
}

// stack.rs:1042-1045 on the (already renormalised) state n, given an estimate q_est in {q - 1, q}:
// new state = (q << 24) | (left + n % prob).  With r = n - q_est * prob in [0, 2 prob) (its low 32 bits carry
// it exactly since prob <= 2^24): if r >= prob the true quotient is q_est + 1 and the true remainder
// r - prob, i.e. the new state is larger by 2^24 - prob.
CTR_HD uint64_t ans_encode_recombine(uint64_t n, uint64_t q_est, uint32_t left, uint32_t prob) {
    const uint32_t r = (uint32_t)n - (uint32_t)q_est * prob;
    const uint32_t bump = r >= prob ? kTotal - prob : 0u;
    return (q_est << kPrecision) + (uint64_t)(left + r + bump);
}

CTR_HD uint64_t ans_encode_update(uint64_t state, uint32_t left, uint32_t prob, uint64_t rcp) {
    return ans_encode_recombine(state, ans_quotient_estimate<false>(state, (uint32_t)rcp, (uint32_t)(rcp >> 32)), left, prob);
}
CTR_HD uint64_t ans_encode_update_f64(uint64_t state, uint32_t left, uint32_t prob, uint64_t rcp_bits) {
    return ans_encode_recombine(state, ans_quotient_estimate<true>(state, (uint32_t)rcp_bits, (uint32_t)(rcp_bits >> 32)),
                                left, prob);
}

// stack.rs:1086
CTR_HD uint32_t ans_peek_quantile(uint64_t state) { return (uint32_t)state & kQuantileMask; }

// stack.rs:1088-1090 (without the refill)
CTR_HD uint64_t ans_decode_update(uint64_t state, uint32_t quantile, uint32_t left, uint32_t prob) {
    return (state >> kPrecision) * (uint64_t)prob + (uint64_t)(quantile - left);
}

// lib.rs:719-730 bit_array_to_chunks_truncated::<u64,u32>: number of words the state occupies.
CTR_HD uint32_t ans_state_words(uint64_t state) { return state == 0 ? 0u : ((state >> 32) ? 2u : 1u); }

// ---------------------------------------------------------------- Range (queue) --------------

struct RangeEncState {
    uint64_t lower;
    uint64_t range;         // u64::MAX when empty (queue.rs:98-106)
    uint32_t num_inverted;  // 0 == EncoderSituation::Normal
    uint32_t first_inverted;
};

CTR_HD RangeEncState range_enc_init() {
    RangeEncState s;
    s.lower = 0;
    s.range = ~0ull;
    s.num_inverted = 0;
    s.first_inverted = 0;
    return s;
}

// Result of one encode step: up to `n_burst` words become final *before* `word` (the held-back
// words of a resolved Inverted situation: `burst_first`, then n_burst-1 copies of `burst_fill`),
// then `emit` says whether `word` is appended too.
struct RangeEmit {
    uint32_t n_burst;
    uint32_t burst_first;
    uint32_t burst_fill;
    bool emit;
    uint32_t word;
};

// queue.rs:612-705.  Returns false for an impossible symbol (range would collapse to zero).
CTR_HD bool range_encode_step(RangeEncState &s, uint32_t left, uint32_t prob, RangeEmit &out) {
    out.n_burst = 0;
    out.emit = false;
    const uint64_t scale = s.range >> kPrecision;
    const uint64_t new_range = scale * (uint64_t)prob;
    if (new_range == 0) return false;
    const uint64_t new_lower = s.lower + scale * (uint64_t)left;  // wrapping
    if (s.num_inverted != 0) {
        if (new_lower + new_range > new_lower) {  // interval no longer wraps: resolve
            const bool carried = new_lower < s.lower;
            out.n_burst = s.num_inverted;
            out.burst_first = s.first_inverted + (carried ? 1u : 0u);
            out.burst_fill = carried ? 0u : 0xffffffffu;
            s.num_inverted = 0;
        }
    }
    s.lower = new_lower;
    s.range = new_range;
    if (s.range < (1ull << 32)) {
        s.range <<= 32;
        const uint32_t lower_word = (uint32_t)(s.lower >> 32);
        s.lower <<= 32;
        if (s.num_inverted != 0) {
            s.num_inverted += 1;
        } else if (s.lower + s.range > s.lower) {
            out.emit = true;
            out.word = lower_word;
        } else {
            s.num_inverted = 1;
            s.first_inverted = lower_word;
        }
    }
    return true;
}

// queue.rs:357-376
CTR_HD uint32_t range_num_seal_words(const RangeEncState &s) {
    if (s.range == ~0ull) return 0;
    const uint32_t point_word = (uint32_t)((s.lower + 0xffffffffull) >> 32);
    const uint32_t upper_word = (uint32_t)((s.lower + s.range) >> 32);
    return (upper_word == point_word ? 2u : 1u) + s.num_inverted;
}

// queue.rs:458-523: the i-th (0-based) of range_num_seal_words(s) seal words.
CTR_HD uint32_t range_seal_word(const RangeEncState &s, uint32_t i) {
    const uint64_t point = s.lower + 0xffffffffull;
    if (i < s.num_inverted) {
        const bool carried = point < s.lower;
        if (i == 0) return s.first_inverted + (carried ? 1u : 0u);
        return carried ? 0u : 0xffffffffu;
    }
    if (i == s.num_inverted) return (uint32_t)(point >> 32);
    return 0u;
}

struct RangeDecState {
    uint64_t lower;
    uint64_t range;
    uint64_t point;
};

// queue.rs:989-993: quantile = (point - lower) / (range >> 24); invalid data if it is >= 2^24.
// The quotient is estimated in double precision (the divisor `scale` < 2^40 is exact in a double,
// the dividend is rounded by at most 2^-53 relative, the quotient of interest is < 2^24, so the
// estimate is off by at most one) and then corrected exactly with integer arithmetic.
CTR_HD bool range_peek_quantile(const RangeDecState &s, uint32_t &quantile) {
    const uint64_t scale = s.range >> kPrecision;
    const uint64_t diff = s.point - s.lower;
    // diff / scale >= 2^24  <=>  diff >= scale << 24 ; (scale << 24) <= range < 2^64: no overflow
    if (diff >= (scale << kPrecision)) return false;
#if defined(__CUDA_ARCH__)
    // approximate reciprocal (relative error <= 2^-23) + one Newton step (-> ~2^-46): far more than the
    // 2^-25 that a quotient below 2^24 needs to be within one of the truth
    const double sd = __ull2double_rz(scale);
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(sd));
    r = __fma_rn(r, __fma_rn(-sd, r, 1.0), r);
    uint64_t q = (uint64_t)__double2uint_rz(__ull2double_rz(diff) * r);
#else
    const double est = (double)diff / (double)scale;
    uint64_t q = (uint64_t)est;
#endif
    if (q > kQuantileMask) q = kQuantileMask;
    // exact correction: find q with q*scale <= diff < (q+1)*scale
    uint64_t prod = q * scale;
    if (prod > diff) {
        q -= 1;
        prod -= scale;
        if (prod > diff) q -= 1;  // cannot happen (|error| <= 1); kept for safety
    } else if (diff - prod >= scale) {
        q += 1;
        prod += scale;
        if (diff - prod >= scale) q += 1;  // cannot happen
    }
    quantile = (uint32_t)q;
    return true;
}

// queue.rs:998-1032 without reading the next word: returns true if the caller must shift in a word
// (`point = point << 32 | word`, or just `point <<= 32` if the stream is exhausted).
CTR_HD bool range_decode_update(RangeDecState &s, uint32_t left, uint32_t prob) {
    const uint64_t scale = s.range >> kPrecision;
    s.lower += scale * (uint64_t)left;
    s.range = scale * (uint64_t)prob;
    if (s.range < (1ull << 32)) {
        s.lower <<= 32;
        s.range <<= 32;
        s.point <<= 32;
        return true;
    }
    return false;
}

}  // namespace ctr
