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
constexpr uint32_t kErrBadModel = 5;          // invalid model parameter (std <= 0)          -> ValueError
constexpr uint32_t kErrOutOfSpace = 7;        // backends.rs:1512 BoundedWriteError::OutOfSpace
constexpr uint32_t kErrBadArgument = 8;       // sym_offsets that do not describe slices of the symbol array -> ValueError

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

// ---- range encoder, "eager words, late carry" formulation --------------------------------------------
//
// The reference's encoder (queue.rs:612-705) holds words back while the coding interval straddles a multiple
// of 2^64 (`EncoderSituation::Inverted(n, first)`) and releases them as `first, 0xffffffff x (n-1)` or, if the
// interval ended up beyond the boundary, as `first + 1, 0 x (n-1)`.  The same words result from appending every
// word the moment the interval is renormalised and, should `lower` later wrap around, adding one to the words
// already written (a carry that ripples through the trailing 0xffffffff words into `first`):
//   * a Normal interval [lower, lower + range) does not wrap and `scale * left < range`, so `lower` cannot wrap;
//   * while Inverted, a renormalisation that leaves the interval wrapped has lower > 2^64 - 2^32, i.e. the
//     word it appends is 0xffffffff -- exactly the fill word the reference will emit if no carry arrives;
//   * `first <= 0xfffffffe` (the interval was not wrapped before its word was taken), so the carry stops there.
// The situation is a function of (lower, range) alone -- Inverted <=> !(lower + range > lower) -- and the
// number of held-back words is the run of trailing 0xffffffff words plus one, so the hot loop carries only
// (lower, range) and the reference's 4-tuple is recovered when a caller asks for the raw state.
struct RangeEncState {
    uint64_t lower;
    uint64_t range;  // u64::MAX when empty (queue.rs:98-106)
};

CTR_HD RangeEncState range_enc_init() {
    RangeEncState s;
    s.lower = 0;
    s.range = ~0ull;
    return s;
}

// EncoderSituation::Inverted <=> the interval wraps (queue.rs:647-666, 684-700)
CTR_HD bool range_enc_inverted(const RangeEncState &s) { return !(s.lower + s.range > s.lower); }

// One encode step (queue.rs:612-705).  Returns bit 0: add one to the words written so far (before `word`);
// bit 1: append `word`.  prob == 0 (impossible symbol) must be rejected by the caller.
CTR_HD uint32_t range_encode_step(RangeEncState &s, uint32_t left, uint32_t prob, uint32_t &word) {
    const uint64_t scale = s.range >> kPrecision;
    uint64_t range = scale * (uint64_t)prob;
    uint64_t lower = s.lower + scale * (uint64_t)left;  // wrapping
    uint32_t flags = lower < s.lower ? 1u : 0u;
    if (range < (1ull << 32)) {
        word = (uint32_t)(lower >> 32);
        lower <<= 32;
        range <<= 32;
        flags |= 2u;
    }
    s.lower = lower;
    s.range = range;
    return flags;
}

// seal (queue.rs:349-376, 458-523): `carry` = add one to the words written so far; then `n` (0..2) words follow:
// the point word, and a zero word iff the word above the interval's end equals the point word.
struct RangeSeal {
    bool carry;
    uint32_t n;
    uint32_t point_word;
};
CTR_HD RangeSeal range_seal(const RangeEncState &s) {
    RangeSeal r;
    r.carry = false;
    r.n = 0;
    r.point_word = 0;
    if (s.range == ~0ull) return r;  // nothing was encoded
    const uint64_t point = s.lower + 0xffffffffull;
    r.carry = point < s.lower;
    r.point_word = (uint32_t)(point >> 32);
    const uint32_t upper_word = (uint32_t)((s.lower + s.range) >> 32);
    r.n = upper_word == r.point_word ? 2u : 1u;
    return r;
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
