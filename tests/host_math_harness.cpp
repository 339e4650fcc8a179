// Host build of the product's `__host__ __device__` coder / model arithmetic
// (constriction_b200/csrc/coder_math.cuh, model_math.cuh) so that it can be fuzzed against the oracle
// on a machine without a GPU.  Test infrastructure only; compiled by tests/host_harness.py with
//   g++ -O2 -ffp-contract=off -shared -fPIC
#include <stdint.h>
#include <stdlib.h>

#include <vector>

#include "../constriction_b200/csrc/coder_math.cuh"
#include "../constriction_b200/csrc/model_math.cuh"

using namespace ctr;

extern "C" {

void h_divmod(uint64_t n, uint32_t d, uint64_t *q, uint32_t *r) {
    divmod_by_reciprocal(n, d, reciprocal_u64(d), *q, *r);
}

// returns number of mismatches vs native / and % over `count` (n, d) pairs
uint64_t h_divmod_check(const uint64_t *n, const uint32_t *d, uint64_t count) {
    uint64_t bad = 0;
    for (uint64_t i = 0; i < count; ++i) {
        uint64_t q;
        uint32_t r;
        divmod_by_reciprocal(n[i], d[i], reciprocal_u64(d[i]), q, r);
        if (q != n[i] / d[i] || r != (uint32_t)(n[i] % d[i])) ++bad;
    }
    return bad;
}

// returns mismatches of the FP64-reciprocal division (estimate + exact correction) against native / and %
// over `count` (n, d) pairs with n < d * 2^40 (the encoder's operand range)
uint64_t h_f64div_check(const uint64_t *n, const uint32_t *d, uint64_t count) {
    uint64_t bad = 0;
    for (uint64_t i = 0; i < count; ++i) {
        const uint64_t got = ans_encode_update_f64(n[i], 0u, d[i], reciprocal_f64_bits(d[i]));
        const uint64_t want = ((n[i] / d[i]) << 24) | (n[i] % d[i]);
        if (got != want) ++bad;
    }
    return bad;
}

// One ANS coder: encode symbols (reverse) with cdf; returns words (bulk ++ state).  out must hold n+2.
uint64_t h_ans_encode(const int32_t *symbols, uint64_t n, const uint32_t *cdf, int32_t min_symbol,
                      uint64_t init_state, uint32_t *out, uint64_t *state_out, int f64) {
    uint64_t state = init_state, len = 0;
    for (uint64_t i = n; i-- > 0;) {
        const uint32_t idx = (uint32_t)symbols[i] - (uint32_t)min_symbol;
        const uint32_t left = cdf[idx], prob = cdf[idx + 1] - cdf[idx];
        if (ans_encode_needs_flush(state, prob)) {
            out[len++] = (uint32_t)state;
            state >>= 32;
        }
        state = f64 ? ans_encode_update_f64(state, left, prob, reciprocal_f64_bits(prob))
                    : ans_encode_update(state, left, prob, reciprocal_u64(prob));
    }
    *state_out = state;
    const uint32_t ns = ans_state_words(state);
    if (ns >= 1) out[len++] = (uint32_t)state;
    if (ns == 2) out[len++] = (uint32_t)(state >> 32);
    return len;
}

void h_ans_decode(const uint32_t *words, uint64_t n_words, int32_t *symbols, uint64_t n, const uint32_t *cdf,
                  uint32_t alphabet, int32_t min_symbol) {
    uint64_t len = n_words, state = 0;
    if (len) {
        state = words[--len];
        if (len) state = (state << 32) | words[--len];
    }
    for (uint64_t i = 0; i < n; ++i) {
        const uint32_t q = ans_peek_quantile(state);
        uint32_t lo = 0, hi = alphabet - 1;
        while (lo < hi) {
            const uint32_t mid = (lo + hi + 1) >> 1;
            if (cdf[mid] <= q) lo = mid; else hi = mid - 1;
        }
        state = ans_decode_update(state, q, cdf[lo], cdf[lo + 1] - cdf[lo]);
        if ((state >> 32) == 0 && len) state = (state << 32) | words[--len];
        symbols[i] = (int32_t)((uint32_t)min_symbol + lo);
    }
}

// One range encoder incl. seal.  out must hold n + 8 words.
static void carry_into(uint32_t *out, uint64_t len) {
    while (len > 0) {
        len -= 1;
        if (++out[len] != 0u) break;
    }
}

// One-shot encode when split >= n; otherwise the coder is suspended after `split` symbols the way the kernels
// do it for a raw-state caller (held-back words stripped and described by (num_inverted, first)), and resumed.
uint64_t h_range_encode_split(const int32_t *symbols, uint64_t n, uint64_t split, const uint32_t *cdf, int32_t min_symbol,
                              uint32_t *out) {
    RangeEncState st = range_enc_init();
    uint64_t len = 0;
    for (uint64_t i = 0; i < n; ++i) {
        if (i == split && range_enc_inverted(st) && st.range != ~0ull) {
            uint64_t held = 1;
            while (out[len - held] == 0xffffffffu) held += 1;
            const uint32_t first = out[len - held];
            len -= held;  // what a raw-state caller receives: the final words only
            // resume: the held-back words are written again, speculatively
            out[len++] = first;
            for (uint64_t j = 1; j < held; ++j) out[len++] = 0xffffffffu;
        }
        const uint32_t idx = (uint32_t)symbols[i] - (uint32_t)min_symbol;
        const uint32_t prob = cdf[idx + 1] - cdf[idx];
        if (prob == 0) return ~0ull;
        uint32_t word = 0;
        const uint32_t flags = range_encode_step(st, cdf[idx], prob, word);
        if (flags & 1u) carry_into(out, len);
        if (flags & 2u) out[len++] = word;
    }
    const RangeSeal seal = range_seal(st);
    if (seal.carry) carry_into(out, len);
    if (seal.n >= 1) out[len++] = seal.point_word;
    if (seal.n == 2) out[len++] = 0u;
    return len;
}

uint64_t h_range_encode(const int32_t *symbols, uint64_t n, const uint32_t *cdf, int32_t min_symbol, uint32_t *out) {
    return h_range_encode_split(symbols, n, ~0ull, cdf, min_symbol, out);
}

int h_range_decode(const uint32_t *words, uint64_t n_words, int32_t *symbols, uint64_t n, const uint32_t *cdf,
                   uint32_t alphabet, int32_t min_symbol) {
    RangeDecState st;
    st.lower = 0;
    st.range = ~0ull;
    st.point = 0;
    uint64_t pos = 0;
    if (n_words >= 1) st.point = (uint64_t)words[pos++] << 32;
    if (n_words >= 2) st.point |= words[pos++];
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t q;
        if (!range_peek_quantile(st, q)) return 2;
        uint32_t lo = 0, hi = alphabet - 1;
        while (lo < hi) {
            const uint32_t mid = (lo + hi + 1) >> 1;
            if (cdf[mid] <= q) lo = mid; else hi = mid - 1;
        }
        if (range_decode_update(st, cdf[lo], cdf[lo + 1] - cdf[lo])) {
            if (pos < n_words) st.point |= words[pos++];
        }
        symbols[i] = (int32_t)((uint32_t)min_symbol + lo);
    }
    return 0;
}

// returns mismatches of range_peek_quantile against native u64 division
uint64_t h_range_quantile_check(const uint64_t *diff, const uint64_t *range, uint64_t count) {
    uint64_t bad = 0;
    for (uint64_t i = 0; i < count; ++i) {
        RangeDecState st;
        st.lower = 0;
        st.range = range[i];
        st.point = diff[i];
        const uint64_t scale = range[i] >> 24;
        const uint64_t q_true = diff[i] / scale;
        uint32_t q;
        const bool ok = range_peek_quantile(st, q);
        if (ok != (q_true < (1ull << 24))) ++bad;
        else if (ok && q != (uint32_t)q_true) ++bad;
    }
    return bad;
}

double h_erf(double x) { return mm::erf_msun(x); }
double h_exp(double x) { return mm::exp_msun(x); }

int h_qgauss_cdf(int32_t min_symbol, int32_t max_symbol, double mean, double std, uint32_t *cdf) {
    double fw;
    if (!mm::leaky_free_weight(min_symbol, max_symbol, fw)) return 5;
    const uint32_t n = (uint32_t)((int64_t)max_symbol - (int64_t)min_symbol) + 1;
    for (uint32_t i = 0; i < n; ++i) cdf[i] = mm::leaky_gaussian_left(fw, min_symbol, mean, std, i);
    cdf[n] = kTotal;
    return 0;
}
}
