/*
 * oracle_generic.c -- the reference's coders and categorical models for ANY preset (Word bits W in {16, 32},
 * State = 2 W bits, PRECISION P <= W), restated once from the reference's generic Rust code.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  The reference writes AnsCoder<Word, State>, RangeEncoder<Word, State>
 * and the categorical models generically over their integer types (stack.rs:1014-1100, queue.rs:612-705,968-1035,
 * categorical.rs:16-177, lookup_contiguous.rs:297-333,564-607); its "Small" preset (stack.rs:153, queue.rs:156,747:
 * u16 / u32 / 12) has NO golden vectors of its own.  This file is therefore pinned through the Default preset: the
 * SAME functions called with (W, P) = (32, 24) must reproduce oracle.c, which reproduces the reference's goldens
 * (tests/test_oracle_generic.py), and the Small preset is the same code path with (16, 12).
 *
 * Words are passed as uint32_t array elements whatever W (for W = 16 every element is < 65536).
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t state_mask(unsigned W) { return W == 32 ? ~0ull : ((1ull << (2 * W)) - 1); }
static inline uint64_t word_mask(unsigned W) { return (1ull << W) - 1; }

typedef struct {
    uint32_t *v;
    size_t len, cap;
} wvec;
static int wpush(wvec *w, uint32_t x) {
    if (w->len == w->cap) {
        size_t nc = w->cap ? w->cap * 2 : 64;
        uint32_t *nv = (uint32_t *)realloc(w->v, nc * sizeof *nv);
        if (!nv) return 0;
        w->v = nv;
        w->cap = nc;
    }
    w->v[w->len++] = x;
    return 1;
}

/* contiguous.rs:673-700 / :628-665 on a CDF row cdf[0..n] with cdf[n] = 2^P */
static int g_left_prob(const uint32_t *cdf, size_t n, int64_t idx, uint32_t *left, uint32_t *prob) {
    if (idx < 0 || (uint64_t)idx >= n) return ORC_ERR_IMPOSSIBLE_SYMBOL;
    *left = cdf[idx];
    *prob = cdf[idx + 1] - cdf[idx];
    return *prob ? ORC_OK : ORC_ERR_IMPOSSIBLE_SYMBOL;
}
static size_t g_quantile(const uint32_t *cdf, size_t n, uint32_t q) {
    size_t lo = 0, hi = n - 1; /* last index with cdf[idx] <= q */
    while (lo < hi) {
        size_t mid = (lo + hi + 1) / 2;
        if (cdf[mid] <= q)
            lo = mid;
        else
            hi = mid - 1;
    }
    return lo;
}

/* lookup_contiguous.rs:297-333: table[q] = index of the symbol whose interval contains quantile q */
int orc_g_lookup_table(unsigned P, const uint32_t *cdf, size_t n, uint32_t *table) {
    if (cdf[0] != 0 || cdf[n] != (1u << P)) return ORC_ERR_BAD_MODEL;
    for (size_t s = 0; s < n; s++) {
        if (cdf[s + 1] < cdf[s]) return ORC_ERR_BAD_MODEL;
        for (uint32_t q = cdf[s]; q < cdf[s + 1]; q++) table[q] = (uint32_t)s;
    }
    return ORC_OK;
}

/* ---- ANS (stack.rs) ---- */
/* encode_iid_symbols_reverse + into_compressed (stack.rs:1014-1048, 835-849, 891-895; lib.rs:719-730) */
int orc_g_ans_encode_iid_reverse(unsigned W, unsigned P, const int32_t *symbols, size_t n, const uint32_t *cdf,
                                 int32_t min_sym, size_t alphabet, uint32_t **words_out, size_t *n_words) {
    if ((W != 16 && W != 32) || P == 0 || P > W) return ORC_ERR_BAD_MODEL;
    const unsigned S = 2 * W;
    wvec out = {0, 0, 0};
    uint64_t state = 0;
    for (size_t i = n; i-- > 0;) {
        uint32_t left, prob;
        int rc = g_left_prob(cdf, alphabet, (int64_t)symbols[i] - (int64_t)min_sym, &left, &prob);
        if (rc) {
            free(out.v);
            return rc;
        }
        if ((state >> (S - P)) >= prob) {
            wpush(&out, (uint32_t)(state & word_mask(W)));
            state >>= W;
        }
        const uint64_t remainder = state % prob, prefix = state / prob;
        state = ((prefix << P) | (left + remainder)) & state_mask(W);
    }
    /* bit_array_to_chunks_truncated(state).rev(): least significant word first, leading zero words dropped */
    for (uint64_t s = state; s != 0; s >>= W) wpush(&out, (uint32_t)(s & word_mask(W)));
    *words_out = out.v;
    *n_words = out.len;
    return ORC_OK;
}

/* from_compressed (stack.rs:299-318,440-462) + decode_iid_symbols (stack.rs:1070-1100); `table` != NULL decodes
 * through the lookup table (lookup_contiguous.rs:564-607) instead of the binary search */
int orc_g_ans_decode_iid(unsigned W, unsigned P, const uint32_t *words, size_t n_words, size_t n, const uint32_t *cdf,
                         int32_t min_sym, size_t alphabet, const uint32_t *table, int32_t *symbols_out) {
    if ((W != 16 && W != 32) || P == 0 || P > W) return ORC_ERR_BAD_MODEL;
    const unsigned S = 2 * W;
    size_t len = n_words;
    uint64_t state = 0;
    if (len) {
        const uint32_t first = words[--len];
        if (first == 0) return ORC_ERR_TRAILING_ZERO;
        state = first;
        while (len) {
            state = (state << W) | words[--len];
            if (state >= (1ull << (S - W))) break;
        }
    }
    for (size_t i = 0; i < n; i++) {
        const uint32_t q = (uint32_t)(state & ((1ull << P) - 1));
        const size_t idx = table ? table[q] : g_quantile(cdf, alphabet, q);
        const uint32_t left = cdf[idx], prob = cdf[idx + 1] - cdf[idx];
        state = ((state >> P) * prob + (q - left)) & state_mask(W);
        if (state < (1ull << (S - W)) && len) state = (state << W) | words[--len];
        symbols_out[i] = (int32_t)((int64_t)min_sym + (int64_t)idx);
    }
    return ORC_OK;
}

/* ---- range coder (queue.rs) ---- */
int orc_g_range_encode_iid(unsigned W, unsigned P, const int32_t *symbols, size_t n, const uint32_t *cdf, int32_t min_sym,
                           size_t alphabet, uint32_t **words_out, size_t *n_words) {
    if ((W != 16 && W != 32) || P == 0 || P > W) return ORC_ERR_BAD_MODEL;
    const unsigned S = 2 * W;
    const uint64_t M = state_mask(W), WM = word_mask(W);
    wvec out = {0, 0, 0};
    uint64_t lower = 0, range = M;
    size_t num_inverted = 0;
    uint32_t first_inverted = 0;
    for (size_t i = 0; i < n; i++) { /* queue.rs:612-705 */
        uint32_t left, prob;
        int rc = g_left_prob(cdf, alphabet, (int64_t)symbols[i] - (int64_t)min_sym, &left, &prob);
        if (rc) {
            free(out.v);
            return rc;
        }
        const uint64_t scale = range >> P;
        range = scale * prob;
        const uint64_t new_lower = (lower + scale * left) & M;
        if (num_inverted) {
            if (((new_lower + range) & M) > new_lower) {
                uint32_t first_word, consecutive;
                if (new_lower < lower) {
                    first_word = (uint32_t)((first_inverted + 1u) & WM);
                    consecutive = 0;
                } else {
                    first_word = first_inverted;
                    consecutive = (uint32_t)WM;
                }
                wpush(&out, first_word);
                for (size_t j = 1; j < num_inverted; j++) wpush(&out, consecutive);
                num_inverted = 0;
            }
        }
        lower = new_lower;
        if (range < (1ull << (S - W))) {
            range = (range << W) & M;
            const uint32_t lower_word = (uint32_t)(lower >> (S - W));
            lower = (lower << W) & M;
            if (num_inverted) {
                num_inverted += 1;
            } else if (((lower + range) & M) > lower) {
                wpush(&out, lower_word);
            } else {
                num_inverted = 1;
                first_inverted = lower_word;
            }
        }
    }
    if (range != M) { /* seal: queue.rs:349-355,458-523 */
        const uint64_t point = (lower + ((1ull << (S - W)) - 1)) & M;
        if (num_inverted) {
            uint32_t first, consecutive;
            if (point >= lower) {
                first = first_inverted;
                consecutive = (uint32_t)WM;
            } else {
                first = (uint32_t)((first_inverted + 1u) & WM);
                consecutive = 0;
            }
            wpush(&out, first);
            for (size_t j = 1; j < num_inverted; j++) wpush(&out, consecutive);
        }
        const uint32_t point_word = (uint32_t)(point >> (S - W));
        wpush(&out, point_word);
        const uint32_t upper_word = (uint32_t)(((lower + range) & M) >> (S - W));
        if (upper_word == point_word) wpush(&out, 0);
    }
    *words_out = out.v;
    *n_words = out.len;
    return ORC_OK;
}

int orc_g_range_decode_iid(unsigned W, unsigned P, const uint32_t *words, size_t n_words, size_t n, const uint32_t *cdf,
                           int32_t min_sym, size_t alphabet, const uint32_t *table, int32_t *symbols_out) {
    if ((W != 16 && W != 32) || P == 0 || P > W) return ORC_ERR_BAD_MODEL;
    const unsigned S = 2 * W;
    const uint64_t M = state_mask(W);
    size_t pos = 0;
    uint64_t lower = 0, range = M, point = 0;
    { /* read_point: queue.rs:847-868 */
        unsigned num_read = 0;
        while (pos < n_words) {
            point = ((point << W) | words[pos++]) & M;
            if (++num_read == 2) break;
        }
        if (num_read == 1) point = (point << W) & M;
    }
    for (size_t i = 0; i < n; i++) { /* queue.rs:968-1035 */
        const uint64_t scale = range >> P;
        const uint64_t quantile = ((point - lower) & M) / scale;
        if (quantile >= (1ull << P)) return ORC_ERR_INVALID_DATA;
        const uint32_t q = (uint32_t)quantile;
        const size_t idx = table ? table[q] : g_quantile(cdf, alphabet, q);
        const uint32_t left = cdf[idx], prob = cdf[idx + 1] - cdf[idx];
        lower = (lower + scale * left) & M;
        range = scale * prob;
        if (range < (1ull << (S - W))) {
            lower = (lower << W) & M;
            range = (range << W) & M;
            point = (point << W) & M;
            if (pos < n_words) point |= words[pos++];
        }
        symbols_out[i] = (int32_t)((int64_t)min_sym + (int64_t)idx);
    }
    return ORC_OK;
}

/* ---- categorical models at precision P ---- */
static inline uint32_t sat_u(double v, unsigned bits) { /* Rust `as uN`: truncating, saturating, NaN -> 0 */
    const double max = bits == 16 ? 65535.0 : 4294967295.0;
    if (!(v > 0.0)) return 0;
    if (v >= max) return (uint32_t)max;
    return (uint32_t)v;
}
static inline uint32_t sat_uf(float v, unsigned bits) {
    if (!(v > 0.0f)) return 0;
    if (bits == 16) return v >= 65535.0f ? 65535u : (uint32_t)v;
    return v >= 4294967296.0f ? 0xffffffffu : (uint32_t)v;
}

/* categorical.rs:16-54 fast_quantized_cdf with Probability = uW', PRECISION = P (W' = 16 for P <= 16 presets, else 32) */
int orc_g_cat_cdf_f64(unsigned P, unsigned prob_bits, const double *pmf, size_t n, uint32_t *cdf) {
    const uint64_t total = 1ull << P;
    if (n < 2 || n >= total - 1) return ORC_ERR_BAD_MODEL;
    const uint32_t free_weight = (uint32_t)(total - n);
    double norm = 0.0;
    for (size_t i = 0; i < n; i++) norm = norm + pmf[i];
    if (!isnormal(norm) || signbit(norm)) return ORC_ERR_BAD_MODEL;
    const double scale = (double)free_weight / norm;
    double cum = 0.0;
    for (size_t i = 0; i < n; i++) {
        cdf[i] = sat_u(cum * scale, prob_bits) + (uint32_t)i;
        cum = cum + pmf[i];
    }
    cdf[n] = (uint32_t)total;
    return ORC_OK;
}
int orc_g_cat_cdf_f32(unsigned P, unsigned prob_bits, const float *pmf, size_t n, uint32_t *cdf) {
    const uint64_t total = 1ull << P;
    if (n < 2 || n >= total - 1) return ORC_ERR_BAD_MODEL;
    const uint32_t free_weight = (uint32_t)(total - n);
    float norm = 0.0f;
    for (size_t i = 0; i < n; i++) norm = norm + pmf[i];
    if (!isnormal(norm) || signbit(norm)) return ORC_ERR_BAD_MODEL;
    const float scale = (float)free_weight / norm;
    float cum = 0.0f;
    for (size_t i = 0; i < n; i++) {
        cdf[i] = sat_uf(cum * scale, prob_bits) + (uint32_t)i;
        cum = cum + pmf[i];
    }
    cdf[n] = (uint32_t)total;
    return ORC_OK;
}

/* categorical.rs:56-177 at precision P (see oracle.c cat_perfect_weights for the order-sensitive details) */
typedef struct {
    size_t original_index;
    double prob;
    uint32_t weight;
    double win, loss;
} gslot;
static void gsort(gslot *a, gslot *tmp, size_t n) {
    for (size_t width = 1; width < n; width *= 2) {
        for (size_t lo = 0; lo < n; lo += 2 * width) {
            size_t mid = lo + width < n ? lo + width : n, hi = lo + 2 * width < n ? lo + 2 * width : n;
            size_t i = lo, j = mid, k = lo;
            while (i < mid && j < hi) tmp[k++] = a[j].win > a[i].win ? a[j++] : a[i++];
            while (i < mid) tmp[k++] = a[i++];
            while (j < hi) tmp[k++] = a[j++];
        }
        memcpy(a, tmp, n * sizeof *a);
    }
}
int orc_g_cat_perfect_cdf_f64(unsigned P, unsigned prob_bits, const double *probs, size_t n, uint32_t *cdf) {
    const uint64_t total = 1ull << P;
    if (n < 2 || n > total) return ORC_ERR_BAD_MODEL;
    uint32_t remaining = (uint32_t)(total - n);
    double norm = 0.0;
    for (size_t i = 0; i < n; i++) norm = norm + probs[i];
    if (!isnormal(norm) || signbit(norm)) return ORC_ERR_BAD_MODEL;
    const double scale = (double)remaining / norm;
    gslot *slots = (gslot *)malloc(2 * n * sizeof *slots);
    if (!slots) return ORC_ERR_BAD_MODEL;
    gslot *tmp = slots + n;
    for (size_t i = 0; i < n; i++) {
        const double prob = probs[i];
        if (prob < 0.0) {
            free(slots);
            return ORC_ERR_BAD_MODEL;
        }
        const uint32_t current = sat_u(prob * scale, prob_bits);
        remaining -= current;
        const uint32_t weight = current + 1u;
        slots[i].original_index = i;
        slots[i].prob = prob;
        slots[i].weight = weight;
        slots[i].win = prob * orc_log1p(1.0 / (double)weight);
        slots[i].loss = weight == 1u ? INFINITY : -prob * orc_log1p(-1.0 / (double)weight);
    }
    while (remaining != 0u) {
        gsort(slots, tmp, n);
        const size_t batch = remaining < n ? remaining : n;
        for (size_t i = 0; i < batch; i++) {
            slots[i].weight += 1u;
            slots[i].win = slots[i].prob * orc_log1p(1.0 / (double)slots[i].weight);
            slots[i].loss = -slots[i].prob * orc_log1p(-1.0 / (double)slots[i].weight);
        }
        remaining -= (uint32_t)batch;
    }
    for (;;) {
        size_t buyer = 0, seller = 0;
        for (size_t i = 1; i < n; i++) {
            if (slots[i].win >= slots[buyer].win) buyer = i;
            if (slots[i].loss < slots[seller].loss) seller = i;
        }
        if (buyer == seller) break;
        if (slots[buyer].win <= slots[seller].loss) break;
        slots[seller].weight -= 1u;
        slots[seller].win = -INFINITY;
        slots[seller].loss = slots[seller].weight == 1u ? INFINITY : -slots[seller].prob * orc_log1p(-1.0 / (double)slots[seller].weight);
        slots[buyer].weight += 1u;
        slots[buyer].loss = INFINITY;
        slots[buyer].win = slots[buyer].prob * orc_log1p(1.0 / (double)slots[buyer].weight);
    }
    uint32_t *w = (uint32_t *)malloc(n * sizeof *w);
    for (size_t i = 0; i < n; i++) w[slots[i].original_index] = slots[i].weight;
    uint64_t acc = 0;
    for (size_t i = 0; i < n; i++) {
        cdf[i] = (uint32_t)acc;
        acc += w[i];
    }
    cdf[n] = (uint32_t)total;
    free(w);
    free(slots);
    return acc == total ? ORC_OK : ORC_ERR_BAD_MODEL;
}
