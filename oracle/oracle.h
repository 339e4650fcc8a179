/*
 * oracle.h -- CPU restatement of constriction's ANS / range-coder hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (`constriction_b200/`,
 * `include/`) may include, link or call this.  Only `tests/`,
 * `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
 * `bench.py` use it, and only as the checker / reported CPU baseline.
 *
 * The reference (bamler-lab/constriction v0.5.0, Rust) cannot be compiled in
 * this image (no cargo/rustc), so this is a restatement in plain C of the
 * algorithm in the cited files, pinned by the reference's own golden vectors
 * (tests/test_oracle_golden.py).
 *
 * Fixed to the "Default" preset, the only one the reference's Python API
 * exposes: Word=u32, State=u64, Probability=u32, PRECISION=24, Symbol=i32
 * (reference: src/pybindings/stream/model/internals.rs:21-39).
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_PRECISION 24u
#define ORC_TOTAL (1u << ORC_PRECISION)

/* status codes */
#define ORC_OK 0
#define ORC_ERR_IMPOSSIBLE_SYMBOL 1 /* reference: lib.rs:376 ImpossibleSymbol -> KeyError      */
#define ORC_ERR_INVALID_DATA 2      /* reference: queue.rs:1401 InvalidData    -> AssertionError */
#define ORC_ERR_TRAILING_ZERO 3     /* reference: stack.rs:1555                -> ValueError     */
#define ORC_ERR_NOT_SEALED 4        /* reference: stack.rs:944-955 into_binary -> AssertionError */
#define ORC_ERR_BAD_MODEL 5         /* not normalisable / std<=0 / support too large             */
#define ORC_ERR_SEEK 6

/* ---------- math: libm-0.2.x erf/exp (FreeBSD msun s_erf.c / e_exp.c) ---------- */
double orc_exp(double x);
double orc_erf(double x);
/* probability-0.20.3 Gaussian::distribution: (1 + erf((x-mu)/(sigma*sqrt2)))/2 */
double orc_gaussian_cdf(double x, double mean, double std);

double orc_log1p(double x); /* libm log1p (categorical.rs:11) */
double orc_atan(double x);
double orc_laplace_cdf(double x, double mu, double b);
double orc_cauchy_cdf(double x, double x0, double gamma);

/* ---------- entropy models ---------- */
/* LeakyQuantizer tables of the other closed-form models of the Python API (pybindings/stream/model.rs:740-960):
 * kind 0 Gaussian(mean, std), 1 Laplace(mean, scale), 2 Cauchy(loc, scale); Binomial(n, p) over {0..n}.
 * PARITY UNPINNED for Laplace / Cauchy / Binomial: their CDFs come from the crate `probability`, absent here, and the
 * reference holds no golden vectors for them. */
int orc_qdist_cdf(int kind, int32_t min_sym, int32_t max_sym, double p0, double p1, uint32_t *cdf);
int orc_binomial_cdf(int32_t n, double p, uint32_t *cdf);
/* categorical.rs:56-177 + contiguous.rs:301-312: `Categorical(p)` / `Bernoulli(p)` with perfect=True (the Python default) */
int orc_cat_perfect_cdf_f32(const float *pmf, size_t n, uint32_t *cdf);
int orc_cat_perfect_cdf_f64(const double *pmf, size_t n, uint32_t *cdf);
/* quantize.rs:525-568 LeakilyQuantizedDistribution::left_cumulative_and_probability */
int orc_qgauss_left_prob(int32_t min_sym, int32_t max_sym, double mean, double std, int32_t symbol,
                         uint32_t *left, uint32_t *prob);
/* table of the same values: cdf[0..n] with cdf[0]=0, cdf[n]=2^24, n = max-min+1 */
int orc_qgauss_cdf(int32_t min_sym, int32_t max_sym, double mean, double std, uint32_t *cdf);
/* quantize.rs:580-779: result is the unique bin with left<=q<right; computed by search */
int orc_qgauss_quantile(int32_t min_sym, int32_t max_sym, double mean, double std, uint32_t quantile,
                        int32_t *symbol, uint32_t *left, uint32_t *prob);

/* categorical.rs:16-54 fast_quantized_cdf (+ contiguous.rs:499-512 terminal 2^24) */
int orc_cat_cdf_f32(const float *pmf, size_t n, uint32_t *cdf);
int orc_cat_cdf_f64(const double *pmf, size_t n, uint32_t *cdf);
/* lazy_contiguous.rs:228-257 / :268-330 (what Python `Categorical(perfect=False)` families use) */
int orc_cat_lazy_left_prob_f32(const float *pmf, size_t n, int32_t symbol, uint32_t *left, uint32_t *prob);
int orc_cat_lazy_left_prob_f64(const double *pmf, size_t n, int32_t symbol, uint32_t *left, uint32_t *prob);
int orc_cat_lazy_quantile_f32(const float *pmf, size_t n, uint32_t q, int32_t *symbol, uint32_t *left, uint32_t *prob);
int orc_cat_lazy_quantile_f64(const double *pmf, size_t n, uint32_t q, int32_t *symbol, uint32_t *left, uint32_t *prob);
/* contiguous.rs:628-700 table model: encoder lookup and binary-search decoder (symbols 0..n-1) */
int orc_cdf_left_prob(const uint32_t *cdf, size_t n, int64_t index, uint32_t *left, uint32_t *prob);
void orc_cdf_quantile(const uint32_t *cdf, size_t n, uint32_t q, size_t *index, uint32_t *left, uint32_t *prob);
/* uniform.rs:91-146 */
int orc_uniform_left_prob(uint32_t size, int32_t symbol, uint32_t *left, uint32_t *prob);
void orc_uniform_quantile(uint32_t size, uint32_t q, int32_t *symbol, uint32_t *left, uint32_t *prob);

/* ---------- ANS coder (stack.rs) ---------- */
typedef struct {
    uint32_t *bulk; /* Vec<u32> backend, backends.rs:470-557 */
    size_t len, cap;
    uint64_t state;
} orc_ans;

void orc_ans_init(orc_ans *c);
void orc_ans_free(orc_ans *c);
void orc_ans_clear(orc_ans *c);
int orc_ans_from_compressed(orc_ans *c, const uint32_t *words, size_t n); /* stack.rs:299-318,440-462 */
void orc_ans_from_binary(orc_ans *c, const uint32_t *words, size_t n);    /* stack.rs:341-360 */
void orc_ans_encode(orc_ans *c, uint32_t left, uint32_t prob);            /* stack.rs:1014-1048 */
uint32_t orc_ans_peek_quantile(const orc_ans *c);                         /* stack.rs:1086 */
void orc_ans_decode_advance(orc_ans *c, uint32_t left, uint32_t prob);    /* stack.rs:1088-1097 */
size_t orc_ans_num_words(const orc_ans *c);                               /* stack.rs:609-615 */
size_t orc_ans_num_valid_bits(const orc_ans *c);                          /* stack.rs:624-630 */
int orc_ans_is_empty(const orc_ans *c);
size_t orc_ans_get_compressed(const orc_ans *c, uint32_t *out);           /* stack.rs:537-547,891-895 */
int orc_ans_get_binary(const orc_ans *c, uint32_t *out, size_t *n_out);   /* stack.rs:549-556,944-955 */
int orc_ans_seek(orc_ans *c, size_t pos, uint64_t state);                 /* stack.rs:1117-1127 */

/* whole-array helpers (used for bulk parity tests and the CPU baseline):
 * encode symbols[n-1..0] (reverse) / decode n symbols with one CDF table */
int orc_ans_encode_iid_reverse(orc_ans *c, const int32_t *symbols, size_t n, const uint32_t *cdf,
                               int32_t min_sym, size_t alphabet);
void orc_ans_decode_iid(orc_ans *c, int32_t *symbols, size_t n, const uint32_t *cdf, int32_t min_sym,
                        size_t alphabet);
/* per-symbol model index into cdfs[model][stride] */
int orc_ans_encode_indexed_reverse(orc_ans *c, const int32_t *symbols, const uint32_t *model_idx, size_t n,
                                   const uint32_t *cdfs, size_t stride, int32_t min_sym, size_t alphabet);
void orc_ans_decode_indexed(orc_ans *c, int32_t *symbols, const uint32_t *model_idx, size_t n,
                            const uint32_t *cdfs, size_t stride, int32_t min_sym, size_t alphabet);
/* reference-shaped lazy Gaussian (2x erf per encoded symbol), stack.rs + quantize.rs:525-568 */
int orc_ans_encode_qgauss_lazy_reverse(orc_ans *c, const int32_t *symbols, size_t n, int32_t min_sym,
                                       int32_t max_sym, const double *means, const double *stds,
                                       int per_symbol_params);

/* quantize.rs:580-779 with the reference's control flow (inverse-CDF guess, doubling steps, bisection) and the
 * lazily evaluated decode loop on top of it: the work stock constriction does per decoded QuantizedGaussian symbol */
int orc_qgauss_quantile_guided(int32_t min_sym, int32_t max_sym, double mean, double std, uint32_t quantile,
                               int32_t *symbol, uint32_t *left, uint32_t *prob);
int orc_ans_decode_qgauss_lazy(orc_ans *c, int32_t *symbols, size_t n, int32_t min_sym, int32_t max_sym,
                               const double *means, const double *stds, int per_symbol_params);

/* ---------- Range coder (queue.rs) ---------- */
typedef struct {
    uint32_t *bulk;
    size_t len, cap;
    uint64_t lower, range;
    size_t num_inverted; /* 0 == EncoderSituation::Normal */
    uint32_t first_inverted;
} orc_renc;

typedef struct {
    const uint32_t *bulk; /* borrowed */
    size_t len, pos;
    uint64_t lower, range, point;
} orc_rdec;

void orc_renc_init(orc_renc *e);
void orc_renc_free(orc_renc *e);
void orc_renc_clear(orc_renc *e);
int orc_renc_encode(orc_renc *e, uint32_t left, uint32_t prob);  /* queue.rs:612-705 */
size_t orc_renc_num_seal_words(const orc_renc *e);               /* queue.rs:357-376 */
size_t orc_renc_num_words(const orc_renc *e);
size_t orc_renc_get_compressed(const orc_renc *e, uint32_t *out); /* queue.rs:349-355,458-523 */
int orc_renc_encode_iid(orc_renc *e, const int32_t *symbols, size_t n, const uint32_t *cdf, int32_t min_sym,
                        size_t alphabet);

void orc_rdec_init(orc_rdec *d, const uint32_t *words, size_t n); /* queue.rs:755-773,847-868 */
int orc_rdec_peek_quantile(const orc_rdec *d, uint32_t *q);       /* queue.rs:989-993 */
void orc_rdec_advance(orc_rdec *d, uint32_t left, uint32_t prob); /* queue.rs:998-1032 */
int orc_rdec_maybe_exhausted(const orc_rdec *d);                  /* queue.rs:872-883 */
int orc_rdec_seek(orc_rdec *d, size_t pos, uint64_t lower, uint64_t range); /* queue.rs:911-928 */
int orc_rdec_decode_iid(orc_rdec *d, int32_t *symbols, size_t n, const uint32_t *cdf, int32_t min_sym,
                        size_t alphabet);

/* ---------- multi-stream helpers (CPU baseline, bulk parity) ----------
 * K independent coders; stream k owns symbols {k, k+K, k+2K, ...} of a flat
 * message ("interleaved" deal) or symbols[off[k]..off[k+1]) ("contiguous").
 * Output: per-stream words concatenated, offsets[K+1]. Uses `threads` pthreads. */
int orc_multi_ans_encode(const int32_t *symbols, uint64_t n_total, uint64_t K, int interleaved,
                         const uint64_t *sym_off, const uint32_t *cdf, int32_t min_sym, size_t alphabet,
                         uint32_t **words_out, uint64_t *offsets_out, int threads);
int orc_multi_ans_decode(const uint32_t *words, const uint64_t *offsets, uint64_t n_total, uint64_t K,
                         int interleaved, const uint64_t *sym_off, const uint32_t *cdf, int32_t min_sym,
                         size_t alphabet, int32_t *symbols_out, int threads);
int orc_multi_range_encode(const int32_t *symbols, uint64_t n_total, uint64_t K, int interleaved,
                           const uint64_t *sym_off, const uint32_t *cdf, int32_t min_sym, size_t alphabet,
                           uint32_t **words_out, uint64_t *offsets_out, int threads);
int orc_multi_range_decode(const uint32_t *words, const uint64_t *offsets, uint64_t n_total, uint64_t K,
                           int interleaved, const uint64_t *sym_off, const uint32_t *cdf, int32_t min_sym,
                           size_t alphabet, int32_t *symbols_out, int threads);
void orc_free(void *p);

/* ---------- any preset (oracle_generic.c): Word bits W in {16, 32}, State = 2 W bits, PRECISION P <= W ----------
 * The reference's coders and categorical models are generic over their integer types; these functions restate
 * that generic code once.  (W, P) = (32, 24) is the Default preset and must equal the functions above (which the
 * reference's golden vectors pin); (16, 12) is the Small preset (stack.rs:153, queue.rs:156,747), for which the
 * reference holds no golden vectors.  Words travel as uint32_t array elements whatever W; `*words_out` is
 * malloc'ed (orc_free).  `table` (2^P entries, orc_g_lookup_table) selects the lookup decoder model
 * (lookup_contiguous.rs:297-333,564-607), NULL the binary search (contiguous.rs:628-665). */
int orc_g_lookup_table(unsigned P, const uint32_t *cdf, size_t n, uint32_t *table);
int orc_g_ans_encode_iid_reverse(unsigned W, unsigned P, const int32_t *symbols, size_t n, const uint32_t *cdf,
                                 int32_t min_sym, size_t alphabet, uint32_t **words_out, size_t *n_words);
int orc_g_ans_decode_iid(unsigned W, unsigned P, const uint32_t *words, size_t n_words, size_t n, const uint32_t *cdf,
                         int32_t min_sym, size_t alphabet, const uint32_t *table, int32_t *symbols_out);
int orc_g_range_encode_iid(unsigned W, unsigned P, const int32_t *symbols, size_t n, const uint32_t *cdf, int32_t min_sym,
                           size_t alphabet, uint32_t **words_out, size_t *n_words);
int orc_g_range_decode_iid(unsigned W, unsigned P, const uint32_t *words, size_t n_words, size_t n, const uint32_t *cdf,
                           int32_t min_sym, size_t alphabet, const uint32_t *table, int32_t *symbols_out);
int orc_g_cat_cdf_f32(unsigned P, unsigned prob_bits, const float *pmf, size_t n, uint32_t *cdf);
int orc_g_cat_cdf_f64(unsigned P, unsigned prob_bits, const double *pmf, size_t n, uint32_t *cdf);
int orc_g_cat_perfect_cdf_f64(unsigned P, unsigned prob_bits, const double *pmf, size_t n, uint32_t *cdf);

#ifdef __cplusplus
}
#endif
#endif
