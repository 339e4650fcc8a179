/*
 * constriction_b200.h -- C ABI of the B200-native entropy-coding engine.
 *
 * This is the drop-in boundary for the ONE hot path of bamler-lab/constriction (v0.5.0) that this
 * project replaces: the per-symbol ANS / range-coder state-update loop, its entropy-model lookup
 * and its word I/O, for the "Default" preset the reference's Python API exposes
 * (Word=u32, State=u64, Probability=u32, PRECISION=24, Symbol=i32;
 *  reference: src/pybindings/stream/model/internals.rs:21-39).
 *
 * The reference has no C ABI on this path (its only FFI is the pyo3 module,
 * src/pybindings/mod.rs:178-184).  Every entry point below therefore cites the reference
 * *functions* a binding would route to it; INTEGRATION.md shows the Rust `extern "C"` block and the
 * pyo3 / ctypes stubs.  One call processes a *batch* of K independent coders ("streams"): stream k
 * of a batch is bit-identical to one reference coder object fed the same symbols and models.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / CUDA types in signatures (`stream` is a
 *     cudaStream_t passed as void*, NULL = legacy default stream);
 *   - `*_dev` pointers are device memory owned by the caller; all work is enqueued on `stream`
 *     without host synchronisation, unless the function name ends in `_host` (those take host
 *     buffers, copy in/out and synchronise before returning);
 *   - every function returns CTR_OK or a CTR_ERR_* code for *call-level* errors; *data-level*
 *     errors (impossible symbol, invalid compressed data, ...) are reported per batch through four
 *     device words `status_dev[4]` = {max error code, 0, index of one failing stream (lo, hi)},
 *     which the caller zeroes before the first call that uses them;
 *   - there is no CPU fallback: without a CUDA device every compute entry point returns
 *     CTR_ERR_CUDA.
 *
 * Symbol layouts of a batch (ctr_layout):
 *   - interleaved (sym_offsets_dev == NULL): stream k owns symbols k, k+K, k+2K, ... of the flat
 *     array ("lane-interleaved rANS": lane = coder, warp reads 128 contiguous bytes per step);
 *   - contiguous: stream k owns symbols[sym_offsets[k] .. sym_offsets[k+1]).
 * Compressed container of a batch: `words` (u32) + `offsets` (u64[K+1]); words[offsets[k] ..
 * offsets[k+1]) is exactly what the reference's `into_compressed()` / `get_compressed()` returns
 * for coder k, so any single stream can be handed to stock constriction.
 */
#ifndef CONSTRICTION_B200_H
#define CONSTRICTION_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTR_ABI_VERSION 3

/* call-level and data-level status codes */
#define CTR_OK 0
#define CTR_ERR_IMPOSSIBLE_SYMBOL 1 /* lib.rs:376 DefaultEncoderFrontendError::ImpossibleSymbol (py: KeyError)  */
#define CTR_ERR_INVALID_DATA 2      /* queue.rs:1401 DecoderFrontendError::InvalidData     (py: AssertionError) */
#define CTR_ERR_TRAILING_ZERO 3     /* stack.rs:1555 CompressedDataEndsWithZeroWord         (py: ValueError)     */
#define CTR_ERR_NOT_SEALED 4        /* stack.rs:944-955 into_binary on a non-sealed state                        */
#define CTR_ERR_BAD_MODEL 5         /* not normalisable / std <= 0 / support too large / malformed CDF          */
#define CTR_ERR_SEEK 6              /* seek past the end / invalid state                                         */
#define CTR_ERR_OUT_OF_SPACE 7      /* backends.rs:1512 BoundedWriteError::OutOfSpace (output capacity)          */
#define CTR_ERR_BAD_ARGUMENT 8
#define CTR_ERR_CUDA 9 /* CUDA runtime error or no device; see ctr_last_cuda_error() */

/* model_index_mode */
#define CTR_INDEX_NONE 0       /* every symbol uses model 0 of the table set (i.i.d.)          */
#define CTR_INDEX_PER_SYMBOL 1 /* model_index[i] belongs to symbols[i] (same layout as symbols) */
#define CTR_INDEX_PER_STREAM 2 /* model_index[k] belongs to stream k                            */

/* flags */
#define CTR_FLAG_RAW 1u /* states are exchanged through states_in/states_out instead of being      \
                           appended to / parsed from the words (Pos/Seek, stack.rs:1107-1139,      \
                           queue.rs:182-196,911-928; from_raw_parts / into_raw_parts)              */

#define CTR_FLAG_CHECKPOINTS 2u /* contiguous layout only.  Encoders also record their position (Pos::pos,    \
                                   stack.rs:1107-1115, queue.rs:182-196) every `checkpoint_every` symbols;      \
                                   decoders use the records to decode every chunk of every stream on its own   \
                                   lane (Seek::seek, stack.rs:1117-1139, queue.rs:911-928): a long stream no    \
                                   longer decodes at the speed of one dependent chain.  The words are           \
                                   unchanged: any stream is still one stock-constriction stream.               */

typedef struct ctr_model_s *ctr_model_t; /* opaque: device-resident 24-bit CDF tables of M models */

typedef struct {
    uint64_t n_streams;              /* K                                                    */
    uint64_t n_symbols;              /* N, total over all streams                            */
    const uint64_t *sym_offsets_dev; /* u64[K+1] or NULL (interleaved deal)                   */
    const uint32_t *model_index_dev; /* u32[N] / u32[K] / NULL according to model_index_mode  */
    int32_t model_index_mode;        /* CTR_INDEX_*                                           */
    uint32_t flags;                  /* CTR_FLAG_*                                            */
    /* CTR_FLAG_CHECKPOINTS (all zero otherwise).  Stream k of n_k symbols has J_k = ceil(n_k / C) records,
     * record j (decode order) describing the coder where the chunk of symbols starting at
     *   ANS:   s_0 = 0, s_j = n_k - (J_k - j) C   (boundaries counted from the stream's end: ANS is a stack)
     *   range: s_j = j C
     * begins: ANS {words pushed so far, state} (2 x u64), range {words pushed so far, lower, range, 0}
     * (4 x u64) -- the values AnsCoder::pos() / RangeEncoder::pos() return at that moment.             */
    uint32_t checkpoint_every;        /* C, a multiple of 32                                           */
    uint32_t reserved;
    const uint64_t *ckpt_offsets_dev; /* u64[K+1]: index of stream k's first record (ctr_checkpoint_offsets) */
    uint64_t *checkpoints_dev;        /* records; capacity ctr_checkpoint_max_records(layout)           */
} ctr_layout;

/* ---- library ------------------------------------------------------------------------------ */
int ctr_abi_version(void);
const char *ctr_status_string(int code);
const char *ctr_last_cuda_error(void); /* text of the last CUDA failure on this thread, or "" */
int ctr_device_count(void);            /* 0 without a usable CUDA device */

/* ---- entropy models (K5: tabulation) --------------------------------------------------------
 * A model set is M models over one alphabet {min_symbol .. min_symbol+alphabet-1}; model m is a
 * CDF row u32[alphabet+1] with cdf[0]=0, cdf[alphabet]=2^24, non-decreasing.
 * Replaces, evaluated once per model instead of once per symbol:
 *   EncoderModel::left_cumulative_and_probability   src/stream/model.rs:333-336
 *   DecoderModel::quantile_function                 src/stream/model.rs:457-464          */

/* QuantizedGaussian: LeakyQuantizer<f64,i32,u32,24>::quantize(Gaussian) for M (mean,std) pairs.
 * Reference: src/stream/model/quantize.rs:284-308,525-568; pybindings/stream/model.rs:645-708.
 * means/stds are HOST arrays of length n_models.  Returns CTR_ERR_BAD_MODEL if a std is not > 0. */
int ctr_model_quantized_gaussian(int32_t min_symbol, int32_t max_symbol, const double *means_host,
                                 const double *stds_host, uint32_t n_models, void *stream, ctr_model_t *out);

/* Categorical, `fast_quantized_cdf` rounding in the caller's float type (what Python
 * `Categorical(probs, perfect=False)` produces).  pmf is a HOST or DEVICE (is_device != 0)
 * row-major array [n_models][alphabet]; symbols are 0..alphabet-1.
 * Reference: src/stream/model/categorical.rs:16-54; categorical/contiguous.rs:203,499-512;
 * categorical/lazy_contiguous.rs:131-167,228-257. */
int ctr_model_categorical_f32(const float *pmf, int is_device, uint32_t n_models, uint32_t alphabet, void *stream,
                              ctr_model_t *out);
int ctr_model_categorical_f64(const double *pmf, int is_device, uint32_t n_models, uint32_t alphabet, void *stream,
                              ctr_model_t *out);

/* Categorical with the reference's "perfect" quantisation: perfectly_quantized_probabilities
 * (src/stream/model/categorical.rs:56-177) + from_floating_point_probabilities_perfect (categorical/contiguous.rs:301-312),
 * what Python `Categorical(probs)` and `Bernoulli(p)` construct BY DEFAULT (pybindings/stream/model.rs:508-523,1010-1050).
 * Computed in f64 whatever the input type; one device thread per model (the optimisation is sequential). */
int ctr_model_categorical_perfect_f32(const float *pmf, int is_device, uint32_t n_models, uint32_t alphabet, void *stream,
                                      ctr_model_t *out);
int ctr_model_categorical_perfect_f64(const double *pmf, int is_device, uint32_t n_models, uint32_t alphabet, void *stream,
                                      ctr_model_t *out);

/* The leaky quantiser over other two-parameter distributions (pybindings/stream/model.rs:740-900): kind 0
 * QuantizedGaussian(mean, std), 1 QuantizedLaplace(mean, scale), 2 QuantizedCauchy(location, scale); parameters are HOST
 * arrays of length n_models.  And Binomial(n, p) over {0..n} (pybindings/stream/model.rs:925-960); rows are padded to the
 * widest model.  The Laplace / Cauchy / Binomial CDFs live in the crate `probability`, which is not part of the
 * reference tree, and the reference holds no golden vectors for them: they follow the textbook definitions and their
 * bit-parity with a Rust build is UNPINNED (encoder and decoder of this library always agree with each other). */
int ctr_model_quantized(int32_t kind, int32_t min_symbol, int32_t max_symbol, const double *p0_host, const double *p1_host,
                        uint32_t n_models, void *stream, ctr_model_t *out);
int ctr_model_binomial(const int32_t *n_host, const double *p_host, uint32_t n_models, void *stream, ctr_model_t *out);

/* From ready-made fixed-point CDF rows u32[n_models][alphabet+1] (HOST or DEVICE memory; copied).
 * Reference: categorical/contiguous.rs:471-520 from_nonzero_fixed_point_probabilities /
 * from_fixed_point_cdf; lookup_contiguous.rs:297-333. */
int ctr_model_from_cdf(const uint32_t *cdf, int is_device, uint32_t n_models, uint32_t alphabet, int32_t min_symbol,
                       void *stream, ctr_model_t *out);

/* Uniform model over {0..size-1}: src/stream/model/uniform.rs:44-146. */
int ctr_model_uniform(uint32_t size, void *stream, ctr_model_t *out);

int ctr_model_destroy(ctr_model_t model);
int ctr_model_info(ctr_model_t model, uint32_t *n_models, uint32_t *alphabet, int32_t *min_symbol);
/* device pointer to the CDF rows, u32[n_models][alphabet+1] (owned by the model) */
const uint32_t *ctr_model_cdf_dev(ctr_model_t model);
/* copies the CDF rows to HOST memory (synchronises `stream`) */
int ctr_model_copy_cdf_host(ctr_model_t model, uint32_t *cdf_host, void *stream);

/* ---- ANS coder, stack semantics (K1/K2) -----------------------------------------------------
 * Replaces AnsCoder::encode_symbol (src/stream/stack.rs:1014-1048) under the driver loops
 * encode_symbols_reverse / encode_iid_symbols_reverse (stack.rs:784-849, stream/mod.rs:592-607,
 * 671-684), decode_symbol (stack.rs:1070-1100) under decode_symbols / decode_iid_symbols
 * (stream/mod.rs:893-910,1016-1031,1274-1297), into_compressed / get_compressed
 * (stack.rs:537-547,891-895,1148-1196; lib.rs:719-730) and from_compressed
 * (stack.rs:299-318,440-462), with Vec<u32> push/pop word I/O (src/backends.rs:470-557). */

/* bytes of device workspace ctr_ans_encode_reverse needs for this layout */
size_t ctr_ans_encode_workspace_bytes(const ctr_layout *layout);
/* upper bound on the total number of compressed words of a batch (capacity for words_out_dev) */
uint64_t ctr_ans_max_compressed_words(const ctr_layout *layout);

/* Every stream k encodes its symbols in REVERSE order onto coder k.
 *   states_in_dev   u64[K] or NULL (= fresh coders, state 0)
 *   words_out_dev   compact output, capacity words_capacity
 *   offsets_out_dev u64[K+1]; offsets[K] = total words
 *   states_out_dev  u64[K] or NULL; final coder states
 * Without CTR_FLAG_RAW stream k's words are bulk ++ state words (= get_compressed());
 * with it they are only the words pushed by this call (append them to the coder's bulk).   */
int ctr_ans_encode_reverse(ctr_model_t model, const int32_t *symbols_dev, const ctr_layout *layout,
                           const uint64_t *states_in_dev, void *workspace_dev, size_t workspace_bytes,
                           uint32_t *words_out_dev, uint64_t words_capacity, uint64_t *offsets_out_dev,
                           uint64_t *states_out_dev, uint32_t *status_dev, void *stream);

/* Every stream k decodes its symbols (forward order) from words[offsets[k]..offsets[k+1]).
 *   words_dev        MUST be 16-byte aligned, and the allocation behind it must extend to a multiple of 16 bytes
 *                    (4 words): the decoders (ANS and range) stage each stream with aligned 16-byte asynchronous
 *                    copies, so the block that holds a stream's last word is read whole -- up to 12 bytes past
 *                    offsets[K] words, never before words_dev.  A misaligned base returns CTR_ERR_BAD_ARGUMENT;
 *                    the padding cannot be checked by the library (cudaMalloc / cudaMallocAsync / framework
 *                    allocators round sizes up to >= 256 bytes, so whole allocations always qualify; a tightly
 *                    sized sub-buffer carved out of a larger one qualifies as long as the bytes behind it are
 *                    mapped).  The padding is never interpreted.
 *   states_in_dev    u64[K]; required with CTR_FLAG_RAW (all words are bulk), else NULL
 *   states_out_dev   u64[K] or NULL
 *   words_left_dev   u64[K] or NULL; bulk words not consumed (Pos::pos().0, stack.rs:1107-1115)
 * Decoding past the end of a stream is allowed and deterministic (stack.rs:1062-1065).       */
int ctr_ans_decode(ctr_model_t model, const uint32_t *words_dev, const uint64_t *offsets_dev,
                   const ctr_layout *layout, const uint64_t *states_in_dev, int32_t *symbols_out_dev,
                   uint64_t *states_out_dev, uint64_t *words_left_dev, uint32_t *status_dev, void *stream);

/* ---- Range coder, queue semantics (K3/K4) ---------------------------------------------------
 * Replaces RangeEncoder::encode_symbol (src/stream/queue.rs:612-705), seal / get_compressed
 * (queue.rs:349-376,458-523), RangeDecoder::from_compressed / read_point (queue.rs:755-773,847-868)
 * and decode_symbol (queue.rs:968-1035).
 * Range state on the wire (states_*): 4 x u64 per stream = {lower, range, num_inverted,
 * first_inverted_word} for encoders, {lower, range, point, unused} for decoders.               */
size_t ctr_range_encode_workspace_bytes(const ctr_layout *layout);
uint64_t ctr_range_max_compressed_words(const ctr_layout *layout);

int ctr_range_encode(ctr_model_t model, const int32_t *symbols_dev, const ctr_layout *layout,
                     const uint64_t *states_in_dev, void *workspace_dev, size_t workspace_bytes,
                     uint32_t *words_out_dev, uint64_t words_capacity, uint64_t *offsets_out_dev,
                     uint64_t *states_out_dev, uint32_t *status_dev, void *stream);

int ctr_range_decode(ctr_model_t model, const uint32_t *words_dev, const uint64_t *offsets_dev,
                     const ctr_layout *layout, const uint64_t *states_in_dev, int32_t *symbols_out_dev,
                     uint64_t *states_out_dev, uint64_t *words_read_dev, uint32_t *status_dev, void *stream);

/* ---- checkpoints -------------------------------------------------------------------------------
 * upper bound on the number of records of a batch: n_symbols / C + n_streams */
uint64_t ctr_checkpoint_max_records(const ctr_layout *layout);
/* fills ckpt_offsets_out_dev (u64[K+1]) from layout->sym_offsets_dev and layout->checkpoint_every */
int ctr_checkpoint_offsets(const ctr_layout *layout, uint64_t *ckpt_offsets_out_dev, void *stream);

/* ---- QuantizedGaussian with per-symbol parameters, evaluated on the device (no tables) -----------
 * What Python callers write as  coder.encode_reverse(symbols, QuantizedGaussian(lo, hi), means, stds)
 * / coder.decode(QuantizedGaussian(lo, hi), means, stds): the reference builds one
 * LeakilyQuantizedDistribution per symbol (src/pybindings/stream/model/internals.rs:188-249) and
 * evaluates it lazily -- encode: left_cumulative_and_probability, two Gaussian CDFs per symbol
 * (src/stream/model/quantize.rs:525-568); decode: quantile_function (quantize.rs:580-779).  Here the
 * encoders run a parallel pre-pass (thread per symbol) that produces the (left, probability) pairs, and
 * the decoders search the symbol's cumulative function inside the coder kernel; results are
 * word-for-word those of the table path (ctr_model_quantized_gaussian with one model per symbol), without
 * its O(alphabet) memory per symbol.
 *   means_dev / stds_dev  f64[N], laid out like the symbols (means[i], stds[i] belong to symbols[i]);
 *                         f32 parameters are widened by the caller, as the reference does
 *                         (pybindings/stream/model/internals.rs:169-174)
 *   layout                model_index_mode must be CTR_INDEX_NONE; N < 2^32
 * Data-level errors: CTR_ERR_BAD_MODEL if a std is not > 0 (pybindings/stream/model.rs:654-657),
 * CTR_ERR_IMPOSSIBLE_SYMBOL for a symbol outside [min_symbol, max_symbol].  Workspace and capacity as
 * for the table entry points (ctr_ans_encode_workspace_bytes, ctr_ans_max_compressed_words); scratch
 * for the pre-pass comes from the stream-ordered CUDA memory pool (no host synchronisation). */
int ctr_ans_encode_reverse_gaussian(int32_t min_symbol, int32_t max_symbol, const double *means_dev,
                                    const double *stds_dev, const int32_t *symbols_dev, const ctr_layout *layout,
                                    const uint64_t *states_in_dev, void *workspace_dev, size_t workspace_bytes,
                                    uint32_t *words_out_dev, uint64_t words_capacity, uint64_t *offsets_out_dev,
                                    uint64_t *states_out_dev, uint32_t *status_dev, void *stream);
int ctr_ans_decode_gaussian(int32_t min_symbol, int32_t max_symbol, const double *means_dev, const double *stds_dev,
                            const uint32_t *words_dev, const uint64_t *offsets_dev, const ctr_layout *layout,
                            const uint64_t *states_in_dev, int32_t *symbols_out_dev, uint64_t *states_out_dev,
                            uint64_t *words_left_dev, uint32_t *status_dev, void *stream);
int ctr_range_encode_gaussian(int32_t min_symbol, int32_t max_symbol, const double *means_dev,
                              const double *stds_dev, const int32_t *symbols_dev, const ctr_layout *layout,
                              const uint64_t *states_in_dev, void *workspace_dev, size_t workspace_bytes,
                              uint32_t *words_out_dev, uint64_t words_capacity, uint64_t *offsets_out_dev,
                              uint64_t *states_out_dev, uint32_t *status_dev, void *stream);
int ctr_range_decode_gaussian(int32_t min_symbol, int32_t max_symbol, const double *means_dev, const double *stds_dev,
                              const uint32_t *words_dev, const uint64_t *offsets_dev, const ctr_layout *layout,
                              const uint64_t *states_in_dev, int32_t *symbols_out_dev, uint64_t *states_out_dev,
                              uint64_t *words_read_dev, uint32_t *status_dev, void *stream);

/* ---- the "Small" preset with lookup decoder models ------------------------------------------
 * Word = u16, State = u32, Probability = u16, PRECISION = 12: SmallAnsCoder (src/stream/stack.rs:153),
 * SmallRangeEncoder / SmallRangeDecoder (src/stream/queue.rs:156,747), SmallContiguousCategoricalEntropyModel
 * (categorical/contiguous.rs:30) for encoding and ContiguousLookupDecoderModel / NonContiguousLookupDecoderModel
 * (categorical/lookup_contiguous.rs:169-333,564-607; lookup_noncontiguous.rs:167,602-645) for decoding: a table
 * with one entry per 12-bit quantile, so a decoded symbol costs one shared-memory load instead of a search.
 * A model set is M CDF rows u16[alphabet + 1] with cdf[0] = 0, cdf[alphabet] = 4096 (alphabet <= 4096).  Batches and
 * containers are described as for the Default preset, with u16 words (offsets count u16 words).  Models are shared
 * (CTR_INDEX_NONE) or per stream (CTR_INDEX_PER_STREAM); flags must be 0.  Stream k's words equal what
 * SmallAnsCoder::into_compressed() / SmallRangeEncoder::into_compressed() return for that stream. */
typedef struct ctr_small_model_s *ctr_small_model_t;
/* from fixed-point CDF rows (from_nonzero_fixed_point_probabilities, lookup_contiguous.rs:405-442).  symbols_host, if not
 * NULL, is i32[alphabet]: index i stands for the symbol symbols_host[i] (non-contiguous alphabet) */
int ctr_small_model_from_cdf(const uint16_t *cdf, int is_device, uint32_t n_models, uint32_t alphabet, int32_t min_symbol,
                             const int32_t *symbols_host, void *stream, ctr_small_model_t *out);
/* from_floating_point_probabilities_fast (perfect = 0; categorical.rs:16-54) / _perfect (categorical.rs:56-177) */
int ctr_small_model_categorical_f32(const float *pmf, int is_device, uint32_t n_models, uint32_t alphabet, int perfect,
                                    void *stream, ctr_small_model_t *out);
int ctr_small_model_categorical_f64(const double *pmf, int is_device, uint32_t n_models, uint32_t alphabet, int perfect,
                                    void *stream, ctr_small_model_t *out);
int ctr_small_model_destroy(ctr_small_model_t model);
int ctr_small_model_copy_cdf_host(ctr_small_model_t model, uint16_t *cdf_host, void *stream);
size_t ctr_small_encode_workspace_bytes(const ctr_layout *layout);
uint64_t ctr_small_max_compressed_words(const ctr_layout *layout);
int ctr_small_ans_encode_reverse(ctr_small_model_t model, const int32_t *symbols_dev, const ctr_layout *layout,
                                 void *workspace_dev, size_t workspace_bytes, uint16_t *words_out_dev, uint64_t words_capacity,
                                 uint64_t *offsets_out_dev, uint32_t *status_dev, void *stream);
int ctr_small_ans_decode(ctr_small_model_t model, const uint16_t *words_dev, const uint64_t *offsets_dev, const ctr_layout *layout,
                         int32_t *symbols_out_dev, uint32_t *status_dev, void *stream);
int ctr_small_range_encode(ctr_small_model_t model, const int32_t *symbols_dev, const ctr_layout *layout, void *workspace_dev,
                           size_t workspace_bytes, uint16_t *words_out_dev, uint64_t words_capacity, uint64_t *offsets_out_dev,
                           uint32_t *status_dev, void *stream);
int ctr_small_range_decode(ctr_small_model_t model, const uint16_t *words_dev, const uint64_t *offsets_dev, const ctr_layout *layout,
                           int32_t *symbols_out_dev, uint32_t *status_dev, void *stream);

/* ---- host-buffer entry points (the reference-facing call: host in, host out) ------------------
 * Same semantics with HOST buffers: what a pyo3 / Rust binding for `AnsCoder::encode_iid_symbols_reverse` /
 * `decode_iid_symbols` (stack.rs:835-849, stream/mod.rs:1016-1031) over many coders calls; bench.py's `e2e` number
 * times them.  A call is PCIe-bound, so the library pipelines it: the batch is cut into chunks of consecutive
 * streams (each a complete batch of its own) that flow through upload -> coder kernel -> download on three CUDA
 * streams, device buffers from the stream-ordered memory pool.  Pinned host buffers (cudaHostAlloc /
 * cudaHostRegister) make every copy an asynchronous DMA; pageable buffers work but serialise the pipeline.
 * The caller owns all buffers: `words_out_host` needs room for `words_capacity` words
 * (ctr_ans_max_compressed_words gives a safe bound; CTR_ERR_OUT_OF_SPACE if the batch needs more), offsets_out_host
 * u64[K+1].  `*data_status` / `*failing_stream` receive the data-level status.  sym_offsets_host that do not
 * describe slices of the symbol array return CTR_ERR_BAD_ARGUMENT. */
int ctr_ans_encode_reverse_host(ctr_model_t model, const int32_t *symbols_host, uint64_t n_symbols,
                                uint64_t n_streams, const uint64_t *sym_offsets_host,
                                const uint32_t *model_index_host, int32_t model_index_mode,
                                uint32_t *words_out_host, uint64_t words_capacity, uint64_t *offsets_out_host,
                                int *data_status, uint64_t *failing_stream);
int ctr_ans_decode_host(ctr_model_t model, const uint32_t *words_host, const uint64_t *offsets_host,
                        uint64_t n_symbols, uint64_t n_streams, const uint64_t *sym_offsets_host,
                        const uint32_t *model_index_host, int32_t model_index_mode, int32_t *symbols_out_host,
                        int *data_status, uint64_t *failing_stream);
int ctr_range_encode_host(ctr_model_t model, const int32_t *symbols_host, uint64_t n_symbols, uint64_t n_streams,
                          const uint64_t *sym_offsets_host, const uint32_t *model_index_host,
                          int32_t model_index_mode, uint32_t *words_out_host, uint64_t words_capacity,
                          uint64_t *offsets_out_host, int *data_status, uint64_t *failing_stream);
int ctr_range_decode_host(ctr_model_t model, const uint32_t *words_host, const uint64_t *offsets_host,
                          uint64_t n_symbols, uint64_t n_streams, const uint64_t *sym_offsets_host,
                          const uint32_t *model_index_host, int32_t model_index_mode, int32_t *symbols_out_host,
                          int *data_status, uint64_t *failing_stream);

/* Asynchronous forms: the same pipeline runs on a thread of the library; the call returns at once with a job handle
 * and every buffer (inputs, outputs, data_status, failing_stream) must stay valid until ctr_host_job_wait, which
 * returns the call's status and frees the handle.  One host thread can so keep both directions of the bus busy:
 * start the encode of batch i+1 (upload-bound) and the decode of batch i (download-bound), then wait for both. */
typedef struct ctr_host_job_s *ctr_host_job_t;
int ctr_ans_encode_reverse_host_async(ctr_model_t model, const int32_t *symbols_host, uint64_t n_symbols,
                                      uint64_t n_streams, const uint64_t *sym_offsets_host,
                                      const uint32_t *model_index_host, int32_t model_index_mode,
                                      uint32_t *words_out_host, uint64_t words_capacity, uint64_t *offsets_out_host,
                                      int *data_status, uint64_t *failing_stream, ctr_host_job_t *job);
int ctr_ans_decode_host_async(ctr_model_t model, const uint32_t *words_host, const uint64_t *offsets_host,
                              uint64_t n_symbols, uint64_t n_streams, const uint64_t *sym_offsets_host,
                              const uint32_t *model_index_host, int32_t model_index_mode, int32_t *symbols_out_host,
                              int *data_status, uint64_t *failing_stream, ctr_host_job_t *job);
int ctr_range_encode_host_async(ctr_model_t model, const int32_t *symbols_host, uint64_t n_symbols, uint64_t n_streams,
                                const uint64_t *sym_offsets_host, const uint32_t *model_index_host,
                                int32_t model_index_mode, uint32_t *words_out_host, uint64_t words_capacity,
                                uint64_t *offsets_out_host, int *data_status, uint64_t *failing_stream,
                                ctr_host_job_t *job);
int ctr_range_decode_host_async(ctr_model_t model, const uint32_t *words_host, const uint64_t *offsets_host,
                                uint64_t n_symbols, uint64_t n_streams, const uint64_t *sym_offsets_host,
                                const uint32_t *model_index_host, int32_t model_index_mode, int32_t *symbols_out_host,
                                int *data_status, uint64_t *failing_stream, ctr_host_job_t *job);
int ctr_host_job_wait(ctr_host_job_t job);

/* ---- wire / on-disk form of a batch container ----------------------------------------------------
 * The reference stores raw native-endian words and leaves framing to the user (src/lib.rs:425-580 describes the
 * (position, state) snapshots -- Pos::pos / Seek::seek -- that make random access possible).  A batch needs framing:
 * little-endian header {magic "CTRB200", version, coder, word bits, precision, K, N, total words, checkpoint_every,
 * flags, n_records}, then [sym_offsets u64[K+1]] offsets u64[K+1] [ckpt_offsets u64[K+1], records] words (csrc/container.cu
 * has the byte layout).  words[offsets[k] .. offsets[k+1]) is one stock constriction stream; a record is what
 * AnsCoder::pos() / RangeEncoder::pos() returned at a chunk boundary, so stock coders can seek() to it.
 * Host memory only, no device work.  pack() copies the arrays of `view` into `out` (ctr_container_size bytes);
 * unpack() validates a buffer (8-byte aligned) and fills `view` with pointers INTO it (no copy). */
typedef struct {
    uint32_t coder;            /* 0 = ANS (stack), 1 = range coder (queue)                               */
    uint32_t word_bits;        /* 32 (Default preset) or 16 (Small preset)                               */
    uint32_t precision;        /* 24 or 12                                                               */
    uint32_t checkpoint_every; /* 0 = no records                                                         */
    uint32_t flags;            /* bit 0: interleaved deal (stream k owns symbols k, k+K, ...; sym_offsets NULL) */
    uint32_t reserved;
    uint64_t n_streams, n_symbols, total_words, n_records;
    const uint64_t *sym_offsets; /* u64[K+1] or NULL                                                     */
    const uint64_t *offsets;     /* u64[K+1], in words                                                   */
    const uint64_t *ckpt_offsets;/* u64[K+1] or NULL                                                     */
    const uint64_t *records;     /* n_records x 2 (ANS) or x 4 (range) u64                               */
    const void *words;           /* total_words words of word_bits bits                                  */
} ctr_container_view;
size_t ctr_container_size(const ctr_container_view *view);
int ctr_container_pack(const ctr_container_view *view, void *out, size_t out_bytes);
int ctr_container_unpack(const void *bytes, size_t n_bytes, ctr_container_view *view);

/* ---- multi-GPU exchange of compressed containers (ctr_gather_compressed) ----------------------
 * Streams are independent coders, so encode and decode shard over the GPUs of a node with no communication; the one
 * exchange step is an all-gather of the per-rank containers (what a host does with the per-shard Vec<u32>s that
 * `into_compressed()` returns, stack.rs:891-895 / queue.rs:349-355, cf. tests/issue52.rs:38-53).  The gathered
 * container is SLOTTED: words u32[n_buffers][world][slot_words], offsets u64[n_buffers][world][slot_streams + 1];
 * slot (turn mod n_buffers, r) holds rank r's container of that turn exactly as its encoder wrote it (offsets relative
 * to the slot), so ctr_*_decode(model, slot words, slot offsets, ...) decodes rank r's streams on any rank.
 *
 * Peer-memory implementation (one process per GPU, one node): `words_bases[r]`, `offsets_bases[r]`, `flags_bases[r]`
 * are rank r's receive buffers mapped into THIS process (symmetric memory over NVLink; allocation and exchange of the
 * mappings is the host's business: torch.distributed._symmetric_memory, or cuMemCreate + cuMemExportToShareableHandle).
 * `flags` is u32[3][n_buffers][world], zeroed before the first turn; slot_words a multiple of 4.  Per turn:
 *   ctr_gather_begin_turn   -> turn number q and MY slot in MY buffers: encode straight into it
 *   (the caller enqueues ctr_*_encode with words_out = slot, offsets_out = slot offsets on `encode_stream`)
 *   ctr_gather_push         enqueues an 8-byte size read-back behind the encode and returns at once; a worker thread of
 *                           the library waits for THIS rank's encode, then pushes the used part of the slot to every
 *                           peer with copy-engine transfers (no SM, no kernel) and raises `arrived` flags in the
 *                           peers' memory with cuStreamWriteValue32 -- no cross-rank host wait, no size exchange
 *   ctr_gather_wait         `consumer_stream` waits (cuStreamWaitValue32) until every peer's slot of turn q has arrived
 *   ctr_gather_release      `consumer_stream` tells every peer that this rank is done reading turn q (a peer's push
 *                           of turn q + n_buffers waits for it); also required before THIS rank encodes turn
 *                           q + n_buffers into the same buffer (stream order on the caller's side)
 *   ctr_gather_sync         host waits until all pushes issued so far have completed; returns a data-level error of
 *                           the worker (CTR_ERR_OUT_OF_SPACE if a container exceeded slot_words)                     */
typedef struct ctr_gather_s *ctr_gather_t;
int ctr_gather_create(uint32_t world, uint32_t rank, uint32_t n_buffers, uint64_t slot_words, uint64_t slot_streams,
                      void *const *words_bases, void *const *offsets_bases, void *const *flags_bases, ctr_gather_t *out);
int ctr_gather_destroy(ctr_gather_t g);
int ctr_gather_begin_turn(ctr_gather_t g, uint32_t *turn, uint32_t **words_slot_dev, uint64_t *words_capacity,
                          uint64_t **offsets_slot_dev);
int ctr_gather_slot(ctr_gather_t g, uint32_t turn, uint32_t src_rank, const uint32_t **words_slot_dev,
                    const uint64_t **offsets_slot_dev);
int ctr_gather_push(ctr_gather_t g, uint32_t turn, uint64_t n_streams, void *encode_stream);
int ctr_gather_wait(ctr_gather_t g, uint32_t turn, void *consumer_stream);
int ctr_gather_release(ctr_gather_t g, uint32_t turn, void *consumer_stream);
int ctr_gather_sync(ctr_gather_t g);

/* NCCL implementation of the same exchange on a caller-supplied communicator (`nccl_comm` is an ncclComm_t; NCCL is
 * resolved at run time with dlsym, CTR_NCCL_LIB names the library if the host has not loaded it).  One buffer:
 * words_out_dev u32[world][slot_words], offsets_out_dev u64[world][slot_streams + 1].  meta_dev is u64[2 * world + 2]
 * of device scratch, meta_host u64[2 * world] receives {total words, streams} of every rank.  Enqueues an
 * ncclAllGather of the sizes, WAITS on the host for them (the one synchronisation), then enqueues grouped
 * ncclBroadcasts that write every rank's words and offset table straight into its slot.  Returns
 * CTR_ERR_OUT_OF_SPACE if a rank's container does not fit its slot. */
int ctr_gather_compressed_nccl(void *nccl_comm, uint32_t world, uint32_t rank, const uint32_t *words_dev,
                               const uint64_t *offsets_dev, uint64_t n_streams, uint64_t slot_words,
                               uint64_t slot_streams, uint32_t *words_out_dev, uint64_t *offsets_out_dev,
                               uint64_t *meta_dev, uint64_t *meta_host, void *stream);

/* ---- stream-ordered signalling without any SM (building blocks of the exchange above) ----
 * cuStreamWriteValue32 / cuStreamWaitValue32 (>=) on `stream`: `addr` is a 4-byte aligned device address, which may
 * be a peer GPU's memory mapped into this process (symmetric memory). */
int ctr_stream_write_value32(void *addr, uint32_t value, void *stream);
int ctr_stream_wait_value32(void *addr, uint32_t value, void *stream);

/* ---- launch accounting and kernel timing (bench.py's `gpu_launches` and `roofline`) ----------
 * With profiling enabled the library brackets every main coder kernel (not the compaction helpers)
 * with CUDA events on the launching stream.  ctr_profile_read synchronises those events and returns
 * the summed durations since the last read: which = 0 ANS encode, 1 ANS decode, 2 range encode,
 * 3 range decode. */
uint64_t ctr_kernel_launch_count(void); /* kernels this library has launched in this process */
void ctr_profile_enable(int on);
int ctr_profile_read(int which, double *total_ms, uint64_t *launches);

#ifdef __cplusplus
}
#endif
#endif /* CONSTRICTION_B200_H */
