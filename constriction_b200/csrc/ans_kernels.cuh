// ans_kernels.cuh -- batched rANS encode / decode kernels (K1 / K2), one lane per independent coder.
//
// Execution model.  A batch holds K independent coders ("streams").  Lane l of warp w owns stream
// k = 32 w + l for the whole kernel and keeps that coder's 64-bit state in registers, so the
// loop-carried dependency of the reference's per-symbol loop (stream/mod.rs:592-607,1274-1297) is
// private to a lane and 32 such chains advance per warp instruction.  Everything that touches HBM is
// warp-cooperative and coalesced:
//   - symbols: interleaved layout -> one 128-byte row per warp step (pointer bumped by K per step);
//     contiguous layout -> 32x32 tiles transposed through shared memory;
//   - compressed words: each lane appends to / pops from a small private ring in shared memory; the ring
//     is drained / topped up 16 bytes at a time with lane-private vector accesses (encoder: LDS.128 +
//     STG.128 into the lane's scratch region; decoder: asynchronous LDGSTS.128 from the lane's stream),
//     so there is no warp-cooperative phase, no shuffle and no divergence in the word path; the L2
//     merges the two 16-byte halves of each 32-byte sector;
//   - the model: for a single shared model the encoder table (left, prob, 64-bit reciprocal) or the
//     decoder table (CDF pairs + 4096-bucket quantile index) is staged into shared memory by the TMA
//     engine (cp.async.bulk); model sets too big for that are read through L1/L2 (`ld.global.nc`).
//
// The kernels are issue-bound (integer pipe), so the per-symbol instruction count is what matters:
//   - hot loops have uniform trip counts (the ragged last row of the interleaved deal is peeled off);
//   - lanes that own no stream run the same instructions on clamped addresses (their pushes are
//     disabled through a shift amount, their stores predicated off);
//   - data errors never branch: an out-of-range symbol is clamped to a sentinel table entry with
//     probability 0 and detected at the end from a running minimum;
//   - shared memory is addressed with 32-bit shared addresses (no generic-pointer arithmetic);
//   - the word rings are inspected once per kCheckEvery symbols with straight-line predicated code.
//
// Per-stream results equal the reference's `AnsCoder` (src/stream/stack.rs:1014-1100) word for word.
#pragma once
#include <cuda.h>

#include <type_traits>

#include "compact.cuh"
#include "device_utils.cuh"
#include "gauss_kernels.cuh"

namespace ctr {

constexpr int kAnsBlock = 256;  // threads per CTA (8 warps)
// Interleaved deal with one model per stream: symbols move between HBM and shared memory as 2-D TMA boxes of
// kBoxRows rows x 32 streams per warp (device_utils.cuh), kEncBoxSlots / kDecBoxSlots boxes in flight per warp.
#ifndef CTR_ENC_BOX_SLOTS
#define CTR_ENC_BOX_SLOTS 2
#endif
#ifndef CTR_DEC_BOX_SLOTS
#define CTR_DEC_BOX_SLOTS 2
#endif
constexpr int kEncBoxSlots = CTR_ENC_BOX_SLOTS;
constexpr int kDecBoxSlots = CTR_DEC_BOX_SLOTS;
constexpr uint32_t kMaxSharedAlphabet = 4095;      // decoder: bigger alphabets use the global-table path
constexpr uint32_t kMaxSharedEncAlphabet = 511;    // encoder: its shared table is replicated 8x (128 B per entry)

struct ModelView {
    const uint32_t *cdf;   // [n_models][alphabet + 1]
    const uint4 *enc;      // [n_models][alphabet + 1] {left, prob, reciprocal lo, hi}; entry [alphabet] is
                           // the all-zero sentinel that out-of-range symbols are clamped to
    const uint4 *enc_rep;  // model 0 only: [alphabet + 1][8] -- each entry 8 times, one copy per 16-byte bank group
    const uint8_t *cidx;   // [n_models][kCoarseSize + 1] u8 (alphabet <= 256) or u16: coarse quantile index for
                           // decoding with global tables; cidx[m][b] = symbol of model m containing quantile b << 16
    const uint32_t *dec;   // model 0 only: quantile index uint2[kLutSize] ++ cdf u32[alphabet + 2] (padded to 16 B)
    const uint32_t *dec_big;  // the same with 2^kBigLutBits buckets (large-batch ANS decoder); nullptr until first used
    uint32_t n_models;
    uint32_t alphabet;
    int32_t min_symbol;
    uint32_t dec_cdf_bytes;    // size of the cdf part of `dec`: (alphabet + 2) * 4 rounded up to 16
    // pool decoders (a whole model set staged in shared memory): sizes of `cdf` and `cidx`, rounded up to 16
    uint32_t pool_cdf_bytes, pool_cidx_bytes;
};

// where a decoder kernel finds its model tables
constexpr int kTableGlobal = 0;  // read through L1/L2
constexpr int kTableLut = 1;     // one model for the batch: quantile index + cdf in shared memory
constexpr int kTablePool = 2;    // several models, all CDF rows + coarse indices in shared memory
constexpr int kTableGauss = 3;   // no table: model_index[i] names the (mean, std) pair of symbol i (gauss_kernels.cuh)

struct AnsParams {
    ModelView model;
    uint64_t K, N;
    const uint64_t *sym_off;      // nullptr -> interleaved
    const uint32_t *model_index;  // per symbol / per stream / nullptr
    int index_mode;
    uint32_t flags;
    const int32_t *symbols_in;    // encode
    int32_t *symbols_out;         // decode
    const uint64_t *states_in;
    uint64_t *states_out;
    uint32_t *status;
    // encode
    uint32_t *scratch;      // strided per-stream regions
    CompactParams compact;  // fused compaction into the dense container
    // decode
    const uint32_t *words;
    const uint64_t *offsets;
    uint64_t *words_left;
    // checkpoints (contiguous layout): the encoders record their coder position every ckpt_every symbols
    // (Pos::pos, stack.rs:1107-1115 / queue.rs:182-196) so that one stream can be decoded by many lanes;
    // ckpt_off[k] = index of stream k's first record in ckpt_out (u64[K+1])
    uint32_t ckpt_every;  // multiple of 32; 0 = none
    const uint64_t *ckpt_off;
    uint64_t *ckpt_out;   // ANS: {words pushed, state} per record; range: {words pushed, lower, range, 0}
    // decoders: stream k's words end at ends[k] instead of offsets[k + 1] (virtual streams of a chunked decode)
    const uint64_t *ends;
    // kTableGauss decoders: per-symbol Gaussian parameters and the quantiser's free weight (quantize.rs:284-308)
    const double *gauss_means, *gauss_stds;
    double gauss_free_weight;
    // interleaved deal: the symbol array as a [full rows][K] int32 tensor (boxes of kBoxRows x 32), if use_tma
    uint32_t use_tma;
    alignas(64) CUtensorMap tmap;
};

// ---- warp-cooperative tile I/O for the contiguous layout and the range kernels (generic pointers) ----

// Write the first count_i words of row i to dst_i, for every lane i in `mask` (rows of kRowStride words).
static __device__ __noinline__ void warp_flush_rows(unsigned mask, const uint32_t *rows, uint32_t *dst, uint32_t count,
                                             int lane) {
    __syncwarp();
    while (mask) {
        const int i = __ffs(mask) - 1;
        mask &= mask - 1;
        uint32_t *d = (uint32_t *)shfl_u64((uint64_t)dst, i);
        const uint32_t c = __shfl_sync(kFullMask, count, i);
        if ((uint32_t)lane < c) st_stream_u32(d + lane, rows[i * kRowStride + lane]);
    }
    __syncwarp();
}

// Fill row i with the `count_i` words at src_i, for every lane i in `mask` (count_i == 0 for the others).
// Eight rows are in flight at a time: the loads of a group are all issued before the first store, so a
// tile costs four global round trips instead of thirty-two.
static __device__ __noinline__ void warp_fill_rows(unsigned mask, uint32_t *rows, const uint32_t *src, uint32_t count,
                                            int lane) {
    __syncwarp();
    for (int i0 = 0; i0 < 32; i0 += 8) {
        if (((mask >> i0) & 0xffu) == 0u) continue;
        uint32_t v[8], c[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t *s = (const uint32_t *)shfl_u64((uint64_t)src, i0 + j);
            c[j] = __shfl_sync(kFullMask, count, i0 + j);
            v[j] = 0u;
            if ((uint32_t)lane < c[j]) v[j] = ld_stream_u32(s + lane);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if ((uint32_t)lane < c[j]) rows[(i0 + j) * kRowStride + lane] = v[j];
    }
    __syncwarp();
}

// Asynchronous fill (LDGSTS, 4 bytes per lane): row i of the tile at shared address `rows` <- the count_i words
// at src_i.  The caller commits the cp.async group and, before reading the tile, waits for it and
// synchronises the warp (every lane writes into every row).
__device__ __forceinline__ void warp_fill_rows_async(uint32_t rows, const uint32_t *src, uint32_t count, int lane) {
    rows += (uint32_t)lane * 4u;
#pragma unroll 8
    for (int i = 0; i < 32; ++i) {
        const uint32_t *s = (const uint32_t *)shfl_u64((uint64_t)src, i);
        const uint32_t c = __shfl_sync(kFullMask, count, i);
        if ((uint32_t)lane < c) cp_async_4(rows + (uint32_t)i * (kRowStride * 4u), s + lane);
    }
}

// ---- model lookups ----------------------------------------------------------------------------------

// Decoder table of a shared model (built by build_dec_table_kernel), staged in shared memory:
//   lut[b] (8 bytes) describes the symbol s that contains the first quantile of bucket b (b = q >> 12):
//       x = cdf[s] | (s & 0xff) << 24,   y = cdf[s+1] | (s >> 8) << 25
//     One 8-byte load resolves every quantile whose bucket does not reach beyond s.  Otherwise the
//     quantile belongs to a later symbol; the (few) lanes concerned read cdf[s+2] and step to s+1, and
//     only if that is not enough either (several tiny-probability symbols in one bucket) a cold binary
//     search runs.
//   cdf[0 .. alphabet] is the plain CDF row, cdf[alphabet + 1] = 2^24 pads the last probe.
static __device__ __noinline__ uint32_t lookup_far_cold(uint32_t cdf_addr, uint32_t alphabet, uint32_t s, uint32_t q) {
    uint32_t lo = s + 1, hi = alphabet - 1;  // cdf[s + 1] <= q is known
    while (lo < hi) {
        const uint32_t mid = (lo + hi + 1) >> 1;
        if (lds_table_u32(cdf_addr + mid * 4u) <= q)
            lo = mid;
        else
            hi = mid - 1;
    }
    return lo;
}

// `word` is any value whose low 24 bits are the quantile q; SMALL: alphabet <= 256 (s fits the top byte of x)
//   BITS    : log2 of the number of buckets of the index at lut_addr
//   UNIFORM : the second probe sits behind a warp-uniform branch (fine indices: no lane needs it in most warp steps)
template <bool SMALL, int BITS = kLutBits, bool UNIFORM = false>
__device__ __forceinline__ uint32_t lookup_shared(uint32_t lut_addr, uint32_t cdf_addr, uint32_t alphabet, uint32_t word,
                                                  uint32_t q, uint32_t &left, uint32_t &right) {
    constexpr int kShift = kPrecision - BITS;
    const uint2 e = lds_table_v2(lut_addr + ((word >> (kShift - 3)) & (((1u << BITS) - 1u) << 3)));
    uint32_t s = e.x >> 24;
    left = e.x & kQuantileMask;
    right = e.y;
    if (!SMALL) {
        s |= (e.y >> 25) << 8;
        right = e.y & 0x1ffffffu;
    }
    const bool beyond = q >= right;  // the bucket straddles the boundary and q lies past it
    if (!UNIFORM || __any_sync(kFullMask, beyond)) {
        uint32_t next = right;
        if (beyond) next = lds_table_u32(cdf_addr + (s + 2) * 4u);
        left = beyond ? right : left;
        right = beyond ? next : right;
        s += beyond ? 1u : 0u;
        if (q >= right) {
            s = lookup_far_cold(cdf_addr, alphabet, s, q);
            left = lds_table_u32(cdf_addr + s * 4u);
            right = lds_table_u32(cdf_addr + s * 4u + 4u);
        }
    }
    return s;
}

// decoder with global tables (model sets that do not fit shared memory, read through L1/L2):
// the last index s with cdf[s] <= q  (categorical/contiguous.rs:628-665: partition point - 1).
// A 257-entry coarse index per model (cidx[b] = symbol containing quantile b << 16) narrows the binary search
// to the symbols that intersect q's bucket, so a lookup costs ~3 dependent loads instead of log2(alphabet) + 2
// -- for a 1 GB pool of CDF rows each of them is a DRAM round trip.
constexpr int kCoarseBits = 8;
constexpr int kCoarseSize = 1 << kCoarseBits;
constexpr int kCoarseShift = kPrecision - kCoarseBits;

__device__ __forceinline__ uint32_t lookup_global(const uint32_t *row, const uint8_t *cidx_row, bool wide,
                                                  uint32_t alphabet, uint32_t q, uint32_t &left, uint32_t &right) {
    uint32_t lo = 0, hi = alphabet - 1;
    if (cidx_row != nullptr) {
        const uint32_t b = q >> kCoarseShift;
        if (wide) {
            const uint16_t *c = reinterpret_cast<const uint16_t *>(cidx_row);
            lo = __ldg(c + b);
            hi = __ldg(c + b + 1);
        } else {
            lo = __ldg(cidx_row + b);
            hi = __ldg(cidx_row + b + 1);
        }
    }
    while (lo < hi) {
        const uint32_t mid = (lo + hi + 1) >> 1;
        if (__ldg(row + mid) <= q)
            lo = mid;
        else
            hi = mid - 1;
    }
    left = __ldg(row + lo);
    right = __ldg(row + lo + 1);
    return lo;
}

// The same search on a model set staged in shared memory (`row` / `cidx_row` are shared addresses).  Two load
// levels in the common case: the coarse index names the first symbol `lo` of q's bucket, then cdf[lo .. lo+2]
// are read together and q is in symbol lo or lo + 1 unless three or more symbols share the bucket (cold).
static __device__ __noinline__ uint32_t lookup_pool_cold(uint32_t row, uint32_t lo, uint32_t hi, uint32_t q) {
    while (lo < hi) {
        const uint32_t mid = (lo + hi + 1) >> 1;
        if (lds_table_u32(row + mid * 4u) <= q)
            lo = mid;
        else
            hi = mid - 1;
    }
    return lo;
}
__device__ __forceinline__ uint32_t lookup_pool(uint32_t row, uint32_t cidx_row, bool wide, uint32_t q, uint32_t &left,
                                                uint32_t &right) {
    const uint32_t b = q >> kCoarseShift;
    uint32_t lo, hi;
    if (wide) {
        lo = lds_table_u16(cidx_row + b * 2u);
        hi = lds_table_u16(cidx_row + b * 2u + 2u);
    } else {
        lo = lds_table_u8(cidx_row + b);
        hi = lds_table_u8(cidx_row + b + 1u);
    }
    // (cdf[lo + 2] may be one word past the row when lo is the last symbol; it is not used then)
    const uint32_t at = row + lo * 4u;
    const uint32_t c0 = lds_table_u32(at), c1 = lds_table_u32(at + 4u), c2 = lds_table_u32(at + 8u);
    const bool second = q >= c1;
    uint32_t s = lo + (second ? 1u : 0u);
    left = second ? c1 : c0;
    right = second ? c2 : c1;
    if (second && hi > lo + 1u && q >= c2) {
        s = lookup_pool_cold(row, lo + 2u, hi, q);
        left = lds_table_u32(row + s * 4u);
        right = lds_table_u32(row + s * 4u + 4u);
    }
    return s;
}

// Geometry of the interleaved deal shared by all kernels.
struct Interleave {
    uint64_t T;     // rows (symbols of the longest stream)
    uint64_t last;  // streams that own a symbol in row T-1 (1..K), 0 if N == 0
};
__device__ __forceinline__ Interleave interleave_of(uint64_t N, uint64_t K) {
    Interleave g;
    g.T = (N + K - 1) / K;
    g.last = g.T ? N - (g.T - 1) * K : 0;
    return g;
}

__device__ __forceinline__ uint64_t warp_max_u64(uint64_t v, int lane) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const uint64_t o = shfl_u64(v, lane ^ d);
        v = o > v ? o : v;
    }
    return v;
}

// =====================================================================================================
// encode
// =====================================================================================================
//   SHARED : model 0's encoder table lives in shared memory (index_mode == NONE, small alphabet)
//   CONTIG : stream k owns symbols[sym_off[k] .. sym_off[k+1]) (else interleaved deal)
//   PERSYM : a model index per symbol (else one model per stream / model 0)
//   F64DIV : the table holds double-precision reciprocals and the quotient estimate uses the FP64 pipe
//   BLOCK  : threads per CTA (kAnsBlock, or kSmallBlock for batches too small to fill the GPU with big CTAs)
template <int BLOCK, bool SHARED, bool CONTIG, bool PERSYM, bool F64DIV>
#ifndef CTR_ENC_MIN_CTAS
#define CTR_ENC_MIN_CTAS 4
#endif
__global__ void __launch_bounds__(BLOCK, BLOCK >= 256 ? CTR_ENC_MIN_CTAS : 8) ans_encode_kernel(const __grid_constant__ AnsParams p) {
    extern __shared__ __align__(128) uint32_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint64_t tma_bar[BLOCK / 32][kEncBoxSlots];

    const int lane = threadIdx.x & 31;
    const int warp_in_cta = threadIdx.x >> 5;
    constexpr int kWarpsPerCta = BLOCK / 32;

    // shared memory carve-up: [lane rings (32 B each)][table][symbol tiles / TMA boxes][index tiles]
    const uint32_t alphabet = p.model.alphabet;
    const uint32_t table_words = SHARED ? (alphabet + 1) * 32 : 0;
    constexpr uint32_t kEncRingBytes = kAnsEncRingWords * 4u;  // (shadows the 8-word constant of the other encoders)
    constexpr uint32_t kRingsWords = BLOCK * kAnsEncRingWords;
    const uint32_t ring = smem_u32_pinned(smem) + threadIdx.x * kEncRingBytes;  // 32-byte aligned
    auto ring_slot = [&](uint32_t byte_pos) -> uint32_t { return ring | (byte_pos & (kEncRingBytes - 1u)); };
    // lane l reads copy (l & 7) of an entry: the 8 lanes of a quarter-warp always hit 8 different 16-byte bank
    // groups, so the random-index LDS.128 is conflict free
    const uint32_t table_addr = smem_u32_pinned(smem + kRingsWords) + (uint32_t)(lane & 7) * 16u;
    // contiguous layout: two symbol tiles per warp (double buffered), then one index tile per warp
    uint32_t *sym_tile = smem + kRingsWords + table_words + warp_in_cta * (2 * kTileWords);
    uint32_t *idx_tile = smem + kRingsWords + table_words + kWarpsPerCta * (2 * kTileWords) + warp_in_cta * kTileWords;

    // (the copy runs during the prologue: it is waited for right before the first symbol is coded)
    if (SHARED) stage_table_begin(smem + kRingsWords, p.model.enc_rep, (alphabet + 1) * 128u, &bar);

    const uint64_t K = p.K, N = p.N;
    const uint32_t tile = take_tile_ticket(p.compact.ticket);  // which BLOCK streams this CTA codes
    const uint64_t k = (uint64_t)tile * BLOCK + threadIdx.x;
    const bool valid = k < K;
    const uint64_t kc = valid ? k : K - 1;  // lanes without a stream shadow the last one (loads only)

    // stream geometry and my scratch region
    uint64_t n_k, o_k;
    if (CONTIG) {
        o_k = p.sym_off[kc];
        n_k = p.sym_off[kc + 1] - o_k;
        if (o_k > N || n_k > N - o_k) {  // offsets outside the symbol array (or decreasing): flag, code nothing
            if (valid) report_error(p.status, kErrBadArgument, k);
            o_k = 0;
            n_k = 0;
        }
    } else {
        n_k = interleaved_len(N, K, kc);
        o_k = interleaved_start(N, K, kc);
    }
    // my scratch region (128-byte aligned start) and its capacity in bytes; the write cursor is gbase + drained
    char *gbase;
    uint32_t cap;
    {
        uint32_t *const gbegin = p.scratch + scratch_start(o_k, k);
        const uint64_t r = valid ? scratch_start(o_k + n_k, k + 1) - scratch_start(o_k, k) : 0;
        cap = r > 0x1ffffff0u ? 0x7fffffc0u : (uint32_t)r * 4u;
        gbase = reinterpret_cast<char *>(gbegin);
    }

    uint64_t state = (valid && p.states_in) ? p.states_in[k] : 0;
    uint32_t pushed = 0;             // bytes pushed into my ring so far (ring slot = pushed & 31); bit 31: words dropped
    uint32_t drained = 0;            // bytes of my ring already written to scratch
    uint32_t min_prob = 0xffffffffu; // running minimum of the probabilities used (0 <=> impossible symbol)
    const uint32_t push_shift = valid ? 8u : 32u;  // (state >> 32) >> 32 == 0: lanes without a stream never push
    const uint32_t stream_model = (p.index_mode == 2) ? p.model_index[kc] : 0u;
    const uint32_t n_models = p.model.n_models;
    const uint32_t min_symbol = (uint32_t)p.model.min_symbol;

    // one reference encode_symbol (stack.rs:1014-1048)
    // symbol -> table index; out-of-range symbols map to the sentinel entry [alphabet]
    auto index_of = [&](int32_t sym) -> uint32_t { return min((uint32_t)sym - min_symbol, alphabet); };
    auto lookup = [&](uint32_t idx, uint32_t m) -> uint4 {
        if (SHARED) return lds_table_v4(table_addr + idx * 128u);
        const bool ok = m < n_models;
        idx = ok ? idx : alphabet;
        m = ok ? m : 0u;
        return __ldg(p.model.enc + (uint64_t)m * (alphabet + 1) + idx);
    };
    auto encode_entry_nomin = [&](const uint4 &e) {
        // stack.rs:1035-1040 as one block of straight-line PTX: the push is predicated, and the renormalised state is
        // selected straight into the register pair the conversion reads (the compiler's own code moves it around)
        uint64_t n;
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b32 a, b, t;\n\t"
            "shr.u32 t, %3, %5;\n\t"
            "setp.ge.u32 p, t, %4;\n\t"
            "@p st.shared.u32 [%6], %2;\n\t"
            "@p add.u32 %1, %1, 4;\n\t"
            "selp.b32 a, %3, %2, p;\n\t"
            "selp.b32 b, 0, %3, p;\n\t"
            "mov.b64 %0, {a, b};\n\t}"
            : "=l"(n), "+r"(pushed)
            : "r"((uint32_t)state), "r"((uint32_t)(state >> 32)), "r"(e.y), "r"(push_shift),
              "r"(ring_slot(pushed))
            : "memory");
        state = ans_encode_recombine(n, ans_quotient_estimate<F64DIV>(n, e.z, e.w), e.x, e.y);
    };
    auto encode_entry = [&](const uint4 &e) {
        min_prob = min(min_prob, e.y);
        encode_entry_nomin(e);
    };
    // two symbols, one three-input minimum for the impossible-symbol check
    auto encode_pair = [&](uint32_t idx0, uint32_t idx1, uint32_t m) {
        const uint4 e0 = lookup(idx0, m), e1 = lookup(idx1, m);
        encode_entry_nomin(e0);
        encode_entry_nomin(e1);
        min_prob = __vimin3_u32(min_prob, e0.y, e1.y);
    };
    auto encode_idx = [&](uint32_t idx, uint32_t m) { encode_entry(lookup(idx, m)); };
    auto encode_one = [&](int32_t sym, uint32_t m) { encode_idx(index_of(sym), m); };

    // every kCheckEvery symbols: a complete 16-byte group of my ring goes to my scratch region.  A lane
    // pushes at most kCheckEvery words in between, so one group per check keeps up with any input.
    // bytes in my ring that are not yet written to scratch
    auto in_ring = [&]() -> uint32_t { return (pushed - drained) & 0x7fffffffu; };
    auto drain_store = [&](const uint4 &v) {
        if (in_ring() >= 16u) {
            if (drained + 16u <= cap)
                st_stream_v4(gbase + drained, v);
            else
                pushed |= 0x80000000u;  // words dropped: the stream is flagged at the end
            drained += 16u;
        }
    };
    // the oldest (possibly incomplete) 16-byte group of my ring; harmless to read when it is not yet complete
    auto drain_load = [&]() -> uint4 {
        return lds_v4(ring | (drained & (kEncRingBytes - 16u)));
    };
    auto drain_ring = [&]() { drain_store(drain_load()); };
    // split form: `full = in_ring() >= 16` and `oldest = drain_load()` are taken at the check, the store is issued
    // a couple of symbols later so that the shared-memory latency is covered by coding work (a group that
    // completes in between waits for the next check; the ring has room for that)
    auto drain_decided = [&](bool full, const uint4 &oldest) {
        if (full) {
            if (drained + 16u <= cap)
                st_stream_v4(gbase + drained, oldest);
            else
                pushed |= 0x80000000u;
            drained += 16u;
        }
    };

    if (!CONTIG) {
        // ---- interleaved deal: row t holds symbols[t*K .. t*K+K); coded from the last row backwards ---
        const Interleave g = interleave_of(N, K);
        // TMA path (one model per stream): the warp's column strip arrives as boxes of kBoxRows rows (one UTMALDG by lane 0
        // per box, completion on a per-slot mbarrier); lanes read their column with LDS.  kEncBoxSlots boxes are in
        // flight per warp, which covers the HBM latency without holding any register.  The first boxes are requested
        // before anything is coded, and the symbols that are not part of a box (the ragged last row, the rows above the
        // highest box) are all loaded before the first of them is used: one HBM round trip instead of up to nine.
        const bool use_boxes = !PERSYM && p.use_tma && g.T > 1;
        const uint64_t rows_total = g.T > 1 ? g.T - 1 : 0;  // full rows T-2 .. 0
        uint32_t nbox = use_boxes ? (uint32_t)(rows_total / kBoxRows) : 0u;
        const uint32_t top_rows = use_boxes ? (uint32_t)(rows_total - (uint64_t)nbox * kBoxRows) : 0u;
        const uint32_t bars = smem_u32(&tma_bar[warp_in_cta][0]);
        const uint32_t boxes = smem_u32_pinned(smem + kRingsWords + table_words) + (uint32_t)warp_in_cta * (kEncBoxSlots * kBoxBytes);
        const int32_t x0 = (int32_t)((uint32_t)tile * BLOCK + (uint32_t)warp_in_cta * 32u);  // my warp's first stream
        uint32_t next = nbox;  // boxes [0, next) are not requested yet; they are taken from the top
        auto request_box = [&](uint32_t slot) {
            if (next != 0u) {
                next -= 1u;
                if (lane == 0) {
                    mbar_expect_tx_addr(bars + 8u * slot, kBoxBytes);
                    tma_load_box(boxes + slot * kBoxBytes, &p.tmap, x0, (int32_t)(next * kBoxRows), bars + 8u * slot);
                }
            }
        };
        if (use_boxes) {
            if (lane == 0) {
#pragma unroll
                for (int sl = 0; sl < kEncBoxSlots; ++sl) mbar_init_addr(bars + 8u * sl, 1);
                fence_mbar_init();
            }
            __syncwarp();
#pragma unroll
            for (int sl = 0; sl < kEncBoxSlots; ++sl) request_box(sl);
        }
        const bool has_last = g.T > 0 && valid && k < g.last;
        int32_t last_sym = 0, top_sym[kBoxRows - 1];
        uint32_t last_model = stream_model;
        if (has_last) {
            const uint64_t i = (g.T - 1) * K + k;
            last_sym = ld_stream_s32(p.symbols_in + i);
            if (PERSYM) last_model = ld_stream_u32(p.model_index + i);
        }
        if (use_boxes) {
            const int32_t *ps = p.symbols_in + (g.T - 2) * K + kc;
#pragma unroll
            for (int j = 0; j < kBoxRows - 1; ++j)
                if (j < (int)top_rows) top_sym[j] = ld_stream_s32(ps - (uint64_t)j * K);
        }
        if (SHARED) stage_table_wait(&bar);
        if (has_last) encode_one(last_sym, last_model);  // ragged last row
        if (use_boxes) {
            // the rows above the highest box
#pragma unroll
            for (int j = 0; j < kBoxRows - 1; ++j) {
                if (j < (int)top_rows) {
                    if ((j & (kCheckEvery - 1)) == 0) drain_ring();
                    encode_one(top_sym[j], stream_model);
                }
            }
            drain_ring();
            const uint32_t my_col = boxes + (uint32_t)lane * 4u;
            // one box: wait for it, code its rows top down, hand the slot back to the TMA engine
            auto code_box = [&](uint32_t slot, uint32_t parity) {
                mbar_wait_addr(bars + 8u * slot, parity);
                const uint32_t box = my_col + slot * kBoxBytes;
#pragma unroll
                for (int half = kBoxRows / kCheckEvery - 1; half >= 0; --half) {  // batches of rows, top down
                    uint32_t idx[kCheckEvery];
#pragma unroll
                    for (int u = 0; u < kCheckEvery; ++u)
                        idx[u] = index_of((int32_t)lds_u32(box + (uint32_t)(half * kCheckEvery + (kCheckEvery - 1 - u)) * 128u));
                    const bool full = in_ring() >= 16u;
                    const uint4 oldest = drain_load();
                    encode_pair(idx[0], idx[1], stream_model);
                    drain_decided(full, oldest);
                    encode_pair(idx[2], idx[3], stream_model);
                }
                __syncwarp();  // every lane has read the box: its slot is requested again
                request_box(slot);
            };
            // the slots in turn with compile-time slot numbers (no slot / parity arithmetic in the loop)
            static_assert(kEncBoxSlots == 2, "the unrolled loop is written for two slots");
            uint32_t parity = 0;
            for (; nbox >= 2u; nbox -= 2u) {
                code_box(0u, parity);
                code_box(1u, parity);
                parity ^= 1u;
            }
            if (nbox != 0u) code_box(0u, parity);
            drain_ring();
        } else
        if (g.T > 1) {
            const uint64_t rows_total = g.T - 1;  // full rows T-2 .. 0
            const char *ps = reinterpret_cast<const char *>(p.symbols_in + (g.T - 2) * K + kc);
            const char *pm = PERSYM ? reinterpret_cast<const char *>(p.model_index + (g.T - 2) * K + kc) : nullptr;
            uint64_t row_bytes = K * 4u;  // distance between consecutive symbols of a stream
            asm volatile("" : "+l"(row_bytes));
            // batches of kCheckEvery symbols; the loads of the next batch are in flight while this one is coded
            int32_t buf[2][kCheckEvery];
            uint32_t mbuf[2][kCheckEvery];
            auto load_batch = [&](int which) {
#pragma unroll
                for (int u = 0; u < kCheckEvery; ++u) {
#ifdef CTR_DBG_NO_LOAD
                    buf[which][u] = (int32_t)(((uint32_t)(uintptr_t)ps >> 7) % 23u) - 8;
#else
                    buf[which][u] = ld_stream_s32(reinterpret_cast<const int32_t *>(ps));
#endif
                    ps -= row_bytes;
                    if (PERSYM) {
                        mbuf[which][u] = ld_stream_u32(reinterpret_cast<const uint32_t *>(pm));
                        pm -= row_bytes;
                    } else {
                        mbuf[which][u] = stream_model;
                    }
                }
            };
            // L2 prefetch kPrefetchBatches batches ahead: one instruction covers the warp's four 128-byte lines
            // of a future batch (lanes 8j..8j+7 address row j of that batch)
#ifndef CTR_PF_BATCHES
#define CTR_PF_BATCHES 6
#endif
            constexpr int kPrefetchBatches = CTR_PF_BATCHES;
            const char *pf = ps - ((uint64_t)kPrefetchBatches * kCheckEvery + (uint32_t)(lane >> 3)) * row_bytes;
            const char *const pf_floor = reinterpret_cast<const char *>(p.symbols_in);
            const uint64_t batch_bytes = row_bytes * kCheckEvery;
            static_assert(kCheckEvery == 4, "code_batch is written for batches of four");
            auto code_batch = [&](int which, bool load_next) {
                // Consume this batch's loads first (they were issued a whole batch ago), and only then
                // issue the next batch's: a scoreboard wait on "my" loads must never cover loads that
                // were issued a moment ago.  The empty asm statements pin that order.
                uint32_t idx[kCheckEvery];
#pragma unroll
                for (int u = 0; u < kCheckEvery; ++u) {
                    idx[u] = index_of(buf[which][u]);
                    asm volatile("" : "+r"(idx[u]));
                }
                if (load_next) load_batch(which ^ 1);
                if (kPrefetchBatches > 0) {
                    if (pf >= pf_floor) prefetch_l2(pf);
                    pf -= batch_bytes;
                }
                // drain: the 16-byte group is read from the ring now and stored two symbols later, so the
                // shared-memory latency is covered by coding work.  The decision is taken now (a group that
                // completes during this batch waits for the next check; the ring has room for that).
                const bool full = in_ring() >= 16u;
                const uint4 oldest = drain_load();
                encode_idx(idx[0], mbuf[which][0]);
                encode_idx(idx[1], mbuf[which][1]);
                drain_decided(full, oldest);
                encode_idx(idx[2], mbuf[which][2]);
                encode_idx(idx[3], mbuf[which][3]);
            };
            // (a stream of >= 2^34 symbols is split by the caller; 32-bit counters keep the loop lean)
            uint32_t batches = (uint32_t)(rows_total / kCheckEvery);
            uint32_t rows_left = (uint32_t)(rows_total - (uint64_t)batches * kCheckEvery);
            if (batches > 0) {
                load_batch(0);
                while (batches > 2) {
                    code_batch(0, true);
                    code_batch(1, true);
                    batches -= 2;
                }
                if (batches == 2) {
                    code_batch(0, true);
                    code_batch(1, false);
                } else {
                    code_batch(0, false);
                }
            }
            drain_ring();
            while (rows_left > 0) {  // at most kCheckEvery-1 more symbols
                const int32_t sym = ld_stream_s32(reinterpret_cast<const int32_t *>(ps));
                ps -= row_bytes;
                uint32_t m = stream_model;
                if (PERSYM) {
                    m = ld_stream_u32(reinterpret_cast<const uint32_t *>(pm));
                    pm -= row_bytes;
                }
                encode_one(sym, m);
                rows_left -= 1;
            }
        }
    } else {
        if (SHARED) stage_table_wait(&bar);
        // ---- contiguous: 32x32 tiles, transposed through shared memory ---------------------------
        // The tile of the next round is filled asynchronously (LDGSTS) while this one is coded.  Within a
        // tile, groups of four symbols are looked up first (independent of the coder state) and then coded,
        // so that the loop-carried chain holds the state update only.
        const uint32_t tiles_addr = smem_u32(sym_tile);
        const uint32_t my_row = (uint32_t)lane * (kRowStride * 4u);
        const uint32_t idx_row = smem_u32(idx_tile) + my_row;
        uint64_t remaining = n_k;  // symbols of my stream not yet requested (I consume from the end)
        const uint32_t ckpt_every = valid ? p.ckpt_every : 0u;
        const uint64_t ckpt_base = ckpt_every ? p.ckpt_off[k] : 0;
        uint32_t to_ckpt = ckpt_every;  // symbols until the next checkpoint
        const uint64_t rounds = (warp_max_u64(n_k, lane) + 31) / 32;
        uint32_t c_next = remaining < 32 ? (uint32_t)remaining : 32u;
        remaining -= c_next;
        warp_fill_rows_async(tiles_addr, reinterpret_cast<const uint32_t *>(p.symbols_in + o_k + remaining), c_next, lane);
        cp_async_commit();
        for (uint64_t r = 0; r < rounds; ++r) {
            const uint32_t c = c_next;
            const uint64_t first = remaining;  // my symbols of this round are [o_k + first, o_k + first + c)
            const uint32_t row = tiles_addr + (uint32_t)(r & 1) * (kTileWords * 4u) + my_row;
            c_next = remaining < 32 ? (uint32_t)remaining : 32u;
            remaining -= c_next;
            if (r + 1 < rounds)
                warp_fill_rows_async(tiles_addr + (uint32_t)((r + 1) & 1) * (kTileWords * 4u),
                                     reinterpret_cast<const uint32_t *>(p.symbols_in + o_k + remaining), c_next, lane);
            cp_async_commit();
            if (PERSYM) {
                warp_fill_rows_async(smem_u32(idx_tile), p.model_index + o_k + first, c, lane);
                cp_async_commit();
                cp_async_wait_group<0>();
            } else {
                cp_async_wait_group<1>();
            }
            __syncwarp();
            const uint32_t cmin = __reduce_min_sync(kFullMask, c), cmax = __reduce_max_sync(kFullMask, c);
            uint32_t s = 0;
            for (; s + kCheckEvery <= cmin; s += kCheckEvery) {  // every lane owns all four symbols
                const bool full = in_ring() >= 16u;
                const uint4 oldest = drain_load();
                uint4 e[kCheckEvery];
#pragma unroll
                for (int u = 0; u < kCheckEvery; ++u) {
                    const uint32_t at = (c - 1u - s - (uint32_t)u) * 4u;
                    e[u] = lookup(index_of((int32_t)lds_u32(row + at)), PERSYM ? lds_u32(idx_row + at) : stream_model);
                }
                encode_entry(e[0]);
                encode_entry(e[1]);
                drain_decided(full, oldest);
                encode_entry(e[2]);
                encode_entry(e[3]);
            }
            for (; s < cmax; ++s) {  // ragged end of the round
                if ((s & (kCheckEvery - 1)) == 0) drain_ring();
                if (s < c) {
                    const uint32_t at = (c - 1u - s) * 4u;
                    encode_one((int32_t)lds_u32(row + at), PERSYM ? lds_u32(idx_row + at) : stream_model);
                }
            }
            // checkpoint: every symbol from `first` on is coded.  Record j (decode order) starts the chunk of
            // symbols [first, ...): j = ceil(first / ckpt_every); boundaries are counted from the stream's end.
            if (ckpt_every != 0u && c != 0u) {
                to_ckpt -= c;
                if (to_ckpt == 0u || first == 0u) {
                    uint64_t *rec = p.ckpt_out + 2u * (ckpt_base + (first + ckpt_every - 1u) / ckpt_every);
                    rec[0] = (pushed & 0x7fffffffu) >> 2;
                    rec[1] = state;
                    to_ckpt = ckpt_every;
                }
            }
            __syncwarp();  // the tile is refilled by the next round's asynchronous fill
        }
    }

    // ---- finalize: state words (lib.rs:719-730, low word first), then the rest of my ring --------------
    drain_ring();  // <= 3 words left in the ring
    const bool raw = (p.flags & 1u) != 0;
    const uint32_t n_state = (valid && !raw) ? ans_state_words(state) : 0u;
    if (n_state >= 1) {
        sts_u32(ring_slot(pushed), (uint32_t)state);
        pushed += 4u;
    }
    if (n_state == 2) {
        sts_u32(ring_slot(pushed), (uint32_t)(state >> 32));
        pushed += 4u;
    }
    drain_ring();
    bool overflow = (pushed & 0x80000000u) != 0u;
    while (in_ring() != 0u) {  // < 4 words, one at a time
        if (drained + 4u <= cap)
            *reinterpret_cast<uint32_t *>(gbase + drained) = lds_u32(ring_slot(drained));
        else
            overflow = true;
        drained += 4u;
    }
    if (valid) {
        if (p.states_out) p.states_out[k] = state;
        if (min_prob == 0u) report_error(p.status, kErrImpossibleSymbol, k);
        if (overflow) report_error(p.status, kErrOutOfSpace, k);
    }
    // ---- K6: place my stream in the dense container ---------------------------------------------------
    const uint32_t *gbegin = reinterpret_cast<const uint32_t *>(gbase);
#ifdef CTR_DBG_NO_TAIL
    if (threadIdx.x == 0x7fffffff)
#endif
    compact_tail<BLOCK>(p.compact, tile, k, K, valid, gbegin,
                        (valid && !overflow) ? drained >> 2 : 0u, p.status);
}

// =====================================================================================================
// decode
// =====================================================================================================
//   SMALL : (SHARED only) alphabet <= 256
// With a shared model and the interleaved layout the CTA is 1024 threads so that the 32 KB quantile index is
// staged once per SM; otherwise (transposition tiles, or global tables: nothing to amortise) 256 threads.
constexpr int kDecBlockShared = 1024;
//   BLOCK : threads per CTA; 0 = decided at launch (pool decoders: the CTA is sized so that the grid is one wave)
//   TABLE : kTableGlobal / kTableLut / kTablePool
template <int BLOCK, int TABLE, bool CONTIG, bool PERSYM, bool SMALL>
__global__ void __launch_bounds__(BLOCK ? BLOCK : 1024, BLOCK == 0 || BLOCK >= 1024 ? 1 : (BLOCK >= 256 ? 2 : 8))
    ans_decode_kernel(const __grid_constant__ AnsParams p) {
    extern __shared__ __align__(128) uint32_t smem[];
    __shared__ uint64_t bar;

    constexpr bool SHARED = TABLE == kTableLut, POOL = TABLE == kTablePool, GAUSS = TABLE == kTableGauss;
    // one CTA per SM: room for the finer quantile index (p.model.dec_big)
    constexpr bool BIG_LUT = SHARED && BLOCK == kDecBlockShared;
    constexpr int kIndexBits = BIG_LUT ? kBigLutBits : kLutBits;
    constexpr uint32_t kIndexBytes = 8u << kIndexBits;
    const uint32_t kBlock = BLOCK ? (uint32_t)BLOCK : blockDim.x;
    const int lane = threadIdx.x & 31;
    const int warp_in_cta = threadIdx.x >> 5;
    const uint32_t kWarpsPerCta = kBlock / 32;

    // shared memory carve-up: [lane rings (64 B each)][quantile index + cdf | model pool][symbol tiles][index tiles]
    const uint32_t alphabet = p.model.alphabet;
    const uint32_t dec_table_bytes = kIndexBytes + p.model.dec_cdf_bytes;
    const uint32_t table_words = SHARED ? dec_table_bytes / 4
                                        : (POOL ? (p.model.pool_cdf_bytes + p.model.pool_cidx_bytes) / 4 : 0);
    const uint32_t kRingsWords = kBlock * kDecRingWords;
    const uint32_t ring = smem_u32_pinned(smem) + threadIdx.x * kDecRingBytes;  // 64-byte aligned
    const uint32_t lut_addr = smem_u32_pinned(smem + kRingsWords);
    uint32_t cdf_addr = lut_addr + (POOL ? 0u : kIndexBytes);  // POOL: [n_models][alphabet + 1] starts the table area
    asm volatile("" : "+r"(cdf_addr));
    uint32_t *sym_tile = smem + kRingsWords + table_words + warp_in_cta * kTileWords;
    uint32_t *idx_tile = sym_tile + kWarpsPerCta * kTileWords;

    // (the index is needed by the first decode_one only: the copy runs during the prologue)
    if (SHARED) stage_table_begin(smem + kRingsWords, BIG_LUT ? p.model.dec_big : p.model.dec, dec_table_bytes, &bar);
    if (POOL)
        stage_tables(smem + kRingsWords, p.model.cdf, p.model.pool_cdf_bytes, smem + kRingsWords + p.model.pool_cdf_bytes / 4,
                     p.model.cidx, p.model.pool_cidx_bytes, &bar);

    const uint64_t K = p.K, N = p.N;
    const uint64_t k = (uint64_t)blockIdx.x * kBlock + threadIdx.x;
    const bool valid = k < K;
    const uint64_t kc = valid ? k : K - 1;
    const bool raw = (p.flags & 1u) != 0;

    // My stream is words[begin, end); I pop from the end.  The ring slot of a word is its global address
    // mod 64, so aligned 16-byte blocks of the stream map to aligned 16-byte groups of the ring
    // (the words buffer is 16-byte aligned: checked by the host).
    uint64_t n_k = 0, o_k = 0;
    uint64_t begin = 0, end = 0;
    if (valid) {
        if (CONTIG) {
            o_k = p.sym_off[k];
            n_k = p.sym_off[k + 1] - o_k;
            if (o_k > N || n_k > N - o_k) {  // offsets outside the symbol array (or decreasing): flag, decode nothing
                report_error(p.status, kErrBadArgument, k);
                o_k = 0;
                n_k = 0;
            }
        }
        begin = p.offsets[k];
        end = p.ends ? p.ends[k] : p.offsets[k + 1];
    }
    // low 32 bits of the global byte address just past the next word to pop
    uint32_t pop_off = (uint32_t)(uintptr_t)(p.words + end);
    // ... and of the lowest word of my stream that has landed in the ring: (pop_off - landed_off) / 4 words are unread
    // (the per-symbol path then only moves pop_off)
    uint32_t landed_off = pop_off;
    uint32_t pending = 0;                   // words of the 16-byte block that is in flight
    // (a stream of >= 2^32 words does not exist: the encoder's lengths are 32-bit)
    uint32_t unstaged = (uint32_t)(end - begin);  // words of my stream not yet requested
    const char *gblock = reinterpret_cast<const char *>(p.words) + ((end * 4u) & ~(uint64_t)15);  // block holding word end
    if ((end & 3u) == 0) gblock -= 16;      // ... or rather the block holding word end-1
    auto avail_bytes = [&]() -> uint32_t { return pop_off - landed_off; };

    // request the next block below (asynchronously); the first block of a stream may be partial at the top,
    // the last one at the bottom
    auto request_block = [&](uint32_t top_words) {
        const uint32_t n = unstaged < top_words ? unstaged : top_words;
        cp_async_16(ring | ((uint32_t)(uintptr_t)gblock & (kDecRingBytes - 1u)), gblock);
        gblock -= 16;
        unstaged -= n;
        pending = n;
    };
    // every kCheckEvery symbols: the block requested TWO checks ago has landed (a request has two check intervals,
    // ~3000 cycles, to arrive: with one interval 10 % of the kernel's stall samples were this wait); request the next
    // one if the ring has room for it next to the block still in flight.  With S = unread + in-flight words, a
    // request whenever S <= 12 keeps S >= 9 and therefore >= 5 unread words after every check while the stream has
    // words left: the coder never finds its ring empty before the stream really is.
    uint32_t pending_old = 0;  // words of the block requested at the previous check (in flight)
    auto top_up = [&]() {
        cp_async_wait_group<1>();
        landed_off -= pending_old * 4u;
        pending_old = pending;
        pending = 0;
        if (avail_bytes() + pending_old * 4u <= (uint32_t)(kDecRingWords - 4) * 4u && unstaged != 0u) request_block(4u);
        cp_async_commit();
    };
    auto land_all = [&]() {
        cp_async_wait_all();
        landed_off -= (pending_old + pending) * 4u;
        pending_old = 0;
        pending = 0;
    };
    // start-up: stage at least 8 words (or the whole stream) synchronously
    {
        const uint32_t first = (end & 3u) ? (uint32_t)(end & 3u) : 4u;  // words of my stream in the top block
        if (unstaged != 0u) request_block(first);
        cp_async_commit();
#pragma unroll 1
        for (int i = 0; i < 3; ++i) top_up();
        land_all();
    }

    auto pop_word = [&]() -> uint32_t {
        pop_off -= 4u;
        return lds_u32(ring | (pop_off & (kDecRingBytes - 1u)));
    };

    // ---- initial state: stack.rs:299-318, 440-462 (from_compressed) or the caller's raw state ------
    uint32_t lo = 0, hi = 0;  // the coder state
    bool trailing_zero = false;
    if (raw) {
        if (valid && p.states_in) {
            const uint64_t s = p.states_in[k];
            lo = (uint32_t)s;
            hi = (uint32_t)(s >> 32);
        }
    } else {
        if (avail_bytes() != 0u) {
            lo = pop_word();
            trailing_zero = lo == 0u;
            if (avail_bytes() != 0u && lo != 0u) {
                hi = lo;
                lo = pop_word();
            }
        }
    }
    top_up();
    if (SHARED) stage_table_wait(&bar);

    const uint32_t n_models = p.model.n_models;
    uint32_t min_symbol = (uint32_t)p.model.min_symbol;
    asm volatile("" : "+r"(min_symbol));
    const uint32_t stream_model = (p.index_mode == 2) ? p.model_index[kc] : 0u;
    const uint32_t pool_row_bytes = (alphabet + 1) * 4u;
    const uint32_t pool_cidx_stride = (alphabet > 256 ? 2u : 1u) * (kCoarseSize + 1);
    const uint32_t pool_cidx_addr = cdf_addr + p.model.pool_cdf_bytes;

    bool bad_model = false;  // GAUSS: a std that is not > 0 (the symbols decoded with it are garbage)
    // one reference decode_symbol (stack.rs:1070-1100)
    // `converged`: std::true_type where every lane of the warp makes this call together (the hot loops), so that the
    // lookup may vote; std::false_type at the ragged ends, where only some lanes still own a symbol
    auto decode_one = [&](uint32_t m, auto converged) -> int32_t {
        constexpr bool kVote = BIG_LUT && decltype(converged)::value;
        const uint32_t q = lo & kQuantileMask;
        uint32_t left, right, s;
        if (SHARED) {
            s = lookup_shared<SMALL, kIndexBits, kVote>(lut_addr, cdf_addr, alphabet, lo, q, left, right);
        } else if (GAUSS) {
            m = m < n_models ? m : n_models - 1;
            const double mean = __ldg(p.gauss_means + m), std = __ldg(p.gauss_stds + m);
            bad_model |= !(std > 0.0);
            s = gauss_quantile(q, mean, std, p.gauss_free_weight, p.model.min_symbol, alphabet, left, right);
        } else if (POOL) {
            m = m < n_models ? m : n_models - 1;
            s = lookup_pool(cdf_addr + m * pool_row_bytes, pool_cidx_addr + m * pool_cidx_stride, alphabet > 256, q, left, right);
        } else {
            m = m < n_models ? m : n_models - 1;  // decoding cannot fail (stack.rs:1062-1065)
            const uint32_t cstride = (alphabet > 256 ? 2u : 1u) * (kCoarseSize + 1);
            s = lookup_global(p.model.cdf + (uint64_t)m * (alphabet + 1),
                              p.model.cidx ? p.model.cidx + (uint64_t)m * cstride : nullptr, alphabet > 256, alphabet, q, left,
                              right);
        }
        // state = (state >> 24) * prob + (q - left), in 32-bit pieces (state >> 24 has 40 bits)
        const uint32_t prob = right - left;
        const uint64_t t = (uint64_t)__funnelshift_r(lo, hi, kPrecision) * prob + (uint64_t)(q - left);
        hi = (uint32_t)(t >> 32) + (hi >> kPrecision) * prob;
        lo = (uint32_t)t;
        if (hi == 0u && pop_off != landed_off) {  // stack.rs:1091-1097 (a word is left to pop)
            hi = lo;
            lo = pop_word();
        }
        return (int32_t)(min_symbol + s);
    };

    if (!CONTIG && !PERSYM && p.use_tma) {
        // ---- TMA path: decoded symbols are collected in shared-memory boxes of kBoxRows rows x 32 streams and
        // leave as one UTMASTG per box (columns beyond K are clipped by the tensor map)
        const Interleave g = interleave_of(N, K);
        const uint64_t rows_total = g.T - 1;  // full rows 0 .. T-2
        const uint32_t nbox = (uint32_t)(rows_total / kBoxRows);
        // (the tables end on a 16-byte boundary; boxes need 128)
        const uint32_t boxes = ((smem_u32_pinned(smem + kRingsWords + table_words) + 127u) & ~127u) + (uint32_t)warp_in_cta * (kDecBoxSlots * kBoxBytes);
        const uint32_t my_col = boxes + (uint32_t)lane * 4u;
        const int32_t x0 = (int32_t)(blockIdx.x * kBlock + (uint32_t)warp_in_cta * 32u);
        uint32_t slot = 0;
        for (uint32_t b = 0; b < nbox; ++b) {
            // the store that last read this slot (kDecBoxSlots boxes ago) must have read it
            if (lane == 0) tma_store_wait_read<kDecBoxSlots - 1>();
            __syncwarp();
            const uint32_t box = my_col + slot * kBoxBytes;
#pragma unroll
            for (int half = 0; half < kBoxRows / kCheckEvery; ++half) {
#pragma unroll
                for (int u = 0; u < kCheckEvery; ++u)
                    sts_u32(box + (uint32_t)(half * kCheckEvery + u) * 128u, (uint32_t)decode_one(stream_model, std::true_type{}));
                top_up();
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                tma_store_box(&p.tmap, x0, (int32_t)(b * kBoxRows), boxes + slot * kBoxBytes);
                tma_store_commit();
            }
            if (++slot == kDecBoxSlots) slot = 0;
        }
        {  // the rows after the last box, then the ragged last row
            int32_t *po = p.symbols_out + (uint64_t)nbox * kBoxRows * K + kc;
            for (uint64_t r = (uint64_t)nbox * kBoxRows; r < rows_total; ++r) {
                const int32_t sym = decode_one(stream_model, std::true_type{});
                if (valid) st_stream_s32(po, sym);
                po += K;
                top_up();
            }
            if (valid && k < g.last) st_stream_s32(p.symbols_out + (g.T - 1) * K + k, decode_one(stream_model, std::false_type{}));
        }
        if (lane == 0) tma_store_wait_all();
    } else if (!CONTIG) {
        const Interleave g = interleave_of(N, K);
        if (g.T > 1) {
            char *po = reinterpret_cast<char *>(p.symbols_out + kc);
            const char *pm = PERSYM ? reinterpret_cast<const char *>(p.model_index + kc) : nullptr;
            uint64_t row_bytes = K * 4u;  // distance between consecutive symbols of a stream
            asm volatile("" : "+l"(row_bytes));
            const uint64_t rows_total = g.T - 1;  // full rows 0 .. T-2
            // FULL: every lane of the warp owns a stream, so nothing in the loop is predicated on `valid`
            auto run_rows = [&](auto full_tag) {
                constexpr bool FULL = decltype(full_tag)::value;
                uint32_t batches = (uint32_t)(rows_total / kCheckEvery);
                uint32_t rows_left = (uint32_t)(rows_total - (uint64_t)batches * kCheckEvery);
                for (; batches > 0; --batches) {
                    uint32_t mbuf[kCheckEvery];
#pragma unroll
                    for (int u = 0; u < kCheckEvery; ++u) {
                        mbuf[u] = stream_model;
                        if (PERSYM) {
                            mbuf[u] = ld_stream_u32(reinterpret_cast<const uint32_t *>(pm));
                            pm += row_bytes;
                        }
                    }
#pragma unroll
                    for (int u = 0; u < kCheckEvery; ++u) {
                        const int32_t sym = decode_one(mbuf[u], std::true_type{});
                        if (FULL || valid) st_stream_s32(reinterpret_cast<int32_t *>(po), sym);
                        po += row_bytes;
                    }
                    top_up();
                }
                while (rows_left > 0) {  // at most kCheckEvery-1 more symbols
                    uint32_t m = stream_model;
                    if (PERSYM) {
                        m = ld_stream_u32(reinterpret_cast<const uint32_t *>(pm));
                        pm += row_bytes;
                    }
                    const int32_t sym = decode_one(m, std::true_type{});
                    if (FULL || valid) st_stream_s32(reinterpret_cast<int32_t *>(po), sym);
                    po += row_bytes;
                    rows_left -= 1;
                }
                top_up();
            };
            if (__all_sync(kFullMask, valid))
                run_rows(std::true_type{});
            else
                run_rows(std::false_type{});
        }
        if (g.T > 0) {  // ragged last row
            if (valid && k < g.last) {
                const uint64_t i = (g.T - 1) * K + k;
                const int32_t sym = decode_one(PERSYM ? ld_stream_u32(p.model_index + i) : stream_model, std::false_type{});
                st_stream_s32(p.symbols_out + i, sym);
            }
        }
    } else {
        uint64_t done = 0;  // symbols of my stream already produced
        const uint32_t my_row = (uint32_t)lane * (kRowStride * 4u);
        const uint32_t row = smem_u32(sym_tile) + my_row, idx_row = smem_u32(idx_tile) + my_row;
        const uint64_t rounds = (warp_max_u64(n_k, lane) + 31) / 32;
        for (uint64_t r = 0; r < rounds; ++r) {
            const uint64_t left_n = n_k - done;
            const uint32_t c = left_n < 32 ? (uint32_t)left_n : 32u;
            const unsigned have = __ballot_sync(kFullMask, c > 0);
            if (PERSYM) warp_fill_rows(have, idx_tile, p.model_index + o_k + done, c, lane);
            const uint32_t cmin = __reduce_min_sync(kFullMask, c), cmax = __reduce_max_sync(kFullMask, c);
            uint32_t s = 0;
            for (; s + kCheckEvery <= cmin; s += kCheckEvery) {  // every lane owns all four symbols
                top_up();
#pragma unroll
                for (int u = 0; u < kCheckEvery; ++u) {
                    const uint32_t at = (s + (uint32_t)u) * 4u;
                    sts_u32(row + at, (uint32_t)decode_one(PERSYM ? lds_u32(idx_row + at) : stream_model, std::true_type{}));
                }
            }
            for (; s < cmax; ++s) {  // ragged end of the round
                if ((s & (kCheckEvery - 1)) == 0) top_up();
                if (s < c) sts_u32(row + s * 4u, (uint32_t)decode_one(PERSYM ? lds_u32(idx_row + s * 4u) : stream_model, std::false_type{}));
            }
            warp_flush_rows(have, sym_tile, reinterpret_cast<uint32_t *>(p.symbols_out + o_k + done), c, lane);
            done += c;
        }
    }

    cp_async_wait_all();
    if (valid) {
        if (p.states_out) p.states_out[k] = ((uint64_t)hi << 32) | lo;
        if (p.words_left) p.words_left[k] = (uint64_t)unstaged + pending + pending_old + (avail_bytes() >> 2);
        if (trailing_zero) report_error(p.status, kErrTrailingZero, k);
        if (GAUSS && bad_model) report_error(p.status, kErrBadModel, k);
    }
}

}  // namespace ctr
