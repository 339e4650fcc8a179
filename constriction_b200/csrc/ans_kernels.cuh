// ans_kernels.cuh -- batched rANS encode / decode kernels (K1 / K2), one lane per independent coder.
//
// Execution model.  A batch holds K independent coders ("streams").  Lane l of warp w owns stream
// k = 32 w + l for the whole kernel and keeps that coder's 64-bit state in registers, so the
// loop-carried dependency of the reference's per-symbol loop (stream/mod.rs:592-607,1274-1297) is
// private to a lane and 32 such chains advance per warp instruction.  Everything that touches HBM is
// warp-cooperative and coalesced:
//   - symbols: interleaved layout -> one 128-byte row per warp step; contiguous layout -> 32x32 tiles
//     transposed through shared memory;
//   - compressed words: each lane appends to / pops from a private 32-word shared-memory row; a full
//     (empty) row is written (refilled) by the whole warp as one 128-byte transaction;
//   - the model: for a single shared model the encoder table (left, prob, 64-bit reciprocal) or the
//     decoder table (CDF pairs + 4096-bucket quantile index) is staged into shared memory by the TMA
//     engine (cp.async.bulk); model sets too big for that are read through L1/L2 (`ld.global.nc`).
//
// Per-stream results equal the reference's `AnsCoder` (src/stream/stack.rs:1014-1100) word for word.
#pragma once
#include "device_utils.cuh"

namespace ctr {

constexpr int kAnsBlock = 128;  // threads per CTA (4 warps)
constexpr int kLutBits = 12;
constexpr int kLutSize = 1 << kLutBits;            // quantile buckets of the decoder index
constexpr int kLutShift = kPrecision - kLutBits;   // q >> 12
constexpr uint32_t kMaxSharedAlphabet = 4096;      // bigger alphabets use the global-table path

struct ModelView {
    const uint32_t *cdf;   // [n_models][alphabet + 1]
    const uint4 *enc;      // [n_models][alphabet]   {left, prob, rcp_lo, rcp_hi}
    const uint32_t *dec;   // model 0 only: pairs uint2[alphabet_padded] ++ lut u32[kLutSize]
    uint32_t n_models;
    uint32_t alphabet;
    int32_t min_symbol;
    uint32_t dec_pairs_bytes;  // alphabet * 8 rounded up to 16
};

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
    uint32_t *scratch;   // strided per-stream regions
    uint32_t *lengths;   // u32[K] words written per stream
    // decode
    const uint32_t *words;
    const uint64_t *offsets;
    uint64_t *words_left;
};

// ---- warp-cooperative row I/O ---------------------------------------------------------------------

// Write row i (its first count_i words) of the warp's row buffer to dst_i, for every lane i in `mask`.
__device__ __forceinline__ void warp_flush_rows(unsigned mask, const uint32_t *rows, uint32_t *dst, uint32_t count,
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

// Fill row i with the `count_i` words at src_i, for every lane i in `mask`.
template <typename T>
__device__ __forceinline__ void warp_fill_rows(unsigned mask, T *rows, const T *src, uint32_t count, int lane) {
    __syncwarp();
    while (mask) {
        const int i = __ffs(mask) - 1;
        mask &= mask - 1;
        const T *s = (const T *)shfl_u64((uint64_t)src, i);
        const uint32_t c = __shfl_sync(kFullMask, count, i);
        if ((uint32_t)lane < c) rows[i * kRowStride + lane] = (T)ld_stream_u32((const uint32_t *)(s + lane));
    }
    __syncwarp();
}

// ---- model lookups ----------------------------------------------------------------------------------

// decoder: quantile -> (symbol index, left, right) in a shared-memory table with bucket index
__device__ __forceinline__ uint32_t lookup_shared(const uint2 *pairs, const uint32_t *lut, uint32_t q, uint32_t &left,
                                                  uint32_t &right) {
    const uint32_t lh = lut[q >> kLutShift];
    uint32_t lo = lh & 0xffffu, hi = lh >> 16;
    while (lo < hi) {  // almost always zero iterations: the bucket lies inside one symbol's interval
        const uint32_t mid = (lo + hi + 1) >> 1;
        if (pairs[mid].x <= q)
            lo = mid;
        else
            hi = mid - 1;
    }
    const uint2 pr = pairs[lo];
    left = pr.x;
    right = pr.y;
    return lo;
}

// decoder: binary search of a CDF row in global memory (through L1/L2):
// the last index s with cdf[s] <= q  (categorical/contiguous.rs:628-665 partition point - 1)
__device__ __forceinline__ uint32_t lookup_global(const uint32_t *row, uint32_t alphabet, uint32_t q, uint32_t &left,
                                                  uint32_t &right) {
    uint32_t lo = 0, hi = alphabet - 1;
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

// =====================================================================================================
// encode
// =====================================================================================================
//   SHARED : model 0's encoder table lives in shared memory (index_mode == NONE, small alphabet)
//   CONTIG : stream k owns symbols[sym_off[k] .. sym_off[k+1]) (else interleaved deal)
template <bool SHARED, bool CONTIG>
__global__ void __launch_bounds__(kAnsBlock) ans_encode_kernel(const AnsParams p) {
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ uint64_t bar;

    const int lane = threadIdx.x & 31;
    const int warp_in_cta = threadIdx.x >> 5;
    constexpr int kWarpsPerCta = kAnsBlock / 32;

    // shared memory carve-up: [table][word rows][symbol tiles][index tiles]
    const uint32_t table_words = SHARED ? p.model.alphabet * 4 : 0;
    const uint4 *s_enc = reinterpret_cast<const uint4 *>(smem);
    uint32_t *rows = smem + table_words + warp_in_cta * kTileWords;
    int32_t *sym_tile = reinterpret_cast<int32_t *>(smem + table_words + kWarpsPerCta * kTileWords) + warp_in_cta * kTileWords;
    uint32_t *idx_tile = smem + table_words + 2 * kWarpsPerCta * kTileWords + warp_in_cta * kTileWords;

    if (SHARED) stage_table(smem, p.model.enc, p.model.alphabet * 16u, &bar);

    const uint64_t k = (uint64_t)blockIdx.x * kAnsBlock + threadIdx.x;
    const bool valid = k < p.K;
    const uint64_t K = p.K, N = p.N;

    // stream geometry
    uint64_t n_k = 0, o_k = 0;
    if (valid) {
        if (CONTIG) {
            o_k = p.sym_off[k];
            n_k = p.sym_off[k + 1] - o_k;
        } else {
            n_k = interleaved_len(N, K, k);
            o_k = interleaved_start(N, K, k);
        }
    }
    uint32_t *const region = p.scratch + scratch_start(o_k, k);
    const uint64_t capacity = valid ? scratch_start(o_k + n_k, k + 1) - scratch_start(o_k, k) : 0;

    uint64_t state = (valid && p.states_in) ? p.states_in[k] : 0;
    uint32_t cnt = 0;        // words in my row
    uint64_t flushed = 0;    // words already written to my scratch region
    bool alive = valid;
    const uint32_t stream_model = (p.index_mode == 2 && valid) ? p.model_index[k] : 0u;
    const uint32_t alphabet = p.model.alphabet;
    const int32_t min_symbol = p.model.min_symbol;

    auto fail = [&](uint32_t code) {
        report_error(p.status, code, k);
        alive = false;
    };

    // one reference encode_symbol (stack.rs:1014-1048)
    auto encode_one = [&](int32_t sym, uint32_t m) {
        const uint32_t idx = (uint32_t)sym - (uint32_t)min_symbol;
        if (idx >= alphabet || (!SHARED && m >= p.model.n_models)) {
            fail(kErrImpossibleSymbol);
            return;
        }
        const uint4 e = SHARED ? s_enc[idx] : __ldg(p.model.enc + (uint64_t)m * alphabet + idx);
        if (e.y == 0) {
            fail(kErrImpossibleSymbol);
            return;
        }
        if (ans_encode_needs_flush(state, e.y)) {
            rows[lane * kRowStride + cnt] = (uint32_t)state;
            cnt += 1;
            state >>= 32;
        }
        state = ans_encode_update(state, e.x, e.y, ((uint64_t)e.w << 32) | e.z);
    };

    // after each step: write out the rows that filled up
    auto flush_full = [&]() {
        const bool full = cnt == kRowWords;
        const unsigned mask = __ballot_sync(kFullMask, full);
        if (mask) {
            bool ok = true;
            if (full && flushed + kRowWords > capacity) ok = false;
            const unsigned okmask = __ballot_sync(kFullMask, full && ok);
            warp_flush_rows(okmask, rows, region + flushed, cnt, lane);
            if (full) {
                if (ok)
                    flushed += kRowWords;
                else
                    fail(kErrOutOfSpace);
                cnt = 0;
            }
        }
    };

    if (!CONTIG) {
        // ---- interleaved deal: step t touches symbols[t*K + k]; one coalesced row per warp --------
        const uint64_t T = K ? (N + K - 1) / K : 0;
        constexpr int U = 4;  // symbols prefetched per lane
        int32_t buf[U];
        uint32_t mbuf[U];
        auto load_batch = [&](uint64_t t_hi) {  // loads steps t_hi-1 .. t_hi-U (those >= 0)
#pragma unroll
            for (int u = 0; u < U; ++u) {
                buf[u] = 0;
                mbuf[u] = stream_model;
                if (t_hi >= (uint64_t)(u + 1)) {
                    const uint64_t i = (t_hi - 1 - u) * K + k;
                    if (valid && i < N) {
                        buf[u] = ld_stream_s32(p.symbols_in + i);
                        if (p.index_mode == 1) mbuf[u] = ld_stream_u32(p.model_index + i);
                    }
                }
            }
        };
        uint64_t t_hi = T;
        while (t_hi > 0) {
            load_batch(t_hi);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (t_hi >= (uint64_t)(u + 1)) {
                    const uint64_t i = (t_hi - 1 - u) * K + k;
                    if (alive && i < N) encode_one(buf[u], mbuf[u]);
                    flush_full();
                }
            }
            t_hi = t_hi > U ? t_hi - U : 0;
        }
    } else {
        // ---- contiguous: 32x32 tiles, transposed through shared memory ---------------------------
        uint64_t remaining = n_k;  // symbols of my stream not yet loaded (I consume from the end)
        uint64_t max_n = n_k;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const uint64_t o = shfl_u64(max_n, lane ^ d);
            max_n = o > max_n ? o : max_n;
        }
        const uint64_t rounds = (max_n + 31) / 32;
        for (uint64_t r = 0; r < rounds; ++r) {
            const uint32_t c = remaining < 32 ? (uint32_t)remaining : 32u;
            remaining -= c;
            const int32_t *src = p.symbols_in + o_k + remaining;
            const unsigned have = __ballot_sync(kFullMask, c > 0);
            warp_fill_rows<int32_t>(have, sym_tile, src, c, lane);
            if (p.index_mode == 1) warp_fill_rows<uint32_t>(have, idx_tile, p.model_index + o_k + remaining, c, lane);
            for (uint32_t s = 0; s < 32; ++s) {
                if (alive && s < c) {
                    const int32_t sym = sym_tile[lane * kRowStride + (c - 1 - s)];
                    const uint32_t m = p.index_mode == 1 ? idx_tile[lane * kRowStride + (c - 1 - s)] : stream_model;
                    encode_one(sym, m);
                }
                flush_full();
            }
        }
    }

    // ---- finalize: state words (lib.rs:719-730, low word first), remaining partial rows --------------
    const bool raw = (p.flags & 1u) != 0;
    uint32_t n_state = (valid && !raw) ? ans_state_words(state) : 0u;
    {
        // make room for the state words
        const bool tight = cnt + n_state > kRowWords;
        bool ok = !(tight && flushed + cnt > capacity);
        const unsigned mask = __ballot_sync(kFullMask, tight && ok);
        warp_flush_rows(mask, rows, region + flushed, cnt, lane);
        if (tight) {
            if (ok) {
                flushed += cnt;
                cnt = 0;
            } else {
                fail(kErrOutOfSpace);
                cnt = 0;
                n_state = 0;
            }
        }
    }
    if (n_state >= 1) rows[lane * kRowStride + cnt++] = (uint32_t)state;
    if (n_state == 2) rows[lane * kRowStride + cnt++] = (uint32_t)(state >> 32);
    {
        bool ok = flushed + cnt <= capacity;
        const unsigned mask = __ballot_sync(kFullMask, cnt > 0 && ok);
        warp_flush_rows(mask, rows, region + flushed, cnt, lane);
        if (cnt > 0 && !ok) {
            report_error(p.status, kErrOutOfSpace, k);
            cnt = 0;
        }
    }
    if (valid) {
        p.lengths[k] = (uint32_t)(flushed + cnt);
        if (p.states_out) p.states_out[k] = state;
    }
}

// =====================================================================================================
// decode
// =====================================================================================================
template <bool SHARED, bool CONTIG>
__global__ void __launch_bounds__(kAnsBlock) ans_decode_kernel(const AnsParams p) {
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ uint64_t bar;

    const int lane = threadIdx.x & 31;
    const int warp_in_cta = threadIdx.x >> 5;
    constexpr int kWarpsPerCta = kAnsBlock / 32;

    const uint32_t table_words = SHARED ? (p.model.dec_pairs_bytes / 4 + kLutSize) : 0;
    const uint2 *s_pairs = reinterpret_cast<const uint2 *>(smem);
    const uint32_t *s_lut = smem + (SHARED ? p.model.dec_pairs_bytes / 4 : 0);
    uint32_t *rows = smem + table_words + warp_in_cta * kTileWords;
    int32_t *sym_tile = reinterpret_cast<int32_t *>(smem + table_words + kWarpsPerCta * kTileWords) + warp_in_cta * kTileWords;
    uint32_t *idx_tile = smem + table_words + 2 * kWarpsPerCta * kTileWords + warp_in_cta * kTileWords;

    if (SHARED) stage_table(smem, p.model.dec, p.model.dec_pairs_bytes + kLutSize * 4u, &bar);

    const uint64_t k = (uint64_t)blockIdx.x * kAnsBlock + threadIdx.x;
    const bool valid = k < p.K;
    const uint64_t K = p.K, N = p.N;
    const bool raw = (p.flags & 1u) != 0;

    uint64_t n_k = 0, o_k = 0;
    const uint32_t *base = p.words;
    uint64_t rem = 0;  // words of my stream not yet staged into my row
    if (valid) {
        if (CONTIG) {
            o_k = p.sym_off[k];
            n_k = p.sym_off[k + 1] - o_k;
        } else {
            n_k = interleaved_len(N, K, k);
        }
        const uint64_t b = p.offsets[k];
        base = p.words + b;
        rem = p.offsets[k + 1] - b;
    }
    uint32_t cnt = 0;
    const uint32_t alphabet = p.model.alphabet;
    const int32_t min_symbol = p.model.min_symbol;
    const uint32_t stream_model = (p.index_mode == 2 && valid) ? p.model_index[k] : 0u;

    // stage the next (up to) 32 words below my cursor; chunks end on 128-byte boundaries of the
    // global address space so that every refill after the first is one aligned line
    auto refill = [&]() {
        const bool need = cnt == 0 && rem > 0;
        const unsigned mask = __ballot_sync(kFullMask, need);
        if (mask) {
            const uint32_t *top = base + rem;  // one past the highest unstaged word
            uint64_t lo_addr = ((uint64_t)(top - 1)) & ~(uint64_t)127;
            const uint32_t *lo = (const uint32_t *)lo_addr;
            if (lo < base) lo = base;
            const uint32_t c = need ? (uint32_t)(top - lo) : 0u;
            warp_fill_rows<uint32_t>(mask, rows, lo, c, lane);
            if (need) {
                cnt = c;
                rem -= c;
            }
        }
    };
    auto pop = [&]() -> uint32_t { return rows[lane * kRowStride + (--cnt)]; };

    // ---- initial state: stack.rs:299-318, 440-462 (from_compressed) or the caller's raw state ------
    uint64_t state = 0;
    bool alive = valid;
    if (raw) {
        if (valid && p.states_in) state = p.states_in[k];
        refill();
    } else {
        refill();
        if (valid && cnt > 0) {
            const uint32_t w = pop();
            if (w == 0) {
                report_error(p.status, kErrTrailingZero, k);
                alive = false;
            }
            state = w;
        }
        refill();
        if (valid && alive && cnt > 0 && state != 0) state = (state << 32) | pop();
        refill();
    }

    // one reference decode_symbol (stack.rs:1070-1100)
    auto decode_one = [&](uint32_t m) -> int32_t {
        const uint32_t q = ans_peek_quantile(state);
        uint32_t left, right, s;
        if (SHARED) {
            s = lookup_shared(s_pairs, s_lut, q, left, right);
        } else {
            if (m >= p.model.n_models) m = p.model.n_models - 1;  // cannot report through Infallible
            s = lookup_global(p.model.cdf + (uint64_t)m * (alphabet + 1), alphabet, q, left, right);
        }
        state = ans_decode_update(state, q, left, right - left);
        if ((state >> 32) == 0 && cnt > 0) state = (state << 32) | pop();
        return (int32_t)((uint32_t)min_symbol + s);
    };

    if (!CONTIG) {
        const uint64_t T = K ? (N + K - 1) / K : 0;
        for (uint64_t t = 0; t < T; ++t) {
            const uint64_t i = t * K + k;
            const bool act = alive && i < N;
            uint32_t m = stream_model;
            if (act && p.index_mode == 1) m = ld_stream_u32(p.model_index + i);
            if (act) st_stream_s32(p.symbols_out + i, decode_one(m));
            refill();
        }
    } else {
        uint64_t done = 0;  // symbols of my stream already produced
        uint64_t max_n = n_k;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const uint64_t o = shfl_u64(max_n, lane ^ d);
            max_n = o > max_n ? o : max_n;
        }
        const uint64_t rounds = (max_n + 31) / 32;
        for (uint64_t r = 0; r < rounds; ++r) {
            const uint64_t left_n = n_k - done;
            const uint32_t c = left_n < 32 ? (uint32_t)left_n : 32u;
            const unsigned have = __ballot_sync(kFullMask, c > 0);
            if (p.index_mode == 1) warp_fill_rows<uint32_t>(have, idx_tile, p.model_index + o_k + done, c, lane);
            for (uint32_t s = 0; s < 32; ++s) {
                if (alive && s < c) {
                    const uint32_t m = p.index_mode == 1 ? idx_tile[lane * kRowStride + s] : stream_model;
                    sym_tile[lane * kRowStride + s] = decode_one(m);
                }
                refill();
            }
            warp_flush_rows(have, reinterpret_cast<const uint32_t *>(sym_tile),
                            reinterpret_cast<uint32_t *>(p.symbols_out + o_k + done), alive ? c : 0u, lane);
            done += c;
        }
    }

    if (valid) {
        if (p.states_out) p.states_out[k] = state;
        if (p.words_left) p.words_left[k] = rem + cnt;
    }
}

}  // namespace ctr
