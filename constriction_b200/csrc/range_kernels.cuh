// range_kernels.cuh -- batched range encode / decode kernels (K3 / K4), one lane per independent coder.
//
// Same execution model as ans_kernels.cuh (lane = coder, shared-memory rows for coalesced word I/O,
// TMA-staged tables), with queue semantics: symbols are coded in forward order and words are read
// from the front.  The encoder's lazy carry ("Inverted" situation, queue.rs:647-702) can release a
// burst of held-back words in one step; bursts are drained by a warp-uniform loop so that the
// cooperative row flush stays convergent.
//
// Per-stream results equal the reference's RangeEncoder / RangeDecoder (src/stream/queue.rs) word
// for word, including the seal words (queue.rs:349-376,458-523).
//
// Coder state on the wire (CTR_FLAG_RAW, states_in / states_out): 4 x u64 per stream,
//   encoder {lower, range, num_inverted, first_inverted_word}, decoder {lower, range, point, 0}.
#pragma once
#include "ans_kernels.cuh"

namespace ctr {

template <bool SHARED, bool CONTIG>
__global__ void __launch_bounds__(kAnsBlock) range_encode_kernel(const AnsParams p) {
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ uint64_t bar;

    const int lane = threadIdx.x & 31;
    const int warp_in_cta = threadIdx.x >> 5;
    constexpr int kWarpsPerCta = kAnsBlock / 32;

    const uint32_t table_words = SHARED ? p.model.alphabet * 4 : 0;
    const uint4 *s_enc = reinterpret_cast<const uint4 *>(smem);
    uint32_t *rows = smem + table_words + warp_in_cta * kTileWords;
    int32_t *sym_tile = reinterpret_cast<int32_t *>(smem + table_words + kWarpsPerCta * kTileWords) + warp_in_cta * kTileWords;
    uint32_t *idx_tile = smem + table_words + 2 * kWarpsPerCta * kTileWords + warp_in_cta * kTileWords;

    if (SHARED) stage_table(smem, p.model.enc, p.model.alphabet * 16u, &bar);

    const uint64_t k = (uint64_t)blockIdx.x * kAnsBlock + threadIdx.x;
    const bool valid = k < p.K;
    const uint64_t K = p.K, N = p.N;

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

    RangeEncState st = range_enc_init();
    if (valid && p.states_in) {
        st.lower = p.states_in[4 * k];
        st.range = p.states_in[4 * k + 1];
        st.num_inverted = (uint32_t)p.states_in[4 * k + 2];
        st.first_inverted = (uint32_t)p.states_in[4 * k + 3];
    }
    uint32_t cnt = 0;
    uint64_t flushed = 0;
    bool alive = valid;
    const uint32_t stream_model = (p.index_mode == 2 && valid) ? p.model_index[k] : 0u;
    const uint32_t alphabet = p.model.alphabet;
    const int32_t min_symbol = p.model.min_symbol;

    auto fail = [&](uint32_t code) {
        report_error(p.status, code, k);
        alive = false;
    };

    auto flush_full = [&]() {
        const bool full = cnt == kRowWords;
        const unsigned mask = __ballot_sync(kFullMask, full);
        if (mask) {
            const bool ok = !(full && flushed + kRowWords > capacity);
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

    // push `pending` words word_at(0..pending) of every lane; warp-uniform
    auto drain = [&](uint32_t pending, auto word_at) {
        uint32_t j = 0;
        while (__any_sync(kFullMask, j < pending)) {
            if (j < pending) {
                rows[lane * kRowStride + cnt] = word_at(j);
                cnt += 1;
                j += 1;
            }
            flush_full();
        }
    };

    // one reference encode_symbol (queue.rs:612-705); returns the words that became final
    auto encode_step = [&](bool act, int32_t sym, uint32_t m) {
        RangeEmit em;
        em.n_burst = 0;
        em.emit = false;
        em.burst_first = em.burst_fill = em.word = 0;
        if (act) {
            const uint32_t idx = (uint32_t)sym - (uint32_t)min_symbol;
            if (idx >= alphabet || (!SHARED && m >= p.model.n_models)) {
                fail(kErrImpossibleSymbol);
            } else {
                const uint4 e = SHARED ? s_enc[idx] : __ldg(p.model.enc + (uint64_t)m * alphabet + idx);
                if (e.y == 0 || !range_encode_step(st, e.x, e.y, em)) {
                    fail(kErrImpossibleSymbol);
                    em.n_burst = 0;
                    em.emit = false;
                }
            }
        }
        const uint32_t pending = em.n_burst + (em.emit ? 1u : 0u);
        drain(pending, [&](uint32_t j) {
            return j < em.n_burst ? (j == 0 ? em.burst_first : em.burst_fill) : em.word;
        });
    };

    if (!CONTIG) {
        const uint64_t T = K ? (N + K - 1) / K : 0;
        constexpr int U = 4;
        int32_t buf[U];
        uint32_t mbuf[U];
        for (uint64_t t0 = 0; t0 < T; t0 += U) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                buf[u] = 0;
                mbuf[u] = stream_model;
                const uint64_t i = (t0 + u) * K + k;
                if (valid && t0 + u < T && i < N) {
                    buf[u] = ld_stream_s32(p.symbols_in + i);
                    if (p.index_mode == 1) mbuf[u] = ld_stream_u32(p.model_index + i);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (t0 + u < T) {
                    const uint64_t i = (t0 + u) * K + k;
                    encode_step(alive && i < N, buf[u], mbuf[u]);
                }
            }
        }
    } else {
        uint64_t done = 0;
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
            warp_fill_rows<int32_t>(have, sym_tile, p.symbols_in + o_k + done, c, lane);
            if (p.index_mode == 1) warp_fill_rows<uint32_t>(have, idx_tile, p.model_index + o_k + done, c, lane);
            for (uint32_t s = 0; s < 32; ++s) {
                const bool act = alive && s < c;
                const int32_t sym = act ? sym_tile[lane * kRowStride + s] : 0;
                const uint32_t m = (act && p.index_mode == 1) ? idx_tile[lane * kRowStride + s] : stream_model;
                encode_step(act, sym, m);
            }
            done += c;
        }
    }

    // ---- seal (queue.rs:349-355, 458-523) unless the caller keeps the raw state -----------------------
    const bool raw = (p.flags & 1u) != 0;
    const uint32_t n_seal = (alive && !raw) ? range_num_seal_words(st) : 0u;
    drain(n_seal, [&](uint32_t j) { return range_seal_word(st, j); });
    {
        const bool ok = flushed + cnt <= capacity;
        const unsigned mask = __ballot_sync(kFullMask, cnt > 0 && ok);
        warp_flush_rows(mask, rows, region + flushed, cnt, lane);
        if (cnt > 0 && !ok) {
            report_error(p.status, kErrOutOfSpace, k);
            cnt = 0;
        }
    }
    if (valid) {
        p.lengths[k] = (uint32_t)(flushed + cnt);
        if (p.states_out) {
            p.states_out[4 * k] = st.lower;
            p.states_out[4 * k + 1] = st.range;
            p.states_out[4 * k + 2] = st.num_inverted;
            p.states_out[4 * k + 3] = st.first_inverted;
        }
    }
}

template <bool SHARED, bool CONTIG>
__global__ void __launch_bounds__(kAnsBlock) range_decode_kernel(const AnsParams p) {
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
    const uint32_t *next = p.words;  // next word of my stream that is not yet staged
    uint64_t rem = 0, total_words = 0;
    if (valid) {
        if (CONTIG) {
            o_k = p.sym_off[k];
            n_k = p.sym_off[k + 1] - o_k;
        } else {
            n_k = interleaved_len(N, K, k);
        }
        const uint64_t b = p.offsets[k];
        next = p.words + b;
        rem = total_words = p.offsets[k + 1] - b;
    }
    uint32_t filled = 0, ridx = 0;  // my row holds rows[ridx .. filled)
    const uint32_t alphabet = p.model.alphabet;
    const int32_t min_symbol = p.model.min_symbol;
    const uint32_t stream_model = (p.index_mode == 2 && valid) ? p.model_index[k] : 0u;

    // stage the next (up to) 32 words; chunks end on 128-byte boundaries of the global address space
    auto refill = [&]() {
        const bool need = ridx == filled && rem > 0;
        const unsigned mask = __ballot_sync(kFullMask, need);
        if (mask) {
            const uint64_t line_end = (((uint64_t)next) & ~(uint64_t)127) + 128;
            uint64_t c64 = (line_end - (uint64_t)next) / 4;
            if (c64 > rem) c64 = rem;
            const uint32_t c = need ? (uint32_t)c64 : 0u;
            warp_fill_rows<uint32_t>(mask, rows, next, c, lane);
            if (need) {
                filled = c;
                ridx = 0;
                next += c;
                rem -= c;
            }
        }
    };
    auto have_word = [&]() { return ridx < filled; };
    auto take = [&]() -> uint32_t { return rows[lane * kRowStride + (ridx++)]; };

    // ---- initial state: queue.rs:755-773 + read_point :847-868, or the caller's raw state -----------
    RangeDecState st;
    st.lower = 0;
    st.range = ~0ull;
    st.point = 0;
    bool alive = valid;
    refill();
    if (raw) {
        if (valid) {
            st.lower = p.states_in[4 * k];
            st.range = p.states_in[4 * k + 1];
            st.point = p.states_in[4 * k + 2];
        }
    } else {
        uint32_t got = 0;
        if (valid && have_word()) {
            st.point = take();
            got = 1;
        }
        refill();
        if (valid && got == 1 && have_word()) {
            st.point = (st.point << 32) | take();
            got = 2;
        }
        refill();
        if (got == 1) st.point <<= 32;
    }

    // one reference decode_symbol (queue.rs:968-1035)
    auto decode_one = [&](uint32_t m, int32_t &sym) -> bool {
        uint32_t q;
        if (!range_peek_quantile(st, q)) return false;
        uint32_t left, right, s;
        if (SHARED) {
            s = lookup_shared(s_pairs, s_lut, q, left, right);
        } else {
            if (m >= p.model.n_models) m = p.model.n_models - 1;
            s = lookup_global(p.model.cdf + (uint64_t)m * (alphabet + 1), alphabet, q, left, right);
        }
        if (range_decode_update(st, left, right - left)) {
            if (have_word()) st.point |= take();
        }
        sym = (int32_t)((uint32_t)min_symbol + s);
        return true;
    };

    if (!CONTIG) {
        const uint64_t T = K ? (N + K - 1) / K : 0;
        for (uint64_t t = 0; t < T; ++t) {
            const uint64_t i = t * K + k;
            const bool act = alive && i < N;
            uint32_t m = stream_model;
            if (act && p.index_mode == 1) m = ld_stream_u32(p.model_index + i);
            if (act) {
                int32_t sym;
                if (decode_one(m, sym)) {
                    st_stream_s32(p.symbols_out + i, sym);
                } else {
                    report_error(p.status, kErrInvalidData, k);
                    alive = false;
                }
            }
            refill();
        }
    } else {
        uint64_t done = 0;
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
            uint32_t produced = 0;
            for (uint32_t s = 0; s < 32; ++s) {
                if (alive && s < c) {
                    const uint32_t m = p.index_mode == 1 ? idx_tile[lane * kRowStride + s] : stream_model;
                    int32_t sym;
                    if (decode_one(m, sym)) {
                        sym_tile[lane * kRowStride + s] = sym;
                        produced = s + 1;
                    } else {
                        report_error(p.status, kErrInvalidData, k);
                        alive = false;
                    }
                }
                refill();
            }
            warp_flush_rows(have, reinterpret_cast<const uint32_t *>(sym_tile),
                            reinterpret_cast<uint32_t *>(p.symbols_out + o_k + done), produced, lane);
            done += c;
        }
    }

    if (valid) {
        if (p.states_out) {
            p.states_out[4 * k] = st.lower;
            p.states_out[4 * k + 1] = st.range;
            p.states_out[4 * k + 2] = st.point;
            p.states_out[4 * k + 3] = 0;
        }
        if (p.words_left) p.words_left[k] = total_words - rem - (filled - ridx);  // words consumed (Pos::pos().0)
    }
}

}  // namespace ctr
