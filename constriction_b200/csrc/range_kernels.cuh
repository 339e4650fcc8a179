// range_kernels.cuh -- batched range encode / decode kernels (K3 / K4), one lane per independent coder.
//
// Same execution model as ans_kernels.cuh (lane = coder, shared-memory rows for coalesced word I/O,
// TMA-staged tables, uniform hot loops with the ragged last row peeled off), with queue semantics:
// symbols are coded in forward order and words are read from the front.  The encoder's lazy carry
// ("Inverted" situation, queue.rs:647-702) can release a burst of held-back words in one step; the
// first word of a step goes through the normal (predicated) push, anything beyond it is drained by a
// cold warp-uniform loop so that the cooperative row flush stays convergent.
//
// Per-stream results equal the reference's RangeEncoder / RangeDecoder (src/stream/queue.rs) word
// for word, including the seal words (queue.rs:349-376,458-523).
//
// Coder state on the wire (CTR_FLAG_RAW, states_in / states_out): 4 x u64 per stream,
//   encoder {lower, range, num_inverted, first_inverted_word}, decoder {lower, range, point, 0}.
#pragma once
#include "ans_kernels.cuh"

namespace ctr {

constexpr int kSymBatch = 4;  // symbols loaded per lane before they are coded

template <bool SHARED, bool CONTIG, bool PERSYM>
__global__ void __launch_bounds__(kAnsBlock) range_encode_kernel(const AnsParams p) {
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ uint64_t bar;

    const int lane = threadIdx.x & 31;
    const int warp_in_cta = threadIdx.x >> 5;
    constexpr int kWarpsPerCta = kAnsBlock / 32;

    const uint32_t table_words = SHARED ? (p.model.alphabet + 1) * 4 : 0;
    const uint4 *s_enc = reinterpret_cast<const uint4 *>(smem);
    uint32_t *rows = smem + table_words + warp_in_cta * kTileWords;
    uint32_t *sym_tile = smem + table_words + (kWarpsPerCta + warp_in_cta) * kTileWords;
    uint32_t *idx_tile = smem + table_words + (2 * kWarpsPerCta + warp_in_cta) * kTileWords;

    if (SHARED) stage_table(smem, p.model.enc, (p.model.alphabet + 1) * 16u, &bar);

    const uint64_t K = p.K, N = p.N;
    const uint32_t tile = take_tile_ticket(p.compact.ticket);  // which 256 streams this CTA codes
    const uint64_t k = (uint64_t)tile * kAnsBlock + threadIdx.x;
    const bool valid = k < K;
    const uint64_t kc = valid ? k : K - 1;

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
    uint32_t *gptr = p.scratch + scratch_start(o_k, k);
    uint32_t *const gbegin = gptr;
    uint32_t *const gend = valid ? p.scratch + scratch_start(o_k + n_k, k + 1) : gptr;

    RangeEncState st = range_enc_init();
    if (valid && p.states_in) {
        st.lower = p.states_in[4 * k];
        st.range = p.states_in[4 * k + 1];
        st.num_inverted = (uint32_t)p.states_in[4 * k + 2];
        st.first_inverted = (uint32_t)p.states_in[4 * k + 3];
    }
    uint32_t *const myrow = rows + lane * kRowStride;
    uint32_t *wptr = myrow;
    bool bad = false, overflow = false;
    const uint32_t stream_model = (p.index_mode == 2) ? p.model_index[kc] : 0u;
    const uint32_t alphabet = p.model.alphabet;
    const uint32_t n_models = p.model.n_models;
    const int32_t min_symbol = p.model.min_symbol;

    auto flush_full = [&]() {
        const bool full = wptr == myrow + kRowWords;
        const unsigned mask = __ballot_sync(kFullMask, full);
        if (mask) {
            const bool ok = !(full && gptr + kRowWords > gend);
            const unsigned okmask = __ballot_sync(kFullMask, full && ok);
            warp_flush_rows(okmask, rows, gptr, kRowWords, lane);
            if (full) {
                if (ok)
                    gptr += kRowWords;
                else
                    overflow = true;
                wptr = myrow;
            }
        }
    };

    // push word_at(first .. pending) of every lane, one word per lane per round (cold unless pending > 1)
    auto drain_from = [&](uint32_t first, uint32_t pending, auto word_at) {
        uint32_t j = first;
        while (__any_sync(kFullMask, j < pending)) {
            if (j < pending) {
                *wptr++ = word_at(j);
                j += 1;
            }
            flush_full();
        }
    };

    // one reference encode_symbol (queue.rs:612-705).  `act` is false for lanes that have no symbol in
    // this step; impossible symbols are skipped and flagged.
    auto encode_step = [&](bool act, int32_t sym, uint32_t m) {
        uint32_t idx = (uint32_t)sym - (uint32_t)min_symbol;
        bool ok = idx < alphabet;
        uint4 e;
        if (SHARED) {
            idx = ok ? idx : 0u;
            e = s_enc[idx];
        } else {
            ok = ok && m < n_models;
            idx = ok ? idx : 0u;
            m = ok ? m : 0u;
            e = __ldg(p.model.enc + (uint64_t)m * (alphabet + 1) + idx);
        }
        ok = ok && e.y != 0u;
        RangeEmit em;
        em.n_burst = 0;
        em.emit = false;
        em.burst_first = em.burst_fill = em.word = 0;
        if (act) {
            if (ok)
                ok = range_encode_step(st, e.x, e.y, em);
            bad |= !ok;
        }
        const uint32_t pending = em.n_burst + (em.emit ? 1u : 0u);
        auto word_at = [&](uint32_t j) { return j < em.n_burst ? (j == 0 ? em.burst_first : em.burst_fill) : em.word; };
        if (pending) *wptr++ = word_at(0);
        flush_full();
        if (__any_sync(kFullMask, pending > 1)) drain_from(1, pending, word_at);
    };

    if (!CONTIG) {
        const Interleave g = interleave_of(N, K);
        if (g.T > 1) {
            uint64_t rows_left = g.T - 1;  // full rows 0 .. T-2
            const int32_t *ps = p.symbols_in + kc;
            const uint32_t *pm = PERSYM ? p.model_index + kc : nullptr;
            while (rows_left >= (uint64_t)kSymBatch) {
                int32_t buf[kSymBatch];
                uint32_t mbuf[kSymBatch];
#pragma unroll
                for (int u = 0; u < kSymBatch; ++u) {
                    buf[u] = ld_stream_s32(ps);
                    ps += K;
                    if (PERSYM) {
                        mbuf[u] = ld_stream_u32(pm);
                        pm += K;
                    } else {
                        mbuf[u] = stream_model;
                    }
                }
#pragma unroll
                for (int u = 0; u < kSymBatch; ++u) encode_step(valid, buf[u], mbuf[u]);
                rows_left -= kSymBatch;
            }
            while (rows_left > 0) {
                const int32_t sym = ld_stream_s32(ps);
                ps += K;
                uint32_t m = stream_model;
                if (PERSYM) {
                    m = ld_stream_u32(pm);
                    pm += K;
                }
                encode_step(valid, sym, m);
                rows_left -= 1;
            }
        }
        if (g.T > 0) {  // ragged last row
            const bool has = valid && k < g.last;
            const uint64_t i = (g.T - 1) * K + (has ? k : 0);
            const int32_t sym = has ? ld_stream_s32(p.symbols_in + i) : 0;
            const uint32_t m = (has && PERSYM) ? ld_stream_u32(p.model_index + i) : stream_model;
            encode_step(has, sym, m);
        }
    } else {
        uint64_t done = 0;
        const uint64_t rounds = (warp_max_u64(n_k, lane) + 31) / 32;
        for (uint64_t r = 0; r < rounds; ++r) {
            const uint64_t left_n = n_k - done;
            const uint32_t c = left_n < 32 ? (uint32_t)left_n : 32u;
            const unsigned have = __ballot_sync(kFullMask, c > 0);
            warp_fill_rows(have, sym_tile, reinterpret_cast<const uint32_t *>(p.symbols_in + o_k + done), c, lane);
            if (PERSYM) warp_fill_rows(have, idx_tile, p.model_index + o_k + done, c, lane);
            const uint32_t cmax = __reduce_max_sync(kFullMask, c);
            for (uint32_t s = 0; s < cmax; ++s) {
                const bool act = s < c;
                const int32_t sym = act ? (int32_t)sym_tile[lane * kRowStride + s] : 0;
                const uint32_t m = (act && PERSYM) ? idx_tile[lane * kRowStride + s] : stream_model;
                encode_step(act, sym, m);
            }
            done += c;
        }
    }

    // ---- seal (queue.rs:349-355, 458-523) unless the caller keeps the raw state -----------------------
    const bool raw = (p.flags & 1u) != 0;
    const uint32_t n_seal = (valid && !bad && !raw) ? range_num_seal_words(st) : 0u;
    drain_from(0, n_seal, [&](uint32_t j) { return range_seal_word(st, j); });
    uint32_t cnt = (uint32_t)(wptr - myrow);
    {
        const bool ok = gptr + cnt <= gend;
        const unsigned mask = __ballot_sync(kFullMask, cnt > 0 && ok);
        if (mask) warp_flush_rows(mask, rows, gptr, cnt, lane);
        if (cnt > 0 && !ok) {
            overflow = true;
            cnt = 0;
        }
    }
    if (valid) {
        if (p.states_out) {
            p.states_out[4 * k] = st.lower;
            p.states_out[4 * k + 1] = st.range;
            p.states_out[4 * k + 2] = st.num_inverted;
            p.states_out[4 * k + 3] = st.first_inverted;
        }
        if (bad) report_error(p.status, kErrImpossibleSymbol, k);
        if (overflow) report_error(p.status, kErrOutOfSpace, k);
    }
    // ---- K6: place my stream in the dense container ---------------------------------------------------
    compact_tail<kAnsBlock>(p.compact, tile, k, K, valid, gbegin, valid ? (uint32_t)(gptr - gbegin) + cnt : 0u, p.status);
}

template <bool SHARED, bool CONTIG, bool PERSYM>
__global__ void __launch_bounds__(kAnsBlock) range_decode_kernel(const AnsParams p) {
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ uint64_t bar;

    const int lane = threadIdx.x & 31;
    const int warp_in_cta = threadIdx.x >> 5;
    constexpr int kWarpsPerCta = kAnsBlock / 32;

    const uint32_t table_words = SHARED ? (kLutBytes + p.model.dec_cdf_bytes) / 4 : 0;
    const uint32_t lut_addr = smem_u32_pinned(smem);
    const uint32_t cdf_addr = lut_addr + kLutBytes;
    uint32_t *rows = smem + table_words + warp_in_cta * kTileWords;
    uint32_t *sym_tile = smem + table_words + (kWarpsPerCta + warp_in_cta) * kTileWords;
    uint32_t *idx_tile = smem + table_words + (2 * kWarpsPerCta + warp_in_cta) * kTileWords;

    if (SHARED) stage_table(smem, p.model.dec, kLutBytes + p.model.dec_cdf_bytes, &bar);

    const uint64_t K = p.K, N = p.N;
    const uint64_t k = (uint64_t)blockIdx.x * kAnsBlock + threadIdx.x;
    const bool valid = k < K;
    const uint64_t kc = valid ? k : K - 1;
    const bool raw = (p.flags & 1u) != 0;

    uint64_t n_k = 0, o_k = 0;
    const uint32_t *gnext = p.words;  // next word of my stream that is not yet staged
    const uint32_t *gend = p.words;   // end of my stream
    const uint32_t *gfirst = p.words;
    if (valid) {
        if (CONTIG) {
            o_k = p.sym_off[k];
            n_k = p.sym_off[k + 1] - o_k;
        }
        gfirst = gnext = p.words + p.offsets[k];
        gend = p.words + p.offsets[k + 1];
    }
    uint32_t *const myrow = rows + lane * kRowStride;
    uint32_t *rptr = myrow, *rend = myrow;  // my row holds the unread words [rptr, rend)
    const uint32_t alphabet = p.model.alphabet;
    const uint32_t n_models = p.model.n_models;
    const int32_t min_symbol = p.model.min_symbol;
    const uint32_t stream_model = (p.index_mode == 2) ? p.model_index[kc] : 0u;

    // stage the next (up to) 32 words; chunks end on 128-byte boundaries of the global address space (cold)
    auto refill = [&]() {
        const bool need = rptr == rend && gnext != gend;
        const unsigned mask = __ballot_sync(kFullMask, need);
        if (mask) {
            const uint32_t *line_end = (const uint32_t *)((((uint64_t)gnext) & ~(uint64_t)127) + 128);
            const uint32_t *hi = line_end < gend ? line_end : gend;
            const uint32_t c = need ? (uint32_t)(hi - gnext) : 0u;
            warp_fill_rows(mask, rows, gnext, c, lane);
            if (need) {
                rptr = myrow;
                rend = myrow + c;
                gnext += c;
            }
        }
    };

    // ---- initial state: queue.rs:755-773 + read_point :847-868, or the caller's raw state -----------
    RangeDecState st;
    st.lower = 0;
    st.range = ~0ull;
    st.point = 0;
    bool invalid_data = false;
    refill();
    if (raw) {
        if (valid) {
            st.lower = p.states_in[4 * k];
            st.range = p.states_in[4 * k + 1];
            st.point = p.states_in[4 * k + 2];
        }
    } else {
        uint32_t got = 0;
        if (rptr != rend) {
            st.point = *rptr++;
            got = 1;
        }
        refill();
        if (got == 1 && rptr != rend) {
            st.point = (st.point << 32) | *rptr++;
            got = 2;
        }
        refill();
        if (got == 1) st.point <<= 32;
    }

    // one reference decode_symbol (queue.rs:968-1035); after invalid data the lane keeps running on a
    // clamped quantile (its symbols are garbage and the stream is flagged)
    auto decode_one = [&](uint32_t m) -> int32_t {
        uint32_t q = kQuantileMask;
        invalid_data |= !range_peek_quantile(st, q);
        uint32_t left, right, s;
        if (SHARED) {
            s = lookup_shared<false>(lut_addr, cdf_addr, alphabet, q, q, left, right);
        } else {
            m = m < n_models ? m : n_models - 1;
            const uint32_t cstride = (alphabet > 256 ? 2u : 1u) * (kCoarseSize + 1);
            s = lookup_global(p.model.cdf + (uint64_t)m * (alphabet + 1),
                              p.model.cidx ? p.model.cidx + (uint64_t)m * cstride : nullptr, alphabet > 256, alphabet, q, left,
                              right);
        }
        if (range_decode_update(st, left, right - left)) {
            if (rptr != rend) st.point |= *rptr++;
        }
        return (int32_t)((uint32_t)min_symbol + s);
    };

    if (!CONTIG) {
        const Interleave g = interleave_of(N, K);
        if (g.T > 1) {
            int32_t *po = p.symbols_out + kc;
            const uint32_t *pm = PERSYM ? p.model_index + kc : nullptr;
            for (uint64_t t = 0; t + 1 < g.T; ++t) {
                uint32_t m = stream_model;
                if (PERSYM) {
                    m = ld_stream_u32(pm);
                    pm += K;
                }
                const int32_t sym = decode_one(m);
                if (valid) st_stream_s32(po, sym);
                po += K;
                refill();
            }
        }
        if (g.T > 0) {
            if (valid && k < g.last) {
                const uint64_t i = (g.T - 1) * K + k;
                const int32_t sym = decode_one(PERSYM ? ld_stream_u32(p.model_index + i) : stream_model);
                st_stream_s32(p.symbols_out + i, sym);
            }
            refill();
        }
    } else {
        uint64_t done = 0;
        const uint64_t rounds = (warp_max_u64(n_k, lane) + 31) / 32;
        for (uint64_t r = 0; r < rounds; ++r) {
            const uint64_t left_n = n_k - done;
            const uint32_t c = left_n < 32 ? (uint32_t)left_n : 32u;
            const unsigned have = __ballot_sync(kFullMask, c > 0);
            if (PERSYM) warp_fill_rows(have, idx_tile, p.model_index + o_k + done, c, lane);
            const uint32_t cmax = __reduce_max_sync(kFullMask, c);
            for (uint32_t s = 0; s < cmax; ++s) {
                if (s < c) {
                    const uint32_t m = PERSYM ? idx_tile[lane * kRowStride + s] : stream_model;
                    sym_tile[lane * kRowStride + s] = (uint32_t)decode_one(m);
                }
                refill();
            }
            warp_flush_rows(have, sym_tile, reinterpret_cast<uint32_t *>(p.symbols_out + o_k + done), c, lane);
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
        // words consumed so far (Pos::pos().0, queue.rs:182-196)
        if (p.words_left) p.words_left[k] = (uint64_t)(gnext - gfirst) - (uint64_t)(rend - rptr);
        if (invalid_data) report_error(p.status, kErrInvalidData, k);
    }
}

}  // namespace ctr
