// range_kernels.cuh -- batched range encode / decode kernels (K3 / K4), one lane per independent coder.
//
// Same execution model as ans_kernels.cuh: lane = coder with its state in registers, TMA-staged tables in
// shared memory for a shared model, lane-private word rings that are drained (encoder: LDS.128 + STG.128
// into the lane's scratch region) or topped up (decoder: asynchronous LDGSTS.128) 16 bytes at a time, uniform
// hot loops with the ragged last row of the interleaved deal peeled off, fused compaction in the encoder's
// tail.  Queue semantics: symbols are coded in forward order and words are read from the front.
//
// The encoder's lazy carry ("Inverted" situation, queue.rs:647-702) can release a burst of held-back words in
// one step; because the word path is lane-private, the (rare) burst is simply pushed by a lane-local loop that
// drains the ring as it goes.
//
// Per-stream results equal the reference's RangeEncoder / RangeDecoder (src/stream/queue.rs) word for word,
// including the seal words (queue.rs:349-376,458-523).
//
// Coder state on the wire (CTR_FLAG_RAW, states_in / states_out): 4 x u64 per stream,
//   encoder {lower, range, num_inverted, first_inverted_word}, decoder {lower, range, point, 0}.
#pragma once
#include "ans_kernels.cuh"

namespace ctr {

template <bool SHARED, bool CONTIG, bool PERSYM>
__global__ void __launch_bounds__(kAnsBlock, 4) range_encode_kernel(const AnsParams p) {
    extern __shared__ __align__(128) uint32_t smem[];
    __shared__ uint64_t bar;

    const int lane = threadIdx.x & 31;
    const int warp_in_cta = threadIdx.x >> 5;
    constexpr int kWarpsPerCta = kAnsBlock / 32;

    // shared memory carve-up as in ans_encode_kernel: [rings + parking slots][replicated table][tiles]
    const uint32_t alphabet = p.model.alphabet;
    const uint32_t table_words = SHARED ? (alphabet + 1) * 32 : 0;
    constexpr uint32_t kRingsWords = kAnsBlock * (kEncRingWords + 4);
    const uint32_t ring = smem_u32_pinned(smem) + threadIdx.x * kEncRingBytes;
    const uint32_t park = smem_u32(smem) + kAnsBlock * kEncRingBytes + threadIdx.x * 16u;
    const uint32_t table_addr = smem_u32_pinned(smem + kRingsWords) + (uint32_t)(lane & 7) * 16u;
    uint32_t *sym_tile = smem + kRingsWords + table_words + warp_in_cta * kTileWords;
    uint32_t *idx_tile = sym_tile + kWarpsPerCta * kTileWords;

    if (SHARED) stage_table(smem + kRingsWords, p.model.enc_rep, (alphabet + 1) * 128u, &bar);

    const uint64_t K = p.K, N = p.N;
    const uint32_t tile = take_tile_ticket(p.compact.ticket);
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
    char *gw;
    uint32_t room;
    {
        uint32_t *const gbegin = p.scratch + scratch_start(o_k, k);
        const uint64_t r = valid ? scratch_start(o_k + n_k, k + 1) - scratch_start(o_k, k) : 0;
        room = r > 0x3ffffff0u ? 0xffffffc0u : (uint32_t)r * 4u;
        gw = reinterpret_cast<char *>(gbegin);
        const uint64_t gb = (uint64_t)gbegin;
        asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(park), "r"((uint32_t)gb), "r"((uint32_t)(gb >> 32)) : "memory");
    }

    RangeEncState st = range_enc_init();
    if (valid && p.states_in) {
        st.lower = p.states_in[4 * k];
        st.range = p.states_in[4 * k + 1];
        st.num_inverted = (uint32_t)p.states_in[4 * k + 2];
        st.first_inverted = (uint32_t)p.states_in[4 * k + 3];
    }
    uint32_t pushed = 0, pending = 0;  // bytes pushed into my ring / not yet written to scratch
    uint32_t min_prob = 0xffffffffu;
    bool overflow = false;
    const uint32_t stream_model = (p.index_mode == 2) ? p.model_index[kc] : 0u;
    const uint32_t n_models = p.model.n_models;
    const uint32_t min_symbol = (uint32_t)p.model.min_symbol;

    auto push = [&](uint32_t w) {
        sts_u32(ring | (pushed & (kEncRingBytes - 1u)), w);
        pushed += 4u;
        pending += 4u;
    };
    auto drain_ring = [&]() {
        if (pending >= 16u) {
            const uint4 v = lds_v4(ring | ((pushed - pending) & (kEncRingBytes - 16u)));
            if (room >= 16u) {
                st_stream_v4(gw, v);
                gw += 16;
                room -= 16u;
            } else {
                room = 0u;
                overflow = true;
            }
            pending -= 16u;
        }
    };

    // one reference encode_symbol (queue.rs:612-705); impossible symbols are skipped and flagged
    auto encode_one = [&](int32_t sym, uint32_t m) {
        uint32_t idx = min((uint32_t)sym - min_symbol, alphabet);  // out of range -> sentinel entry (prob 0)
        uint4 e;
        if (SHARED) {
            e = lds_table_v4(table_addr + idx * 128u);
        } else {
            const bool ok = m < n_models;
            idx = ok ? idx : alphabet;
            m = ok ? m : 0u;
            e = __ldg(p.model.enc + (uint64_t)m * (alphabet + 1) + idx);
        }
        min_prob = min(min_prob, e.y);
        if (valid && e.y != 0u) {
            RangeEmit em;
            range_encode_step(st, e.x, e.y, em);
            if (em.n_burst != 0u) {  // rare: a resolved Inverted situation releases its held-back words
                for (uint32_t j = 0; j < em.n_burst; ++j) {
                    push(j == 0 ? em.burst_first : em.burst_fill);
                    drain_ring();
                }
            }
            if (em.emit) push(em.word);
        }
    };

    if (!CONTIG) {
        const Interleave g = interleave_of(N, K);
        if (g.T > 1) {
            const uint64_t rows_total = g.T - 1;  // full rows 0 .. T-2
            const char *ps = reinterpret_cast<const char *>(p.symbols_in + kc);
            const char *pm = PERSYM ? reinterpret_cast<const char *>(p.model_index + kc) : nullptr;
            uint64_t row_bytes = K * 4u;
            asm volatile("" : "+l"(row_bytes));
            int32_t buf[2][kCheckEvery];
            uint32_t mbuf[2][kCheckEvery];
            auto load_batch = [&](int which) {
#pragma unroll
                for (int u = 0; u < kCheckEvery; ++u) {
                    buf[which][u] = ld_stream_s32(reinterpret_cast<const int32_t *>(ps));
                    ps += row_bytes;
                    if (PERSYM) {
                        mbuf[which][u] = ld_stream_u32(reinterpret_cast<const uint32_t *>(pm));
                        pm += row_bytes;
                    } else {
                        mbuf[which][u] = stream_model;
                    }
                }
            };
            auto code_batch = [&](int which, bool load_next) {
                if (load_next) load_batch(which ^ 1);
                drain_ring();
#pragma unroll
                for (int u = 0; u < kCheckEvery; ++u) encode_one(buf[which][u], mbuf[which][u]);
            };
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
            while (rows_left > 0) {
                const int32_t sym = ld_stream_s32(reinterpret_cast<const int32_t *>(ps));
                ps += row_bytes;
                uint32_t m = stream_model;
                if (PERSYM) {
                    m = ld_stream_u32(reinterpret_cast<const uint32_t *>(pm));
                    pm += row_bytes;
                }
                encode_one(sym, m);
                rows_left -= 1;
            }
            drain_ring();
        }
        if (g.T > 0) {  // ragged last row
            if (valid && k < g.last) {
                const uint64_t i = (g.T - 1) * K + k;
                encode_one(ld_stream_s32(p.symbols_in + i), PERSYM ? ld_stream_u32(p.model_index + i) : stream_model);
            }
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
                if ((s & (kCheckEvery - 1)) == 0) drain_ring();
                if (s < c) {
                    const int32_t sym = (int32_t)sym_tile[lane * kRowStride + s];
                    const uint32_t m = PERSYM ? idx_tile[lane * kRowStride + s] : stream_model;
                    encode_one(sym, m);
                }
            }
            done += c;
        }
    }

    // ---- seal (queue.rs:349-355, 458-523) unless the caller keeps the raw state -----------------------
    drain_ring();
    const bool raw = (p.flags & 1u) != 0;
    const bool bad = min_prob == 0u;
    const uint32_t n_seal = (valid && !bad && !raw) ? range_num_seal_words(st) : 0u;
    for (uint32_t j = 0; j < n_seal; ++j) {
        push(range_seal_word(st, j));
        drain_ring();
    }
    while (pending != 0u) {  // < 4 words, one at a time
        if (room >= 4u) {
            *reinterpret_cast<uint32_t *>(gw) = lds_u32(ring | ((pushed - pending) & (kEncRingBytes - 1u)));
            gw += 4;
            room -= 4u;
        } else {
            overflow = true;
        }
        pending -= 4u;
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
    uint32_t gb_lo, gb_hi;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(gb_lo), "=r"(gb_hi) : "r"(park) : "memory");
    const uint32_t *gbegin = reinterpret_cast<const uint32_t *>(((uint64_t)gb_hi << 32) | gb_lo);
    compact_tail<kAnsBlock>(p.compact, tile, k, K, valid, gbegin,
                            (valid && !overflow) ? (uint32_t)((reinterpret_cast<const uint32_t *>(gw)) - gbegin) : 0u, p.status);
}

template <bool SHARED, bool CONTIG, bool PERSYM, bool SMALL>
__global__ void __launch_bounds__((SHARED && !CONTIG) ? kDecBlockShared : kAnsBlock, (SHARED && !CONTIG) ? 1 : 2)
    range_decode_kernel(const AnsParams p) {
    extern __shared__ __align__(128) uint32_t smem[];
    __shared__ uint64_t bar;

    constexpr int kBlock = (SHARED && !CONTIG) ? kDecBlockShared : kAnsBlock;
    const int lane = threadIdx.x & 31;
    const int warp_in_cta = threadIdx.x >> 5;
    constexpr int kWarpsPerCta = kBlock / 32;

    const uint32_t alphabet = p.model.alphabet;
    const uint32_t table_words = SHARED ? (kLutBytes + p.model.dec_cdf_bytes) / 4 : 0;
    constexpr uint32_t kRingsWords = kBlock * kDecRingWords;
    const uint32_t ring = smem_u32_pinned(smem) + threadIdx.x * kDecRingBytes;  // 64-byte aligned
    const uint32_t lut_addr = smem_u32_pinned(smem + kRingsWords);
    uint32_t cdf_addr = lut_addr + kLutBytes;
    asm volatile("" : "+r"(cdf_addr));
    uint32_t *sym_tile = smem + kRingsWords + table_words + warp_in_cta * kTileWords;
    uint32_t *idx_tile = sym_tile + kWarpsPerCta * kTileWords;

    if (SHARED) stage_table(smem + kRingsWords, p.model.dec, kLutBytes + p.model.dec_cdf_bytes, &bar);

    const uint64_t K = p.K, N = p.N;
    const uint64_t k = (uint64_t)blockIdx.x * kBlock + threadIdx.x;
    const bool valid = k < K;
    const uint64_t kc = valid ? k : K - 1;
    const bool raw = (p.flags & 1u) != 0;

    // My stream is words[begin, end); I read from the front.  Ring slot of a word = its global address mod 64.
    uint64_t n_k = 0, o_k = 0;
    uint64_t begin = 0, end = 0;
    if (valid) {
        if (CONTIG) {
            o_k = p.sym_off[k];
            n_k = p.sym_off[k + 1] - o_k;
        }
        begin = p.offsets[k];
        end = p.offsets[k + 1];
    }
    uint32_t pop_off = (uint32_t)(uintptr_t)(p.words + begin);  // low address bits of the next word to read
    uint32_t avail = 0, pending = 0;
    const uint32_t total_words = (uint32_t)(end - begin);
    uint32_t unstaged = total_words;
    const char *gblock = reinterpret_cast<const char *>(p.words) + ((begin * 4u) & ~(uint64_t)15);  // block holding word begin

    auto request_block = [&](uint32_t block_words) {
        const uint32_t n = unstaged < block_words ? unstaged : block_words;
        cp_async_16(ring | ((uint32_t)(uintptr_t)gblock & (kDecRingBytes - 1u)), gblock);
        gblock += 16;
        unstaged -= n;
        pending = n;
    };
    auto top_up = [&]() {
        cp_async_wait_all();
        avail += pending;
        pending = 0;
        if (avail <= (uint32_t)(kDecRingWords - 4) && unstaged != 0u) request_block(4u);
        cp_async_commit();
    };
    {
        const uint32_t first = 4u - (uint32_t)(begin & 3u);  // words of my stream in the bottom block
        if (unstaged != 0u) request_block(first);
        cp_async_commit();
#pragma unroll 1
        for (int i = 0; i < 3; ++i) top_up();
        cp_async_wait_all();
        avail += pending;
        pending = 0;
    }
    auto pop_word = [&]() -> uint32_t {
        const uint32_t w = lds_u32(ring | (pop_off & (kDecRingBytes - 1u)));
        pop_off += 4u;
        avail -= 1u;
        return w;
    };

    // ---- initial state: queue.rs:755-773 + read_point :847-868, or the caller's raw state -----------
    RangeDecState st;
    st.lower = 0;
    st.range = ~0ull;
    st.point = 0;
    bool invalid_data = false;
    if (raw) {
        if (valid) {
            st.lower = p.states_in[4 * k];
            st.range = p.states_in[4 * k + 1];
            st.point = p.states_in[4 * k + 2];
        }
    } else {
        if (avail != 0u) {
            st.point = (uint64_t)pop_word() << 32;
            if (avail != 0u) st.point |= pop_word();
        }
    }
    top_up();

    const uint32_t n_models = p.model.n_models;
    uint32_t min_symbol = (uint32_t)p.model.min_symbol;
    asm volatile("" : "+r"(min_symbol));
    const uint32_t stream_model = (p.index_mode == 2) ? p.model_index[kc] : 0u;

    // one reference decode_symbol (queue.rs:968-1035); after invalid data the lane keeps running on a
    // clamped quantile (its symbols are garbage and the stream is flagged)
    auto decode_one = [&](uint32_t m) -> int32_t {
        uint32_t q = kQuantileMask;
        invalid_data |= !range_peek_quantile(st, q);
        uint32_t left, right, s;
        if (SHARED) {
            s = lookup_shared<SMALL>(lut_addr, cdf_addr, alphabet, q, q, left, right);
        } else {
            m = m < n_models ? m : n_models - 1;
            const uint32_t cstride = (alphabet > 256 ? 2u : 1u) * (kCoarseSize + 1);
            s = lookup_global(p.model.cdf + (uint64_t)m * (alphabet + 1),
                              p.model.cidx ? p.model.cidx + (uint64_t)m * cstride : nullptr, alphabet > 256, alphabet, q, left,
                              right);
        }
        if (range_decode_update(st, left, right - left)) {
            if (avail != 0u) st.point |= pop_word();
        }
        return (int32_t)(min_symbol + s);
    };

    if (!CONTIG) {
        const Interleave g = interleave_of(N, K);
        if (g.T > 1) {
            char *po = reinterpret_cast<char *>(p.symbols_out + kc);
            const char *pm = PERSYM ? reinterpret_cast<const char *>(p.model_index + kc) : nullptr;
            uint64_t row_bytes = K * 4u;
            asm volatile("" : "+l"(row_bytes));
            const uint64_t rows_total = g.T - 1;
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
                        const int32_t sym = decode_one(mbuf[u]);
                        if (FULL || valid) st_stream_s32(reinterpret_cast<int32_t *>(po), sym);
                        po += row_bytes;
                    }
                    top_up();
                }
                while (rows_left > 0) {
                    uint32_t m = stream_model;
                    if (PERSYM) {
                        m = ld_stream_u32(reinterpret_cast<const uint32_t *>(pm));
                        pm += row_bytes;
                    }
                    const int32_t sym = decode_one(m);
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
        if (g.T > 0) {
            if (valid && k < g.last) {
                const uint64_t i = (g.T - 1) * K + k;
                const int32_t sym = decode_one(PERSYM ? ld_stream_u32(p.model_index + i) : stream_model);
                st_stream_s32(p.symbols_out + i, sym);
            }
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
                if ((s & (kCheckEvery - 1)) == 0) top_up();
                if (s < c) {
                    const uint32_t m = PERSYM ? idx_tile[lane * kRowStride + s] : stream_model;
                    sym_tile[lane * kRowStride + s] = (uint32_t)decode_one(m);
                }
            }
            warp_flush_rows(have, sym_tile, reinterpret_cast<uint32_t *>(p.symbols_out + o_k + done), c, lane);
            done += c;
        }
    }

    cp_async_wait_all();
    if (valid) {
        if (p.states_out) {
            p.states_out[4 * k] = st.lower;
            p.states_out[4 * k + 1] = st.range;
            p.states_out[4 * k + 2] = st.point;
            p.states_out[4 * k + 3] = 0;
        }
        // words consumed so far (Pos::pos().0, queue.rs:182-196)
        if (p.words_left) p.words_left[k] = (uint64_t)(total_words - unstaged - pending - avail);
        if (invalid_data) report_error(p.status, kErrInvalidData, k);
    }
}

}  // namespace ctr
