// range_kernels.cuh -- batched range encode / decode kernels (K3 / K4), one lane per independent coder.
//
// Same execution model as ans_kernels.cuh: lane = coder with its state in registers, TMA-staged tables in
// shared memory for a shared model, lane-private word rings that are drained (encoder: LDS.128 + STG.128
// into the lane's scratch region) or topped up (decoder: asynchronous LDGSTS.128) 16 bytes at a time, uniform
// hot loops with the ragged last row of the interleaved deal peeled off, fused compaction in the encoder's
// tail.  Queue semantics: symbols are coded in forward order and words are read from the front.
//
// The encoder uses the "eager words, late carry" formulation of coder_math.cuh: a word is appended whenever
// the interval is renormalised, and the reference's lazy carry ("Inverted" situation, queue.rs:647-702) becomes
// an increment of the words already written -- a cold path that touches the lane's ring or, if the words
// have been drained already, its scratch region.  The hot loop carries (lower, range) only and is branch-free
// apart from that carry.
//
// Per-stream results equal the reference's RangeEncoder / RangeDecoder (src/stream/queue.rs) word for word,
// including the seal words (queue.rs:349-376,458-523).
//
// Coder state on the wire (CTR_FLAG_RAW, states_in / states_out): 4 x u64 per stream,
//   encoder {lower, range, num_inverted, first_inverted_word}, decoder {lower, range, point, 0}.
// A raw-state encoder call returns the final words only: the held-back words of an unresolved Inverted
// situation are stripped from the output and described by (num_inverted, first_inverted_word), exactly what
// the reference's `pos()` reports; the next call writes them again before it continues.
#pragma once
#include "ans_kernels.cuh"

namespace ctr {

template <int BLOCK, bool SHARED, bool CONTIG, bool PERSYM>
__global__ void __launch_bounds__(BLOCK, BLOCK >= 256 ? 4 : 8) range_encode_kernel(const __grid_constant__ AnsParams p) {
    extern __shared__ __align__(128) uint32_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint64_t tma_bar[BLOCK / 32][kEncBoxSlots];

    const int lane = threadIdx.x & 31;
    const int warp_in_cta = threadIdx.x >> 5;
    constexpr int kWarpsPerCta = BLOCK / 32;

    // shared memory carve-up as in ans_encode_kernel: [rings + parking slots][replicated table][tiles]
    const uint32_t alphabet = p.model.alphabet;
    const uint32_t table_words = SHARED ? (alphabet + 1) * 32 : 0;
    constexpr uint32_t kRingsWords = BLOCK * (kEncRingWords + 4);
    const uint32_t ring = smem_u32_pinned(smem) + threadIdx.x * kEncRingBytes;
    const uint32_t park = smem_u32(smem) + BLOCK * kEncRingBytes + threadIdx.x * 16u;
    const uint32_t table_addr = smem_u32_pinned(smem + kRingsWords) + (uint32_t)(lane & 7) * 16u;
    uint32_t *sym_tile = smem + kRingsWords + table_words + warp_in_cta * (2 * kTileWords);
    uint32_t *idx_tile = smem + kRingsWords + table_words + kWarpsPerCta * (2 * kTileWords) + warp_in_cta * kTileWords;

    if (SHARED) stage_table(smem + kRingsWords, p.model.enc_rep, (alphabet + 1) * 128u, &bar);

    const uint64_t K = p.K, N = p.N;
    const uint32_t tile = take_tile_ticket(p.compact.ticket);
    const uint64_t k = (uint64_t)tile * BLOCK + threadIdx.x;
    const bool valid = k < K;
    const uint64_t kc = valid ? k : K - 1;  // lanes without a stream shadow the last one; nothing they push is stored

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

    uint64_t lower = 0, range = ~0ull;
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
    // The word pushed when `pushed` was `pos`: still in my ring, or already in my scratch region.
    auto word_in_ring = [&](uint32_t pos) { return pushed - pos <= pending; };
    auto scratch_word = [&](uint32_t pos) { return reinterpret_cast<uint32_t *>(gw - (pushed - pending - pos)); };
    // late carry (cold): add one to the words written so far -- the trailing 0xffffffff words wrap to zero and
    // the word before them absorbs the carry (it is <= 0xfffffffe, see coder_math.cuh)
    auto propagate_carry = [&]() {
        uint32_t pos = pushed;
        while (pos != 0u) {
            pos -= 4u;
            uint32_t w;
            if (word_in_ring(pos)) {
                const uint32_t a = ring | (pos & (kEncRingBytes - 1u));
                w = lds_u32(a) + 1u;
                sts_u32(a, w);
            } else {
                if (overflow) break;  // words were dropped: the stream is flagged anyway
                uint32_t *g = scratch_word(pos);
                w = __ldcg(g) + 1u;
                __stcg(g, w);
            }
            if (w != 0u) break;
        }
    };

    if (valid && p.states_in) {  // resume (queue.rs:182-196): the held-back words are written again
        lower = p.states_in[4 * k];
        range = p.states_in[4 * k + 1];
        const uint64_t held = p.states_in[4 * k + 2];
        const uint32_t first = (uint32_t)p.states_in[4 * k + 3];
        for (uint64_t j = 0; j < held; ++j) {
            push(j == 0 ? first : 0xffffffffu);
            drain_ring();
        }
    }

    auto lookup = [&](int32_t sym, uint32_t m) -> uint2 {
        uint32_t idx = min((uint32_t)sym - min_symbol, alphabet);  // out of range -> sentinel entry (prob 0)
        if (SHARED) return lds_table_v2(table_addr + idx * 128u);
        const bool ok = m < n_models;
        idx = ok ? idx : alphabet;
        m = ok ? m : 0u;
        const uint4 e = __ldg(p.model.enc + (uint64_t)m * (alphabet + 1) + idx);
        return make_uint2(e.x, e.y);
    };
    // one reference encode_symbol (queue.rs:612-705) on a looked-up (left, prob).  An impossible symbol
    // (prob 0) collapses the range to zero: the words from there on are garbage and the stream is flagged.
    // The most recent word is HELD in a register instead of being pushed right away: `lower` wraps for about one
    // symbol in fifty (the interval is often a sizeable fraction of 2^64) and the carry then is one add on that
    // register.  Only if there is no held word yet (a resumed coder before its first new word) or the held word itself
    // overflows (it was 0xffffffff: probability 2^-32 per carry) the carry ripples into words that have already left
    // for the ring or the scratch region -- the cold path.
    uint32_t held = 0u;
    bool has_held = false;
    auto encode_entry = [&](const uint2 &e) {
        min_prob = min(min_prob, e.y);
        const uint64_t scale = range >> kPrecision;
        const uint64_t nr = scale * (uint64_t)e.y;
        const uint64_t nl = lower + scale * (uint64_t)e.x;
        const bool wrap = nl < lower;
        held += (wrap && has_held) ? 1u : 0u;
        if (wrap && (!has_held || held == 0u)) propagate_carry();
        const bool renorm = (uint32_t)(nr >> 32) == 0u;
        const bool spill = renorm && has_held;  // a new word arrives: the held one goes to the ring
        if (spill) sts_u32(ring | (pushed & (kEncRingBytes - 1u)), held);
        pushed += spill ? 4u : 0u;
        pending += spill ? 4u : 0u;
        held = renorm ? (uint32_t)(nl >> 32) : held;
        has_held = has_held || renorm;
        lower = renorm ? nl << 32 : nl;
        range = renorm ? nr << 32 : nr;
    };
    auto encode_one = [&](int32_t sym, uint32_t m) { encode_entry(lookup(sym, m)); };

    if (!CONTIG && !PERSYM && p.use_tma) {
        // ---- TMA path (see ans_encode_kernel): the warp's column strip arrives as boxes of kBoxRows rows, queue order
        const Interleave g = interleave_of(N, K);
        const uint64_t rows_total = g.T - 1;  // full rows 0 .. T-2
        const uint32_t nbox = (uint32_t)(rows_total / kBoxRows);
        const uint32_t bars = smem_u32(&tma_bar[warp_in_cta][0]);
        const uint32_t boxes = smem_u32_pinned(smem + kRingsWords + table_words) + (uint32_t)warp_in_cta * (kEncBoxSlots * kBoxBytes);
        const int32_t x0 = (int32_t)((uint32_t)tile * BLOCK + (uint32_t)warp_in_cta * 32u);
        if (lane == 0) {
#pragma unroll
            for (int sl = 0; sl < kEncBoxSlots; ++sl) mbar_init_addr(bars + 8u * sl, 1);
            fence_mbar_init();
        }
        __syncwarp();
        uint32_t next = 0;  // boxes [next, nbox) are not requested yet
        auto request_box = [&](uint32_t slot) {
            if (next < nbox) {
                if (lane == 0) {
                    mbar_expect_tx_addr(bars + 8u * slot, kBoxBytes);
                    tma_load_box(boxes + slot * kBoxBytes, &p.tmap, x0, (int32_t)(next * kBoxRows), bars + 8u * slot);
                }
                next += 1u;
            }
        };
#pragma unroll
        for (int sl = 0; sl < kEncBoxSlots; ++sl) request_box(sl);
        uint32_t slot = 0, parity = 0;
        const uint32_t my_col = boxes + (uint32_t)lane * 4u;
        for (uint32_t b = 0; b < nbox; ++b) {
            mbar_wait_addr(bars + 8u * slot, parity);
            const uint32_t box = my_col + slot * kBoxBytes;
#pragma unroll
            for (int half = 0; half < kBoxRows / kCheckEvery; ++half) {
                uint2 e[kCheckEvery];
#pragma unroll
                for (int u = 0; u < kCheckEvery; ++u)
                    e[u] = lookup((int32_t)lds_u32(box + (uint32_t)(half * kCheckEvery + u) * 128u), stream_model);
                drain_ring();
#pragma unroll
                for (int u = 0; u < kCheckEvery; ++u) encode_entry(e[u]);
            }
            __syncwarp();  // every lane has read the box: its slot is requested again
            request_box(slot);
            if (++slot == kEncBoxSlots) {
                slot = 0;
                parity ^= 1u;
            }
        }
        {  // the rows after the last box, then the ragged last row
            const int32_t *ps = p.symbols_in + (uint64_t)nbox * kBoxRows * K + kc;
            uint32_t j = 0;
            for (uint64_t r = (uint64_t)nbox * kBoxRows; r < rows_total; ++r, ++j) {
                if ((j & (kCheckEvery - 1)) == 0) drain_ring();
                encode_one(ld_stream_s32(ps), stream_model);
                ps += K;
            }
            drain_ring();
            if (valid && k < g.last) encode_one(ld_stream_s32(p.symbols_in + (g.T - 1) * K + k), stream_model);
        }
    } else if (!CONTIG) {
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
            // L2 prefetch a few batches ahead, as in ans_encode_kernel
            constexpr int kPrefetchBatches = CTR_PF_BATCHES;
            const char *pf = ps + ((uint64_t)kPrefetchBatches * kCheckEvery + (uint32_t)(lane >> 3)) * row_bytes;
            const char *const pf_end = reinterpret_cast<const char *>(p.symbols_in + N);
            const uint64_t batch_bytes = row_bytes * kCheckEvery;
            auto code_batch = [&](int which, bool load_next) {
                // consume this batch's loads (issued a whole batch ago) before the next batch's are issued
                uint2 e[kCheckEvery];
#pragma unroll
                for (int u = 0; u < kCheckEvery; ++u) e[u] = lookup(buf[which][u], mbuf[which][u]);
                if (load_next) load_batch(which ^ 1);
                if (kPrefetchBatches > 0) {
                    if (pf < pf_end) prefetch_l2(pf);
                    pf += batch_bytes;
                }
                drain_ring();
#pragma unroll
                for (int u = 0; u < kCheckEvery; ++u) encode_entry(e[u]);
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
        // contiguous: double-buffered asynchronous tiles, lookups of four symbols ahead of their coding
        // (see ans_encode_kernel); queue order: symbols are consumed from the front
        const uint32_t tiles_addr = smem_u32(sym_tile);
        const uint32_t my_row = (uint32_t)lane * (kRowStride * 4u);
        const uint32_t idx_row = smem_u32(idx_tile) + my_row;
        uint64_t done = 0;  // symbols of my stream already requested
        const uint64_t rounds = (warp_max_u64(n_k, lane) + 31) / 32;
        uint32_t c_next = n_k < 32 ? (uint32_t)n_k : 32u;
        warp_fill_rows_async(tiles_addr, reinterpret_cast<const uint32_t *>(p.symbols_in + o_k), c_next, lane);
        cp_async_commit();
        const uint32_t ckpt_every = valid ? p.ckpt_every : 0u;
        const uint64_t ckpt_base = ckpt_every ? p.ckpt_off[k] : 0;
        uint32_t to_ckpt = 0;  // symbols until the next checkpoint (the first one is at symbol 0)
        for (uint64_t r = 0; r < rounds; ++r) {
            const uint32_t c = c_next;
            const uint64_t first = done;
            done += c;
            // checkpoint j = first / ckpt_every: the coder's position before symbol `first` (queue.rs:182-196)
            if (ckpt_every != 0u && c != 0u) {
                if (to_ckpt == 0u) {
                    uint64_t *rec = p.ckpt_out + 4u * (ckpt_base + first / ckpt_every);
                    rec[0] = (pushed >> 2) + (has_held ? 1u : 0u);
                    rec[1] = lower;
                    rec[2] = range;
                    rec[3] = 0;
                    to_ckpt = ckpt_every;
                }
                to_ckpt -= c;
            }
            const uint32_t row = tiles_addr + (uint32_t)(r & 1) * (kTileWords * 4u) + my_row;
            const uint64_t left_n = n_k - done;
            c_next = left_n < 32 ? (uint32_t)left_n : 32u;
            if (r + 1 < rounds)
                warp_fill_rows_async(tiles_addr + (uint32_t)((r + 1) & 1) * (kTileWords * 4u),
                                     reinterpret_cast<const uint32_t *>(p.symbols_in + o_k + done), c_next, lane);
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
                uint2 e[kCheckEvery];
#pragma unroll
                for (int u = 0; u < kCheckEvery; ++u) {
                    const uint32_t at = (s + (uint32_t)u) * 4u;
                    e[u] = lookup((int32_t)lds_u32(row + at), PERSYM ? lds_u32(idx_row + at) : stream_model);
                }
                drain_ring();
#pragma unroll
                for (int u = 0; u < kCheckEvery; ++u) encode_entry(e[u]);
            }
            for (; s < cmax; ++s) {  // ragged end of the round
                if ((s & (kCheckEvery - 1)) == 0) drain_ring();
                if (s < c) encode_one((int32_t)lds_u32(row + s * 4u), PERSYM ? lds_u32(idx_row + s * 4u) : stream_model);
            }
            __syncwarp();  // the tile is refilled by the next round's asynchronous fill
        }
    }

    // ---- seal (queue.rs:349-355, 458-523), or hand the raw state back to the caller -----------------------
    drain_ring();
    if (has_held) push(held);
    drain_ring();
    const bool raw = (p.flags & 1u) != 0;
    const bool bad = min_prob == 0u;
    RangeEncState st;
    st.lower = lower;
    st.range = range;
    if (valid && !bad && !raw) {
        const RangeSeal seal = range_seal(st);
        if (seal.carry) propagate_carry();
        if (seal.n >= 1u) push(seal.point_word);
        drain_ring();
        if (seal.n == 2u) push(0u);
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
    uint32_t gb_lo, gb_hi;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(gb_lo), "=r"(gb_hi) : "r"(park) : "memory");
    const uint32_t *gbegin = reinterpret_cast<const uint32_t *>(((uint64_t)gb_hi << 32) | gb_lo);
    uint32_t n_words = (valid && !overflow) ? (uint32_t)((reinterpret_cast<const uint32_t *>(gw)) - gbegin) : 0u;
    if (valid) {
        if (p.states_out) {
            // EncoderSituation (queue.rs:98-106) recovered from the words: Inverted(n, first) <=> the interval
            // wraps; its held-back words are `first` followed by n-1 words 0xffffffff at the end of the stream
            uint32_t held = 0, first = 0;
            if (raw && !bad && !overflow && st.range != ~0ull && range_enc_inverted(st) && n_words != 0u) {
                do {
                    held += 1u;
                    first = __ldcg(gbegin + (n_words - held));
                } while (first == 0xffffffffu && held < n_words);
                n_words -= held;
            }
            p.states_out[4 * k] = st.lower;
            p.states_out[4 * k + 1] = st.range;
            p.states_out[4 * k + 2] = held;
            p.states_out[4 * k + 3] = first;
        }
        if (bad)
            report_error(p.status, kErrImpossibleSymbol, k);
        else if (overflow)
            report_error(p.status, kErrOutOfSpace, k);
    }
    compact_tail<BLOCK>(p.compact, tile, k, K, valid, gbegin, n_words, p.status);
}

template <int BLOCK, int TABLE, bool CONTIG, bool PERSYM, bool SMALL>
__global__ void __launch_bounds__(BLOCK ? BLOCK : 1024, BLOCK == 0 || BLOCK >= 1024 ? 1 : (BLOCK >= 256 ? 2 : 8))
    range_decode_kernel(const AnsParams p) {
    extern __shared__ __align__(128) uint32_t smem[];
    __shared__ uint64_t bar;

    constexpr bool SHARED = TABLE == kTableLut, POOL = TABLE == kTablePool, GAUSS = TABLE == kTableGauss;
    // one CTA per SM: room for the finer quantile index (p.model.dec_big), as in ans_decode_kernel
    constexpr bool BIG_LUT = SHARED && BLOCK == kDecBlockShared;
    constexpr int kIndexBits = BIG_LUT ? kBigLutBits : kLutBits;
    constexpr uint32_t kIndexBytes = 8u << kIndexBits;
    const uint32_t kBlock = BLOCK ? (uint32_t)BLOCK : blockDim.x;
    const int lane = threadIdx.x & 31;
    const int warp_in_cta = threadIdx.x >> 5;
    const uint32_t kWarpsPerCta = kBlock / 32;

    // shared memory carve-up as in ans_decode_kernel
    const uint32_t alphabet = p.model.alphabet;
    const uint32_t table_words = SHARED ? (kIndexBytes + p.model.dec_cdf_bytes) / 4
                                        : (POOL ? (p.model.pool_cdf_bytes + p.model.pool_cidx_bytes) / 4 : 0);
    const uint32_t kRingsWords = kBlock * kDecRingWords;
    const uint32_t ring = smem_u32_pinned(smem) + threadIdx.x * kDecRingBytes;  // 64-byte aligned
    const uint32_t lut_addr = smem_u32_pinned(smem + kRingsWords);
    uint32_t cdf_addr = lut_addr + (POOL ? 0u : kIndexBytes);
    asm volatile("" : "+r"(cdf_addr));
    uint32_t *sym_tile = smem + kRingsWords + table_words + warp_in_cta * kTileWords;
    uint32_t *idx_tile = sym_tile + kWarpsPerCta * kTileWords;

    if (SHARED) stage_table(smem + kRingsWords, BIG_LUT ? p.model.dec_big : p.model.dec, kIndexBytes + p.model.dec_cdf_bytes, &bar);
    if (POOL)
        stage_tables(smem + kRingsWords, p.model.cdf, p.model.pool_cdf_bytes, smem + kRingsWords + p.model.pool_cdf_bytes / 4,
                     p.model.cidx, p.model.pool_cidx_bytes, &bar);

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
            if (o_k > N || n_k > N - o_k) {  // offsets outside the symbol array (or decreasing): flag, decode nothing
                report_error(p.status, kErrBadArgument, k);
                o_k = 0;
                n_k = 0;
            }
        }
        begin = p.offsets[k];
        end = p.ends ? p.ends[k] : p.offsets[k + 1];
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
    // (two blocks in flight: a request has two check intervals to arrive; see ans_decode_kernel for the invariant)
    uint32_t pending_old = 0;
    auto top_up = [&]() {
        cp_async_wait_group<1>();
        avail += pending_old;
        pending_old = pending;
        pending = 0;
        if (avail + pending_old <= (uint32_t)(kDecRingWords - 4) && unstaged != 0u) request_block(4u);
        cp_async_commit();
    };
    {
        const uint32_t first = 4u - (uint32_t)(begin & 3u);  // words of my stream in the bottom block
        if (unstaged != 0u) request_block(first);
        cp_async_commit();
#pragma unroll 1
        for (int i = 0; i < 3; ++i) top_up();
        cp_async_wait_all();
        avail += pending + pending_old;
        pending = 0;
        pending_old = 0;
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
    const uint32_t pool_row_bytes = (alphabet + 1) * 4u;
    const uint32_t pool_cidx_stride = (alphabet > 256 ? 2u : 1u) * (kCoarseSize + 1);
    const uint32_t pool_cidx_addr = cdf_addr + p.model.pool_cdf_bytes;

    // one reference decode_symbol (queue.rs:968-1035); after invalid data the lane keeps running on a
    // clamped quantile (its symbols are garbage and the stream is flagged)
    bool bad_model = false;  // GAUSS: a std that is not > 0
    // `converged`: std::true_type where every lane of the warp makes this call together (see ans_decode_kernel)
    auto decode_one = [&](uint32_t m, auto converged) -> int32_t {
        constexpr bool kVote = BIG_LUT && decltype(converged)::value;
        uint32_t q = kQuantileMask;
        invalid_data |= !range_peek_quantile(st, q);
        uint32_t left, right, s;
        if (SHARED) {
            s = lookup_shared<SMALL, kIndexBits, kVote>(lut_addr, cdf_addr, alphabet, q, q, left, right);
        } else if (GAUSS) {
            m = m < n_models ? m : n_models - 1;
            const double mean = __ldg(p.gauss_means + m), std = __ldg(p.gauss_stds + m);
            bad_model |= !(std > 0.0);
            s = gauss_quantile(q, mean, std, p.gauss_free_weight, p.model.min_symbol, alphabet, left, right);
        } else if (POOL) {
            m = m < n_models ? m : n_models - 1;
            s = lookup_pool(cdf_addr + m * pool_row_bytes, pool_cidx_addr + m * pool_cidx_stride, alphabet > 256, q, left, right);
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
                        const int32_t sym = decode_one(mbuf[u], std::true_type{});
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
        if (g.T > 0) {
            if (valid && k < g.last) {
                const uint64_t i = (g.T - 1) * K + k;
                const int32_t sym = decode_one(PERSYM ? ld_stream_u32(p.model_index + i) : stream_model, std::false_type{});
                st_stream_s32(p.symbols_out + i, sym);
            }
        }
    } else {
        uint64_t done = 0;
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
        if (p.states_out) {
            p.states_out[4 * k] = st.lower;
            p.states_out[4 * k + 1] = st.range;
            p.states_out[4 * k + 2] = st.point;
            p.states_out[4 * k + 3] = 0;
        }
        // words consumed so far (Pos::pos().0, queue.rs:182-196)
        if (p.words_left) p.words_left[k] = (uint64_t)(total_words - unstaged - pending - pending_old - avail);
        if (invalid_data) report_error(p.status, kErrInvalidData, k);
        if (GAUSS && bad_model) report_error(p.status, kErrBadModel, k);
    }
}

}  // namespace ctr
