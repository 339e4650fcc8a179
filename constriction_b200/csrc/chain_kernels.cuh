// chain_kernels.cuh -- coder kernels for FEW, LONG streams (contiguous layout), e.g. BASELINE configs[3]: 1024 streams
// of 122,070 symbols per GPU.
//
// With few streams a lane-per-stream kernel has about one warp per SM sub-partition, so nothing hides latency: the
// time per symbol is the length of the dependent instruction chain of ONE warp, and in the general kernels that warp
// also loads and transposes symbols, looks models up, walks its word ring and stores results.  Here a CTA is one CODER
// warp (lane = stream, 32 streams) plus helper warps, and the coder warp's loop contains the loop-carried state update
// and nothing else:
//   encoders: kChainProducers PRODUCER warps fetch symbols ahead of the coder (coalesced 128-byte reads of each
//             stream), look up (left, probability[, reciprocal]) -- work that does not depend on the coder state -- and
//             hand tiles of 32 streams x 32 entries to the coder through a ring in shared memory (mbarrier full /
//             empty pairs, XOR-swizzled so that both sides are bank-conflict free);
//   (decoders cannot look ahead: the next symbol depends on the state.)
// Results are word for word those of the general kernels (same arithmetic from coder_math.cuh, same fused compaction).
#pragma once
#include "range_kernels.cuh"

namespace ctr {

constexpr int kChainBlock = 128;    // warp 0 codes, warps 1..3 produce
constexpr int kChainProducers = 3;
constexpr int kChainSlots = 4;      // tiles in the ring
constexpr uint32_t kChainTileEntries = 32u * 32u;

__device__ __forceinline__ void mbar_arrive_addr(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint2 lds_v2(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_v2(uint32_t addr, const uint2 &v) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void sts_v4(uint32_t addr, const uint4 &v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ANS: entries are {left, prob, reciprocal lo, hi} (16 bytes), symbols are consumed from the END of every stream;
// range: entries are {left, prob} (8 bytes), symbols are consumed from the front.
//   SHARED : model 0's encoder entries staged in shared memory (index_mode NONE, small alphabet); else global table
//            with one model per stream (p.model_index, mode 2) or model 0
template <bool ANS, bool SHARED, bool F64DIV>
__global__ void __launch_bounds__(kChainBlock) encode_chain_kernel(const __grid_constant__ AnsParams p) {
    extern __shared__ __align__(128) uint32_t smem[];
    __shared__ uint64_t bar_full[kChainSlots], bar_empty[kChainSlots];
    __shared__ uint64_t s_off[32], s_len[32];
    __shared__ uint32_t s_model[32];
    __shared__ uint64_t s_rounds;

    constexpr uint32_t kEntryBytes = ANS ? 16u : 8u;
    constexpr uint32_t kTileBytes = kChainTileEntries * kEntryBytes;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t alphabet = p.model.alphabet;
    // shared memory: [32 lane rings + parking][table (SHARED)][kChainSlots tiles]
    constexpr uint32_t kRingsWords = 32 * (kEncRingWords + 4);
    const uint32_t table_bytes = SHARED ? (alphabet + 1) * 16u : 0u;
    const uint32_t ring = smem_u32_pinned(smem) + (uint32_t)lane * kEncRingBytes;
    const uint32_t table_addr = smem_u32(smem + kRingsWords);
    const uint32_t tiles_addr = ((table_addr + table_bytes + 127u) & ~127u);

    const uint64_t K = p.K, N = p.N;
    const uint32_t tile = take_tile_ticket(p.compact.ticket);  // which 32 streams this CTA codes (includes a CTA barrier)
    const uint64_t k = (uint64_t)tile * 32 + (uint32_t)lane;
    const bool valid = warp == 0 && k < K;

    if (warp == 0) {
        uint64_t o = 0, n = 0;
        if (k < K) {
            o = p.sym_off[k];
            n = p.sym_off[k + 1] - o;
            if (o > N || n > N - o) {
                report_error(p.status, kErrBadArgument, k);
                o = 0;
                n = 0;
            }
        }
        s_off[lane] = o;
        s_len[lane] = n;
        s_model[lane] = (k < K && p.index_mode == 2) ? p.model_index[k] : 0u;
        const uint64_t longest = warp_max_u64(n, lane);
        if (lane == 0) {
            s_rounds = (longest + 31) / 32;
            for (int s = 0; s < kChainSlots; ++s) {
                mbar_init(&bar_full[s], 1);
                mbar_init(&bar_empty[s], 1);
            }
            fence_mbar_init();
        }
    }
    if (SHARED) {  // model 0's entries {left, prob, rcp lo, rcp hi}, 16 bytes each
        const uint4 *src = p.model.enc;
        for (uint32_t i = threadIdx.x; i <= alphabet; i += kChainBlock) sts_v4(table_addr + i * 16u, __ldg(src + i));
    }
    __syncthreads();
    const uint64_t rounds = s_rounds;
    const uint32_t full0 = smem_u32(&bar_full[0]), empty0 = smem_u32(&bar_empty[0]);

    if (warp != 0) {
        // ---------------- producers: tile t = symbols [32 t, 32 t + 32) of every stream, counted in coding order ------
        const uint32_t min_symbol = (uint32_t)p.model.min_symbol;
        const uint32_t n_models = p.model.n_models;
        for (uint64_t t = (uint64_t)(warp - 1); t < rounds; t += kChainProducers) {
            const uint32_t slot = (uint32_t)(t % kChainSlots);
            const uint64_t use = t / kChainSlots;
            if (use > 0) mbar_wait_addr(empty0 + 8u * slot, (uint32_t)((use - 1) & 1));
            const uint32_t base = tiles_addr + slot * kTileBytes;
#pragma unroll 1
            for (int s0 = 0; s0 < 32; s0 += 8) {
                int32_t sym[8];
                bool have[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {  // eight streams' loads in flight
                    const uint64_t n = s_len[s0 + u], pos = t * 32 + (uint32_t)lane;
                    have[u] = pos < n;
                    // ANS codes a stream backwards: coding position `pos` is symbol n - 1 - pos
                    const uint64_t at = s_off[s0 + u] + (ANS ? n - 1 - pos : pos);
                    sym[u] = have[u] ? ld_stream_s32(p.symbols_in + at) : 0;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int s = s0 + u;
                    uint32_t idx = min((uint32_t)sym[u] - min_symbol, alphabet);  // out of range -> sentinel (prob 0)
                    uint4 e;
                    if (SHARED) {
                        e = lds_table_v4(table_addr + idx * 16u);
                    } else {
                        uint32_t m = s_model[s];
                        const bool ok = m < n_models;
                        idx = ok ? idx : alphabet;
                        m = ok ? m : 0u;
                        e = __ldg(p.model.enc + (uint64_t)m * (alphabet + 1) + idx);
                    }
                    if (!have[u]) e = make_uint4(0u, 1u, 0u, 0u);  // never coded; must not look impossible
                    // entry (stream s, position lane) lives at row `lane`, column s ^ lane: both the producers (fixed s,
                    // lanes = rows) and the coder (fixed row, lanes = streams) touch 32 different columns
                    const uint32_t a = base + ((uint32_t)lane * 32u + ((uint32_t)s ^ (uint32_t)lane)) * kEntryBytes;
                    if (ANS)
                        sts_v4(a, e);
                    else
                        sts_v2(a, make_uint2(e.x, e.y));
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_addr(full0 + 8u * slot);
        }
    }

    // ---------------- the coder warp ----------------------------------------------------------------------------
    const uint64_t n_k = warp == 0 ? s_len[lane] : 0;
    const uint64_t o_k = warp == 0 ? s_off[lane] : 0;
    char *gw = nullptr;
    uint32_t room = 0;
    const uint32_t *gbegin = nullptr;
    uint32_t n_words = 0;
    if (warp == 0) {
        uint32_t *const gb = p.scratch + scratch_start(o_k, k < K ? k : 0);
        gbegin = gb;
        const uint64_t r = valid ? scratch_start(o_k + n_k, k + 1) - scratch_start(o_k, k) : 0;
        room = r > 0x3ffffff0u ? 0xffffffc0u : (uint32_t)r * 4u;
        gw = reinterpret_cast<char *>(gb);
        uint32_t pushed = 0, pending = 0;  // bytes pushed into my ring / not yet written to scratch
        uint32_t min_prob = 0xffffffffu;
        bool overflow = false;
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
        // the same without a branch (with one warp per scheduler every branch is a pipeline bubble)
        auto drain_ring_flat = [&]() {
            const bool full = pending >= 16u;
            const bool fits = room >= 16u;
            const uint4 v = lds_v4(ring | ((pushed - pending) & (kEncRingBytes - 16u)));
            if (full && fits) st_stream_v4(gw, v);
            overflow |= full && !fits;
            const uint32_t step = (full && fits) ? 16u : 0u;
            gw += step;
            room = (full && !fits) ? 0u : room - step;
            pending -= full ? 16u : 0u;
        };
        const uint32_t ckpt_every = valid ? p.ckpt_every : 0u;
        const uint64_t ckpt_base = ckpt_every ? p.ckpt_off[k] : 0;

        if (ANS) {
            uint64_t state = 0;
            for (uint64_t t = 0; t < rounds; ++t) {
                const uint32_t slot = (uint32_t)(t % kChainSlots);
                mbar_wait_addr(full0 + 8u * slot, (uint32_t)((t / kChainSlots) & 1));
                const uint64_t done = t * 32;
                const uint32_t c = n_k > done ? (uint32_t)min((uint64_t)32, n_k - done) : 0u;
                const uint32_t base = tiles_addr + slot * kTileBytes;
                const uint32_t cmin = __reduce_min_sync(kFullMask, c), cmax = __reduce_max_sync(kFullMask, c);
                auto code = [&](const uint4 &e) {
                    min_prob = min(min_prob, e.y);
                    uint32_t lo = (uint32_t)state, hi = (uint32_t)(state >> 32);
                    const bool flush = (hi >> 8) >= e.y;  // stack.rs:1035-1040
                    if (flush) sts_u32(ring | (pushed & (kEncRingBytes - 1u)), lo);
                    pushed += flush ? 4u : 0u;
                    pending += flush ? 4u : 0u;
                    lo = flush ? hi : lo;
                    hi = flush ? 0u : hi;
                    const uint64_t n = ((uint64_t)hi << 32) | lo;
                    state = ans_encode_recombine(n, ans_quotient_estimate<F64DIV>(n, e.z, e.w), e.x, e.y);
                };
                uint32_t j = 0;
                for (; j + 4 <= cmin; j += 4) {
                    uint4 e[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) e[u] = lds_v4(base + ((j + u) * 32u + ((uint32_t)lane ^ (j + u))) * 16u);
                    drain_ring_flat();
#pragma unroll
                    for (int u = 0; u < 4; ++u) code(e[u]);
                }
                for (; j < cmax; ++j) {
                    if ((j & 3u) == 0u) drain_ring();
                    if (j < c) code(lds_v4(base + (j * 32u + ((uint32_t)lane ^ j)) * 16u));
                }
                // checkpoint: the symbols [n_k - done - c, n_k) are coded.  Record j (decode order) starts the chunk
                // at symbol first = n_k - done - c when that is a multiple of ckpt_every counted from the end, or 0.
                if (ckpt_every != 0u && c != 0u) {
                    const uint64_t first = n_k - done - c, coded = done + c;
                    if (first == 0 || coded % ckpt_every == 0) {
                        uint64_t *rec = p.ckpt_out + 2u * (ckpt_base + (first + ckpt_every - 1u) / ckpt_every);
                        rec[0] = pushed >> 2;
                        rec[1] = state;
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive_addr(empty0 + 8u * slot);
            }
            drain_ring();
            const uint32_t n_state = valid ? ans_state_words(state) : 0u;  // lib.rs:719-730, low word first
            if (n_state >= 1) push((uint32_t)state);
            if (n_state == 2) push((uint32_t)(state >> 32));
            drain_ring();
        } else {
            uint64_t lower = 0, range = ~0ull;
            auto word_in_ring = [&](uint32_t pos) { return pushed - pos <= pending; };
            // late carry (cold), see range_encode_kernel: add one to the words pushed before byte position `from`
            auto propagate_carry_from = [&](uint32_t from) {
                uint32_t pos = from;
                while (pos != 0u) {
                    pos -= 4u;
                    uint32_t w;
                    if (word_in_ring(pos)) {
                        const uint32_t a = ring | (pos & (kEncRingBytes - 1u));
                        w = lds_u32(a) + 1u;
                        sts_u32(a, w);
                    } else {
                        if (overflow) break;
                        uint32_t *g = reinterpret_cast<uint32_t *>(gw - (pushed - pending - pos));
                        w = __ldcg(g) + 1u;
                        __stcg(g, w);
                    }
                    if (w != 0u) break;
                }
            };
            auto propagate_carry = [&]() { propagate_carry_from(pushed); };
            // One symbol (queue.rs:612-705), straight-line: with one warp per scheduler every branch is a pipeline
            // bubble, so a wrap of `lower` is only RECORDED here (bit u of `wraps`, and where the words ended at that
            // moment) and the carry is applied after the group of four -- additions into disjoint word prefixes commute.
            // The most recent word is HELD in a register instead of being pushed: `lower` wraps for about one symbol in
            // fifty (the interval is often a sizeable fraction of 2^64), and the carry then is one add on that register.
            // Only if the held word itself overflows (it was 0xffffffff: probability 2^-32 per carry) the carry has to
            // ripple into words that already left for the ring -- the cold path below.
            uint32_t held = 0u;
            bool has_held = false;
            uint32_t wraps = 0u, wrap_pos[4];
            auto code_at = [&](const uint2 &e, int u) {
                min_prob = min(min_prob, e.y);
                const uint64_t scale = range >> kPrecision;
                const uint64_t nr = scale * (uint64_t)e.y;
                const uint64_t nl = lower + scale * (uint64_t)e.x;
                const bool wrap = nl < lower;
                held += wrap ? 1u : 0u;
                wraps |= ((wrap && held == 0u) ? 1u : 0u) << u;  // the held word overflowed: ripple into earlier words
                wrap_pos[u] = pushed;
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
            auto settle_wraps = [&]() {
                if (wraps != 0u) {
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if ((wraps >> u) & 1u) propagate_carry_from(wrap_pos[u]);
                    wraps = 0u;
                }
            };
            auto code = [&](const uint2 &e) {
                code_at(e, 0);
                settle_wraps();
            };
            for (uint64_t t = 0; t < rounds; ++t) {
                const uint32_t slot = (uint32_t)(t % kChainSlots);
                mbar_wait_addr(full0 + 8u * slot, (uint32_t)((t / kChainSlots) & 1));
                const uint64_t done = t * 32;
                const uint32_t c = n_k > done ? (uint32_t)min((uint64_t)32, n_k - done) : 0u;
                if (ckpt_every != 0u && c != 0u && done % ckpt_every == 0) {  // the coder's position before symbol `done`
                    uint64_t *rec = p.ckpt_out + 4u * (ckpt_base + done / ckpt_every);
                    rec[0] = (pushed >> 2) + (has_held ? 1u : 0u);
                    rec[1] = lower;
                    rec[2] = range;
                    rec[3] = 0;
                }
                const uint32_t base = tiles_addr + slot * kTileBytes;
                const uint32_t cmin = __reduce_min_sync(kFullMask, c), cmax = __reduce_max_sync(kFullMask, c);
                uint32_t j = 0;
                for (; j + 4 <= cmin; j += 4) {
                    uint2 e[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) e[u] = lds_v2(base + ((j + u) * 32u + ((uint32_t)lane ^ (j + u))) * 8u);
                    drain_ring_flat();
#pragma unroll
                    for (int u = 0; u < 4; ++u) code_at(e[u], u);
                    settle_wraps();
                }
                for (; j < cmax; ++j) {
                    if ((j & 3u) == 0u) drain_ring();
                    if (j < c) code(lds_v2(base + (j * 32u + ((uint32_t)lane ^ j)) * 8u));
                }
                __syncwarp();
                if (lane == 0) mbar_arrive_addr(empty0 + 8u * slot);
            }
            if (has_held) push(held);
            drain_ring();
            if (valid && min_prob != 0u) {  // seal (queue.rs:349-355, 458-523)
                RangeEncState st;
                st.lower = lower;
                st.range = range;
                const RangeSeal seal = range_seal(st);
                if (seal.carry) propagate_carry();
                if (seal.n >= 1u) push(seal.point_word);
                drain_ring();
                if (seal.n == 2u) push(0u);
                drain_ring();
            }
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
        n_words = (valid && !overflow) ? (uint32_t)(reinterpret_cast<const uint32_t *>(gw) - gbegin) : 0u;
        if (valid) {
            if (min_prob == 0u)
                report_error(p.status, kErrImpossibleSymbol, k);
            else if (overflow)
                report_error(p.status, kErrOutOfSpace, k);
        }
    }
    // K6: the CTA's 32 streams go to their place in the dense container (threads of the helper warps own no stream)
    compact_tail<kChainBlock>(p.compact, tile, k, K, valid, gbegin, n_words, p.status);
}

// =====================================================================================================================
// decoders: warp 0 decodes (lane = stream), warp 1 writes finished 32 x 32 symbol tiles to global memory with one
// 128-byte store per stream and tile.  The coder warp's loop is straight-line: the next compressed word is fetched
// from the lane's ring BEFORE it is known whether the state needs it (its address does not depend on the state), the
// refill is a pair of selects, and symbols go to shared memory.
//   TABLE : kTableLut (one shared model: quantile index + cdf in shared memory) or kTablePool (model set in shared
//           memory, one model per stream)
// =====================================================================================================================
// u64 -> float within a few ulp and 1 / x to 22 bits: three and one instruction (the results only seed an estimate)
__device__ __forceinline__ float u64_to_float_cheap(uint64_t v) {
    return fmaf(__uint2float_rz((uint32_t)(v >> 32)), 4294967296.0f, __uint2float_rz((uint32_t)v));
}
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

constexpr int kDecChainBlock = 64;
constexpr int kDecChainTiles = 4;

template <bool RANGE, int TABLE, bool SMALL>
__global__ void __launch_bounds__(kDecChainBlock) decode_chain_kernel(const __grid_constant__ AnsParams p) {
    extern __shared__ __align__(128) uint32_t smem[];
    __shared__ uint64_t bar, bar_full[kDecChainTiles], bar_empty[kDecChainTiles];
    __shared__ uint64_t s_off[32], s_len[32];
    __shared__ uint64_t s_rounds;
    constexpr bool SHARED = TABLE == kTableLut, POOL = TABLE == kTablePool;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t alphabet = p.model.alphabet;
    // shared memory: [32 lane rings (64 B each)][tables][kDecChainTiles symbol tiles of 32 x 32 words]
    constexpr uint32_t kRingsWords = 32 * kDecRingWords;
    // SHARED: the finer quantile index (p.model.dec_big): its rarely taken second probe is off the coder's chain
    constexpr uint32_t kIndexBytes = kBigLutBytes;
    const uint32_t table_words = SHARED ? (kIndexBytes + p.model.dec_cdf_bytes) / 4 : (p.model.pool_cdf_bytes + p.model.pool_cidx_bytes) / 4;
    const uint32_t ring = smem_u32_pinned(smem) + (uint32_t)lane * kDecRingBytes;
    const uint32_t lut_addr = smem_u32_pinned(smem + kRingsWords);
    uint32_t cdf_addr = lut_addr + (POOL ? 0u : kIndexBytes);
    asm volatile("" : "+r"(cdf_addr));
    const uint32_t tiles_addr = smem_u32(smem + kRingsWords + table_words);

    if (SHARED) stage_table(smem + kRingsWords, p.model.dec_big, kIndexBytes + p.model.dec_cdf_bytes, &bar);
    if (POOL)
        stage_tables(smem + kRingsWords, p.model.cdf, p.model.pool_cdf_bytes, smem + kRingsWords + p.model.pool_cdf_bytes / 4,
                     p.model.cidx, p.model.pool_cidx_bytes, &bar);

    const uint64_t K = p.K, N = p.N;
    const uint64_t k = (uint64_t)blockIdx.x * 32 + (uint32_t)lane;
    const bool valid = warp == 0 && k < K;
    if (warp == 0) {
        uint64_t o = 0, n = 0;
        if (k < K) {
            o = p.sym_off[k];
            n = p.sym_off[k + 1] - o;
            if (o > N || n > N - o) {
                report_error(p.status, kErrBadArgument, k);
                o = 0;
                n = 0;
            }
        }
        s_off[lane] = o;
        s_len[lane] = n;
        const uint64_t longest = warp_max_u64(n, lane);
        if (lane == 0) {
            s_rounds = (longest + 31) / 32;
            for (int s = 0; s < kDecChainTiles; ++s) {
                mbar_init(&bar_full[s], 1);
                mbar_init(&bar_empty[s], 1);
            }
            fence_mbar_init();
        }
    }
    __syncthreads();
    const uint64_t rounds = s_rounds;
    const uint32_t full0 = smem_u32(&bar_full[0]), empty0 = smem_u32(&bar_empty[0]);

    if (warp == 1) {
        // ---------------- writer: tile t = symbols [32 t, 32 t + 32) of the CTA's 32 streams ---------------------------
        for (uint64_t t = 0; t < rounds; ++t) {
            const uint32_t slot = (uint32_t)(t % kDecChainTiles);
            mbar_wait_addr(full0 + 8u * slot, (uint32_t)((t / kDecChainTiles) & 1));
            const uint32_t base = tiles_addr + slot * 4096u;
#pragma unroll 8
            for (int s = 0; s < 32; ++s) {
                const uint64_t n = s_len[s], pos = t * 32 + (uint32_t)lane;
                const uint32_t v = lds_u32(base + ((uint32_t)lane * 32u + ((uint32_t)s ^ (uint32_t)lane)) * 4u);
                if (pos < n) st_stream_s32(p.symbols_out + s_off[s] + pos, (int32_t)v);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_addr(empty0 + 8u * slot);
        }
        return;
    }

    // ---------------- the coder warp ------------------------------------------------------------------------------------
    const uint64_t n_k = s_len[lane];
    const bool raw = (p.flags & 1u) != 0;
    uint64_t begin = 0, end = 0;
    if (valid) {
        begin = p.offsets[k];
        end = p.ends ? p.ends[k] : p.offsets[k + 1];
    }
    // word staging exactly as in ans_decode_kernel / range_decode_kernel: ring slot = global address mod 64
    uint32_t avail = 0, pending = 0;
    uint32_t unstaged = (uint32_t)(end - begin);
    uint32_t pop_off;
    const char *gblock;
    if (RANGE) {
        pop_off = (uint32_t)(uintptr_t)(p.words + begin);
        gblock = reinterpret_cast<const char *>(p.words) + ((begin * 4u) & ~(uint64_t)15);
    } else {
        pop_off = (uint32_t)(uintptr_t)(p.words + end);
        gblock = reinterpret_cast<const char *>(p.words) + ((end * 4u) & ~(uint64_t)15);
        if ((end & 3u) == 0) gblock -= 16;
    }
    auto request_block = [&](uint32_t block_words) {
        const uint32_t n = unstaged < block_words ? unstaged : block_words;
        cp_async_16(ring | ((uint32_t)(uintptr_t)gblock & (kDecRingBytes - 1u)), gblock);
        gblock += RANGE ? 16 : -16;
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
        const uint32_t first = RANGE ? 4u - (uint32_t)(begin & 3u) : ((end & 3u) ? (uint32_t)(end & 3u) : 4u);
        if (unstaged != 0u) request_block(first);
        cp_async_commit();
#pragma unroll 1
        for (int i = 0; i < 3; ++i) top_up();
        cp_async_wait_all();
        avail += pending + pending_old;
        pending = 0;
        pending_old = 0;
    }
    // the word the next refill would take (valid only while avail != 0)
    auto peek_word = [&]() -> uint32_t { return lds_u32(ring | ((RANGE ? pop_off : pop_off - 4u) & (kDecRingBytes - 1u))); };

    const uint32_t n_models = p.model.n_models;
    uint32_t min_symbol = (uint32_t)p.model.min_symbol;
    asm volatile("" : "+r"(min_symbol));
    uint32_t m = (valid && p.index_mode == 2) ? p.model_index[k] : 0u;
    m = m < n_models ? m : n_models - 1;
    const uint32_t pool_row = cdf_addr + m * ((alphabet + 1) * 4u);
    const uint32_t pool_cidx = cdf_addr + p.model.pool_cdf_bytes + m * ((alphabet > 256 ? 2u : 1u) * (kCoarseSize + 1));
    // `converged`: std::true_type where the whole coder warp makes the call together (the lookup may then vote)
    auto lookup = [&](uint32_t word, uint32_t q, uint32_t &left, uint32_t &right, auto converged) -> uint32_t {
        if (SHARED)
            return lookup_shared<SMALL, kBigLutBits, decltype(converged)::value>(lut_addr, cdf_addr, alphabet, word, q, left, right);
        return lookup_pool(pool_row, pool_cidx, alphabet > 256, q, left, right);
    };

    uint32_t lo = 0, hi = 0;  // ANS state
    RangeDecState st;         // range state
    st.lower = 0;
    st.range = ~0ull;
    st.point = 0;
    bool trailing_zero = false, invalid_data = false;
    if (raw) {
        if (valid && p.states_in) {
            if (RANGE) {
                st.lower = p.states_in[4 * k];
                st.range = p.states_in[4 * k + 1];
                st.point = p.states_in[4 * k + 2];
            } else {
                const uint64_t s = p.states_in[k];
                lo = (uint32_t)s;
                hi = (uint32_t)(s >> 32);
            }
        }
    } else if (RANGE) {  // queue.rs:755-773,847-868
        if (avail != 0u) {
            st.point = (uint64_t)peek_word() << 32;
            pop_off += 4u;
            avail -= 1u;
            if (avail != 0u) {
                st.point |= peek_word();
                pop_off += 4u;
                avail -= 1u;
            }
        }
    } else {  // stack.rs:299-318,440-462
        if (avail != 0u) {
            lo = peek_word();
            pop_off -= 4u;
            avail -= 1u;
            trailing_zero = lo == 0u;
            if (avail != 0u && lo != 0u) {
                hi = lo;
                lo = peek_word();
                pop_off -= 4u;
                avail -= 1u;
            }
        }
    }
    top_up();
    uint32_t nxt = peek_word();
    float rscale = rcp_approx(u64_to_float_cheap(st.range >> kPrecision));  // 1 / scale, single precision (RANGE)

    // one symbol; `act` = this lane still owns a symbol at this position
    auto decode_one = [&](bool act_in, auto all_tag) -> uint32_t {
        const bool act = decltype(all_tag)::value ? true : act_in;  // ALL: every lane owns this position (full tile)
        uint32_t left, right, s;
        if (RANGE) {
            // queue.rs:989-993 needs q = (point - lower) / (range >> 24) only to find the symbol with left <= q < right,
            // i.e. scale * left <= point - lower < scale * right.  So the division is ESTIMATED in single precision
            // (reciprocal of the scale computed right after the previous update, off the critical path; error of a few
            // quantiles), the symbol is looked up for the estimate, and the exact test is the decoder's own invariant
            // after the update, (point - new_lower) < new_range.  Only if it fails (the estimate fell on the wrong side
            // of a symbol boundary: ~1e-4 per symbol) the exact division runs.
            const uint64_t scale = st.range >> kPrecision;
            const uint64_t diff = st.point - st.lower;
            invalid_data |= act && diff >= (scale << kPrecision);
            uint32_t q = __float2uint_rz(u64_to_float_cheap(diff) * rscale);
            q = min(q, kQuantileMask);
            s = lookup(q, q, left, right, std::true_type{});
            uint64_t nl = st.lower + scale * (uint64_t)left;
            uint64_t nr = scale * (uint64_t)(right - left);
            if (st.point - nl >= nr) {  // cold: exact quantile
                q = kQuantileMask;
                range_peek_quantile(st, q);
                s = lookup(q, q, left, right, std::false_type{});  // (only the lanes whose estimate missed are here)
                nl = st.lower + scale * (uint64_t)left;
                nr = scale * (uint64_t)(right - left);
            }
            const bool renorm = (uint32_t)(nr >> 32) == 0u;  // queue.rs:1018-1032
            const bool pop = act && renorm && avail != 0u;
            const uint64_t np = renorm ? ((st.point << 32) | (pop ? nxt : 0u)) : st.point;
            st.lower = act ? (renorm ? nl << 32 : nl) : st.lower;
            st.range = act ? (renorm ? nr << 32 : nr) : st.range;
            st.point = act ? np : st.point;
            rscale = rcp_approx(u64_to_float_cheap(st.range >> kPrecision));
            pop_off += pop ? 4u : 0u;
            avail -= pop ? 1u : 0u;
        } else {
            const uint32_t q = lo & kQuantileMask;
            s = lookup(lo, q, left, right, std::true_type{});
            const uint32_t prob = right - left;
            const uint64_t t = (uint64_t)__funnelshift_r(lo, hi, kPrecision) * prob + (uint64_t)(q - left);
            const uint32_t nhi = (uint32_t)(t >> 32) + (hi >> kPrecision) * prob, nlo = (uint32_t)t;
            const bool pop = act && nhi == 0u && avail != 0u;  // stack.rs:1091-1097
            hi = act ? (pop ? nlo : nhi) : hi;
            lo = act ? (pop ? nxt : nlo) : lo;
            pop_off -= pop ? 4u : 0u;
            avail -= pop ? 1u : 0u;
        }
        nxt = peek_word();
        return min_symbol + s;
    };

    for (uint64_t t = 0; t < rounds; ++t) {
        const uint32_t slot = (uint32_t)(t % kDecChainTiles);
        const uint64_t use = t / kDecChainTiles;
        if (use > 0) mbar_wait_addr(empty0 + 8u * slot, (uint32_t)((use - 1) & 1));
        const uint64_t done = t * 32;
        const uint32_t c = n_k > done ? (uint32_t)min((uint64_t)32, n_k - done) : 0u;
        const uint32_t base = tiles_addr + slot * 4096u;
        const uint32_t cmax = __reduce_max_sync(kFullMask, c), cmin = __reduce_min_sync(kFullMask, c);
        if (cmin == 32u) {  // the common case: nothing in the loop is predicated on the stream's length
            for (uint32_t j = 0; j < 32u; j += 4) {
                top_up();
                nxt = peek_word();
#pragma unroll
                for (uint32_t u = 0; u < 4; ++u) {
                    const uint32_t jj = j + u;
                    const uint32_t sym = decode_one(true, std::true_type{});
                    sts_u32(base + (jj * 32u + ((uint32_t)lane ^ jj)) * 4u, sym);
                }
            }
        } else {
            for (uint32_t j = 0; j < cmax; j += 4) {
                top_up();
                nxt = peek_word();
#pragma unroll
                for (uint32_t u = 0; u < 4; ++u) {
                    const uint32_t jj = j + u;
                    const uint32_t sym = decode_one(jj < c, std::false_type{});
                    sts_u32(base + (jj * 32u + ((uint32_t)lane ^ jj)) * 4u, sym);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_addr(full0 + 8u * slot);
    }
    cp_async_wait_all();
    if (valid) {
        if (p.states_out) {
            if (RANGE) {
                p.states_out[4 * k] = st.lower;
                p.states_out[4 * k + 1] = st.range;
                p.states_out[4 * k + 2] = st.point;
                p.states_out[4 * k + 3] = 0;
            } else {
                p.states_out[k] = ((uint64_t)hi << 32) | lo;
            }
        }
        if (p.words_left) {
            const uint32_t left_n = unstaged + pending + pending_old + avail;
            p.words_left[k] = RANGE ? (uint64_t)((uint32_t)(end - begin) - left_n) : (uint64_t)left_n;
        }
        if (trailing_zero) report_error(p.status, kErrTrailingZero, k);
        if (invalid_data) report_error(p.status, kErrInvalidData, k);
    }
}

}  // namespace ctr
