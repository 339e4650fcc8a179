// ans_kernels.cuh -- batched rANS encode / decode kernels (K1 / K2), one lane per independent coder.
//
// Execution model.  A batch holds K independent coders ("streams").  Lane l of warp w owns stream
// k = 32 w + l for the whole kernel and keeps that coder's 64-bit state in registers, so the
// loop-carried dependency of the reference's per-symbol loop (stream/mod.rs:592-607,1274-1297) is
// private to a lane and 32 such chains advance per warp instruction.  Everything that touches HBM is
// warp-cooperative and coalesced:
//   - symbols: interleaved layout -> one 128-byte row per warp step (pointer bumped by K per step);
//     contiguous layout -> 32x32 tiles transposed through shared memory;
//   - compressed words: each lane appends to / pops from a private shared-memory row; rows are moved
//     to / from HBM 32 words (128 bytes) at a time by the whole warp;
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
//   - word rows are inspected once per kCheckEvery symbols, and the flush / refill is a cold path.
//
// Per-stream results equal the reference's `AnsCoder` (src/stream/stack.rs:1014-1100) word for word.
#pragma once
#include <type_traits>

#include "compact.cuh"
#include "device_utils.cuh"

namespace ctr {

constexpr int kAnsBlock = 256;  // threads per CTA (8 warps)
constexpr uint32_t kMaxSharedAlphabet = 4095;      // bigger alphabets use the global-table path

struct ModelView {
    const uint32_t *cdf;   // [n_models][alphabet + 1]
    const uint4 *enc;      // [n_models][alphabet + 1] {left, prob, reciprocal lo, hi}; entry [alphabet] is
                           // the all-zero sentinel that out-of-range symbols are clamped to
    const uint32_t *dec;   // model 0 only: quantile index uint2[kLutSize] ++ cdf u32[alphabet + 2] (padded to 16 B)
    uint32_t n_models;
    uint32_t alphabet;
    int32_t min_symbol;
    uint32_t dec_cdf_bytes;    // size of the cdf part of `dec`: (alphabet + 2) * 4 rounded up to 16
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
    uint32_t *scratch;      // strided per-stream regions
    CompactParams compact;  // fused compaction into the dense container
    // decode
    const uint32_t *words;
    const uint64_t *offsets;
    uint64_t *words_left;
};

// ---- warp-cooperative tile I/O for the contiguous layout and the range kernels (generic pointers) ----

// Write the first count_i words of row i to dst_i, for every lane i in `mask` (rows of kRowStride words).
__device__ __noinline__ void warp_flush_rows(unsigned mask, const uint32_t *rows, uint32_t *dst, uint32_t count,
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
__device__ __noinline__ void warp_fill_rows(unsigned mask, uint32_t *rows, const uint32_t *src, uint32_t count,
                                            int lane) {
    __syncwarp();
    while (mask) {
        const int i = __ffs(mask) - 1;
        mask &= mask - 1;
        const uint32_t *s = (const uint32_t *)shfl_u64((uint64_t)src, i);
        const uint32_t c = __shfl_sync(kFullMask, count, i);
        if ((uint32_t)lane < c) rows[i * kRowStride + lane] = ld_stream_u32(s + lane);
    }
    __syncwarp();
}

// ---- word rows of the ANS kernels (32-bit shared addresses, rows of kWordRowStride words) -------------

// Per-thread "cold slot" (16 bytes of shared memory per lane): everything a lane needs only when one of its
// rows is moved to / from HBM lives here, not in registers, so that the hot loops fit 5 CTAs per SM.
//   encoder: {scratch cursor lo, hi, words of scratch capacity left, words flushed so far | overflow bit}
//   decoder: {cursor lo, hi (one past the highest word not yet staged), words not yet staged, unused}
constexpr uint32_t kColdSlotBytes = 16;
constexpr uint32_t kColdOverflowBit = 0x80000000u;

__device__ __forceinline__ void cold_store(uint32_t slot, const void *ptr, uint32_t a, uint32_t b) {
    const uint64_t v = (uint64_t)ptr;
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(slot), "r"((uint32_t)v), "r"((uint32_t)(v >> 32)), "r"(a),
                 "r"(b)
                 : "memory");
}
__device__ __forceinline__ uint4 cold_load(uint32_t slot) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(slot) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t *cold_ptr(const uint4 &v) { return (uint32_t *)(((uint64_t)v.y << 32) | v.x); }

// Encoder, cold: every lane in `mask` has >= 32 words in its row.  The warp writes the first 32 words of
// each such row to that lane's scratch cursor (one 128-byte store), the lane keeps what is left over.
// Returns the lane's new row cursor.
__device__ __noinline__ uint32_t ans_flush_rows_cold(unsigned mask, uint32_t rows_addr, uint32_t slots_addr,
                                                     uint32_t row_addr, uint32_t wptr, int lane) {
    __syncwarp();
    unsigned todo = mask;
    while (todo) {
        const int i = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint4 c = cold_load(slots_addr + (uint32_t)i * kColdSlotBytes);  // broadcast read of lane i's slot
        if (c.z >= (uint32_t)kRowWords)
            st_stream_u32(cold_ptr(c) + lane, lds_u32(rows_addr + (uint32_t)(i * kWordRowStride + lane) * 4u));
    }
    __syncwarp();
    if ((mask >> lane) & 1u) {
        const uint32_t slot = slots_addr + (uint32_t)lane * kColdSlotBytes;
        const uint4 c = cold_load(slot);
        if (c.z >= (uint32_t)kRowWords)
            cold_store(slot, cold_ptr(c) + kRowWords, c.z - kRowWords, c.w + kRowWords);
        else
            cold_store(slot, cold_ptr(c), c.z, c.w | kColdOverflowBit);  // words dropped, stream flagged
        const uint32_t left = (wptr - row_addr) / 4u - kRowWords;  // 0 .. kCheckEvery-1
        for (uint32_t j = 0; j < left; ++j) sts_u32(row_addr + j * 4u, lds_u32(row_addr + (kRowWords + j) * 4u));
        wptr = row_addr + left * 4u;
    }
    __syncwarp();
    return wptr;
}

// Encoder, end of stream: write the remaining cnt_i (< 32) words of every row.
__device__ __noinline__ void ans_flush_tail_cold(uint32_t rows_addr, uint32_t slots_addr, uint32_t cnt, int lane) {
    __syncwarp();
    unsigned todo = __ballot_sync(kFullMask, cnt > 0);
    while (todo) {
        const int i = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint4 c = cold_load(slots_addr + (uint32_t)i * kColdSlotBytes);
        const uint32_t ci = __shfl_sync(kFullMask, cnt, i);
        if (c.z >= ci && (uint32_t)lane < ci)
            st_stream_u32(cold_ptr(c) + lane, lds_u32(rows_addr + (uint32_t)(i * kWordRowStride + lane) * 4u));
    }
    __syncwarp();
    if (cnt > 0) {
        const uint32_t slot = slots_addr + (uint32_t)lane * kColdSlotBytes;
        const uint4 c = cold_load(slot);
        if (c.z >= cnt)
            cold_store(slot, cold_ptr(c) + cnt, c.z - cnt, c.w + cnt);
        else
            cold_store(slot, cold_ptr(c), c.z, c.w | kColdOverflowBit);
    }
    __syncwarp();
}

// Decoder, cold: every lane in `mask` is down to < kCheckEvery staged words and has more in HBM.  The warp
// loads the next (up to) 32 words below each such lane's cursor; chunks end on 128-byte boundaries of the
// global address space, so every refill after a stream's first is one aligned line.
// Returns {new row cursor, new refill mark (0 when nothing is left in HBM)}.
struct RefillResult {
    uint32_t rptr, mark;
};
__device__ __noinline__ RefillResult ans_refill_rows_cold(unsigned mask, uint32_t rows_addr, uint32_t slots_addr,
                                                          uint32_t row_addr, uint32_t rptr, uint32_t mark, int lane) {
    const bool mine = (mask >> lane) & 1u;
    // keep my unread words (they are older than the chunk that is about to arrive, so they go on top)
    const uint32_t left = mine ? (rptr - row_addr) / 4u : 0u;  // 0 .. kCheckEvery-1
    uint32_t keep[kCheckEvery - 1];
#pragma unroll
    for (int j = 0; j < kCheckEvery - 1; ++j) keep[j] = (uint32_t)j < left ? lds_u32(row_addr + j * 4u) : 0u;
    __syncwarp();
    unsigned todo = mask;
    uint32_t my_c = 0;
    while (todo) {
        const int i = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint4 c = cold_load(slots_addr + (uint32_t)i * kColdSlotBytes);
        const uint32_t *top = cold_ptr(c);
        // chunk = [max(stream begin, 128-byte line of the top word), top)
        uint32_t ci = (uint32_t)((((uint64_t)(top - 1)) & 127u) / 4u) + 1u;
        ci = ci < c.z ? ci : c.z;
        if ((uint32_t)lane < ci)
            sts_u32(rows_addr + (uint32_t)(i * kWordRowStride + lane) * 4u, ld_stream_u32(top - ci + lane));
        if (i == lane) my_c = ci;
    }
    __syncwarp();
    RefillResult r;
    r.rptr = rptr;
    r.mark = mark;
    if (mine) {
        const uint32_t slot = slots_addr + (uint32_t)lane * kColdSlotBytes;
        const uint4 c = cold_load(slot);
        const uint32_t *top = cold_ptr(c) - my_c;
        cold_store(slot, top, c.z - my_c, 0u);
        if (c.z - my_c != 0u) prefetch_l2(top - 1);  // the line below: needed ~200 symbols from now
#pragma unroll
        for (int j = 0; j < kCheckEvery - 1; ++j)
            if ((uint32_t)j < left) sts_u32(row_addr + (my_c + j) * 4u, keep[j]);
        r.rptr = row_addr + (my_c + left) * 4u;
        r.mark = (c.z - my_c != 0u) ? row_addr + kCheckEvery * 4u : 0u;
    }
    __syncwarp();
    return r;
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
__device__ __noinline__ uint32_t lookup_far_cold(uint32_t cdf_addr, uint32_t alphabet, uint32_t s, uint32_t q) {
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
template <bool SMALL>
__device__ __forceinline__ uint32_t lookup_shared(uint32_t lut_addr, uint32_t cdf_addr, uint32_t alphabet, uint32_t word,
                                                  uint32_t q, uint32_t &left, uint32_t &right) {
    const uint2 e = lds_table_v2(lut_addr + ((word >> (kLutShift - 3)) & ((kLutSize - 1) << 3)));
    uint32_t s = e.x >> 24;
    left = e.x & kQuantileMask;
    right = e.y;
    if (!SMALL) {
        s |= (e.y >> 25) << 8;
        right = e.y & 0x1ffffffu;
    }
    const bool beyond = q >= right;  // the bucket straddles the boundary and q lies past it
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
    return s;
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
// per-warp staging block in shared memory: word rows, then the 32 cold slots
constexpr int kWarpStageWords = kWordRowsWords + kWarp * (int)(kColdSlotBytes / 4);

//   SHARED : model 0's encoder table lives in shared memory (index_mode == NONE, small alphabet)
//   CONTIG : stream k owns symbols[sym_off[k] .. sym_off[k+1]) (else interleaved deal)
//   PERSYM : a model index per symbol (else one model per stream / model 0)
//   F64DIV : the table holds double-precision reciprocals and the quotient estimate uses the FP64 pipe
template <bool SHARED, bool CONTIG, bool PERSYM, bool F64DIV>
__global__ void __launch_bounds__(kAnsBlock, 5) ans_encode_kernel(const AnsParams p) {
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ uint64_t bar;

    const int lane = threadIdx.x & 31;
    const int warp_in_cta = threadIdx.x >> 5;
    constexpr int kWarpsPerCta = kAnsBlock / 32;

    // shared memory carve-up: [table][per warp: word rows + cold slots][symbol tiles][index tiles]
    const uint32_t alphabet = p.model.alphabet;
    const uint32_t table_words = SHARED ? (alphabet + 1) * 4 : 0;
    const uint32_t table_addr = smem_u32_pinned(smem);
    const uint32_t rows_addr = smem_u32_pinned(smem + table_words + warp_in_cta * kWarpStageWords);
    const uint32_t slots_addr = rows_addr + kWordRowsWords * 4u;
    uint32_t *sym_tile = smem + table_words + kWarpsPerCta * kWarpStageWords + warp_in_cta * kTileWords;
    uint32_t *idx_tile = sym_tile + kWarpsPerCta * kTileWords;

    if (SHARED) stage_table(smem, p.model.enc, (alphabet + 1) * 16u, &bar);

    const uint64_t K = p.K, N = p.N;
    const uint32_t tile = take_tile_ticket(p.compact.ticket);  // which 256 streams this CTA codes
    const uint64_t k = (uint64_t)tile * kAnsBlock + threadIdx.x;
    const bool valid = k < K;
    const uint64_t kc = valid ? k : K - 1;  // lanes without a stream shadow the last one (loads only)

    // stream geometry; the scratch cursor and capacity go to my cold slot
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
    const uint32_t my_slot = slots_addr + (uint32_t)lane * kColdSlotBytes;
    {
        const uint64_t begin = scratch_start(o_k, k);
        const uint64_t room = valid ? scratch_start(o_k + n_k, k + 1) - begin : 0;
        cold_store(my_slot, p.scratch + begin, room > 0x7fffffffu ? 0x7fffffffu : (uint32_t)room, 0u);
    }

    uint64_t state = (valid && p.states_in) ? p.states_in[k] : 0;
    const uint32_t row_addr = rows_addr + (uint32_t)lane * (kWordRowStride * 4u);
    uint32_t wptr = row_addr;           // shared address of the next free slot of my row
    uint32_t min_prob = 0xffffffffu;    // running minimum of the probabilities used (0 <=> impossible symbol)
    const uint32_t push_shift = valid ? 8u : 32u;  // (state >> 32) >> 32 == 0: lanes without a stream never push
    const uint32_t stream_model = (p.index_mode == 2) ? p.model_index[kc] : 0u;
    const uint32_t n_models = p.model.n_models;
    const uint32_t min_symbol = (uint32_t)p.model.min_symbol;

    // one reference encode_symbol (stack.rs:1014-1048)
    auto encode_one = [&](int32_t sym, uint32_t m) {
        uint32_t idx = min((uint32_t)sym - min_symbol, alphabet);  // out of range -> sentinel entry
        uint4 e;
        if (SHARED) {
            e = lds_table_v4(table_addr + idx * 16u);
        } else {
            const bool ok = m < n_models;
            idx = ok ? idx : alphabet;
            m = ok ? m : 0u;
            e = __ldg(p.model.enc + (uint64_t)m * (alphabet + 1) + idx);
        }
        min_prob = min(min_prob, e.y);
        uint32_t lo = (uint32_t)state, hi = (uint32_t)(state >> 32);
        if (shr_clamp(hi, push_shift) >= e.y) {  // stack.rs:1035-1040
            sts_u32(wptr, lo);
            wptr += 4u;
            lo = hi;
            hi = 0u;
        }
        const uint64_t n = ((uint64_t)hi << 32) | lo;
        uint64_t q;
        if (F64DIV) {
            q = __double2ull_rz(__ull2double_rz(n) * __hiloint2double((int)e.w, (int)e.z));
        } else {
            q = __umul64hi(n, ((uint64_t)e.w << 32) | e.z);
        }
        uint32_t r = lo - (uint32_t)q * e.y;  // remainder of the estimate, in [0, 2 prob)
        const bool fix = r >= e.y;
        r -= fix ? e.y : 0u;
        q += fix ? 1u : 0u;
        state = (q << kPrecision) + (uint64_t)(e.x + r);  // stack.rs:1042-1045 (left + r < 2^24)
    };

    // every kCheckEvery symbols: move rows that reached 32 words to HBM (cold)
    auto check_rows = [&]() {
        const unsigned mask = __ballot_sync(kFullMask, wptr >= row_addr + kRowWords * 4u);
        if (mask) wptr = ans_flush_rows_cold(mask, rows_addr, slots_addr, row_addr, wptr, lane);
    };

    if (!CONTIG) {
        // ---- interleaved deal: row t holds symbols[t*K .. t*K+K); coded from the last row backwards ---
        const Interleave g = interleave_of(N, K);
        if (g.T > 0) {  // ragged last row
            if (valid && k < g.last) {
                const uint64_t i = (g.T - 1) * K + k;
                encode_one(ld_stream_s32(p.symbols_in + i), PERSYM ? ld_stream_u32(p.model_index + i) : stream_model);
            }
        }
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
                    buf[which][u] = ld_stream_s32(reinterpret_cast<const int32_t *>(ps));
                    ps -= row_bytes;
                    if (PERSYM) {
                        mbuf[which][u] = ld_stream_u32(reinterpret_cast<const uint32_t *>(pm));
                        pm -= row_bytes;
                    } else {
                        mbuf[which][u] = stream_model;
                    }
                }
            };
            // The row check (a potential call into the cold path) comes first, then the loads of the next
            // batch are issued, then this batch is coded: no load is in flight across the call site, so
            // the register shuffling around it never waits for memory.
            auto code_batch = [&](int which, bool load_next) {
                check_rows();  // room for kCheckEvery more words in every row
                if (load_next) load_batch(which ^ 1);
#pragma unroll
                for (int u = 0; u < kCheckEvery; ++u) encode_one(buf[which][u], mbuf[which][u]);
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
            check_rows();
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
        // ---- contiguous: 32x32 tiles, transposed through shared memory ---------------------------
        uint64_t remaining = n_k;  // symbols of my stream not yet loaded (I consume from the end)
        const uint64_t rounds = (warp_max_u64(n_k, lane) + 31) / 32;
        for (uint64_t r = 0; r < rounds; ++r) {
            const uint32_t c = remaining < 32 ? (uint32_t)remaining : 32u;
            remaining -= c;
            const unsigned have = __ballot_sync(kFullMask, c > 0);
            warp_fill_rows(have, sym_tile, reinterpret_cast<const uint32_t *>(p.symbols_in + o_k + remaining), c, lane);
            if (PERSYM) warp_fill_rows(have, idx_tile, p.model_index + o_k + remaining, c, lane);
            const uint32_t cmax = __reduce_max_sync(kFullMask, c);
            for (uint32_t s = 0; s < cmax; ++s) {
                if ((s & (kCheckEvery - 1)) == 0) check_rows();
                if (s < c) {
                    const int32_t sym = (int32_t)sym_tile[lane * kRowStride + (c - 1 - s)];
                    const uint32_t m = PERSYM ? idx_tile[lane * kRowStride + (c - 1 - s)] : stream_model;
                    encode_one(sym, m);
                }
            }
        }
    }

    // ---- finalize: state words (lib.rs:719-730, low word first), remaining partial rows --------------
    check_rows();
    const bool raw = (p.flags & 1u) != 0;
    const uint32_t n_state = (valid && !raw) ? ans_state_words(state) : 0u;
    if (n_state >= 1) {
        sts_u32(wptr, (uint32_t)state);
        wptr += 4u;
    }
    if (n_state == 2) {
        sts_u32(wptr, (uint32_t)(state >> 32));
        wptr += 4u;
    }
    check_rows();
    ans_flush_tail_cold(rows_addr, slots_addr, (wptr - row_addr) / 4u, lane);  // < 32 words each
    const uint4 fin = cold_load(my_slot);
    const uint32_t len = fin.w & ~kColdOverflowBit;
    if (valid) {
        if (p.states_out) p.states_out[k] = state;
        if (min_prob == 0u) report_error(p.status, kErrImpossibleSymbol, k);
        if (fin.w & kColdOverflowBit) report_error(p.status, kErrOutOfSpace, k);
    }
    // ---- K6: place my stream in the dense container ---------------------------------------------------
    compact_tail<kAnsBlock>(p.compact, tile, k, K, valid, cold_ptr(fin) - len, valid ? len : 0u, p.status);
}

// =====================================================================================================
// decode
// =====================================================================================================
//   SMALL : (SHARED only) alphabet <= 256
// With a shared model and the interleaved layout the CTA is 1024 threads so that the 32 KB quantile index is
// staged once per SM; otherwise (transposition tiles, or global tables: nothing to amortise) 256 threads.
constexpr int kDecBlockShared = 1024;
template <bool SHARED, bool CONTIG, bool PERSYM, bool SMALL>
__global__ void __launch_bounds__((SHARED && !CONTIG) ? kDecBlockShared : kAnsBlock, (SHARED && !CONTIG) ? 1 : 2)
    ans_decode_kernel(const AnsParams p) {
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ uint64_t bar;

    constexpr int kBlock = (SHARED && !CONTIG) ? kDecBlockShared : kAnsBlock;
    const int lane = threadIdx.x & 31;
    const int warp_in_cta = threadIdx.x >> 5;
    constexpr int kWarpsPerCta = kBlock / 32;

    const uint32_t alphabet = p.model.alphabet;
    const uint32_t table_words = SHARED ? (kLutBytes + p.model.dec_cdf_bytes) / 4 : 0;
    const uint32_t lut_addr = smem_u32_pinned(smem);
    uint32_t cdf_addr = lut_addr + kLutBytes;
    asm volatile("" : "+r"(cdf_addr));
    const uint32_t rows_addr = smem_u32_pinned(smem + table_words + warp_in_cta * kWarpStageWords);
    const uint32_t slots_addr = rows_addr + kWordRowsWords * 4u;
    uint32_t *sym_tile = smem + table_words + kWarpsPerCta * kWarpStageWords + warp_in_cta * kTileWords;
    uint32_t *idx_tile = sym_tile + kWarpsPerCta * kTileWords;

    if (SHARED) stage_table(smem, p.model.dec, kLutBytes + p.model.dec_cdf_bytes, &bar);

    const uint64_t K = p.K, N = p.N;
    const uint64_t k = (uint64_t)blockIdx.x * kBlock + threadIdx.x;
    const bool valid = k < K;
    const uint64_t kc = valid ? k : K - 1;
    const bool raw = (p.flags & 1u) != 0;

    uint64_t n_k = 0, o_k = 0;
    const uint32_t my_slot = slots_addr + (uint32_t)lane * kColdSlotBytes;
    const uint32_t row_addr = rows_addr + (uint32_t)lane * (kWordRowStride * 4u);
    uint32_t rptr = row_addr;  // my row holds the unread words [row_addr, rptr)
    uint32_t mark = 0;         // row_addr + 16 while words remain in HBM (refill when rptr drops below), else 0
    {
        uint64_t begin = 0, end = 0;
        if (valid) {
            if (CONTIG) {
                o_k = p.sym_off[k];
                n_k = p.sym_off[k + 1] - o_k;
            }
            begin = p.offsets[k];
            end = p.offsets[k + 1];
        }
        const uint64_t len = end - begin;
        // (a stream of >= 2^32 words does not exist: the encoder's lengths are 32-bit)
        cold_store(my_slot, p.words + end, (uint32_t)len, 0u);
        mark = len ? row_addr + kCheckEvery * 4u : 0u;
    }
    const uint32_t n_models = p.model.n_models;
    uint32_t min_symbol = (uint32_t)p.model.min_symbol;
    asm volatile("" : "+r"(min_symbol));
    const uint32_t stream_model = (p.index_mode == 2) ? p.model_index[kc] : 0u;

    // every kCheckEvery symbols: rows that are down to < kCheckEvery words get the next 32 (cold).  The first
    // chunk of a stream only reaches down to the next 128-byte boundary and may be shorter than
    // kCheckEvery words, hence the loop (it runs twice at most).
    auto check_rows = [&]() {
        unsigned mask = __ballot_sync(kFullMask, rptr < mark);
        while (mask) {
            const RefillResult u = ans_refill_rows_cold(mask, rows_addr, slots_addr, row_addr, rptr, mark, lane);
            rptr = u.rptr;
            mark = u.mark;
            mask = __ballot_sync(kFullMask, rptr < mark);
        }
    };

    // ---- initial state: stack.rs:299-318, 440-462 (from_compressed) or the caller's raw state ------
    uint32_t lo = 0, hi = 0;  // the coder state
    bool trailing_zero = false;
    check_rows();
    if (raw) {
        if (valid && p.states_in) {
            const uint64_t s = p.states_in[k];
            lo = (uint32_t)s;
            hi = (uint32_t)(s >> 32);
        }
    } else {
        if (rptr != row_addr) {
            rptr -= 4u;
            lo = lds_u32(rptr);
            trailing_zero = lo == 0u;
            if (rptr != row_addr && lo != 0u) {
                hi = lo;
                rptr -= 4u;
                lo = lds_u32(rptr);
            }
        }
        check_rows();
    }

    // one reference decode_symbol (stack.rs:1070-1100)
    auto decode_one = [&](uint32_t m) -> int32_t {
        const uint32_t q = lo & kQuantileMask;
        uint32_t left, right, s;
        if (SHARED) {
            s = lookup_shared<SMALL>(lut_addr, cdf_addr, alphabet, lo, q, left, right);
        } else {
            m = m < n_models ? m : n_models - 1;  // decoding cannot fail (stack.rs:1062-1065)
            s = lookup_global(p.model.cdf + (uint64_t)m * (alphabet + 1), alphabet, q, left, right);
        }
        // state = (state >> 24) * prob + (q - left), in 32-bit pieces (state >> 24 has 40 bits)
        const uint32_t prob = right - left;
        const uint64_t t = (uint64_t)__funnelshift_r(lo, hi, kPrecision) * prob + (uint64_t)(q - left);
        hi = (uint32_t)(t >> 32) + (hi >> kPrecision) * prob;
        lo = (uint32_t)t;
        if (hi == 0u && rptr != row_addr) {  // stack.rs:1091-1097
            hi = lo;
            rptr -= 4u;
            lo = lds_u32(rptr);
        }
        return (int32_t)(min_symbol + s);
    };

    if (!CONTIG) {
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
                        const int32_t sym = decode_one(mbuf[u]);
                        if (FULL || valid) st_stream_s32(reinterpret_cast<int32_t *>(po), sym);
                        po += row_bytes;
                    }
                    check_rows();
                }
                while (rows_left > 0) {  // at most kCheckEvery-1 more symbols
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
                check_rows();
            };
            if (__all_sync(kFullMask, valid))
                run_rows(std::true_type{});
            else
                run_rows(std::false_type{});
        }
        if (g.T > 0) {  // ragged last row
            if (valid && k < g.last) {
                const uint64_t i = (g.T - 1) * K + k;
                const int32_t sym = decode_one(PERSYM ? ld_stream_u32(p.model_index + i) : stream_model);
                st_stream_s32(p.symbols_out + i, sym);
            }
        }
    } else {
        uint64_t done = 0;  // symbols of my stream already produced
        const uint64_t rounds = (warp_max_u64(n_k, lane) + 31) / 32;
        for (uint64_t r = 0; r < rounds; ++r) {
            const uint64_t left_n = n_k - done;
            const uint32_t c = left_n < 32 ? (uint32_t)left_n : 32u;
            const unsigned have = __ballot_sync(kFullMask, c > 0);
            if (PERSYM) warp_fill_rows(have, idx_tile, p.model_index + o_k + done, c, lane);
            const uint32_t cmax = __reduce_max_sync(kFullMask, c);
            for (uint32_t s = 0; s < cmax; ++s) {
                if ((s & (kCheckEvery - 1)) == 0) check_rows();
                if (s < c) {
                    const uint32_t m = PERSYM ? idx_tile[lane * kRowStride + s] : stream_model;
                    sym_tile[lane * kRowStride + s] = (uint32_t)decode_one(m);
                }
            }
            warp_flush_rows(have, sym_tile, reinterpret_cast<uint32_t *>(p.symbols_out + o_k + done), c, lane);
            done += c;
        }
    }

    if (valid) {
        if (p.states_out) p.states_out[k] = ((uint64_t)hi << 32) | lo;
        if (p.words_left) p.words_left[k] = (uint64_t)cold_load(my_slot).z + (uint64_t)((rptr - row_addr) / 4u);
        if (trailing_zero) report_error(p.status, kErrTrailingZero, k);
    }
}

}  // namespace ctr
