// compact.cuh -- K6: variable-length compaction, fused into the tail of the encode kernels.
//
// While a stream is being encoded its final length is unknown, so its words first go to a
// worst-case-sized scratch region.  When a CTA has finished its 256 streams it
//   1. prefix-sums their lengths inside the CTA,
//   2. obtains the total length of all preceding streams with a single-pass "decoupled look-back"
//      over per-CTA status words (aggregate published first, inclusive prefix as soon as it is known;
//      the whole CTA looks back, one predecessor per thread and round),
//   3. writes the container's `offsets` (u64[K+1]) and gathers its streams' words from scratch (still
//      L2-resident: this CTA wrote them microseconds ago) into the dense `words` buffer, so that
//      words[offsets[k] .. offsets[k+1]) is stream k's `get_compressed()` (stack.rs:537-547,
//      queue.rs:349-355).
// There is no separate scan / gather launch and no second pass over HBM-cold data.
//
// CTAs take their tile (block of 256 consecutive streams) from an atomic ticket at kernel start, so a
// CTA only ever waits for tiles whose CTAs started earlier and are therefore resident and running:
// the look-back cannot deadlock whatever the grid size or dispatch order.
#pragma once
#include "device_utils.cuh"

namespace ctr {

constexpr uint64_t kTileInvalid = 0ull;          // status not yet published
constexpr uint64_t kTileAggregate = 1ull << 62;  // value = total of this tile only
constexpr uint64_t kTilePrefix = 2ull << 62;     // value = total of this tile and all before it
constexpr uint64_t kTileValueMask = (1ull << 62) - 1;

struct CompactParams {
    uint64_t *tile_status;    // u64[n_tiles], zeroed before the launch
    unsigned int *ticket;     // zeroed before the launch
    uint32_t *words_out;      // dense container (u16 words for the Small preset: see compact_tail's W)
    uint64_t words_capacity;  // capacity of words_out in words
    uint64_t *offsets_out;    // u64[K+1]
};

__device__ __forceinline__ uint64_t ld_volatile_u64(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u64(uint64_t *p, uint64_t v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_cg_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// The tile this CTA works on (all threads get the same value).
__device__ __forceinline__ uint32_t take_tile_ticket(unsigned int *ticket) {
    __shared__ uint32_t s_tile;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    return s_tile;
}

__device__ __forceinline__ uint64_t warp_inclusive_scan_u64(uint64_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint64_t o = shfl_u64(v, lane >= d ? lane - d : lane);
        if (lane >= d) v += o;
    }
    return v;
}

__device__ __forceinline__ uint64_t warp_sum_u64(uint64_t v, int lane) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += shfl_u64(v, lane ^ d);
    return v;
}

// Called by every thread of the CTA once its stream is complete in scratch.
//   tile    : this CTA's tile index (stream k = tile * blockDim.x + threadIdx.x)
//   src/len : my stream's words in scratch (len = 0 for threads without a stream)
//   W       : word type of the container (uint32_t: Default preset, uint16_t: Small preset)
template <int BLOCK, typename W = uint32_t>
__device__ __forceinline__ void compact_tail(const CompactParams &c, uint32_t tile, uint64_t k, uint64_t K, bool valid,
                                             const W *src, uint32_t len, uint32_t *status) {
    constexpr int kWarps = BLOCK / 32;
    __shared__ uint64_t s_warp_totals[kWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // 1. offsets inside the CTA
    const uint64_t inc = warp_inclusive_scan_u64(len, lane);
    if (lane == 31) s_warp_totals[warp] = inc;
    __syncthreads();
    uint64_t before = 0, tile_total = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
        const uint64_t t = s_warp_totals[w];
        if (w < warp) before += t;
        tile_total += t;
    }

    // 2. decoupled look-back by the whole CTA: in round r thread j inspects tile (tile - 1 - j - BLOCK * r), so
    //    the statuses of BLOCK predecessors are fetched with one L2 round trip.  The nearest predecessor that
    //    has published an inclusive prefix ends the walk; nearer ones contribute their aggregates.
    __shared__ uint64_t s_round_sum[kWarps];
    __shared__ uint32_t s_round_has[kWarps];
    if (threadIdx.x == 0) st_volatile_u64(c.tile_status + tile, (tile == 0 ? kTilePrefix : kTileAggregate) | tile_total);
    uint64_t exclusive = 0;
    {
        int64_t idx = (int64_t)tile - 1 - (int64_t)threadIdx.x;
        bool done = tile == 0;
        while (!done) {
            uint64_t v = kTilePrefix;  // tiles before the first one contribute a prefix of 0
            if (idx >= 0) {
                do {
                    v = ld_volatile_u64(c.tile_status + idx);
                } while ((v >> 62) == 0);
            }
            const unsigned has_prefix = __ballot_sync(kFullMask, (v >> 62) == 2);
            const int stop = has_prefix ? __ffs(has_prefix) - 1 : 31;
            const uint64_t wsum = warp_sum_u64(lane <= stop ? (v & kTileValueMask) : 0ull, lane);
            if (lane == 0) {
                s_round_sum[warp] = wsum;
                s_round_has[warp] = has_prefix;
            }
            __syncthreads();
#pragma unroll
            for (int w = 0; w < kWarps; ++w) {
                if (!done) {
                    exclusive += s_round_sum[w];
                    done = s_round_has[w] != 0u;
                }
            }
            __syncthreads();
            idx -= BLOCK;
        }
    }
    if (threadIdx.x == 0 && tile != 0) st_volatile_u64(c.tile_status + tile, kTilePrefix | (exclusive + tile_total));
    const uint64_t my_off = exclusive + before + inc - len;

    // 3. offsets + gather
    if (valid) {
        c.offsets_out[k] = my_off;
        if (k + 1 == K) c.offsets_out[K] = my_off + len;
    }
    const bool fits = my_off + len <= c.words_capacity;
    if (valid && !fits) report_error(status, kErrOutOfSpace, k);
    const uint32_t n = (valid && fits) ? len : 0u;
    const uint32_t n_max = __reduce_max_sync(kFullMask, n);
    if (n_max == 0) return;
#ifdef CTR_DBG_NO_GATHER
    if (n_max != 0xffffffffu) return;
#endif
    W *dst = reinterpret_cast<W *>(c.words_out) + my_off;
    __syncwarp();
    // The copy is latency-bound (one L2 round trip per batch of loads), so the loads of kGroup streams are
    // put in flight together: lane l moves words l, l+32, ... of each stream; kSlots chunks per stream cover
    // streams of up to 32*kSlots words in one pass (longer ones loop).
#ifndef CTR_GATHER_GROUP
#define CTR_GATHER_GROUP 4
#endif
    constexpr int kGroup = CTR_GATHER_GROUP, kSlots = 4;
    for (int i0 = 0; i0 < 32; i0 += kGroup) {
        uint32_t ni[kGroup];
        const W *si[kGroup];
        W *di[kGroup];
        uint32_t group_max = 0;
#pragma unroll
        for (int g = 0; g < kGroup; ++g) {
            ni[g] = __shfl_sync(kFullMask, n, i0 + g);
            si[g] = (const W *)shfl_u64((uint64_t)src, i0 + g);
            di[g] = (W *)shfl_u64((uint64_t)dst, i0 + g);
            group_max = max(group_max, ni[g]);
        }
        for (uint32_t j0 = 0; j0 < group_max; j0 += 32 * kSlots) {
            W v[kGroup][kSlots];
#pragma unroll
            for (int g = 0; g < kGroup; ++g)
#pragma unroll
                for (int u = 0; u < kSlots; ++u) {
                    const uint32_t j = j0 + u * 32 + lane;
                    if (j < ni[g]) v[g][u] = __ldcg(si[g] + j);
                }
#pragma unroll
            for (int g = 0; g < kGroup; ++g)
#pragma unroll
                for (int u = 0; u < kSlots; ++u) {
                    const uint32_t j = j0 + u * 32 + lane;
                    if (j < ni[g]) di[g][j] = v[g][u];
                }
#ifdef CTR_DISCARD_SCRATCH  // measured on B200: DRAM traffic -1 %, encode kernel +1 % (the lines are mostly written back before the tail): off
            // The scratch words just read are dead.  Their cache lines are still dirty in L2 (this kernel wrote them
            // moments ago): drop them instead of letting the L2 write them back to HBM later -- a round trip of the
            // whole compressed size that nobody would ever read.  (Scratch regions start on 128-byte boundaries and
            // belong to one stream each; a line is discarded only if every byte of it has been gathered.)
            if (sizeof(W) == 4) {
#pragma unroll
                for (int g = 0; g < kGroup; ++g)
#pragma unroll
                    for (int u = 0; u < kSlots; ++u) {
                        const uint32_t j = j0 + u * 32;  // words [j, j + 32) of stream g: one 128-byte line
                        if (lane == 0 && j + 32 <= ni[g]) asm volatile("discard.global.L2 [%0], 128;" ::"l"(si[g] + j) : "memory");
                    }
            }
#endif
        }
    }
}

}  // namespace ctr
