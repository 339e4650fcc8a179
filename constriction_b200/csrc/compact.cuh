// compact.cuh -- K6: variable-length compaction.  The encode kernels leave stream k's words in a
// worst-case-sized scratch region; here the per-stream lengths are prefix-summed into the container's
// `offsets` (u64[K+1]) and the words are gathered into one dense buffer, so that
// words[offsets[k] .. offsets[k+1]) is stream k's `get_compressed()` (stack.rs:537-547,
// queue.rs:349-355).
#pragma once
#include "device_utils.cuh"

namespace ctr {

constexpr int kScanBlock = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanBlock * kScanItems;  // lengths per CTA

__device__ __forceinline__ uint64_t warp_inclusive_scan(uint64_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint64_t o = shfl_u64(v, lane >= d ? lane - d : lane);
        if (lane >= d) v += o;
    }
    return v;
}

// exclusive scan of one value per thread across a CTA of kScanBlock threads; returns the CTA total
__device__ __forceinline__ uint64_t block_exclusive_scan(uint64_t v, uint64_t &total) {
    __shared__ uint64_t warp_sums[kScanBlock / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t inc = warp_inclusive_scan(v, lane);
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    uint64_t base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kScanBlock / 32; ++w) {
        const uint64_t s = warp_sums[w];
        if (w < warp) base += s;
        tot += s;
    }
    __syncthreads();
    total = tot;
    return base + inc - v;
}

// pass 1: sum of each tile of lengths
__global__ void __launch_bounds__(kScanBlock) scan_tile_sums_kernel(const uint32_t *lengths, uint64_t K,
                                                                    uint64_t *tile_sums) {
    const uint64_t first = (uint64_t)blockIdx.x * kScanTile;
    uint64_t v = 0;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
        const uint64_t i = first + (uint64_t)j * kScanBlock + threadIdx.x;
        if (i < K) v += lengths[i];
    }
    uint64_t total;
    block_exclusive_scan(v, total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// pass 2 (one CTA): exclusive scan of the tile sums in place; total -> offsets[K]
__global__ void __launch_bounds__(kScanBlock) scan_top_kernel(uint64_t *tile_sums, uint64_t n_tiles, uint64_t *offsets,
                                                              uint64_t K) {
    uint64_t carry = 0;
    for (uint64_t first = 0; first < n_tiles; first += kScanBlock) {
        const uint64_t i = first + threadIdx.x;
        const uint64_t v = i < n_tiles ? tile_sums[i] : 0;
        uint64_t total;
        const uint64_t ex = block_exclusive_scan(v, total);
        if (i < n_tiles) tile_sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) offsets[K] = carry;
}

// pass 3: offsets[k] for every stream.  Thread t of a tile owns the kScanItems consecutive lengths
// starting at first + t*kScanItems.
__global__ void __launch_bounds__(kScanBlock) scan_apply_kernel(const uint32_t *lengths, uint64_t K,
                                                                const uint64_t *tile_offsets, uint64_t *offsets) {
    const uint64_t first = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
    uint32_t local[kScanItems];
    uint64_t v = 0;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
        const uint64_t i = first + j;
        local[j] = i < K ? lengths[i] : 0u;
        v += local[j];
    }
    uint64_t total;
    uint64_t run = tile_offsets[blockIdx.x] + block_exclusive_scan(v, total);
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
        const uint64_t i = first + j;
        if (i < K) offsets[i] = run;
        run += local[j];
    }
}

// gather: one warp per stream copies scratch region -> dense words
__global__ void __launch_bounds__(256) compact_copy_kernel(const uint32_t *scratch, const uint32_t *lengths,
                                                           const uint64_t *offsets, uint64_t K, uint64_t N,
                                                           const uint64_t *sym_off, uint32_t *words,
                                                           uint64_t capacity, uint32_t *status) {
    const uint64_t k = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (k >= K) return;
    const uint64_t o_k = sym_off ? sym_off[k] : interleaved_start(N, K, k);
    const uint32_t *src = scratch + scratch_start(o_k, k);
    const uint64_t dst0 = offsets[k];
    const uint32_t len = lengths[k];
    if (dst0 + len > capacity) {
        if (lane == 0) report_error(status, kErrOutOfSpace, k);
        return;
    }
    uint32_t *dst = words + dst0;
    for (uint32_t i = lane; i < len; i += 32) dst[i] = src[i];
}

}  // namespace ctr
