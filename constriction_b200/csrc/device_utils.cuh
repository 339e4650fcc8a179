// device_utils.cuh -- warp/CTA plumbing shared by the coder kernels (sm_100a).
//
//   * 1-D bulk async copy (TMA engine, `cp.async.bulk`, SASS UBLKCP) of a model's tables from HBM
//     into shared memory, completion tracked by an mbarrier;
//   * per-lane shared-memory "rows" that turn each lane's private, variable-rate word stream into
//     128-byte coalesced global transactions (one warp-wide store/load per 32 words of one lane);
//   * 32x32 tile transposition for contiguous (one-stream-per-lane) symbol arrays;
//   * per-batch status reporting.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "coder_math.cuh"

namespace ctr {

constexpr int kWarp = 32;
constexpr unsigned kFullMask = 0xffffffffu;
// rows/tiles are [32 lanes][32 words], padded to a stride of 33 words so that both the "every lane
// touches its own row" and the "whole warp touches one row" access patterns are bank-conflict free.
constexpr int kRowWords = 32;
constexpr int kRowStride = 33;
constexpr int kTileWords = kWarp * kRowStride;  // 1056 words = 4224 B per warp
// The ANS kernels keep each lane's compressed words in a small lane-private ring in shared memory and move
// them to / from HBM 16 bytes at a time with lane-private vector accesses (no warp-cooperative phase): the
// ring is inspected once per kCheckEvery symbols, during which a lane moves at most kCheckEvery words.
constexpr int kCheckEvery = 4;
constexpr int kEncRingWords = 8;    // encoder: <= 3 leftover + 4 new words
constexpr int kDecRingWords = 16;   // decoder: <= 12 unread + 4 arriving words
constexpr uint32_t kEncRingBytes = kEncRingWords * 4u;
// the ANS encoder's TMA loop can drain its rings 32 bytes at a time (one 256-bit store per full sector, one check per
// box of 8 rows): the ring then holds 16 words
constexpr int kAnsEncRingWords = kEncRingWords;  // (the ANS encoder has no parking slots: see capi.cu)
constexpr uint32_t kDecRingBytes = kDecRingWords * 4u;

// ---- explicit shared-memory accesses by 32-bit shared address ------------------------------------------
// (generic pointers into shared memory cost 64-bit address arithmetic in the hot loops)
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
// logical right shift whose amount may be >= 32 (PTX clamps: the result is then 0; in C++ it would be UB)
__device__ __forceinline__ uint32_t shr_clamp(uint32_t v, uint32_t amount) {
    uint32_t r;
    asm("shr.u32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(amount));
    return r;
}
// table reads: the tables are immutable once staged, so these may be scheduled freely
__device__ __forceinline__ uint32_t lds_table_u32(uint32_t addr) {
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_table_u16(uint32_t addr) {
    uint16_t v;
    asm("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return (uint32_t)v;
}
__device__ __forceinline__ uint32_t lds_table_u8(uint32_t addr) {
    uint16_t v;
    asm("ld.shared.u8 %0, [%1];" : "=h"(v) : "r"(addr));
    return (uint32_t)v;
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void st_stream_v4(void *p, const uint4 &v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}
// 16-byte asynchronous copy global -> shared (LDGSTS), tracked by cp.async groups
__device__ __forceinline__ void cp_async_16(uint32_t dst_smem, const void *src_gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src_gmem) : "memory");
}
// 4-byte variant (any 4-byte aligned source): used to fill transposition tiles asynchronously
__device__ __forceinline__ void cp_async_4(uint32_t dst_smem, const void *src_gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_smem), "l"(src_gmem) : "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait_group() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ uint2 lds_table_v2(uint32_t addr) {
    uint2 v;
    asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 lds_table_v4(uint32_t addr) {
    uint4 v;
    asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

// decoder quantile index of a shared model: 2^kLutBits buckets of 2^kLutShift quantiles, 8 bytes each
constexpr int kLutBits = 12;
constexpr int kLutSize = 1 << kLutBits;
constexpr int kLutShift = kPrecision - kLutBits;
constexpr uint32_t kLutBytes = kLutSize * 8u;
// Decoders with shared memory to spare (the large-batch kernels: one 1024-thread CTA per SM; the chain kernels: one
// coder warp per CTA) stage a finer index of the same format: with 2^13 buckets a quantile falls beyond its bucket's
// first symbol in < 1 % of the lookups, so the second probe becomes a rarely taken warp-uniform branch instead of
// predicated instructions (and a dependent shared-memory load) in every symbol.  (2^14 buckets: 1 % faster still in
// the large-batch ANS decoder, but 128 KB per CTA would halve the chain decoders' CTAs per SM.)
#ifndef CTR_BIG_LUT_BITS
#define CTR_BIG_LUT_BITS 13
#endif
constexpr int kBigLutBits = CTR_BIG_LUT_BITS;
constexpr uint32_t kBigLutBytes = (1u << kBigLutBits) * 8u;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// Same, but opaque to the optimiser: the value stays in a register instead of being rematerialised from the
// CTA's shared-window base (several uniform-datapath instructions) at every use in a hot loop.
__device__ __forceinline__ uint32_t smem_u32_pinned(const void *p) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("" : "+r"(a));
    return a;
}

// ---- mbarrier + bulk copy ---------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_copy_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- 2-D tiled TMA (cp.async.bulk.tensor, SASS UTMALDG / UTMASTG) ------------------------------------
// The interleaved symbol array is a [rows][K] int32 matrix; a warp's 32 streams are a 128-byte wide column
// strip of it.  One instruction moves a box of kBoxRows rows x 32 columns between HBM and shared memory
// (row pitch in shared memory: 128 bytes), instead of one load / store instruction and one 64-bit address
// update per row.  `tmap` is the address of a CUtensorMap (kernel parameter, __grid_constant__).
#ifndef CTR_BOX_ROWS
#define CTR_BOX_ROWS 8
#endif
constexpr int kBoxRows = CTR_BOX_ROWS;            // rows per box (a multiple of kCheckEvery)
constexpr uint32_t kBoxBytes = kBoxRows * 128u;   // 1 KiB
__device__ __forceinline__ void mbar_init_addr(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_addr(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_addr(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// box at column x, row y of the tensor -> shared memory at `dst` (128-byte aligned); completes on `bar`
__device__ __forceinline__ void tma_load_box(uint32_t dst, const void *tmap, int32_t x, int32_t y, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(tmap), "r"(x), "r"(y), "r"(bar)
                 : "memory");
}
// shared memory at `src` -> box at column x, row y of the tensor (columns / rows outside the tensor are clipped)
__device__ __forceinline__ void tma_store_box(const void *tmap, int32_t x, int32_t y, uint32_t src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tmap), "r"(x), "r"(y), "r"(src)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// waits until at most N of this thread's bulk groups still read their shared-memory source
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy writes to shared memory become visible to the async proxy (before a TMA store reads them)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Stage `bytes` (multiple of 16) of table data into shared memory; all threads of the CTA call it
// and may read the data when it returns.  One elected thread drives the TMA engine.
// (stage_table_begin / stage_table_wait: the same in two halves, so that a kernel's prologue overlaps the copy)
__device__ __forceinline__ void stage_table_begin(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, bytes);
        // the bulk-copy size field is limited; split big tables into 32 KiB pieces
        uint32_t done = 0;
        while (done < bytes) {
            uint32_t piece = bytes - done < 32768u ? bytes - done : 32768u;
            bulk_copy_g2s((char *)dst_smem + done, (const char *)src_gmem + done, piece, bar);
            done += piece;
        }
    }
}
__device__ __forceinline__ void stage_table_wait(uint64_t *bar) { mbar_wait(bar, 0); }
__device__ __forceinline__ void stage_table(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    stage_table_begin(dst_smem, src_gmem, bytes, bar);
    stage_table_wait(bar);
}

// Two tables on one barrier (byte counts multiples of 16; a zero-byte table is skipped).
__device__ __forceinline__ void stage_tables(void *dst0, const void *src0, uint32_t bytes0, void *dst1, const void *src1,
                                             uint32_t bytes1, uint64_t *bar) {
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, bytes0 + bytes1);
        for (int t = 0; t < 2; ++t) {
            char *dst = (char *)(t ? dst1 : dst0);
            const char *src = (const char *)(t ? src1 : src0);
            const uint32_t bytes = t ? bytes1 : bytes0;
            for (uint32_t done = 0; done < bytes;) {
                const uint32_t piece = bytes - done < 32768u ? bytes - done : 32768u;
                bulk_copy_g2s(dst + done, src + done, piece, bar);
                done += piece;
            }
        }
    }
    mbar_wait(bar, 0);
}

// ---- streaming global accesses -----------------------------------------------------------------
__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int32_t ld_stream_s32(const int32_t *p) {
    int32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_stream_s32(int32_t *p, int32_t v) {
    asm volatile("st.global.L1::no_allocate.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src) {
    uint32_t lo = __shfl_sync(kFullMask, (uint32_t)v, src);
    uint32_t hi = __shfl_sync(kFullMask, (uint32_t)(v >> 32), src);
    return ((uint64_t)hi << 32) | lo;
}

// ---- status --------------------------------------------------------------------------------------
// status[0] = max error code, status[2..3] = index of one stream that raised it.
__device__ __forceinline__ void report_error(uint32_t *status, uint32_t code, uint64_t stream) {
    if (status == nullptr) return;
    if (atomicMax(&status[0], code) < code) {
        status[2] = (uint32_t)stream;
        status[3] = (uint32_t)(stream >> 32);
    }
}

// ---- geometry ------------------------------------------------------------------------------------
// interleaved deal: stream k owns symbols k, k+K, ...; its length and the offset of its first symbol
// in "stream-major" counting (used only to place scratch regions).
__device__ __host__ __forceinline__ uint64_t interleaved_len(uint64_t N, uint64_t K, uint64_t k) {
    return k < N ? (N - k + K - 1) / K : 0;
}
__device__ __host__ __forceinline__ uint64_t interleaved_start(uint64_t N, uint64_t K, uint64_t k) {
    const uint64_t T = K ? (N + K - 1) / K : 0;            // symbols of the longest stream
    const uint64_t full = T ? N - (T - 1) * K : 0;         // streams that have T symbols (1..K)
    return T ? k * (T - 1) + (k < full ? k : full) : 0;
}
// Scratch region of stream k whose symbols start at stream-major offset `o`: 32-word aligned and
// guaranteed to hold 25n/32 + 32 words.  A stream of n symbols needs at most 24 bits per symbol
// (ANS), or 24 + log2(1/(1-2^-8)) < 24.006 bits per symbol (range coder, truncation of range>>24),
// plus at most a handful of state / seal words, i.e. < 0.7502 n + 4 words; 25/32 = 0.78125.
__device__ __host__ __forceinline__ uint64_t scratch_words_for(uint64_t o) {
    return (o / 32) * 25 + ((o % 32) * 25) / 32;
}
__device__ __host__ __forceinline__ uint64_t scratch_start(uint64_t o, uint64_t k) {
    return 32ull * (scratch_words_for(o) / 32 + 2 * k);
}

}  // namespace ctr
