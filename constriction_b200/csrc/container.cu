// container.cu -- the wire / on-disk form of a batch container.  The reference stores raw
// native-endian u32 words and leaves framing to the user (src/lib.rs:425-580 describes the (position, state) snapshots
// that make random access possible; examples/python/01-hello-world.ipynb byte-swaps by hand).  A batch needs framing: K
// streams, their word offsets, optionally their symbol offsets and the Pos::pos() records the encoders took, so that
// any single stream -- or any chunk of one -- can be cut out and handed to stock constriction.
//
// Layout, all fields little-endian, every section 8-byte aligned:
//   0   char[8]  "CTRB200\0"
//   8   u32 version (1)        u32 coder (0 ANS / stack, 1 range / queue)
//   16  u32 word_bits (32|16)  u32 precision (24|12)
//   24  u64 n_streams K        u64 n_symbols N          u64 total_words
//   48  u32 checkpoint_every   u32 flags (bit 0: interleaved deal, no symbol offsets stored)
//   56  u64 n_records
//   64  [u64 sym_offsets[K+1]]  u64 offsets[K+1]  [u64 ckpt_offsets[K+1]  u64 records[n_records * R]]  words, zero padded
//       R = 2 (ANS: words pushed, state) or 4 (range: words pushed, lower, range, 0)
// Host code only (no device work); the views returned by ctr_container_unpack point into the caller's buffer.
#include <stdint.h>
#include <string.h>

#include "../../include/constriction_b200.h"

namespace {
constexpr char kMagic[8] = {'C', 'T', 'R', 'B', '2', '0', '0', '\0'};
constexpr size_t kHeader = 64;
inline size_t pad8(size_t n) { return (n + 7) / 8 * 8; }
inline bool little_endian() {
    const uint32_t one = 1;
    return *reinterpret_cast<const unsigned char *>(&one) == 1;
}
size_t record_words(uint32_t coder) { return coder == 0 ? 2 : 4; }
bool view_ok(const ctr_container_view *v) {
    if (!v || v->coder > 1 || (v->word_bits != 32 && v->word_bits != 16) || !v->offsets) return false;
    if (!(v->flags & 1u) && !v->sym_offsets && v->n_streams) return false;
    if (v->checkpoint_every && (!v->ckpt_offsets || (!v->records && v->n_records))) return false;
    if (v->total_words && !v->words) return false;
    return true;
}
}  // namespace

extern "C" size_t ctr_container_size(const ctr_container_view *v) {
    if (!view_ok(v)) return 0;
    const size_t k1 = (size_t)v->n_streams + 1;
    size_t n = kHeader + k1 * 8;
    if (!(v->flags & 1u)) n += k1 * 8;
    if (v->checkpoint_every) n += k1 * 8 + (size_t)v->n_records * record_words(v->coder) * 8;
    return n + pad8((size_t)v->total_words * (v->word_bits / 8));
}

extern "C" int ctr_container_pack(const ctr_container_view *v, void *out, size_t out_bytes) {
    if (!view_ok(v) || !out || !little_endian()) return CTR_ERR_BAD_ARGUMENT;
    const size_t need = ctr_container_size(v);
    if (out_bytes < need) return CTR_ERR_OUT_OF_SPACE;
    const size_t k1 = (size_t)v->n_streams + 1;
    if (v->offsets[v->n_streams] != v->total_words) return CTR_ERR_BAD_ARGUMENT;
    unsigned char *p = static_cast<unsigned char *>(out);
    memset(p, 0, kHeader);
    memcpy(p, kMagic, 8);
    const uint32_t h32[] = {1u, v->coder, v->word_bits, v->precision};
    memcpy(p + 8, h32, 16);
    const uint64_t h64[] = {v->n_streams, v->n_symbols, v->total_words};
    memcpy(p + 24, h64, 24);
    const uint32_t c32[] = {v->checkpoint_every, v->flags};
    memcpy(p + 48, c32, 8);
    memcpy(p + 56, &v->n_records, 8);
    size_t at = kHeader;
    if (!(v->flags & 1u)) {
        memcpy(p + at, v->sym_offsets, k1 * 8);
        at += k1 * 8;
    }
    memcpy(p + at, v->offsets, k1 * 8);
    at += k1 * 8;
    if (v->checkpoint_every) {
        memcpy(p + at, v->ckpt_offsets, k1 * 8);
        at += k1 * 8;
        const size_t rb = (size_t)v->n_records * record_words(v->coder) * 8;
        if (rb) memcpy(p + at, v->records, rb);
        at += rb;
    }
    const size_t wb = (size_t)v->total_words * (v->word_bits / 8);
    if (wb) memcpy(p + at, v->words, wb);
    memset(p + at + wb, 0, pad8(wb) - wb);
    return CTR_OK;
}

extern "C" int ctr_container_unpack(const void *bytes, size_t n_bytes, ctr_container_view *out) {
    if (!bytes || !out || !little_endian() || reinterpret_cast<uintptr_t>(bytes) % 8 != 0) return CTR_ERR_BAD_ARGUMENT;
    const unsigned char *p = static_cast<const unsigned char *>(bytes);
    if (n_bytes < kHeader || memcmp(p, kMagic, 8) != 0) return CTR_ERR_INVALID_DATA;
    ctr_container_view v;
    memset(&v, 0, sizeof v);
    uint32_t h32[4], c32[2];
    uint64_t h64[3];
    memcpy(h32, p + 8, 16);
    memcpy(h64, p + 24, 24);
    memcpy(c32, p + 48, 8);
    memcpy(&v.n_records, p + 56, 8);
    if (h32[0] != 1u) return CTR_ERR_INVALID_DATA;
    v.coder = h32[1];
    v.word_bits = h32[2];
    v.precision = h32[3];
    v.n_streams = h64[0];
    v.n_symbols = h64[1];
    v.total_words = h64[2];
    v.checkpoint_every = c32[0];
    v.flags = c32[1];
    if (v.coder > 1 || (v.word_bits != 32 && v.word_bits != 16) || v.n_streams > (1ull << 40) || v.n_records > (1ull << 40) ||
        v.total_words > (1ull << 48))
        return CTR_ERR_INVALID_DATA;
    const size_t k1 = (size_t)v.n_streams + 1;
    size_t need = kHeader + k1 * 8 + ((v.flags & 1u) ? 0 : k1 * 8);
    if (v.checkpoint_every) need += k1 * 8 + (size_t)v.n_records * record_words(v.coder) * 8;
    need += pad8((size_t)v.total_words * (v.word_bits / 8));
    if (n_bytes < need) return CTR_ERR_INVALID_DATA;
    size_t at = kHeader;
    if (!(v.flags & 1u)) {
        v.sym_offsets = reinterpret_cast<const uint64_t *>(p + at);
        at += k1 * 8;
    }
    v.offsets = reinterpret_cast<const uint64_t *>(p + at);
    at += k1 * 8;
    if (v.checkpoint_every) {
        v.ckpt_offsets = reinterpret_cast<const uint64_t *>(p + at);
        at += k1 * 8;
        v.records = reinterpret_cast<const uint64_t *>(p + at);
        at += (size_t)v.n_records * record_words(v.coder) * 8;
    }
    v.words = p + at;
    // offsets must be monotone and end at total_words; symbol offsets monotone and within N
    if (v.offsets[0] != 0 || v.offsets[v.n_streams] != v.total_words) return CTR_ERR_INVALID_DATA;
    for (uint64_t k = 0; k < v.n_streams; ++k)
        if (v.offsets[k + 1] < v.offsets[k]) return CTR_ERR_INVALID_DATA;
    if (v.sym_offsets) {
        if (v.sym_offsets[v.n_streams] > v.n_symbols) return CTR_ERR_INVALID_DATA;
        for (uint64_t k = 0; k < v.n_streams; ++k)
            if (v.sym_offsets[k + 1] < v.sym_offsets[k]) return CTR_ERR_INVALID_DATA;
    }
    if (v.ckpt_offsets && v.ckpt_offsets[v.n_streams] > v.n_records) return CTR_ERR_INVALID_DATA;
    *out = v;
    return CTR_OK;
}
