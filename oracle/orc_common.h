/*
 * oracle/orc_common.h -- shared helpers for the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  The oracle is a plain-C restatement of the
 * reference's (klauspost/compress v1.19.0) block codec algorithms.  It is the
 * checker for the CUDA product path: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  Nothing under
 * compress_b200/ links, imports or calls it.
 *
 * Parity status: DECODE paths are pinned against the reference's own golden
 * vectors (zstd/testdata/decoder.zip, good.zip, bad.zip, s2 KATs; see
 * tests/).  ENCODE byte-parity with the Go binary is "parity unpinned": the
 * reference holds no golden compressed bytes (zstd/README.md:126-128) and no
 * Go toolchain exists in this image; the encode restatement is pinned by
 * round trips through this oracle's decoder AND the system libzstd 1.5.5.
 */
#ifndef ORC_COMMON_H
#define ORC_COMMON_H

#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* error codes shared by all oracle entry points (negative return values) */
enum {
    ORC_OK = 0,
    ORC_ERR_INCOMPRESSIBLE = -1, /* huff0.ErrIncompressible (huff0/huff0.go:31) */
    ORC_ERR_USE_RLE = -2,        /* huff0.ErrUseRLE        (huff0/huff0.go:34) */
    ORC_ERR_TOO_BIG = -3,        /* huff0.ErrTooBig        (huff0/huff0.go:37) */
    ORC_ERR_DST_SMALL = -4,
    ORC_ERR_CORRUPT = -5,
    ORC_ERR_INTERNAL = -6,
    ORC_ERR_MAGIC = -7,
    ORC_ERR_WINDOW = -8,
    ORC_ERR_CRC = -9,
    ORC_ERR_SIZE = -10,
    ORC_ERR_UNSUPPORTED = -11
};

static inline uint32_t orc_highbit32(uint32_t v) { /* bits.Len32(v)-1; v==0 -> 0xffffffff like Go */
    return v ? (uint32_t)(31 - __builtin_clz(v)) : 0xffffffffu;
}
static inline uint16_t orc_ld16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return v; }
static inline uint32_t orc_ld32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t orc_ld64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline void orc_st16(uint8_t *p, uint16_t v) { memcpy(p, &v, 2); }
static inline void orc_st32(uint8_t *p, uint32_t v) { memcpy(p, &v, 4); }

/* ---- forward LSB-first bit writer (zstd/bitwriter.go, huff0/bitwriter.go, fse/bitwriter.go).
 * All three reference writers produce the plain LSB-first concatenation of the
 * added fields; flush granularity does not change the bytes. */
typedef struct {
    uint8_t *out;
    size_t cap;
    size_t pos;     /* bytes written */
    uint64_t acc;
    unsigned nbits; /* < 32 after every add */
    int overflow;
} orc_bw;

static inline void orc_bw_init(orc_bw *b, uint8_t *out, size_t cap) {
    b->out = out; b->cap = cap; b->pos = 0; b->acc = 0; b->nbits = 0; b->overflow = 0;
}
static inline void orc_bw_add(orc_bw *b, uint64_t value, unsigned bits) { /* bits <= 32 */
    if (bits == 0) return;
    value &= ((1ull << bits) - 1);
    b->acc |= value << b->nbits;
    b->nbits += bits;
    if (b->nbits >= 32) { /* flush32 (zstd/bitwriter.go:78-88) */
        if (b->pos + 4 <= b->cap) { uint32_t w = (uint32_t)b->acc; memcpy(b->out + b->pos, &w, 4); }
        else b->overflow = 1;
        b->pos += 4;
        b->acc >>= 32;
        b->nbits -= 32;
    }
}
static inline void orc_bw_add64(orc_bw *b, uint64_t value, unsigned bits) { /* up to 64 bits */
    if (bits > 32) { orc_bw_add(b, value & 0xffffffffull, 32); orc_bw_add(b, value >> 32, bits - 32); }
    else orc_bw_add(b, value, bits);
}
/* close(): end-mark bit then flush to the next byte boundary (bitwriter.go:100-105) */
static inline void orc_bw_close(orc_bw *b) {
    orc_bw_add(b, 1, 1);
    while (b->nbits > 0) {
        if (b->pos < b->cap) b->out[b->pos] = (uint8_t)b->acc; else b->overflow = 1;
        b->pos++;
        b->acc >>= 8;
        b->nbits = b->nbits > 8 ? b->nbits - 8 : 0;
    }
    b->acc = 0;
}
/* flush whole bytes only, no end mark (fse cState.flush -> bw.flush, fse/compress.go:118) */

/* ---- backward bit reader (zstd/bitreader.go, huff0/bitreader.go, fse/bitreader.go).
 * Stream = little-endian integer; last byte holds the end mark in its highest
 * set bit; bits are consumed from just below the mark downwards. */
typedef struct {
    const uint8_t *in;
    size_t len;
    int64_t total; /* payload bits (mark and padding excluded) */
    int64_t pos;   /* bits consumed */
} orc_br;

static inline int orc_br_init(orc_br *b, const uint8_t *in, size_t len) {
    if (len < 1) return ORC_ERR_CORRUPT;            /* "corrupt stream: too short" */
    uint8_t v = in[len - 1];
    if (v == 0) return ORC_ERR_CORRUPT;             /* "did not find end of stream" */
    b->in = in; b->len = len;
    b->total = (int64_t)8 * (int64_t)(len - 1) + (int64_t)orc_highbit32(v);
    b->pos = 0;
    return 0;
}
/* peek n (<=32) bits at current position without advancing; bits past the
 * start of the buffer read as zero (matches value<<bitsRead behaviour). */
static inline uint32_t orc_br_peek(const orc_br *b, unsigned n) {
    if (n == 0) return 0;
    int64_t lo = b->total - b->pos - (int64_t)n; /* index of lowest bit wanted */
    uint64_t v = 0;
    /* gather bits [lo, lo+n) of the LE integer; negative indices are zero */
    int64_t start = lo < 0 ? 0 : lo;
    int64_t end = lo + (int64_t)n; /* exclusive */
    if (end <= 0) return 0;
    size_t byte0 = (size_t)(start >> 3);
    unsigned sh = (unsigned)(start & 7);
    uint64_t acc = 0;
    for (unsigned i = 0; i < 6; i++) {
        size_t bi = byte0 + i;
        if (bi < b->len) acc |= (uint64_t)b->in[bi] << (8 * i);
    }
    acc >>= sh;
    unsigned got = (unsigned)(end - start);
    v = acc & ((got >= 64) ? ~0ull : ((1ull << got) - 1));
    if (lo < 0) v <<= (unsigned)(-lo);
    return (uint32_t)v;
}
static inline uint32_t orc_br_read(orc_br *b, unsigned n) {
    uint32_t v = orc_br_peek(b, n);
    b->pos += n;
    return v;
}
static inline int orc_br_finished(const orc_br *b) { return b->pos >= b->total; }
static inline int orc_br_overread(const orc_br *b) { return b->pos > b->total; }

/* ---- shared entry points implemented across the oracle's .c files ---- */
uint64_t orc_xxh64(const void *data, size_t len, uint64_t seed);

/* standalone FSE, used only for huff0 weight tables (fse/compress.go, fse/decompress.go) */
int64_t orc_fse_compress(const uint8_t *in, size_t n, const uint32_t *count, unsigned symbolLen,
                         unsigned maxCount, unsigned tableLogReq, uint8_t *out, size_t cap);
int64_t orc_fse_decompress(const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t limit);

#endif
