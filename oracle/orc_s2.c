/*
 * oracle/orc_s2.c -- S2 / Snappy block codec oracle.  TEST INFRASTRUCTURE ONLY (see orc_common.h).
 *
 * Restates the reference's generic (pure Go) block codec:
 *   emitLiteral / emitRepeat / emitCopy / emitCopyNoRepeat     s2/encode_go.go:80-289
 *   encodeBlockGo, encodeBlockGo64K, encodeBlockSnappyGo(64K)  s2/encode_all.go:72-898
 *   encodeBlockBetterGo, encodeBlockBetterGo64K                s2/encode_better.go:50-308, 485-731
 *   Encode / EncodeBetter / EncodeSnappy wrappers, MaxEncodedLen   s2/encode.go:29-60, 117-150, 204-240, 389-418
 *   Decode, decodedLen, s2Decode                               s2/decode.go:36-86, s2/decode_other.go:22-287
 * Pinned by the reference's byte-exact KATs (TestEmitLiteral / TestEmitCopy, s2/s2_test.go:827-942), its decode table
 * (TestDecode, s2/s2_test.go:252-470; TestInvalidVarint :214), the golden Snappy block
 * s2/testdata/Mark.Twain-Tom.Sawyer.txt.rawsnappy and pyarrow's Snappy codec (tests/test_oracle_s2.py).
 * The amd64 assembler encoders may choose different matches than the Go code restated here; both are valid
 * S2 and the reference's own tests only require round trips (s2/s2_test.go:93-115).
 */
#include "orc_common.h"
#include <stdlib.h>

#define TAG_LITERAL 0x00
#define TAG_COPY1 0x01
#define TAG_COPY2 0x02
#define TAG_COPY4 0x03
#define INPUT_MARGIN 8          /* s2/encode.go:371 */
#define MIN_NON_LITERAL 32      /* minNonLiteralBlockSize, s2/encode.go:375 */

/* ---- emitters ------------------------------------------------------------------------------------ */
ORC_API int64_t orc_s2_emit_literal(uint8_t *dst, const uint8_t *lit, size_t len) { /* encode_go.go:80-114 */
    if (len == 0) return 0;
    size_t i;
    uint64_t n = len - 1;
    if (n < 60) { dst[0] = (uint8_t)(n << 2 | TAG_LITERAL); i = 1; }
    else if (n < (1u << 8)) { dst[1] = (uint8_t)n; dst[0] = 60 << 2 | TAG_LITERAL; i = 2; }
    else if (n < (1u << 16)) { dst[2] = (uint8_t)(n >> 8); dst[1] = (uint8_t)n; dst[0] = 61 << 2 | TAG_LITERAL; i = 3; }
    else if (n < (1u << 24)) {
        dst[3] = (uint8_t)(n >> 16); dst[2] = (uint8_t)(n >> 8); dst[1] = (uint8_t)n; dst[0] = 62 << 2 | TAG_LITERAL; i = 4;
    } else {
        dst[4] = (uint8_t)(n >> 24); dst[3] = (uint8_t)(n >> 16); dst[2] = (uint8_t)(n >> 8); dst[1] = (uint8_t)n;
        dst[0] = 63 << 2 | TAG_LITERAL; i = 5;
    }
    memcpy(dst + i, lit, len);
    return (int64_t)(i + len);
}

ORC_API int64_t orc_s2_emit_repeat(uint8_t *dst, int64_t offset, int64_t length) { /* encode_go.go:118-163 */
    length -= 4;
    if (length <= 4) { dst[0] = (uint8_t)(length << 2 | TAG_COPY1); dst[1] = 0; return 2; }
    if (length < 8 && offset < 2048) {
        dst[1] = (uint8_t)offset;
        dst[0] = (uint8_t)((offset >> 8) << 5 | length << 2 | TAG_COPY1);
        return 2;
    }
    if (length < (1 << 8) + 4) {
        length -= 4;
        dst[2] = (uint8_t)length; dst[1] = 0; dst[0] = 5 << 2 | TAG_COPY1;
        return 3;
    }
    if (length < (1 << 16) + (1 << 8)) {
        length -= 1 << 8;
        dst[3] = (uint8_t)(length >> 8); dst[2] = (uint8_t)length; dst[1] = 0; dst[0] = 6 << 2 | TAG_COPY1;
        return 4;
    }
    const int64_t maxRepeat = (1 << 24) - 1;
    length -= 1 << 16;
    int64_t left = 0;
    if (length > maxRepeat) { left = length - maxRepeat + 4; length = maxRepeat - 4; }
    dst[4] = (uint8_t)(length >> 16); dst[3] = (uint8_t)(length >> 8); dst[2] = (uint8_t)length; dst[1] = 0;
    dst[0] = 7 << 2 | TAG_COPY1;
    if (left > 0) return 5 + orc_s2_emit_repeat(dst + 5, offset, left);
    return 5;
}

ORC_API int64_t orc_s2_emit_copy(uint8_t *dst, int64_t offset, int64_t length) { /* encode_go.go:172-231 */
    if (offset >= 65536) {
        int64_t i = 0;
        if (length > 64) {
            dst[4] = (uint8_t)(offset >> 24); dst[3] = (uint8_t)(offset >> 16); dst[2] = (uint8_t)(offset >> 8);
            dst[1] = (uint8_t)offset; dst[0] = 63 << 2 | TAG_COPY4;
            length -= 64;
            if (length >= 4) return 5 + orc_s2_emit_repeat(dst + 5, offset, length);
            i = 5;
        }
        if (length == 0) return i;
        dst[i + 0] = (uint8_t)((length - 1) << 2 | TAG_COPY4);
        dst[i + 1] = (uint8_t)offset; dst[i + 2] = (uint8_t)(offset >> 8); dst[i + 3] = (uint8_t)(offset >> 16);
        dst[i + 4] = (uint8_t)(offset >> 24);
        return i + 5;
    }
    if (length > 64) {
        int64_t off = 3;
        if (offset < 2048) {
            dst[1] = (uint8_t)offset;
            dst[0] = (uint8_t)((offset >> 8) << 5 | (8 - 4) << 2 | TAG_COPY1);
            length -= 8;
            off = 2;
        } else {
            dst[2] = (uint8_t)(offset >> 8); dst[1] = (uint8_t)offset; dst[0] = 59 << 2 | TAG_COPY2;
            length -= 60;
        }
        return off + orc_s2_emit_repeat(dst + off, offset, length);
    }
    if (length >= 12 || offset >= 2048) {
        dst[2] = (uint8_t)(offset >> 8); dst[1] = (uint8_t)offset; dst[0] = (uint8_t)((length - 1) << 2 | TAG_COPY2);
        return 3;
    }
    dst[1] = (uint8_t)offset;
    dst[0] = (uint8_t)((offset >> 8) << 5 | (length - 4) << 2 | TAG_COPY1);
    return 2;
}

ORC_API int64_t orc_s2_emit_copy_norepeat(uint8_t *dst, int64_t offset, int64_t length) { /* encode_go.go:241-289 */
    if (offset >= 65536) {
        int64_t i = 0;
        if (length > 64) {
            dst[4] = (uint8_t)(offset >> 24); dst[3] = (uint8_t)(offset >> 16); dst[2] = (uint8_t)(offset >> 8);
            dst[1] = (uint8_t)offset; dst[0] = 63 << 2 | TAG_COPY4;
            length -= 64;
            if (length >= 4) return 5 + orc_s2_emit_copy_norepeat(dst + 5, offset, length);
            i = 5;
        }
        if (length == 0) return i;
        dst[i + 0] = (uint8_t)((length - 1) << 2 | TAG_COPY4);
        dst[i + 1] = (uint8_t)offset; dst[i + 2] = (uint8_t)(offset >> 8); dst[i + 3] = (uint8_t)(offset >> 16);
        dst[i + 4] = (uint8_t)(offset >> 24);
        return i + 5;
    }
    if (length > 64) {
        dst[2] = (uint8_t)(offset >> 8); dst[1] = (uint8_t)offset; dst[0] = 59 << 2 | TAG_COPY2;
        length -= 60;
        return 3 + orc_s2_emit_copy_norepeat(dst + 3, offset, length);
    }
    if (length >= 12 || offset >= 2048) {
        dst[2] = (uint8_t)(offset >> 8); dst[1] = (uint8_t)offset; dst[0] = (uint8_t)((length - 1) << 2 | TAG_COPY2);
        return 3;
    }
    dst[1] = (uint8_t)offset;
    dst[0] = (uint8_t)((offset >> 8) << 5 | (length - 4) << 2 | TAG_COPY1);
    return 2;
}

/* ---- hashes (s2/encode_all.go:27-30, s2/encode_better.go:16-40) ---------------------------------- */
static inline uint32_t hash6(uint64_t u, unsigned h) { return (uint32_t)(((u << 16) * 227718039650203ull) >> (64 - h)); }
static inline uint32_t hash4(uint64_t u, unsigned h) { return ((uint32_t)u * 2654435761u) >> (32 - h); }
static inline uint32_t hash7(uint64_t u, unsigned h) { return (uint32_t)(((u << 8) * 58295818150454627ull) >> (64 - h)); }

/* ---- fast encoder: encodeBlockGo / encodeBlockGo64K / encodeBlockSnappyGo(64K) ----------------------
 * skipShift 6 = large-block variant, 5 = 64K variant (the u16 table only changes storage); snappy selects
 * emitCopyNoRepeat everywhere (encode_all.go:502-898). Returns bytes written, 0 = "not compressible". */
static int64_t encode_block_fast(uint8_t *dst, const uint8_t *src, int64_t n, int skipShift, int snappy) {
    const unsigned tableBits = 14;
    uint32_t *table = (uint32_t *)calloc(1u << tableBits, sizeof(uint32_t));
    if (!table) return ORC_ERR_INTERNAL;
    const int64_t sLimit = n - INPUT_MARGIN;
    const int64_t dstLimit = n - (n >> 5) - 5;
    int64_t nextEmit = 0, s = 1, d = 0, repeat = 1;
    uint64_t cv = orc_ld64(src + s);
#define RET(v) do { free(table); return (v); } while (0)
    for (;;) {
        int64_t candidate = 0;
        for (;;) {
            int64_t nextS = s + ((s - nextEmit) >> skipShift) + 4;
            if (nextS > sLimit) goto emitRemainder;
            uint32_t hash0 = hash6(cv, tableBits), hash1 = hash6(cv >> 8, tableBits);
            candidate = table[hash0];
            int64_t candidate2 = table[hash1];
            table[hash0] = (uint32_t)s;
            table[hash1] = (uint32_t)(s + 1);
            uint32_t hash2 = hash6(cv >> 16, tableBits);
            /* repeat check at s+1 (encode_all.go:117-167) */
            if ((uint32_t)(cv >> 8) == orc_ld32(src + s - repeat + 1)) {
                int64_t base = s + 1;
                for (int64_t i = base - repeat; base > nextEmit && i > 0 && src[i - 1] == src[base - 1];) { i--; base--; }
                if (d + (base - nextEmit) > dstLimit) RET(0);
                d += orc_s2_emit_literal(dst + d, src + nextEmit, (size_t)(base - nextEmit));
                int64_t cand = s - repeat + 4 + 1;
                s += 4 + 1;
                while (s <= sLimit) {
                    uint64_t diff = orc_ld64(src + s) ^ orc_ld64(src + cand);
                    if (diff != 0) { s += __builtin_ctzll(diff) >> 3; break; }
                    s += 8; cand += 8;
                }
                if (snappy) d += orc_s2_emit_copy_norepeat(dst + d, repeat, s - base);
                else if (nextEmit > 0) d += orc_s2_emit_repeat(dst + d, repeat, s - base);
                else d += orc_s2_emit_copy(dst + d, repeat, s - base);
                nextEmit = s;
                if (s >= sLimit) goto emitRemainder;
                cv = orc_ld64(src + s);
                continue;
            }
            if ((uint32_t)cv == orc_ld32(src + candidate)) break;
            candidate = table[hash2];
            if ((uint32_t)(cv >> 8) == orc_ld32(src + candidate2)) {
                table[hash2] = (uint32_t)(s + 2);
                candidate = candidate2;
                s++;
                break;
            }
            table[hash2] = (uint32_t)(s + 2);
            if ((uint32_t)(cv >> 16) == orc_ld32(src + candidate)) { s += 2; break; }
            cv = orc_ld64(src + nextS);
            s = nextS;
        }
        while (candidate > 0 && s > nextEmit && src[candidate - 1] == src[s - 1]) { candidate--; s--; }
        if (d + (s - nextEmit) > dstLimit) RET(0);
        d += orc_s2_emit_literal(dst + d, src + nextEmit, (size_t)(s - nextEmit));
        for (;;) {
            int64_t base = s;
            repeat = base - candidate;
            s += 4; candidate += 4;
            while (s <= n - 8) {
                uint64_t diff = orc_ld64(src + s) ^ orc_ld64(src + candidate);
                if (diff != 0) { s += __builtin_ctzll(diff) >> 3; break; }
                s += 8; candidate += 8;
            }
            d += snappy ? orc_s2_emit_copy_norepeat(dst + d, repeat, s - base) : orc_s2_emit_copy(dst + d, repeat, s - base);
            nextEmit = s;
            if (s >= sLimit) goto emitRemainder;
            if (d > dstLimit) RET(0);
            uint64_t x = orc_ld64(src + s - 2);
            uint32_t m2Hash = hash6(x, tableBits), currHash = hash6(x >> 16, tableBits);
            candidate = table[currHash];
            table[m2Hash] = (uint32_t)(s - 2);
            table[currHash] = (uint32_t)s;
            if ((uint32_t)(x >> 16) != orc_ld32(src + candidate)) { cv = orc_ld64(src + s + 1); s++; break; }
        }
    }
emitRemainder:
    if (nextEmit < n) {
        if (d + n - nextEmit > dstLimit) RET(0);
        d += orc_s2_emit_literal(dst + d, src + nextEmit, (size_t)(n - nextEmit));
    }
    RET(d);
#undef RET
}

/* ---- better encoder: encodeBlockBetterGo / ...Go64K (encode_better.go:50-308, 485-731) ------------- */
static int64_t encode_block_better(uint8_t *dst, const uint8_t *src, int64_t n, unsigned lBits, unsigned sBits,
                                   int skipShift) {
    const int64_t sLimit = n - INPUT_MARGIN;
    if (n < MIN_NON_LITERAL) return 0;
    uint32_t *lTable = (uint32_t *)calloc((size_t)1 << lBits, sizeof(uint32_t));
    uint32_t *sTable = (uint32_t *)calloc((size_t)1 << sBits, sizeof(uint32_t));
    if (!lTable || !sTable) { free(lTable); free(sTable); return ORC_ERR_INTERNAL; }
    const int64_t dstLimit = n - (n >> 5) - 6;
    int64_t nextEmit = 0, s = 1, d = 0, repeat = 0;
    uint64_t cv = orc_ld64(src + s);
#define RET(v) do { free(lTable); free(sTable); return (v); } while (0)
    for (;;) {
        int64_t candidateL = 0, nextS = 0;
        for (;;) {
            nextS = s + ((s - nextEmit) >> skipShift) + 1;
            if (nextS > sLimit) goto emitRemainder;
            uint32_t hashL = hash7(cv, lBits), hashS = hash4(cv, sBits);
            candidateL = lTable[hashL];
            int64_t candidateS = sTable[hashS];
            lTable[hashL] = (uint32_t)s;
            sTable[hashS] = (uint32_t)s;
            uint64_t valLong = orc_ld64(src + candidateL), valShort = orc_ld64(src + candidateS);
            if (cv == valLong) break;
            if (cv == valShort) { candidateL = candidateS; break; }
            /* (the repeat probe at encode_better.go:107-158 is compiled out in the reference: `if false && ...`) */
            if ((uint32_t)cv == (uint32_t)valLong) break;
            if ((uint32_t)cv == (uint32_t)valShort) {
                hashL = hash7(cv >> 8, lBits);
                candidateL = lTable[hashL];
                lTable[hashL] = (uint32_t)(s + 1);
                if ((uint32_t)(cv >> 8) == orc_ld32(src + candidateL)) { s++; break; }
                candidateL = candidateS;
                break;
            }
            cv = orc_ld64(src + nextS);
            s = nextS;
        }
        while (candidateL > 0 && s > nextEmit && src[candidateL - 1] == src[s - 1]) { candidateL--; s--; }
        if (d + (s - nextEmit) > dstLimit) RET(0);
        int64_t base = s, offset = base - candidateL;
        s += 4; candidateL += 4;
        while (s < n) {
            if (n - s < 8) {
                if (src[s] == src[candidateL]) { s++; candidateL++; continue; }
                break;
            }
            uint64_t diff = orc_ld64(src + s) ^ orc_ld64(src + candidateL);
            if (diff != 0) { s += __builtin_ctzll(diff) >> 3; break; }
            s += 8; candidateL += 8;
        }
        if (offset > 65535 && s - base <= 5 && repeat != offset) { /* encode_better.go:236-244 (large blocks only) */
            s = nextS + 1;
            if (s >= sLimit) goto emitRemainder;
            cv = orc_ld64(src + s);
            continue;
        }
        d += orc_s2_emit_literal(dst + d, src + nextEmit, (size_t)(base - nextEmit));
        if (repeat == offset) d += orc_s2_emit_repeat(dst + d, offset, s - base);
        else { d += orc_s2_emit_copy(dst + d, offset, s - base); repeat = offset; }
        nextEmit = s;
        if (s >= sLimit) goto emitRemainder;
        if (d > dstLimit) RET(0);
        int64_t index0 = base + 1, index1 = s - 2;
        uint64_t cv0 = orc_ld64(src + index0), cv1 = orc_ld64(src + index1);
        lTable[hash7(cv0, lBits)] = (uint32_t)index0;
        sTable[hash4(cv0 >> 8, sBits)] = (uint32_t)(index0 + 1);
        lTable[hash7(cv1, lBits)] = (uint32_t)index1;
        sTable[hash4(cv1 >> 8, sBits)] = (uint32_t)(index1 + 1);
        index0 += 1; index1 -= 1;
        cv = orc_ld64(src + s);
        int64_t index2 = (index0 + index1 + 1) >> 1;
        while (index2 < index1) {
            lTable[hash7(orc_ld64(src + index0), lBits)] = (uint32_t)index0;
            lTable[hash7(orc_ld64(src + index2), lBits)] = (uint32_t)index2;
            index0 += 2; index2 += 2;
        }
    }
emitRemainder:
    if (nextEmit < n) {
        if (d + n - nextEmit > dstLimit) RET(0);
        d += orc_s2_emit_literal(dst + d, src + nextEmit, (size_t)(n - nextEmit));
    }
    RET(d);
#undef RET
}

/* ---- wrappers ------------------------------------------------------------------------------------- */
ORC_API int64_t orc_s2_max_encoded_len(int64_t srcLen) { /* s2/encode.go:389-418, 64-bit platform */
    if (srcLen < 0) return -1;
    uint64_t n = (uint64_t)srcLen;
    if (n > 0xffffffffull) return -1;
    unsigned bl = n ? 64 - (unsigned)__builtin_clzll(n) : 0;
    n = n + (bl + 7) / 7;
    uint64_t extra = 0; /* literalExtraSize, s2/s2.go:129-145 */
    if (srcLen != 0) extra = srcLen < 60 ? 1 : srcLen < (1 << 8) ? 2 : srcLen < (1 << 16) ? 3 : srcLen < (1 << 24) ? 4 : 5;
    n += extra;
    if (n > 0xffffffffull) return -1;
    return (int64_t)n;
}

static size_t put_uvarint(uint8_t *dst, uint64_t v) {
    size_t i = 0;
    while (v >= 0x80) { dst[i++] = (uint8_t)v | 0x80; v >>= 7; }
    dst[i++] = (uint8_t)v;
    return i;
}

/* mode: 0 Encode, 1 EncodeBetter, 2 EncodeSnappy.  The block body only (encodeBlock*), 0 = not compressible. */
ORC_API int64_t orc_s2_encode_block(uint8_t *dst, const uint8_t *src, int64_t n, int mode) {
    if (n < MIN_NON_LITERAL) return 0;
    /* the encoders read 8 bytes at positions up to n-8 only, so no padding is required */
    if (mode == 1) return n <= (64 << 10) ? encode_block_better(dst, src, n, 16, 13, 6) : encode_block_better(dst, src, n, 17, 14, 7);
    int snappy = mode == 2;
    return n <= (64 << 10) ? encode_block_fast(dst, src, n, 5, snappy) : encode_block_fast(dst, src, n, 6, snappy);
}

/* Encode / EncodeBetter / EncodeSnappy (s2/encode.go:29-60): uvarint length + body, or the input as one literal */
ORC_API int64_t orc_s2_encode(uint8_t *dst, size_t cap, const uint8_t *src, int64_t n, int mode) {
    int64_t need = orc_s2_max_encoded_len(n);
    if (need < 0) return ORC_ERR_TOO_BIG;
    if ((int64_t)cap < need) return ORC_ERR_DST_SMALL;
    int64_t d = (int64_t)put_uvarint(dst, (uint64_t)n);
    if (n == 0) return d;
    if (n < MIN_NON_LITERAL) return d + orc_s2_emit_literal(dst + d, src, (size_t)n);
    int64_t b = orc_s2_encode_block(dst + d, src, n, mode);
    if (b < 0) return b;
    if (b > 0) return d + b;
    return d + orc_s2_emit_literal(dst + d, src, (size_t)n);
}

/* decodedLen (s2/decode.go:36-49): returns header bytes, -1 on ErrCorrupt */
ORC_API int orc_s2_decoded_len(const uint8_t *src, size_t n, uint64_t *out) {
    uint64_t v = 0;
    unsigned shift = 0;
    for (size_t i = 0; i < n && i < 10; i++) { /* binary.Uvarint */
        uint8_t b = src[i];
        if (b < 0x80) {
            if (i == 9 && b > 1) return -1; /* overflow */
            v |= (uint64_t)b << shift;
            if (i + 1 > 5 || v > 0xffffffffull) return -1;
            *out = v;
            return (int)(i + 1);
        }
        v |= (uint64_t)(b & 0x7f) << shift;
        shift += 7;
    }
    return -1; /* n <= 0: buffer too small or overflow */
}

/* s2Decode (s2/decode_other.go:22-287): 0 ok, 1 corrupt.  One loop with the bounds checks of the reference's tail loop;
 * its unchecked fast loop accepts exactly the same streams. */
static int s2_decode_body(uint8_t *dst, size_t dlen, const uint8_t *src, size_t slen) {
    size_t d = 0, s = 0;
    int64_t offset = 0, length = 0;
    while (s < slen) {
        uint8_t tag = src[s];
        switch (tag & 3) {
        case TAG_LITERAL: {
            uint32_t x = tag >> 2;
            if (x < 60) s++;
            else if (x == 60) { s += 2; if (s > slen) return 1; x = src[s - 1]; }
            else if (x == 61) { s += 3; if (s > slen) return 1; x = (uint32_t)src[s - 2] | (uint32_t)src[s - 1] << 8; }
            else if (x == 62) { s += 4; if (s > slen) return 1; x = (uint32_t)src[s - 3] | (uint32_t)src[s - 2] << 8 | (uint32_t)src[s - 1] << 16; }
            else { s += 5; if (s > slen) return 1; x = (uint32_t)src[s - 4] | (uint32_t)src[s - 3] << 8 | (uint32_t)src[s - 2] << 16 | (uint32_t)src[s - 1] << 24; }
            length = (int64_t)x + 1;
            if ((uint64_t)length > dlen - d || (uint64_t)length > slen - s) return 1;
            memcpy(dst + d, src + s, (size_t)length);
            d += (size_t)length; s += (size_t)length;
            continue;
        }
        case TAG_COPY1: {
            s += 2;
            if (s > slen) return 1;
            length = (src[s - 2] >> 2) & 7;
            int64_t toffset = (int64_t)(((uint32_t)src[s - 2] & 0xe0) << 3 | (uint32_t)src[s - 1]);
            if (toffset == 0) { /* repeat: keep last offset (decode_other.go:74-101) */
                if (length == 5) { s += 1; if (s > slen) return 1; length = (int64_t)src[s - 1] + 4; }
                else if (length == 6) { s += 2; if (s > slen) return 1; length = (int64_t)((uint32_t)src[s - 2] | (uint32_t)src[s - 1] << 8) + (1 << 8); }
                else if (length == 7) {
                    s += 3; if (s > slen) return 1;
                    length = (int64_t)((uint32_t)src[s - 3] | (uint32_t)src[s - 2] << 8 | (uint32_t)src[s - 1] << 16) + (1 << 16);
                }
            } else offset = toffset;
            length += 4;
            break;
        }
        case TAG_COPY2:
            s += 3;
            if (s > slen) return 1;
            length = 1 + (src[s - 3] >> 2);
            offset = (int64_t)((uint32_t)src[s - 2] | (uint32_t)src[s - 1] << 8);
            break;
        default:
            s += 5;
            if (s > slen) return 1;
            length = 1 + (src[s - 5] >> 2);
            offset = (int64_t)((uint32_t)src[s - 4] | (uint32_t)src[s - 3] << 8 | (uint32_t)src[s - 2] << 16 | (uint32_t)src[s - 1] << 24);
            break;
        }
        if (offset <= 0 || (int64_t)d < offset || (uint64_t)length > dlen - d) return 1;
        for (int64_t i = 0; i < length; i++) dst[d + (size_t)i] = dst[d - (size_t)offset + (size_t)i]; /* forward copy */
        d += (size_t)length;
    }
    return d != dlen;
}

/* Decode (s2/decode.go:58-86): returns the decoded length or ORC_ERR_CORRUPT / ORC_ERR_DST_SMALL */
ORC_API int64_t orc_s2_decode(uint8_t *dst, size_t cap, const uint8_t *src, size_t n) {
    uint64_t dlen = 0;
    int h = orc_s2_decoded_len(src, n, &dlen);
    if (h < 0) return ORC_ERR_CORRUPT;
    if (dlen > cap) return ORC_ERR_DST_SMALL;
    if (s2_decode_body(dst, (size_t)dlen, src + h, n - (size_t)h)) return ORC_ERR_CORRUPT;
    return (int64_t)dlen;
}
