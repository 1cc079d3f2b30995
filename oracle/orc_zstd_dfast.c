/*
 * oracle/orc_zstd_dfast.c -- CPU restatement of the zstd level-2 ("default") match finder.
 * TEST INFRASTRUCTURE ONLY -- see orc_common.h.  Never linked into libb200comp.so.
 *
 * Follows doubleFastEncoder in /root/reference/zstd/enc_dfast.go:
 *   constants :9-22, Encode (history) :38-367, EncodeNoHist :372-675.
 * Parity unpinned at byte level (the reference holds no encoder golden vectors and Go cannot run here);
 * pinned functionally: every stream decodes with the pinned decoder oracle (tests/test_oracle_dfast.py).
 */
#include "orc_zstd.h"
#include <stdlib.h>
#include <string.h>

#define DF_LONG_BITS 17  /* dFastLongTableBits, enc_dfast.go:10 */
#define DF_SHORT_BITS 15 /* dFastShortTableBits = tableBits, enc_dfast.go:18, enc_fast.go:13 */

typedef struct { uint32_t val; int32_t offset; } df_entry; /* tableEntry, enc_fast.go:21-24 */

typedef struct {
    df_entry shortTab[1 << DF_SHORT_BITS];
    df_entry longTab[1 << DF_LONG_BITS];
    int32_t maxMatchOff;
    int32_t cur;
} df_state;

/* hashLen(u, bits, 8) and hashLen(u, bits, 5), zstd/hash.go:7-35 */
static inline uint32_t df_hash_long(uint64_t u) { return (uint32_t)((u * 0xcf1bbcdcb7a56463ull) >> (64 - DF_LONG_BITS)); }
static inline uint32_t df_hash_short(uint64_t u) { return (uint32_t)(((u << 24) * 889523592379ull) >> (64 - DF_SHORT_BITS)); }

/* matchLen(src[s:end], src[t:]), zstd/matchlen_generic.go:16 */
static int32_t df_match_len(const uint8_t *src, int32_t s, int32_t t, int32_t end) {
    int32_t n = 0;
    while (s + n + 8 <= end) {
        uint64_t x = orc_ld64(src + s + n) ^ orc_ld64(src + t + n);
        if (x) return n + (int32_t)(__builtin_ctzll(x) >> 3);
        n += 8;
    }
    while (s + n < end && src[s + n] == src[t + n]) n++;
    return n;
}

static inline void df_put(df_entry *tab, uint32_t h, int32_t off, uint32_t val) { tab[h].offset = off; tab[h].val = val; }

static void df_lits(orc_blockenc *b, const uint8_t *src, int32_t from, int32_t until) {
    if (until > from) orc_blockenc_add_literals(b, src + from, (size_t)(until - from));
}

/* One block [s0,end) of the buffer src (= e.hist in the history variant).  The two variants differ in:
 *  - when "can use repeats" is sampled (live vs once per outer iteration, :118 vs :453),
 *  - the repIndex >= 0 guard (:141),
 *  - the match-length caps while extending backwards (:156, :253),
 *  - which value is hashed into the short table in the offset-2 loop (:326 vs :637: the no-history
 *    code hashes cv1>>8 there, cv1 being the already shifted end-2 word), and
 *  - writing back recentOffsets (:361-362, not done by EncodeNoHist). */
static void df_encode_block(df_state *e, orc_blockenc *blk, const uint8_t *src, int32_t s0, int32_t end, int nohist) {
    const int32_t inputMargin = 8 + 2;
    const int32_t minNonLiteralBlockSize = 16;
    int32_t s = s0;
    blk->size = (size_t)(end - s0);
    if (end - s0 < minNonLiteralBlockSize) {
        blk->extraLits = (size_t)(end - s0);
        blk->nlit = 0;
        orc_blockenc_add_literals(blk, src + s0, (size_t)(end - s0));
        return;
    }
    const int32_t sLimit = end - inputMargin;
    const int32_t stepSize = 1;
    const int kSearchStrength = 8;
    int32_t nextEmit = s;
    uint64_t cv = orc_ld64(src + s);
    int32_t offset1 = (int32_t)blk->recentOffsets[0];
    int32_t offset2 = (int32_t)blk->recentOffsets[1];

    for (;;) {
        int32_t t = 0;
        const int canRepeatOuter = blk->nseq > 2;
        for (;;) {
            uint32_t hL = df_hash_long(cv), hS = df_hash_short(cv);
            df_entry candL = e->longTab[hL], candS = e->shortTab[hS];
            const int32_t repOff = 1;
            int32_t repIndex = s - offset1 + repOff;
            df_put(e->longTab, hL, s + e->cur, (uint32_t)cv);
            df_put(e->shortTab, hS, s + e->cur, (uint32_t)cv);

            int repOK = nohist ? (blk->nseq > 2) : (canRepeatOuter && repIndex >= 0);
            if (repOK && orc_ld32(src + repIndex) == (uint32_t)(cv >> (repOff * 8))) {
                int32_t length = 4 + df_match_len(src, s + 4 + repOff, repIndex + 4, end);
                uint32_t mlen = (uint32_t)(length - ORC_ZSTD_MINMATCH);
                int32_t start = s + repOff;
                int32_t startLimit = nextEmit + 1;
                int32_t tMin = s - e->maxMatchOff; if (tMin < 0) tMin = 0;
                while (repIndex > tMin && start > startLimit && src[repIndex - 1] == src[start - 1] &&
                       (nohist || mlen < ORC_ZSTD_MAX_MATCHLEN - ORC_ZSTD_MINMATCH - 1)) {
                    repIndex--; start--; mlen++;
                }
                uint32_t litLen = start != nextEmit ? (uint32_t)(start - nextEmit) : 0;
                df_lits(blk, src, nextEmit, start);
                orc_blockenc_add_seq(blk, litLen, mlen, 1);
                s += length + repOff;
                nextEmit = s;
                if (s >= sLimit) goto done;
                cv = orc_ld64(src + s);
                continue;
            }
            int32_t coffsetL = s - (candL.offset - e->cur);
            int32_t coffsetS = s - (candS.offset - e->cur);
            if (coffsetL < e->maxMatchOff && (uint32_t)cv == candL.val) { /* long match, :175 / :487 */
                t = candL.offset - e->cur;
                break;
            }
            if (coffsetS < e->maxMatchOff && (uint32_t)cv == candS.val) { /* short match; lazy long probe at s+1 */
                const int32_t checkAt = 1;
                uint64_t cv1 = orc_ld64(src + s + checkAt);
                hL = df_hash_long(cv1);
                candL = e->longTab[hL];
                coffsetL = s - (candL.offset - e->cur) + checkAt;
                df_put(e->longTab, hL, s + checkAt + e->cur, (uint32_t)cv1);
                if (coffsetL < e->maxMatchOff && (uint32_t)cv1 == candL.val) {
                    t = candL.offset - e->cur;
                    s += checkAt;
                    break;
                }
                t = candS.offset - e->cur;
                break;
            }
            s += stepSize + ((s - nextEmit) >> (kSearchStrength - 1));
            if (s >= sLimit) goto done;
            cv = orc_ld64(src + s);
        }

        offset2 = offset1;
        offset1 = s - t;
        int32_t l = df_match_len(src, s + 4, t + 4, end) + 4;
        int32_t tMin = s - e->maxMatchOff; if (tMin < 0) tMin = 0;
        while (t > tMin && s > nextEmit && src[t - 1] == src[s - 1] && (nohist || l < ORC_ZSTD_MAX_MATCHLEN)) {
            s--; t--; l++;
        }
        df_lits(blk, src, nextEmit, s);
        orc_blockenc_add_seq(blk, (uint32_t)(s - nextEmit), (uint32_t)(l - ORC_ZSTD_MINMATCH), (uint32_t)(s - t) + 3);
        s += l;
        nextEmit = s;
        if (s >= sLimit) goto done;

        /* index match start+1 / end-2 (long) and start+2 / end-1 (short), :291-309 / :602-620 */
        int32_t index0 = s - l + 1, index1 = s - 2;
        uint64_t cv0 = orc_ld64(src + index0), cv1 = orc_ld64(src + index1);
        df_put(e->longTab, df_hash_long(cv0), index0 + e->cur, (uint32_t)cv0);
        df_put(e->longTab, df_hash_long(cv1), index1 + e->cur, (uint32_t)cv1);
        cv0 >>= 8; cv1 >>= 8;
        df_put(e->shortTab, df_hash_short(cv0), index0 + 1 + e->cur, (uint32_t)cv0);
        df_put(e->shortTab, df_hash_short(cv1), index1 + 1 + e->cur, (uint32_t)cv1);

        cv = orc_ld64(src + s);
        if (nohist ? (blk->nseq <= 2) : !canRepeatOuter) continue;

        for (;;) { /* offset-2 loop, :318-355 / :629-669 */
            int32_t o2 = s - offset2;
            if (orc_ld32(src + o2) != (uint32_t)cv) break;
            uint32_t hS = df_hash_short(nohist ? (cv1 >> 8) : cv);
            uint32_t hL = df_hash_long(cv);
            int32_t l2 = 4 + df_match_len(src, s + 4, o2 + 4, end);
            df_put(e->longTab, hL, s + e->cur, (uint32_t)cv);
            df_put(e->shortTab, hS, s + e->cur, (uint32_t)cv);
            orc_blockenc_add_seq(blk, 0, (uint32_t)l2 - ORC_ZSTD_MINMATCH, 1);
            s += l2;
            nextEmit = s;
            int32_t tmp = offset1; offset1 = offset2; offset2 = tmp;
            if (s >= sLimit) goto done;
            cv = orc_ld64(src + s);
        }
    }
done:
    if (nextEmit < end) {
        orc_blockenc_add_literals(blk, src + nextEmit, (size_t)(end - nextEmit));
        blk->extraLits = (size_t)(end - nextEmit);
    }
    if (!nohist) {
        blk->recentOffsets[0] = (uint32_t)offset1;
        blk->recentOffsets[1] = (uint32_t)offset2;
    }
}

static df_state *df_state_new(int32_t window) {
    df_state *e = (df_state *)calloc(1, sizeof(*e));
    e->maxMatchOff = window;
    e->cur = window; /* first Reset of a fresh encoder: cur += maxMatchOff, enc_base.go:183-187 */
    return e;
}

ORC_API void orc_enc_dfast_nohist(orc_blockenc *b, const uint8_t *src, size_t n) {
    df_state *e = df_state_new(8 << 20);
    df_encode_block(e, b, src, 0, (int32_t)n, 1);
    free(e);
}

/* Block loop of Encoder.EncodeAll for the level-2 encoder (zstd/encoder.go:775-825); the frame header and
 * checksum are written by the caller (orc_zstd_enc.c).  The whole input is the history buffer: for inputs
 * below window + block size the reference never slides e.hist, and beyond that the window test keeps the
 * result the same. */
/* Pooled encoder state (zstd/encoder.go:90-99 keeps encoders in a pool; Reset does not clear the tables but bumps
 * e.cur past every stored offset: fastBase.resetBase, enc_base.go:168-199). */
void *orc_dfast_state_new(void) { return df_state_new(8 << 20); }
void orc_dfast_state_free(void *st) { free(st); }
void orc_dfast_state_reset(void *st, int32_t lastLen) {
    df_state *e = (df_state *)st;
    if (e->cur >= (1 << 30) - e->maxMatchOff - lastLen) {   /* bufferReset guard: start over with cleared tables */
        memset(e->shortTab, 0, sizeof(e->shortTab));
        memset(e->longTab, 0, sizeof(e->longTab));
        e->cur = e->maxMatchOff;
    } else {
        e->cur += e->maxMatchOff + lastLen;
    }
}

void orc_dfast_encode_all_blocks_st(void *st, orc_blockenc *blk, const uint8_t *src, size_t n, size_t blockSize,
                                    uint8_t *dst, size_t cap, size_t *pos, int *err);

void orc_dfast_encode_all_blocks(orc_blockenc *blk, const uint8_t *src, size_t n, size_t blockSize,
                                 uint8_t *dst, size_t cap, size_t *pos, int *err) {
    df_state *e = df_state_new(8 << 20);
    orc_dfast_encode_all_blocks_st(e, blk, src, n, blockSize, dst, cap, pos, err);
    free(e);
}

void orc_dfast_encode_all_blocks_st(void *st, orc_blockenc *blk, const uint8_t *src, size_t n, size_t blockSize,
                                    uint8_t *dst, size_t cap, size_t *pos, int *err) {
    df_state *e = (df_state *)st;
    *err = 0;
    if (n <= blockSize) {
        orc_blockenc_reset(blk);
        blk->last = 1;
        df_encode_block(e, blk, src, 0, (int32_t)n, 1);
        *err = orc_blockenc_encode(blk, src, n, 0, 1, dst, cap, pos);
    } else {
        size_t off = 0;
        while (off < n && !*err) {
            size_t todo = n - off; if (todo > blockSize) todo = blockSize;
            memcpy(blk->prevRecentOffsets, blk->recentOffsets, sizeof(blk->recentOffsets)); /* pushOffsets */
            df_encode_block(e, blk, src, (int32_t)off, (int32_t)(off + todo), 0);
            if (off + todo == n) blk->last = 1;
            *err = orc_blockenc_encode(blk, src + off, todo, 0, 1, dst, cap, pos);
            orc_blockenc_reset(blk);
            off += todo;
        }
    }
}
