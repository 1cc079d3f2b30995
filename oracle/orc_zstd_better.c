/*
 * oracle/orc_zstd_better.c -- CPU restatement of the zstd level-3 ("better compression") match finder.
 * TEST INFRASTRUCTURE ONLY -- see orc_common.h.  Never linked into libb200comp.so.
 *
 * Follows betterFastEncoder in /root/reference/zstd/enc_better.go:
 *   constants :9-27, prevEntry :29-32, Encode :56-568, EncodeNoHist :573-576 (= ensureHist + Encode).
 * Parity unpinned at byte level (the reference holds no encoder golden vectors and Go cannot run here);
 * pinned functionally: every stream decodes with the pinned decoder oracle (tests/test_oracle_better.py).
 */
#include "orc_zstd.h"
#include <stdlib.h>
#include <string.h>

#define BT_LONG_BITS 19  /* betterLongTableBits, enc_better.go:10 */
#define BT_SHORT_BITS 13 /* betterShortTableBits, enc_better.go:18 */

typedef struct { uint32_t val; int32_t offset; } bt_short; /* tableEntry */
typedef struct { int32_t offset, prev; } bt_long;           /* prevEntry, enc_better.go:29-32 */

typedef struct {
    bt_short shortTab[1 << BT_SHORT_BITS];
    bt_long longTab[1 << BT_LONG_BITS];
    int32_t maxMatchOff;
    int32_t cur;
} bt_state;

static inline uint32_t bt_hash_long(uint64_t u) { return (uint32_t)((u * 0xcf1bbcdcb7a56463ull) >> (64 - BT_LONG_BITS)); }
static inline uint32_t bt_hash_short(uint64_t u) { return (uint32_t)(((u << 24) * 889523592379ull) >> (64 - BT_SHORT_BITS)); }

static int32_t bt_match_len(const uint8_t *src, int32_t s, int32_t t, int32_t end) { /* fastBase.matchlen, enc_base.go:110 */
    int32_t n = 0;
    while (s + n + 8 <= end) {
        uint64_t x = orc_ld64(src + s + n) ^ orc_ld64(src + t + n);
        if (x) return n + (int32_t)(__builtin_ctzll(x) >> 3);
        n += 8;
    }
    while (s + n < end && src[s + n] == src[t + n]) n++;
    return n;
}

static inline void bt_put_long(bt_state *e, uint32_t h, int32_t off) { /* {offset: off, prev: table[h].offset} */
    int32_t p = e->longTab[h].offset;
    e->longTab[h].offset = off;
    e->longTab[h].prev = p;
}
static inline void bt_put_short(bt_state *e, uint32_t h, int32_t off, uint32_t val) {
    e->shortTab[h].offset = off;
    e->shortTab[h].val = val;
}
/* the "index every second position" loops, enc_better.go:208-216, :503-512 */
static void bt_index_range(bt_state *e, const uint8_t *src, int32_t index0, int32_t until) {
    while (index0 < until) {
        uint64_t cv0 = orc_ld64(src + index0), cv1 = cv0 >> 8;
        int32_t off = index0 + e->cur;
        bt_put_long(e, bt_hash_long(cv0), off);
        bt_put_short(e, bt_hash_short(cv1), off + 1, (uint32_t)cv1);
        index0 += 2;
    }
}
static void bt_lits(orc_blockenc *b, const uint8_t *src, int32_t from, int32_t until) {
    if (until > from) orc_blockenc_add_literals(b, src + from, (size_t)(until - from));
}
/* window test + 8-byte verify of a long-table candidate at absolute position c */
static inline int bt_long_ok(const bt_state *e, const uint8_t *src, int32_t s, int32_t c, uint64_t cv) {
    return s - c < e->maxMatchOff && cv == orc_ld64(src + c);
}

/* One block [s0,end) of the history buffer src. */
static void bt_encode_block(bt_state *e, orc_blockenc *blk, const uint8_t *src, int32_t s0, int32_t end) {
    const int32_t inputMargin = 8 + 2;
    const int32_t minNonLiteralBlockSize = 16;
    const int32_t n = end - s0;
    int32_t s = s0;
    blk->size = (size_t)n;

    if (n > ORC_ZSTD_MINMATCH) { /* RLE check, :108-117: the block against itself shifted by one */
        if (bt_match_len(src, s0 + 1, s0, end) == n - 1) {
            orc_blockenc_add_literals(blk, src + s0, 1);
            orc_blockenc_add_seq(blk, 1, (uint32_t)(n - 1) - ORC_ZSTD_MINMATCH, 1 + 3);
            return;
        }
    }
    if (n < minNonLiteralBlockSize) {
        blk->extraLits = (size_t)n;
        blk->nlit = 0;
        orc_blockenc_add_literals(blk, src + s0, (size_t)n);
        return;
    }
    const int32_t sLimit = end - inputMargin;
    const int32_t stepSize = 1;
    const int kSearchStrength = 9;
    int32_t nextEmit = s;
    uint64_t cv = orc_ld64(src + s);
    int32_t offset1 = (int32_t)blk->recentOffsets[0];
    int32_t offset2 = (int32_t)blk->recentOffsets[1];

    for (;;) {
        int32_t t = 0;
        const int canRepeat = blk->nseq > 2;
        int32_t matched = 0, index0 = 0;

        for (;;) {
            uint32_t hL = bt_hash_long(cv), hS = bt_hash_short(cv);
            bt_long candL = e->longTab[hL];
            bt_short candS = e->shortTab[hS];
            const int32_t repOff = 1;
            int32_t repIndex = s - offset1 + repOff;
            int32_t off = s + e->cur;
            e->longTab[hL].offset = off; e->longTab[hL].prev = candL.offset;
            bt_put_short(e, hS, off, (uint32_t)cv);
            index0 = s + 1;

            if (canRepeat && repIndex >= 0 && orc_ld32(src + repIndex) == (uint32_t)(cv >> (repOff * 8))) { /* :170-219 */
                int32_t length = 4 + bt_match_len(src, s + 4 + repOff, repIndex + 4, end);
                uint32_t mlen = (uint32_t)(length - ORC_ZSTD_MINMATCH);
                int32_t start = s + repOff;
                int32_t startLimit = nextEmit + 1;
                int32_t tMin = s - e->maxMatchOff; if (tMin < 0) tMin = 0;
                while (repIndex > tMin && start > startLimit && src[repIndex - 1] == src[start - 1] &&
                       mlen < ORC_ZSTD_MAX_MATCHLEN - ORC_ZSTD_MINMATCH - 1) {
                    repIndex--; start--; mlen++;
                }
                uint32_t litLen = start != nextEmit ? (uint32_t)(start - nextEmit) : 0;
                bt_lits(blk, src, nextEmit, start);
                orc_blockenc_add_seq(blk, litLen, mlen, 1);
                int32_t idx = s + repOff;
                s += length + repOff;
                nextEmit = s;
                if (s >= sLimit) goto done;
                bt_index_range(e, src, idx, s - 1);
                cv = orc_ld64(src + s);
                continue;
            }
            /* (the offset-2 repeat probe at :221-268 is compiled out in the reference: `if false && ...`) */

            int32_t coffsetL = candL.offset - e->cur;
            int32_t coffsetLP = candL.prev - e->cur;
            if (bt_long_ok(e, src, s, coffsetL, cv)) { /* :275-316: long match, maybe the chained one is longer */
                matched = bt_match_len(src, s + 8, coffsetL + 8, end) + 8;
                t = coffsetL;
                if (bt_long_ok(e, src, s, coffsetLP, cv)) {
                    int32_t prevMatch = bt_match_len(src, s + 8, coffsetLP + 8, end) + 8;
                    if (prevMatch > matched) { matched = prevMatch; t = coffsetLP; }
                }
                break;
            }
            if (bt_long_ok(e, src, s, coffsetLP, cv)) { /* :319-335 */
                matched = bt_match_len(src, s + 8, coffsetLP + 8, end) + 8;
                t = coffsetLP;
                break;
            }
            int32_t coffsetS = candS.offset - e->cur;
            if (s - coffsetS < e->maxMatchOff && (uint32_t)cv == candS.val) { /* :339-400: short match + lazy long */
                matched = bt_match_len(src, s + 4, coffsetS + 4, end) + 4;
                const int32_t checkAt = 1;
                uint64_t cvn = orc_ld64(src + s + checkAt);
                uint32_t hN = bt_hash_long(cvn);
                bt_long cN = e->longTab[hN];
                int32_t cL = cN.offset - e->cur;
                e->longTab[hN].offset = s + checkAt + e->cur; e->longTab[hN].prev = cN.offset;
                if (bt_long_ok(e, src, s, cL, cvn)) {
                    int32_t matchedNext = bt_match_len(src, s + 8 + checkAt, cL + 8, end) + 8;
                    if (matchedNext > matched) { t = cL; s += checkAt; matched = matchedNext; break; }
                }
                cL = cN.prev - e->cur;
                if (bt_long_ok(e, src, s, cL, cvn)) {
                    int32_t matchedNext = bt_match_len(src, s + 8 + checkAt, cL + 8, end) + 8;
                    if (matchedNext > matched) { t = cL; s += checkAt; matched = matchedNext; break; }
                }
                t = coffsetS;
                break;
            }
            s += stepSize + ((s - nextEmit) >> (kSearchStrength - 1));
            if (s >= sLimit) goto done;
            cv = orc_ld64(src + s);
        }

        /* "match at the end of the match" probe, :419-460 */
        if (s + matched < sLimit) {
            const int32_t skipBeginning = 3;
            uint32_t hN = bt_hash_long(orc_ld64(src + s + matched));
            int32_t s2 = s + skipBeginning;
            uint32_t cv4 = orc_ld32(src + s2);
            bt_long cN = e->longTab[hN];
            int32_t cL = cN.offset - e->cur - matched + skipBeginning;
            if (cL >= 0 && cL < s2 && s2 - cL < e->maxMatchOff && cv4 == orc_ld32(src + cL)) {
                int32_t matchedNext = bt_match_len(src, s2 + 4, cL + 4, end) + 4;
                if (matchedNext > matched) { t = cL; s = s2; matched = matchedNext; }
            }
            cL = cN.prev - e->cur - matched + skipBeginning;
            if (cL >= 0 && cL < s2 && s2 - cL < e->maxMatchOff && cv4 == orc_ld32(src + cL)) {
                int32_t matchedNext = bt_match_len(src, s2 + 4, cL + 4, end) + 4;
                if (matchedNext > matched) { t = cL; s = s2; matched = matchedNext; }
            }
        }
        offset2 = offset1;
        offset1 = s - t;
        int32_t l = matched;
        int32_t tMin = s - e->maxMatchOff; if (tMin < 0) tMin = 0;
        while (t > tMin && s > nextEmit && src[t - 1] == src[s - 1] && l < ORC_ZSTD_MAX_MATCHLEN) { s--; t--; l++; }
        bt_lits(blk, src, nextEmit, s);
        orc_blockenc_add_seq(blk, (uint32_t)(s - nextEmit), (uint32_t)(l - ORC_ZSTD_MINMATCH), (uint32_t)(s - t) + 3);
        s += l;
        nextEmit = s;
        if (s >= sLimit) goto done;

        bt_index_range(e, src, index0, s - 1); /* :503-512 */
        cv = orc_ld64(src + s);
        if (!canRepeat) continue;

        for (;;) { /* offset-2 loop, :520-557 */
            int32_t o2 = s - offset2;
            if (orc_ld32(src + o2) != (uint32_t)cv) break;
            uint32_t hL = bt_hash_long(cv), hS = bt_hash_short(cv);
            int32_t l2 = 4 + bt_match_len(src, s + 4, o2 + 4, end);
            bt_put_long(e, hL, s + e->cur);
            bt_put_short(e, hS, s + e->cur, (uint32_t)cv);
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
    blk->recentOffsets[0] = (uint32_t)offset1;
    blk->recentOffsets[1] = (uint32_t)offset2;
}

/* Block loop of Encoder.EncodeAll for the level-3 encoder (zstd/encoder.go:775-825); frame header and checksum
 * are written by the caller.  allLitEntropy is on above SpeedDefault (encoder_options.go:262), so blockEnc.encode
 * runs with rawAllLits = false. */
void orc_better_encode_all_blocks(orc_blockenc *blk, const uint8_t *src, size_t n, size_t blockSize,
                                  uint8_t *dst, size_t cap, size_t *pos, int *err) {
    bt_state *e = (bt_state *)calloc(1, sizeof(*e));
    e->maxMatchOff = 8 << 20;
    e->cur = e->maxMatchOff; /* first Reset of a fresh encoder, enc_base.go:183-187 */
    *err = 0;
    if (n <= blockSize) {
        orc_blockenc_reset(blk);
        blk->last = 1;
        bt_encode_block(e, blk, src, 0, (int32_t)n);
        *err = orc_blockenc_encode(blk, src, n, 0, 0, dst, cap, pos);
    } else {
        size_t off = 0;
        while (off < n && !*err) {
            size_t todo = n - off; if (todo > blockSize) todo = blockSize;
            memcpy(blk->prevRecentOffsets, blk->recentOffsets, sizeof(blk->recentOffsets)); /* pushOffsets */
            bt_encode_block(e, blk, src, (int32_t)off, (int32_t)(off + todo));
            if (off + todo == n) blk->last = 1;
            *err = orc_blockenc_encode(blk, src + off, todo, 0, 0, dst, cap, pos);
            orc_blockenc_reset(blk);
            off += todo;
        }
    }
    free(e);
}
