/*
 * oracle/orc_zstd_enc.c -- zstd block + frame encoder, restating
 *   zstd/blockenc.go (blockEnc.encode/encodeLits/encodeRLE/genCodes, headers)
 *   zstd/seqenc.go (llCode/mlCode/ofCode, bit tables, seqCoders.setPrev)
 *   zstd/fse_encoder.go (optimalTableLog, normalizeCount, approxSize, bitCost, setRLE, setBits)
 *   zstd/fse_predefined.go (default distributions)
 *   zstd/enc_fast.go (fastEncoder.Encode / EncodeNoHist), zstd/hash.go, zstd/matchlen_generic.go
 *   zstd/frameenc.go (frameHeader.appendTo), zstd/encoder.go:731-873 (encodeAll, MaxEncodedSize)
 * TEST INFRASTRUCTURE ONLY -- see orc_common.h.  Encode byte-parity with the Go
 * binary is "parity unpinned" (no golden compressed bytes exist upstream).
 */
#include <stdlib.h>
#include "orc_zstd.h"

/* ------------------------------------------------------------------ codes */
static const uint8_t llCodeTable[64] = {0,  1,  2,  3,  4,  5,  6,  7,  8,  9,  10, 11, 12, 13, 14, 15,
                                        16, 16, 17, 17, 18, 18, 19, 19, 20, 20, 20, 20, 21, 21, 21, 21,
                                        22, 22, 22, 22, 22, 22, 22, 22, 23, 23, 23, 23, 23, 23, 23, 23,
                                        24, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24};
static const uint8_t llBitsTable[36] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,  0,  0,  0,  0,  0,  1,  1,
                                        1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
static const uint8_t mlCodeTable[128] = {
    0,  1,  2,  3,  4,  5,  6,  7,  8,  9,  10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25,
    26, 27, 28, 29, 30, 31, 32, 32, 33, 33, 34, 34, 35, 35, 36, 36, 36, 36, 37, 37, 37, 37, 38, 38, 38, 38,
    38, 38, 38, 38, 39, 39, 39, 39, 39, 39, 39, 39, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40,
    40, 40, 41, 41, 41, 41, 41, 41, 41, 41, 41, 41, 41, 41, 41, 41, 41, 41, 42, 42, 42, 42, 42, 42, 42, 42,
    42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42};
static const uint8_t mlBitsTable[53] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,  0,  0,  0,  0,  0, 0,
                                        0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,  0,  0,  1,  1,  1, 1,
                                        2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};

static inline uint8_t ll_code(uint32_t litLength) { /* seqenc.go:69-75 */
    if (litLength <= 63) return llCodeTable[litLength & 63];
    return (uint8_t)(orc_highbit32(litLength) + 19);
}
static inline uint8_t ml_code(uint32_t mlBase) { /* seqenc.go:101-107 */
    if (mlBase <= 127) return mlCodeTable[mlBase & 127];
    return (uint8_t)(orc_highbit32(mlBase) + 36);
}
static inline uint8_t of_code(uint32_t offset) { return (uint8_t)orc_highbit32(offset); } /* seqenc.go:109 */

/* ------------------------------------------------------------ fseEncoder */
static void fenc_set_bits(orc_fse_enc *s, const uint8_t *transform) { /* fse_encoder.go:225-255 */
    if (s->reUsed || s->preDefined) return;
    if (s->useRLE) {
        if (!transform) { s->ct.tt[s->rleVal].outBits = s->rleVal; s->maxBits = s->rleVal; return; }
        s->maxBits = transform[s->rleVal];
        s->ct.tt[s->rleVal].outBits = s->maxBits;
        return;
    }
    if (!transform) {
        for (unsigned i = 0; i < s->symbolLen; i++) s->ct.tt[i].outBits = (uint8_t)i;
        s->maxBits = (uint8_t)(s->symbolLen - 1);
        return;
    }
    s->maxBits = 0;
    for (unsigned i = 0; i < s->symbolLen; i++) {
        s->ct.tt[i].outBits = transform[i];
        if (transform[i] > s->maxBits) s->maxBits = transform[i];
    }
}

static orc_fse_enc predefEnc[3];
static int predefReady = 0;
enum { T_LL = 0, T_OF = 1, T_ML = 2 };

static void init_predef(void) { /* fse_predefined.go:75-158 */
    if (predefReady) return;
    static const int16_t llN[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2,
                                    2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
    static const int16_t ofN[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1,
                                    1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};
    static const int16_t mlN[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                                    1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                                    1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
    const int16_t *norms[3] = {llN, ofN, mlN};
    const unsigned lens[3] = {36, 29, 53};
    const unsigned logs[3] = {6, 5, 6};
    const uint8_t *bitsT[3] = {llBitsTable, NULL, mlBitsTable};
    for (int i = 0; i < 3; i++) {
        orc_fse_enc *e = &predefEnc[i];
        memset(e, 0, sizeof(*e));
        memcpy(e->norm, norms[i], lens[i] * sizeof(int16_t));
        e->symbolLen = lens[i];
        e->actualTableLog = logs[i];
        orc_fse_build_ctable(e->norm, e->symbolLen, e->actualTableLog, &e->ct);
        fenc_set_bits(e, bitsT[i]);
        e->preDefined = 1;
    }
    predefReady = 1;
}

static void fenc_optimal_tablelog(orc_fse_enc *s, int length) { /* fse_encoder.go:429-455 */
    uint8_t tableLog = 8; /* maxEncTableLog */
    uint32_t minBitsSrc = orc_highbit32((uint32_t)length) + 1;
    uint32_t minBitsSymbols = orc_highbit32((uint32_t)(s->symbolLen - 1)) + 2;
    uint8_t minBits = (uint8_t)minBitsSymbols;
    if (minBitsSrc < minBitsSymbols) minBits = (uint8_t)minBitsSrc;
    uint8_t maxBitsSrc = (uint8_t)((uint8_t)orc_highbit32((uint32_t)(length - 1)) - 2);
    if (maxBitsSrc < tableLog) tableLog = maxBitsSrc;
    if (minBits > tableLog) tableLog = minBits;
    if (tableLog < 5) tableLog = 5;
    if (tableLog > 8) tableLog = 8;
    s->actualTableLog = tableLog;
}

static int fenc_normalize_count(orc_fse_enc *s, int length) { /* fse_encoder.go:259-330 */
    if (s->reUsed) return 0;
    fenc_optimal_tablelog(s, length);
    if (s->maxCount == length) { s->useRLE = 1; return 0; }
    s->useRLE = 0;
    int err = orc_fse_normalize(s->count, s->symbolLen, (uint32_t)length, s->actualTableLog, s->norm);
    if (err) return err;
    return orc_fse_build_ctable(s->norm, s->symbolLen, s->actualTableLog, &s->ct);
}

static void fenc_set_rle(orc_fse_enc *s, uint8_t val) { /* fse_encoder.go:208-221 */
    s->actualTableLog = 0;
    s->ct.tt[val].deltaFindState = 0;
    s->ct.tt[val].deltaNbBits = 0;
    s->ct.tt[val].outBits = 0;
    s->ct.stateTable[0] = 0;
    s->rleVal = val;
    s->useRLE = 1;
}

static uint32_t fenc_bit_cost(const orc_fse_enc *s, uint8_t sym, uint32_t accuracyLog) { /* :603-628 */
    uint32_t minNbBits = s->ct.tt[sym].deltaNbBits >> 16;
    uint32_t threshold = (minNbBits + 1) << 16;
    uint32_t tableSize = 1u << s->actualTableLog;
    uint32_t deltaFromThreshold = threshold - (s->ct.tt[sym].deltaNbBits + tableSize);
    uint32_t normalizedDelta = (deltaFromThreshold << accuracyLog) >> s->actualTableLog;
    uint32_t bitMultiplier = 1u << accuracyLog;
    return (minNbBits + 1) * bitMultiplier - normalizedDelta;
}

static uint32_t fenc_approx_size(const orc_fse_enc *s, const uint32_t *hist, unsigned histLen) { /* :634-660 */
    if (s->symbolLen < histLen) return 0xffffffffu;
    if (s->useRLE) return 0xffffffffu;
    const uint32_t kAccuracyLog = 8;
    uint32_t badCost = ((uint32_t)s->actualTableLog + 1) << kAccuracyLog;
    uint32_t cost = 0;
    for (unsigned i = 0; i < histLen; i++) {
        if (hist[i] == 0) continue;
        if (s->norm[i] == 0) return 0xffffffffu;
        uint32_t bc = fenc_bit_cost(s, (uint8_t)i, kAccuracyLog);
        if (bc > badCost) return 0xffffffffu;
        cost += hist[i] * bc;
    }
    return cost >> kAccuracyLog;
}

static uint32_t fenc_max_header_size(const orc_fse_enc *s) { /* :664-672 */
    if (s->preDefined) return 0;
    if (s->useRLE) return 8;
    return ((((uint32_t)s->symbolLen * (uint32_t)s->actualTableLog) >> 3) + 3) * 8;
}

static int64_t fenc_write_count(const orc_fse_enc *s, uint8_t *out, size_t cap) { /* :488-598 */
    if (s->useRLE) {
        if (cap < 1) return ORC_ERR_DST_SMALL;
        out[0] = s->rleVal;
        return 1;
    }
    if (s->preDefined || s->reUsed) return 0;
    return orc_fse_write_ncount(s->norm, s->symbolLen, s->actualTableLog, out, cap);
}

static uint16_t fenc_cstate_init(const orc_fse_enc *e, orc_symtt first) { /* cState.init :676-690 */
    if (e->useRLE) return 0; /* len(stateTable)==1 */
    return orc_fse_cstate_init(&e->ct, first);
}

/* -------------------------------------------------------------- blockEnc */
ORC_API orc_blockenc *orc_blockenc_new(void) {
    init_predef();
    orc_blockenc *b = (orc_blockenc *)calloc(1, sizeof(*b));
    if (!b) return NULL;
    b->lit_cap = ORC_ZSTD_MAX_BLOCK + 64;
    b->literals = (uint8_t *)malloc(b->lit_cap);
    b->tmp = (uint8_t *)malloc(ORC_ZSTD_MAX_BLOCK + 1024);
    b->seq_cap = 2000;
    b->seqs = (orc_seq *)malloc(b->seq_cap * sizeof(orc_seq));
    b->llEnc = &b->store[0]; b->llPrev = &b->store[1];
    b->ofEnc = &b->store[2]; b->ofPrev = &b->store[3];
    b->mlEnc = &b->store[4]; b->mlPrev = &b->store[5];
    orc_huf_scratch_init(&b->litEnc, 4, ORC_HUF_REUSE_NONE); /* WantLogLess: 4, blockenc.go:72 */
    orc_blockenc_init_new_encode(b);
    return b;
}
ORC_API void orc_blockenc_free(orc_blockenc *b) {
    if (!b) return;
    free(b->literals); free(b->seqs); free(b->tmp); free(b);
}

static void compare_swap(orc_fse_enc *used, orc_fse_enc **current, orc_fse_enc **prev) { /* seqenc.go:22-38 */
    if (*current == used) {
        orc_fse_enc *t = *prev; *prev = *current; *current = t;
        (*current)->reUsed = 0;
        (*prev)->reUsed = 1;
        return;
    }
    if (used == *prev) return;
    (*prev)->symbolLen = 0;
}
static void coders_set_prev(orc_blockenc *b, orc_fse_enc *ll, orc_fse_enc *ml, orc_fse_enc *of) {
    compare_swap(ll, &b->llEnc, &b->llPrev);
    compare_swap(ml, &b->mlEnc, &b->mlPrev);
    compare_swap(of, &b->ofEnc, &b->ofPrev);
}

ORC_API void orc_blockenc_init_new_encode(orc_blockenc *b) {
    b->recentOffsets[0] = 1; b->recentOffsets[1] = 4; b->recentOffsets[2] = 8;
    b->litEnc.reuse = ORC_HUF_REUSE_NONE;
    coders_set_prev(b, NULL, NULL, NULL);
}
ORC_API void orc_blockenc_reset(orc_blockenc *b) {
    b->extraLits = 0; b->nlit = 0; b->size = 0; b->nseq = 0; b->last = 0;
}
ORC_API void orc_blockenc_add_literals(orc_blockenc *b, const uint8_t *p, size_t n) {
    if (b->nlit + n > b->lit_cap) {
        b->lit_cap = (b->nlit + n) * 2;
        b->literals = (uint8_t *)realloc(b->literals, b->lit_cap);
    }
    memcpy(b->literals + b->nlit, p, n);
    b->nlit += n;
}
ORC_API void orc_blockenc_add_seq(orc_blockenc *b, uint32_t litLen, uint32_t matchLenMinus3, uint32_t offset) {
    if (b->nseq == b->seq_cap) {
        b->seq_cap *= 2;
        b->seqs = (orc_seq *)realloc(b->seqs, b->seq_cap * sizeof(orc_seq));
    }
    orc_seq s = {litLen, matchLenMinus3, offset, 0, 0, 0};
    b->seqs[b->nseq++] = s;
}

#define PUT(bytep, len_)                                                                               \
    do {                                                                                               \
        if (*pos + (len_) > cap) return ORC_ERR_DST_SMALL;                                             \
        memcpy(dst + *pos, (bytep), (len_));                                                           \
        *pos += (len_);                                                                                \
    } while (0)

static inline uint32_t block_header(int last, unsigned type, uint32_t size) { /* blockenc.go:109-136 */
    return (last ? 1u : 0u) | (type << 1) | (size << 3);
}
static int put_block_header(uint8_t *dst, size_t cap, size_t *pos, uint32_t bh) {
    uint8_t h[3] = {(uint8_t)bh, (uint8_t)(bh >> 8), (uint8_t)(bh >> 16)};
    PUT(h, 3);
    return 0;
}

typedef struct { uint64_t v; int size; } lit_hdr;
static lit_hdr lh_set_size(unsigned type, int regenLen) { /* literalsHeader.setSize, blockenc.go:153-175 */
    lit_hdr h;
    uint64_t lh = type & 3;
    int inBits = regenLen ? (int)orc_highbit32((uint32_t)regenLen) + 1 : 0;
    if (inBits < 5) { lh |= ((uint64_t)regenLen << 3); h.size = 1; }
    else if (inBits < 12) { lh |= (1 << 2) | ((uint64_t)regenLen << 4); h.size = 2; }
    else { lh |= (3 << 2) | ((uint64_t)regenLen << 4); h.size = 3; }
    h.v = lh;
    return h;
}
static lit_hdr lh_set_sizes(unsigned type, int compLen, int inLen, int single) { /* setSizes, :178-213 */
    lit_hdr h;
    uint64_t lh = type & 3;
    int compBits = compLen ? (int)orc_highbit32((uint32_t)compLen) + 1 : 0;
    int inBits = inLen ? (int)orc_highbit32((uint32_t)inLen) + 1 : 0;
    if (compBits <= 10 && inBits <= 10) {
        if (!single) lh |= 1 << 2;
        lh |= ((uint64_t)inLen << 4) | ((uint64_t)compLen << (10 + 4));
        h.size = 3;
    } else if (compBits <= 14 && inBits <= 14) {
        lh |= (2 << 2) | ((uint64_t)inLen << 4) | ((uint64_t)compLen << (14 + 4));
        h.size = 4;
    } else {
        lh |= (3 << 2) | ((uint64_t)inLen << 4) | ((uint64_t)compLen << (18 + 4));
        h.size = 5;
    }
    h.v = lh;
    return h;
}
static int put_lit_hdr(uint8_t *dst, size_t cap, size_t *pos, lit_hdr h) {
    uint8_t b[5];
    for (int i = 0; i < 5; i++) b[i] = (uint8_t)(h.v >> (8 * i));
    PUT(b, (size_t)h.size);
    return 0;
}

/* encodeLits, blockenc.go:337-430 */
static int encode_lits(orc_blockenc *b, const uint8_t *lits, size_t n, int raw, uint8_t *dst, size_t cap,
                       size_t *pos) {
    if (n < 8 || n < 32 /* dictLitEnc == nil */ || raw) {
        int e = put_block_header(dst, cap, pos, block_header(b->last, 0, (uint32_t)n));
        if (e) return e;
        PUT(lits, n);
        return 0;
    }
    uint8_t *tmp = b->tmp;
    int reUsed = 0, single = 0;
    int64_t out;
    if (n >= 1024) out = orc_huf_compress(&b->litEnc, lits, n, 1, tmp, n + 512, &reUsed);
    else if (n > 16) { single = 1; out = orc_huf_compress(&b->litEnc, lits, n, 0, tmp, n + 512, &reUsed); }
    else out = ORC_ERR_INCOMPRESSIBLE;
    if (out == ORC_ERR_DST_SMALL) out = ORC_ERR_INCOMPRESSIBLE; /* bigger than input => incompressible */
    if (out >= 0 && (size_t)out + 5 > n) {
        lit_hdr lh = lh_set_sizes(0, (int)out, (int)n, single);
        if ((size_t)out + (size_t)lh.size >= n) out = ORC_ERR_INCOMPRESSIBLE;
    }
    int e = 0;
    if (out == ORC_ERR_INCOMPRESSIBLE) {
        e = put_block_header(dst, cap, pos, block_header(b->last, 0, (uint32_t)n));
        if (!e) { if (*pos + n > cap) e = ORC_ERR_DST_SMALL; else { memcpy(dst + *pos, lits, n); *pos += n; } }
    } else if (out == ORC_ERR_USE_RLE) {
        e = put_block_header(dst, cap, pos, block_header(b->last, 1, (uint32_t)n));
        if (!e) { if (*pos + 1 > cap) e = ORC_ERR_DST_SMALL; else dst[(*pos)++] = lits[0]; }
    } else if (out < 0) {
        e = (int)out;
    } else {
        b->litEnc.reuse = ORC_HUF_REUSE_ALLOW;
        lit_hdr lh = lh_set_sizes(reUsed ? 3 : 2, (int)out, (int)n, single);
        e = put_block_header(dst, cap, pos, block_header(b->last, 2, (uint32_t)((size_t)out + (size_t)lh.size + 1)));
        if (!e) e = put_lit_hdr(dst, cap, pos, lh);
        if (!e) {
            if (*pos + (size_t)out + 1 > cap) e = ORC_ERR_DST_SMALL;
            else { memcpy(dst + *pos, tmp, (size_t)out); *pos += (size_t)out; dst[(*pos)++] = 0; }
        }
    }
    return e;
}

/* genCodes, blockenc.go:831-893 */
static void gen_codes(orc_blockenc *b) {
    uint32_t *llH = b->llEnc->count, *ofH = b->ofEnc->count, *mlH = b->mlEnc->count;
    memset(llH, 0, 256 * 4); memset(ofH, 0, 256 * 4); memset(mlH, 0, 256 * 4);
    uint8_t llMax = 0, ofMax = 0, mlMax = 0;
    for (size_t i = 0; i < b->nseq; i++) {
        orc_seq *s = &b->seqs[i];
        uint8_t v = ll_code(s->litLen); s->llCode = v; llH[v]++; if (v > llMax) llMax = v;
        v = of_code(s->offset); s->ofCode = v; ofH[v]++; if (v > ofMax) ofMax = v;
        v = ml_code(s->matchLen); s->mlCode = v; mlH[v]++; if (v > mlMax) mlMax = v;
    }
#define FINISH(enc, H, mx)                                                                             \
    do {                                                                                               \
        uint32_t m_ = 0;                                                                               \
        for (unsigned k = 0; k <= (mx); k++) if ((H)[k] > m_) m_ = (H)[k];                             \
        (enc)->maxCount = (int)m_; (enc)->symbolLen = (unsigned)(mx) + 1;                              \
    } while (0)
    FINISH(b->mlEnc, mlH, mlMax);
    FINISH(b->ofEnc, ofH, ofMax);
    FINISH(b->llEnc, llH, llMax);
#undef FINISH
}

/* chooseComp closure, blockenc.go:633-661 */
static orc_fse_enc *choose_comp(orc_fse_enc *cur, orc_fse_enc *prev, orc_fse_enc *preDef, unsigned *mode) {
    const uint32_t *hist = cur->count;
    unsigned hl = cur->symbolLen;
    uint32_t nSize = fenc_approx_size(cur, hist, hl) + fenc_max_header_size(cur);
    uint32_t predefSize = fenc_approx_size(preDef, hist, hl);
    uint32_t prevSize = fenc_approx_size(prev, hist, hl);
    nSize = nSize + ((nSize + 2 * 8 * 16) >> 4);
    if (predefSize <= prevSize && predefSize <= nSize) { *mode = 0; return preDef; }
    if (prevSize <= nSize) { *mode = 3; return prev; }
    *mode = 2;
    return cur;
}

ORC_API int orc_blockenc_encode(orc_blockenc *b, const uint8_t *org, size_t orgLen, int raw, int rawAllLits,
                                uint8_t *dst, size_t cap, size_t *pos) {
    init_predef();
    if (b->nseq == 0) return encode_lits(b, b->literals, b->nlit, rawAllLits, dst, cap, pos);
    if (b->nseq == 1 && orgLen > 0 && b->nlit <= 1) {
        orc_seq s = b->seqs[0];
        if (s.litLen == (uint32_t)b->nlit && s.offset - 3 == 1) {
            /* encodeRLE, blockenc.go:433-440 */
            int e = put_block_header(dst, cap, pos, block_header(b->last, 1, s.matchLen + ORC_ZSTD_MINMATCH + s.litLen));
            if (e) return e;
            PUT(org, 1);
            return 0;
        }
    }
    int saved = (int)b->size - (int)b->nlit - (int)(b->size >> 6);
    if (saved < 16) {
        if (org == NULL) return ORC_ERR_INCOMPRESSIBLE;
        memcpy(b->recentOffsets, b->prevRecentOffsets, sizeof(b->recentOffsets)); /* popOffsets */
        return encode_lits(b, org, orgLen, rawAllLits, dst, cap, pos);
    }
    size_t bhOffset = *pos;
    {
        int e = put_block_header(dst, cap, pos, block_header(b->last, 2, 0));
        if (e) return e;
    }
    /* literals */
    {
        size_t n = b->nlit;
        uint8_t *tmp = b->tmp;
        int reUsed = 0, single = 0;
        int64_t out;
        if (n >= 1024 && !raw) out = orc_huf_compress(&b->litEnc, b->literals, n, 1, tmp, n + 512, &reUsed);
        else if (n > 16 && !raw) { single = 1; out = orc_huf_compress(&b->litEnc, b->literals, n, 0, tmp, n + 512, &reUsed); }
        else out = ORC_ERR_INCOMPRESSIBLE;
        if (out == ORC_ERR_DST_SMALL) out = ORC_ERR_INCOMPRESSIBLE;
        if (out >= 0 && (size_t)out + 5 > n) {
            lit_hdr r = lh_set_size(0, (int)n);
            lit_hdr c = lh_set_sizes(0, (int)out, (int)n, single);
            if ((size_t)out + (size_t)c.size >= n + (size_t)r.size) out = ORC_ERR_INCOMPRESSIBLE;
        }
        int e = 0;
        if (out == ORC_ERR_INCOMPRESSIBLE) {
            e = put_lit_hdr(dst, cap, pos, lh_set_size(0, (int)n));
            if (!e) { if (*pos + n > cap) e = ORC_ERR_DST_SMALL; else { memcpy(dst + *pos, b->literals, n); *pos += n; } }
        } else if (out == ORC_ERR_USE_RLE) {
            e = put_lit_hdr(dst, cap, pos, lh_set_size(1, (int)n));
            if (!e) { if (*pos + 1 > cap) e = ORC_ERR_DST_SMALL; else dst[(*pos)++] = b->literals[0]; }
        } else if (out < 0) {
            e = (int)out;
        } else {
            e = put_lit_hdr(dst, cap, pos, lh_set_sizes(reUsed ? 3 : 2, (int)out, (int)n, single));
            if (!e) { if (*pos + (size_t)out > cap) e = ORC_ERR_DST_SMALL; else { memcpy(dst + *pos, tmp, (size_t)out); *pos += (size_t)out; } }
            b->litEnc.reuse = ORC_HUF_REUSE_ALLOW;
        }
        if (e) return e;
    }
    /* number of sequences, blockenc.go:601-610 */
    {
        uint8_t h[3];
        size_t n = b->nseq;
        if (n < 128) { h[0] = (uint8_t)n; PUT(h, 1); }
        else if (n < 0x7f00) { h[0] = (uint8_t)(128 + (n >> 8)); h[1] = (uint8_t)n; PUT(h, 2); }
        else { n -= 0x7f00; h[0] = 255; h[1] = (uint8_t)n; h[2] = (uint8_t)(n >> 8); PUT(h, 3); }
    }
    gen_codes(b);
    orc_fse_enc *llEnc = b->llEnc, *ofEnc = b->ofEnc, *mlEnc = b->mlEnc;
    int err;
    if ((err = fenc_normalize_count(llEnc, (int)b->nseq))) return err;
    if ((err = fenc_normalize_count(ofEnc, (int)b->nseq))) return err;
    if ((err = fenc_normalize_count(mlEnc, (int)b->nseq))) return err;

    uint8_t mode = 0;
    unsigned m;
    if (llEnc->useRLE) { mode |= 1 << 6; fenc_set_rle(llEnc, b->seqs[0].llCode); }
    else { llEnc = choose_comp(llEnc, b->llPrev, &predefEnc[T_LL], &m); mode |= (uint8_t)(m << 6); }
    if (ofEnc->useRLE) { mode |= 1 << 4; fenc_set_rle(ofEnc, b->seqs[0].ofCode); }
    else { ofEnc = choose_comp(ofEnc, b->ofPrev, &predefEnc[T_OF], &m); mode |= (uint8_t)(m << 4); }
    if (mlEnc->useRLE) { mode |= 1 << 2; fenc_set_rle(mlEnc, b->seqs[0].mlCode); }
    else { mlEnc = choose_comp(mlEnc, b->mlPrev, &predefEnc[T_ML], &m); mode |= (uint8_t)(m << 2); }
    PUT(&mode, 1);
    {
        int64_t w;
        if ((w = fenc_write_count(llEnc, dst + *pos, cap - *pos)) < 0) return (int)w; *pos += (size_t)w;
        if ((w = fenc_write_count(ofEnc, dst + *pos, cap - *pos)) < 0) return (int)w; *pos += (size_t)w;
        if ((w = fenc_write_count(mlEnc, dst + *pos, cap - *pos)) < 0) return (int)w; *pos += (size_t)w;
    }
    /* sequence bitstream, blockenc.go:725-808 */
    orc_bw wr;
    orc_bw_init(&wr, dst + *pos, cap - *pos);
    fenc_set_bits(llEnc, llBitsTable);
    fenc_set_bits(mlEnc, mlBitsTable);
    fenc_set_bits(ofEnc, NULL);
    const orc_symtt *llTT = llEnc->ct.tt, *ofTT = ofEnc->ct.tt, *mlTT = mlEnc->ct.tt;
    int64_t seq = (int64_t)b->nseq - 1;
    orc_seq s = b->seqs[seq];
    orc_symtt llB = llTT[s.llCode], ofB = ofTT[s.ofCode], mlB = mlTT[s.mlCode];
    uint16_t ll = fenc_cstate_init(llEnc, llB);
    uint16_t of = fenc_cstate_init(ofEnc, ofB);
    uint16_t ml = fenc_cstate_init(mlEnc, mlB);
    orc_bw_add(&wr, s.litLen, llB.outBits & 31);
    orc_bw_add(&wr, s.matchLen, mlB.outBits & 31);
    orc_bw_add(&wr, s.offset, ofB.outBits & 31);
    seq--;
    while (seq >= 0) {
        s = b->seqs[seq];
        ofB = ofTT[s.ofCode];
        uint32_t nb = ((uint32_t)of + ofB.deltaNbBits) >> 16;
        int32_t ds = (int32_t)(of >> (nb & 15)) + ofB.deltaFindState;
        orc_bw_add(&wr, of, nb);
        of = ofEnc->ct.stateTable[ds];
        unsigned outBits = ofB.outBits & 31;
        uint64_t extraBits = (uint64_t)(s.offset & (outBits ? ((1ull << outBits) - 1) : 0));
        unsigned extraBitsN = outBits;

        mlB = mlTT[s.mlCode];
        nb = ((uint32_t)ml + mlB.deltaNbBits) >> 16;
        ds = (int32_t)(ml >> (nb & 15)) + mlB.deltaFindState;
        orc_bw_add(&wr, ml, nb);
        ml = mlEnc->ct.stateTable[ds];
        outBits = mlB.outBits & 31;
        extraBits = (extraBits << outBits) | (uint64_t)(s.matchLen & (outBits ? ((1ull << outBits) - 1) : 0));
        extraBitsN += outBits;

        llB = llTT[s.llCode];
        nb = ((uint32_t)ll + llB.deltaNbBits) >> 16;
        ds = (int32_t)(ll >> (nb & 15)) + llB.deltaFindState;
        orc_bw_add(&wr, ll, nb);
        ll = llEnc->ct.stateTable[ds];
        outBits = llB.outBits & 31;
        extraBits = (extraBits << outBits) | (uint64_t)(s.litLen & (outBits ? ((1ull << outBits) - 1) : 0));
        extraBitsN += outBits;

        orc_bw_add64(&wr, extraBits, extraBitsN);
        seq--;
    }
    orc_bw_add(&wr, ml, mlEnc->actualTableLog);
    orc_bw_add(&wr, of, ofEnc->actualTableLog);
    orc_bw_add(&wr, ll, llEnc->actualTableLog);
    orc_bw_close(&wr);
    if (wr.overflow) {
        /* the compressed form no longer fits: that can only happen when it is
         * larger than the raw block, so take the raw path below. */
        *pos = cap + 1;
    } else {
        *pos += wr.pos;
    }
    if (*pos > cap || *pos - 3 - bhOffset >= b->size) {
        /* Discard and encode as raw block, blockenc.go:811-817 */
        *pos = bhOffset;
        int e = put_block_header(dst, cap, pos, block_header(b->last, 0, (uint32_t)orgLen));
        if (e) return e;
        PUT(org, orgLen);
        memcpy(b->recentOffsets, b->prevRecentOffsets, sizeof(b->recentOffsets));
        b->litEnc.reuse = ORC_HUF_REUSE_NONE;
        return 0;
    }
    {
        uint32_t bh = block_header(b->last, 2, (uint32_t)(*pos - bhOffset) - 3);
        dst[bhOffset] = (uint8_t)bh; dst[bhOffset + 1] = (uint8_t)(bh >> 8); dst[bhOffset + 2] = (uint8_t)(bh >> 16);
    }
    coders_set_prev(b, llEnc, mlEnc, ofEnc);
    return 0;
}

/* One-shot helper for entropy-stage parity tests: fresh blockEnc (as after
 * Reset(nil,true)), literals + (litLen, matchLen-3, offset) triples -> block bytes. */
ORC_API int64_t orc_zstd_encode_block(const uint8_t *org, size_t orgLen, const uint8_t *lits, size_t nlits,
                                      const uint32_t *seqTriples, size_t nseq, int last, uint8_t *dst,
                                      size_t cap) {
    orc_blockenc *b = orc_blockenc_new();
    if (!b) return ORC_ERR_INTERNAL;
    orc_blockenc_reset(b);
    b->size = orgLen;
    b->last = last;
    orc_blockenc_add_literals(b, lits, nlits);
    for (size_t i = 0; i < nseq; i++)
        orc_blockenc_add_seq(b, seqTriples[3 * i], seqTriples[3 * i + 1], seqTriples[3 * i + 2]);
    size_t pos = 0;
    int e = orc_blockenc_encode(b, org, orgLen, 0, 1, dst, cap, &pos);
    orc_blockenc_free(b);
    return e ? e : (int64_t)pos;
}

/* ------------------------------------------------------------ match finder */
static const uint64_t prime5bytes = 889523592379ull;
static const uint64_t prime6bytes = 227718039650203ull;
static const uint64_t prime8bytes = 0xcf1bbcdcb7a56463ull;
static inline uint32_t hash6(uint64_t u, unsigned bits) { return (uint32_t)(((u << 16) * prime6bytes) >> (64 - bits)); }
static inline uint32_t hash5(uint64_t u, unsigned bits) { return (uint32_t)(((u << 24) * prime5bytes) >> (64 - bits)); }
static inline uint32_t hash8(uint64_t u, unsigned bits) { return (uint32_t)((u * prime8bytes) >> (64 - bits)); }

/* load6432 at the buffer tail: the Go code slices src[i:i+8] so it may never read
 * past len(src); sLimit = len-8 guarantees that for every call below. */

/* matchLen(a=src[s:end], b=src[t:]) -- zstd/matchlen_generic.go:16 */
static inline int32_t match_len(const uint8_t *src, int32_t s, int32_t t, int32_t end) {
    int32_t n = 0;
    while (s + n + 8 <= end) {
        uint64_t diff = orc_ld64(src + s + n) ^ orc_ld64(src + t + n);
        if (diff) return n + (int32_t)(__builtin_ctzll(diff) >> 3);
        n += 8;
    }
    while (s + n < end && src[s + n] == src[t + n]) n++;
    return n;
}

typedef struct { uint32_t val; int32_t offset; } tentry;
#define FAST_TABLE_BITS 15
#define FAST_TABLE_SIZE (1 << FAST_TABLE_BITS)

typedef struct {
    tentry table[FAST_TABLE_SIZE];
    int32_t maxMatchOff;
    int32_t cur; /* fastBase.cur: table offsets are stored as position + cur so stale entries fall out of the window */
} fast_state;

static void add_lits(orc_blockenc *b, const uint8_t *src, int32_t from, int32_t until) {
    if (until > from) orc_blockenc_add_literals(b, src + from, (size_t)(until - from));
}

/* fastEncoder.Encode / EncodeNoHist (zstd/enc_fast.go:39-289, 294-531).
 * src = whole history buffer (e.hist), block = [s0, end).  With nohist the
 * block starts at 0 and the repeat test uses len(sequences) > 2 without the
 * repIndex >= 0 guard (enc_fast.go:370 vs :133). */
static void fast_encode_block(fast_state *e, orc_blockenc *blk, const uint8_t *src, int32_t s0, int32_t end,
                              int nohist) {
    const int32_t inputMargin = 8;
    const int32_t minNonLiteralBlockSize = 1 + 1 + inputMargin;
    int32_t s = s0;
    blk->size = (size_t)(end - s0);
    if (end - s0 < minNonLiteralBlockSize) {
        blk->extraLits = (size_t)(end - s0);
        blk->nlit = 0;
        orc_blockenc_add_literals(blk, src + s0, (size_t)(end - s0));
        return;
    }
    int32_t sLimit = end - inputMargin;
    const int32_t stepSize = 2;
    const unsigned hashLog = FAST_TABLE_BITS;
    const int kSearchStrength = 6;
    int32_t nextEmit = s;
    uint64_t cv = orc_ld64(src + s);
    int32_t offset1 = (int32_t)blk->recentOffsets[0];
    int32_t offset2 = (int32_t)blk->recentOffsets[1];

    for (;;) {
        int32_t t = 0;
        int canRepeat = blk->nseq > 2;
        int found = 0;
        for (;;) {
            uint32_t nextHash = hash6(cv, hashLog);
            uint32_t nextHash2 = hash6(cv >> 8, hashLog);
            tentry candidate = e->table[nextHash];
            tentry candidate2 = e->table[nextHash2];
            int32_t repIndex = s - offset1 + 2;
            e->table[nextHash].offset = s + e->cur; e->table[nextHash].val = (uint32_t)cv;
            e->table[nextHash2].offset = s + 1 + e->cur; e->table[nextHash2].val = (uint32_t)(cv >> 8);

            int repOK = nohist ? (blk->nseq > 2) : (canRepeat && repIndex >= 0);
            if (repOK && orc_ld32(src + repIndex) == (uint32_t)(cv >> 16)) {
                int32_t length = 4 + match_len(src, s + 6, repIndex + 4, end);
                uint32_t mlen = (uint32_t)(length - ORC_ZSTD_MINMATCH);
                int32_t start = s + 2;
                int32_t startLimit = nextEmit + 1;
                int32_t sMin = s - e->maxMatchOff; if (sMin < 0) sMin = 0;
                while (repIndex > sMin && start > startLimit && src[repIndex - 1] == src[start - 1] &&
                       (nohist || mlen < ORC_ZSTD_MAX_MATCHLEN - ORC_ZSTD_MINMATCH)) {
                    repIndex--; start--; mlen++;
                }
                uint32_t litLen = (start != nextEmit) ? (uint32_t)(start - nextEmit) : 0;
                add_lits(blk, src, nextEmit, start);
                orc_blockenc_add_seq(blk, litLen, mlen, 1);
                s += length + 2;
                nextEmit = s;
                if (s >= sLimit) goto done;
                cv = orc_ld64(src + s);
                continue;
            }
            int32_t coffset0 = s - (candidate.offset - e->cur);
            int32_t coffset1 = s - (candidate2.offset - e->cur) + 1;
            if (coffset0 < e->maxMatchOff && (uint32_t)cv == candidate.val) {
                t = candidate.offset - e->cur; found = 1; break;
            }
            if (coffset1 < e->maxMatchOff && (uint32_t)(cv >> 8) == candidate2.val) {
                t = candidate2.offset - e->cur; s++; found = 1; break;
            }
            s += stepSize + ((s - nextEmit) >> (kSearchStrength - 1));
            if (s >= sLimit) goto done;
            cv = orc_ld64(src + s);
        }
        (void)found;
        offset2 = offset1;
        offset1 = s - t;
        int32_t l = match_len(src, s + 4, t + 4, end) + 4;
        int32_t tMin = s - e->maxMatchOff; if (tMin < 0) tMin = 0;
        while (t > tMin && s > nextEmit && src[t - 1] == src[s - 1] && (nohist || l < ORC_ZSTD_MAX_MATCHLEN)) {
            s--; t--; l++;
        }
        add_lits(blk, src, nextEmit, s);
        orc_blockenc_add_seq(blk, (uint32_t)(s - nextEmit), (uint32_t)(l - ORC_ZSTD_MINMATCH), (uint32_t)(s - t) + 3);
        s += l;
        nextEmit = s;
        if (s >= sLimit) goto done;
        cv = orc_ld64(src + s);

        /* Check offset 2 */
        {
            int32_t o2 = s - offset2;
            int can2 = nohist ? (blk->nseq > 2) : canRepeat;
            if (can2 && orc_ld32(src + o2) == (uint32_t)cv) {
                int32_t l2 = 4 + match_len(src, s + 4, o2 + 4, end);
                uint32_t nextHash = hash6(cv, hashLog);
                e->table[nextHash].offset = s + e->cur; e->table[nextHash].val = (uint32_t)cv;
                orc_blockenc_add_seq(blk, 0, (uint32_t)l2 - ORC_ZSTD_MINMATCH, 1);
                s += l2;
                nextEmit = s;
                int32_t tmp = offset1; offset1 = offset2; offset2 = tmp;
                if (s >= sLimit) goto done;
                cv = orc_ld64(src + s);
            }
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

static fast_state *fast_state_new(int32_t window) {
    fast_state *e = (fast_state *)calloc(1, sizeof(*e));
    e->maxMatchOff = window;
    e->cur = window; /* Reset on a fresh encoder: cur += maxMatchOff (enc_base.go:183-187) */
    return e;
}
/* fastBase.resetBase + the cur bump at the end of EncodeNoHist: everything already in the table ends
 * up >= maxMatchOff away.  histLen = bytes the previous use put behind cur. */
static void fast_state_reset(fast_state *e, int32_t histLen) {
    const int32_t bufferReset = 0x7fffffff - 2 * e->maxMatchOff; /* encoder_options.go:51-73 */
    if (e->cur >= bufferReset) {
        memset(e->table, 0, sizeof(e->table));
        e->cur = e->maxMatchOff;
        return;
    }
    e->cur += e->maxMatchOff + histLen;
}

ORC_API void orc_enc_fast_nohist(orc_blockenc *b, const uint8_t *src, size_t n) {
    fast_state *e = fast_state_new(4 << 20);
    fast_encode_block(e, b, src, 0, (int32_t)n, 1);
    free(e);
}

/* Reusable encoder (one per host thread), the shape of the reference's pooled encoders
 * (zstd/encoder.go:90-99,722-729): no allocation and no table clearing per EncodeAll call. */
void *orc_dfast_state_new(void);
void orc_dfast_state_free(void *st);
void orc_dfast_state_reset(void *st, int32_t lastLen);
void orc_dfast_encode_all_blocks_st(void *st, orc_blockenc *blk, const uint8_t *src, size_t n, size_t blockSize,
                                    uint8_t *dst, size_t cap, size_t *pos, int *err);

typedef struct {
    orc_blockenc *blk;
    fast_state *fast;
    int32_t lastLen;
    void *dfast;          /* pooled level-2 state, created on first use */
    int32_t dfastLastLen;
} orc_zstd_cctx;

ORC_API orc_zstd_cctx *orc_zstd_cctx_new(void) {
    init_predef();
    orc_zstd_cctx *c = (orc_zstd_cctx *)calloc(1, sizeof(*c));
    c->blk = orc_blockenc_new();
    c->fast = fast_state_new(4 << 20);
    c->fast->cur = 0;
    c->lastLen = 0;
    return c;
}
ORC_API void orc_zstd_cctx_free(orc_zstd_cctx *c) {
    if (!c) return;
    orc_blockenc_free(c->blk); free(c->fast); if (c->dfast) orc_dfast_state_free(c->dfast); free(c);
}

/* ----------------------------------------------------------------- frames */
static size_t frame_header(uint8_t *dst, uint64_t contentSize, uint32_t windowSize, int single, int checksum) {
    /* frameHeader.appendTo, zstd/frameenc.go:25-92 (DictID = 0) */
    size_t o = 0;
    dst[o++] = 0x28; dst[o++] = 0xB5; dst[o++] = 0x2F; dst[o++] = 0xFD;
    uint8_t fhd = 0;
    if (checksum) fhd |= 1 << 2;
    if (single) fhd |= 1 << 5;
    uint8_t fcs = 0;
    if (contentSize >= 256) fcs++;
    if (contentSize >= 65536 + 256) fcs++;
    if (contentSize >= 0xffffffffull) fcs++;
    fhd |= (uint8_t)(fcs << 6);
    dst[o++] = fhd;
    if (!single) {
        unsigned len32 = (windowSize - 1) ? orc_highbit32(windowSize - 1) + 1 : 0;
        dst[o++] = (uint8_t)((len32 - 10) << 3);
    }
    switch (fcs) {
    case 0: if (single) dst[o++] = (uint8_t)contentSize; break;
    case 1: contentSize -= 256; dst[o++] = (uint8_t)contentSize; dst[o++] = (uint8_t)(contentSize >> 8); break;
    case 2: for (int i = 0; i < 4; i++) dst[o++] = (uint8_t)(contentSize >> (8 * i)); break;
    default: for (int i = 0; i < 8; i++) dst[o++] = (uint8_t)(contentSize >> (8 * i)); break;
    }
    return o;
}

static int32_t window_size_for(int64_t size, int32_t maxMatchOff) { /* fastBase.WindowSize, enc_base.go:42-50 */
    if (size > 0 && size < (int64_t)maxMatchOff) {
        int blen = 0; { uint64_t v = (uint64_t)size; while (v) { blen++; v >>= 1; } }
        int32_t b = (int32_t)1 << blen;
        return b > 1024 ? b : 1024;
    }
    return maxMatchOff;
}

void orc_dfast_encode_all_blocks(orc_blockenc *blk, const uint8_t *src, size_t n, size_t blockSize,
                                 uint8_t *dst, size_t cap, size_t *pos, int *err);
void orc_better_encode_all_blocks(orc_blockenc *blk, const uint8_t *src, size_t n, size_t blockSize,
                                  uint8_t *dst, size_t cap, size_t *pos, int *err);

static int64_t encode_all_impl(orc_zstd_cctx *cc, const uint8_t *src, size_t n, int level, int crc, uint8_t *dst, size_t cap);

ORC_API int64_t orc_zstd_encode_all(const uint8_t *src, size_t n, int level, int crc, uint8_t *dst, size_t cap) {
    return encode_all_impl(NULL, src, n, level, crc, dst, cap);
}
/* EncodeAll on a reused encoder (level 1 only reuses state; other levels fall back to fresh state) */
ORC_API int64_t orc_zstd_encode_all_ctx(orc_zstd_cctx *cc, const uint8_t *src, size_t n, int level, int crc,
                                        uint8_t *dst, size_t cap) {
    return encode_all_impl(level <= 2 ? cc : NULL, src, n, level, crc, dst, cap);
}

/* Benchmark helper (bench.py cpu_baseline / --impl reference): one EncodeAll per `chunk` bytes of src on a
 * pooled encoder, repeated over the sample until `seconds` of wall time have passed (checked once per pass).
 * Returns encoded bytes of the last pass; *in_bytes receives the total input bytes encoded. */
#include <time.h>
ORC_API int64_t orc_zstd_bench_chunks(orc_zstd_cctx *cc, const uint8_t *src, size_t chunk, size_t nchunks, int level,
                                      int crc, uint8_t *dst, size_t cap, double seconds, uint64_t *in_bytes) {
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    uint64_t done = 0;
    int64_t outb = 0;
    for (;;) {
        outb = 0;
        for (size_t i = 0; i < nchunks; i++) {
            int64_t r = orc_zstd_encode_all_ctx(cc, src + i * chunk, chunk, level, crc, dst, cap);
            if (r < 0) return r;
            outb += r;
        }
        done += (uint64_t)chunk * nchunks;
        clock_gettime(CLOCK_MONOTONIC, &t1);
        double el = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
        if (el >= seconds) break;
    }
    if (in_bytes) *in_bytes = done;
    return outb;
}

static int64_t encode_all_impl(orc_zstd_cctx *cc, const uint8_t *src, size_t n, int level, int crc, uint8_t *dst, size_t cap) {
    init_predef();
    if (level < 1 || level > 3) return ORC_ERR_UNSUPPORTED; /* SpeedFastest, SpeedDefault, SpeedBetterCompression */
    const size_t blockSize = (level == 1) ? (1u << 16) : ORC_ZSTD_MAX_BLOCK; /* encoder_options.go:41,248-252 */
    const int32_t windowSize = (level == 1) ? (4 << 20) : (8 << 20);
    size_t pos = 0;
    if (cap < 18) return ORC_ERR_DST_SMALL;
    if (n == 0) { /* WithZeroFrames default true: header + empty raw last block, encoder.go:732-751 */
        pos = frame_header(dst, 0, 1024, 1, 0);
        dst[pos++] = 1; dst[pos++] = 0; dst[pos++] = 0;
        return (int64_t)pos;
    }
    int single = (n <= (size_t)windowSize) && (n > 1024);
    pos = frame_header(dst, n, (uint32_t)window_size_for((int64_t)n, windowSize), single, crc);
    orc_blockenc *blk = cc ? cc->blk : orc_blockenc_new();
    if (cc) { orc_blockenc_reset(blk); orc_blockenc_init_new_encode(blk); }
    int err = 0;
    if (level == 2) {
        if (cc) {
            if (!cc->dfast) cc->dfast = orc_dfast_state_new();
            else orc_dfast_state_reset(cc->dfast, cc->dfastLastLen);
            cc->dfastLastLen = (int32_t)n;
            orc_dfast_encode_all_blocks_st(cc->dfast, blk, src, n, blockSize, dst, cap, &pos, &err);
        } else {
            orc_dfast_encode_all_blocks(blk, src, n, blockSize, dst, cap, &pos, &err);
        }
    } else if (level == 3) {
        orc_better_encode_all_blocks(blk, src, n, blockSize, dst, cap, &pos, &err);
    } else if (n <= blockSize) {
        orc_blockenc_reset(blk);
        blk->last = 1;
        if (cc) {
            fast_state_reset(cc->fast, cc->lastLen);
            fast_encode_block(cc->fast, blk, src, 0, (int32_t)n, 1);
            cc->lastLen = (int32_t)n;
        } else {
            orc_enc_fast_nohist(blk, src, n);
        }
        err = orc_blockenc_encode(blk, src, n, 0, 1, dst, cap, &pos);
    } else {
        fast_state *e = cc ? cc->fast : fast_state_new(windowSize);
        if (cc) { fast_state_reset(e, cc->lastLen); cc->lastLen = (int32_t)n; }
        size_t off = 0;
        while (off < n && !err) {
            size_t todo = n - off; if (todo > blockSize) todo = blockSize;
            memcpy(blk->prevRecentOffsets, blk->recentOffsets, sizeof(blk->recentOffsets)); /* pushOffsets */
            fast_encode_block(e, blk, src, (int32_t)off, (int32_t)(off + todo), 0);
            if (off + todo == n) blk->last = 1;
            err = orc_blockenc_encode(blk, src + off, todo, 0, 1, dst, cap, &pos);
            orc_blockenc_reset(blk);
            off += todo;
        }
        if (!cc) free(e);
    }
    if (!cc) orc_blockenc_free(blk);
    if (err) return err;
    if (crc) {
        if (pos + 4 > cap) return ORC_ERR_DST_SMALL;
        uint32_t c = (uint32_t)orc_xxh64(src, n, 0);
        orc_st32(dst + pos, c);
        pos += 4;
    }
    return (int64_t)pos;
}

ORC_API size_t orc_zstd_max_encoded_size(size_t size, int level, int crc) { /* encoder.go:843-873 */
    size_t blockSize = (level == 1) ? (1u << 16) : ORC_ZSTD_MAX_BLOCK;
    size_t fh = 4 + 2;
    if (size < 256) fh++;
    else if (size < 65536 + 256) fh += 2;
    else if (size < 0x7fffffff) fh += 4;
    else fh += 8;
    if (crc) fh += 4;
    size_t blocks = (size + blockSize) / blockSize;
    return fh + 3 * blocks + size;
}
