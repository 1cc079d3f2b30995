/*
 * oracle/orc_zstd_dec.c -- zstd frame/block decoder, restating
 *   zstd/framedec.go (frameDec.reset/runDecoder/checkCRC), zstd/decoder.go:319 (DecodeAll)
 *   zstd/blockdec.go (blockDec.reset/decodeBuf/decodeLiterals/decodeCompressed/prepareSequences)
 *   zstd/seqdec.go (sequenceDecs.initialize/decodeSync/next/adjustOffset), zstd/bitreader.go
 *   zstd/fse_decoder.go + fse_decoder_generic.go (readNCount/buildDtable/transform/setRLE)
 *   zstd/fse_predefined.go (default tables, baselines), zstd/history.go
 * including the reference's stricter-than-libzstd validity rules.
 * TEST INFRASTRUCTURE ONLY -- see orc_common.h.  Pinned by the reference's
 * golden vectors (decoder.zip / good.zip / bad.zip), see tests/.
 */
#include <stdlib.h>
#include "orc_zstd.h"

#define MAX_TABLELOG_DEC 9 /* tablelogAbsoluteMax, fse_decoder.go:15 */
#define MAX_WINDOW_SIZE (1ull << 29) /* MaxWindowSize, framedec.go:43 */
#define MIN_WINDOW_SIZE (1ull << 10)
#define MAX_DECODED_SIZE (64ull << 30)
#define MAX_OFFSET_BITS 30

typedef struct {
    uint8_t nbBits;
    uint8_t addBits;
    uint16_t newState;
    uint32_t baseline;
} dsym; /* decSymbol, fse_decoder.go:216-263 */

typedef struct {
    dsym dt[1 << MAX_TABLELOG_DEC];
    unsigned actualTableLog;
    int valid;
} fse_dec;

typedef struct { uint32_t baseLine; uint8_t addBits; } base_off;
static base_off symLL[36], symOF[MAX_OFFSET_BITS + 1], symML[53];
static fse_dec predefDec[3];
static int tablesReady = 0;

static void fill_base(base_off *dst, int n, uint32_t base, const uint8_t *bits) { /* fse_predefined.go:58-73 */
    for (int i = 0; i < n; i++) {
        dst[i].baseLine = base;
        dst[i].addBits = bits[i];
        base += 1u << bits[i];
    }
}

static int build_dtable(fse_dec *f, const int16_t *norm, unsigned symbolLen, unsigned tableLog,
                        const base_off *t, unsigned tlen) {
    static __thread orc_fse_dsym tmp[1 << MAX_TABLELOG_DEC];
    int err = orc_fse_build_dtable(norm, symbolLen, tableLog, tmp);
    if (err) return err;
    unsigned ts = 1u << tableLog;
    for (unsigned i = 0; i < ts; i++) { /* transform, fse_decoder.go:282-299 */
        unsigned sym = tmp[i].symbol;
        if (sym >= tlen) return ORC_ERR_CORRUPT; /* "invalid decoding table entry" */
        f->dt[i].nbBits = tmp[i].nbBits;
        f->dt[i].newState = tmp[i].newState;
        f->dt[i].addBits = t[sym].addBits;
        f->dt[i].baseline = t[sym].baseLine;
    }
    f->actualTableLog = tableLog;
    f->valid = 1;
    return 0;
}

static void init_tables(void) {
    if (tablesReady) return;
    static const uint8_t llb[20] = {1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
    static const uint8_t mlb[21] = {1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
    uint8_t ofb[29];
    for (int i = 0; i < 16; i++) { symLL[i].baseLine = (uint32_t)i; symLL[i].addBits = 0; }
    fill_base(symLL + 16, 20, 16, llb);
    for (int i = 0; i < 32; i++) { symML[i].baseLine = (uint32_t)i + 3; symML[i].addBits = 0; }
    fill_base(symML + 32, 21, 35, mlb);
    symOF[0].baseLine = 0; symOF[0].addBits = 0;
    symOF[1].baseLine = 1; symOF[1].addBits = 1;
    for (int i = 0; i < 29; i++) ofb[i] = (uint8_t)(i + 2);
    fill_base(symOF + 2, 29, 1, ofb);

    static const int16_t llN[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2,
                                    2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
    static const int16_t ofN[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1,
                                    1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};
    static const int16_t mlN[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                                    1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                                    1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
    build_dtable(&predefDec[0], llN, 36, 6, symLL, 36);
    build_dtable(&predefDec[1], ofN, 29, 5, symOF, MAX_OFFSET_BITS + 1);
    build_dtable(&predefDec[2], mlN, 53, 6, symML, 53);
    tablesReady = 1;
}

/* Exposed for TestPredefTables-style checks (zstd/decoder_test.go:1938-2047):
 * writes {nbBits, addBits, newState, baseline} per state; returns table size. */
ORC_API int orc_zstd_predef_table(int which, uint32_t *out /* 4 per entry */, int capEntries) {
    init_tables();
    if (which < 0 || which > 2) return -1;
    int ts = 1 << predefDec[which].actualTableLog;
    if (capEntries < ts) return -1;
    for (int i = 0; i < ts; i++) {
        out[4 * i] = predefDec[which].dt[i].nbBits;
        out[4 * i + 1] = predefDec[which].dt[i].addBits;
        out[4 * i + 2] = predefDec[which].dt[i].newState;
        out[4 * i + 3] = predefDec[which].dt[i].baseline;
    }
    return ts;
}

typedef struct {
    orc_huf_dtable huff;
    int haveHuff;
    fse_dec tables[3]; /* LL, OF, ML owned tables */
    fse_dec *cur[3];   /* current decoder per type (NULL = not defined) */
    int64_t recent[3];
    uint64_t windowSize;
    uint8_t *litBuf;
} frame_state;

/* Test_seqdec_decoder hook (zstd/seqdec_test.go:199-302): when set, decode_sequences stores (mo, ml, ll) per sequence
 * here instead of executing the sequences -- sequenceDecs.decode (seqdec.go:119) vs decodeSync/execute. */
static int64_t *seqDump = NULL;

/* sequence section + execution; out = frame output start, *outLen = bytes so far */
static int decode_sequences(frame_state *fs, const uint8_t *in, size_t inLen, int nSeqs, const uint8_t *literals,
                            size_t nLit, uint8_t *out, size_t *outLen, size_t outCap) {
    orc_br br;
    int err = orc_br_init(&br, in, inLen); /* bitReader.init, bitreader.go:26-45 */
    if (err) return err;
    fse_dec *llT = fs->cur[0], *ofT = fs->cur[1], *mlT = fs->cur[2];
    if (!llT || !ofT || !mlT) return ORC_ERR_CORRUPT; /* "sequence decoder not defined" */
    /* sequenceDecs.initialize, seqdec.go:80-100: LL, OF, ML */
    dsym llS = llT->dt[orc_br_read(&br, llT->actualTableLog)];
    dsym ofS = ofT->dt[orc_br_read(&br, ofT->actualTableLog)];
    dsym mlS = mlT->dt[orc_br_read(&br, mlT->actualTableLog)];
    size_t startSize = *outLen;
    size_t o = *outLen;
    uint64_t maxBlockSize = fs->windowSize < ORC_ZSTD_MAX_BLOCK ? fs->windowSize : ORC_ZSTD_MAX_BLOCK;
    size_t litPos = 0;
    for (int i = nSeqs - 1; i >= 0; i--) {
        if (orc_br_overread(&br)) return ORC_ERR_CORRUPT; /* io.ErrUnexpectedEOF */
        int64_t ll = llS.baseline, ml = mlS.baseline, mo = ofS.baseline;
        unsigned moB = ofS.addBits;
        mo += orc_br_read(&br, moB);
        ml += orc_br_read(&br, mlS.addBits);
        ll += orc_br_read(&br, llS.addBits);
        /* adjustOffset, seqdec.go:463-500 */
        if (moB > 1) {
            fs->recent[2] = fs->recent[1]; fs->recent[1] = fs->recent[0]; fs->recent[0] = mo;
        } else {
            if (ll == 0) mo++;
            if (mo == 0) {
                mo = fs->recent[0];
            } else {
                int64_t temp;
                if (mo == 3) temp = fs->recent[0] - 1;
                else temp = fs->recent[mo];
                if (temp == 0) temp = 1; /* corrupted input: force offset to 1 */
                if (mo != 1) fs->recent[2] = fs->recent[1];
                fs->recent[1] = fs->recent[0];
                fs->recent[0] = temp;
                mo = temp;
            }
        }
        if ((size_t)ll > nLit - litPos) return ORC_ERR_CORRUPT;       /* "unexpected literal count" */
        if (seqDump) {   /* decode only (seqdec.go:119-218): same value rules, no output */
            int64_t *d = seqDump + 3 * (size_t)(nSeqs - 1 - i);
            d[0] = mo; d[1] = ml; d[2] = ll;
            litPos += (size_t)ll;
            if (ml > ORC_ZSTD_MAX_MATCHLEN) return ORC_ERR_CORRUPT;
            if (mo == 0 && ml > 0) return ORC_ERR_CORRUPT;
            goto next_state;
        }
        size_t size = (size_t)ll + (size_t)ml + o;
        if (size - startSize > maxBlockSize) return ORC_ERR_CORRUPT;  /* "output bigger than max block size" */
        if (ml > ORC_ZSTD_MAX_MATCHLEN) return ORC_ERR_CORRUPT;       /* "match len bigger than max allowed" */
        if (size > outCap) return ORC_ERR_DST_SMALL;
        memcpy(out + o, literals + litPos, (size_t)ll);
        o += (size_t)ll; litPos += (size_t)ll;
        if (mo == 0 && ml > 0) return ORC_ERR_CORRUPT;                /* "zero matchoff and matchlen > 0" */
        if ((uint64_t)mo > o || (uint64_t)mo > fs->windowSize) return ORC_ERR_CORRUPT; /* offset > history */
        if (ml > 0) {
            const uint8_t *from = out + o - mo;
            for (int64_t k = 0; k < ml; k++) out[o + k] = from[k]; /* overlap-safe byte copy */
            o += (size_t)ml;
        }
    next_state:
        if (i == 0) break;
        /* state update, seqdec.go:405-424: LL bits first, then ML, then OF */
        {
            unsigned nl = llS.nbBits, nm = mlS.nbBits, no = ofS.nbBits;
            uint32_t bl = orc_br_read(&br, nl);
            uint32_t bm = orc_br_read(&br, nm);
            uint32_t bo = orc_br_read(&br, no);
            llS = llT->dt[(llS.newState + bl) & ((1 << MAX_TABLELOG_DEC) - 1)];
            mlS = mlT->dt[(mlS.newState + bm) & ((1 << MAX_TABLELOG_DEC) - 1)];
            ofS = ofT->dt[(ofS.newState + bo) & ((1 << MAX_TABLELOG_DEC) - 1)];
        }
    }
    if (seqDump) return (br.pos != br.total) ? ORC_ERR_CORRUPT : 0;
    size_t rest = nLit - litPos;
    if (rest + o - startSize > maxBlockSize) return ORC_ERR_CORRUPT;
    if (o + rest > outCap) return ORC_ERR_DST_SMALL;
    memcpy(out + o, literals + litPos, rest);
    o += rest;
    *outLen = o;
    /* br.close(): stream must be consumed exactly, bitreader.go:120-131 */
    if (br.pos != br.total) return ORC_ERR_CORRUPT;
    return 0;
}

static int decode_compressed_block(frame_state *fs, const uint8_t *in, size_t len, uint8_t *out, size_t *outLen,
                                   size_t outCap) {
    /* decodeLiterals, blockdec.go:275-474 */
    if (len < 2) return ORC_ERR_CORRUPT; /* ErrBlockTooSmall */
    unsigned litType = in[0] & 3;
    unsigned sizeFormat = (in[0] >> 2) & 3;
    size_t litRegenSize = 0, litCompSize = 0;
    int fourStreams = 0;
    if (litType == 0 || litType == 1) {
        switch (sizeFormat) {
        case 0: case 2: litRegenSize = in[0] >> 3; in += 1; len -= 1; break;
        case 1: litRegenSize = (size_t)(in[0] >> 4) + ((size_t)in[1] << 4); in += 2; len -= 2; break;
        default:
            if (len < 3) return ORC_ERR_CORRUPT;
            litRegenSize = (size_t)(in[0] >> 4) + ((size_t)in[1] << 4) + ((size_t)in[2] << 12);
            in += 3; len -= 3; break;
        }
    } else {
        uint64_t n;
        switch (sizeFormat) {
        case 0: case 1:
            if (len < 3) return ORC_ERR_CORRUPT;
            n = (uint64_t)(in[0] >> 4) + ((uint64_t)in[1] << 4) + ((uint64_t)in[2] << 12);
            litRegenSize = n & 1023; litCompSize = n >> 10; fourStreams = sizeFormat == 1;
            in += 3; len -= 3; break;
        case 2:
            fourStreams = 1;
            if (len < 4) return ORC_ERR_CORRUPT;
            n = (uint64_t)(in[0] >> 4) + ((uint64_t)in[1] << 4) + ((uint64_t)in[2] << 12) + ((uint64_t)in[3] << 20);
            litRegenSize = n & 16383; litCompSize = n >> 14;
            in += 4; len -= 4; break;
        default:
            fourStreams = 1;
            if (len < 5) return ORC_ERR_CORRUPT;
            n = (uint64_t)(in[0] >> 4) + ((uint64_t)in[1] << 4) + ((uint64_t)in[2] << 12) + ((uint64_t)in[3] << 20) +
                ((uint64_t)in[4] << 28);
            litRegenSize = n & 262143; litCompSize = n >> 18;
            in += 5; len -= 5; break;
        }
    }
    if (litRegenSize > fs->windowSize || litRegenSize > ORC_ZSTD_MAX_BLOCK) return ORC_ERR_WINDOW;
    const uint8_t *literals = NULL;
    switch (litType) {
    case 0:
        if (len < litRegenSize) return ORC_ERR_CORRUPT;
        literals = in; in += litRegenSize; len -= litRegenSize;
        break;
    case 1:
        if (len < 1) return ORC_ERR_CORRUPT;
        memset(fs->litBuf, in[0], litRegenSize);
        literals = fs->litBuf; in += 1; len -= 1;
        break;
    case 3: { /* treeless */
        if (len < litCompSize) return ORC_ERR_CORRUPT;
        if (!fs->haveHuff) return ORC_ERR_CORRUPT; /* "treeless, but no history was defined" */
        int e = fourStreams ? orc_huf_decompress4x(&fs->huff, in, litCompSize, fs->litBuf, litRegenSize)
                            : orc_huf_decompress1x(&fs->huff, in, litCompSize, fs->litBuf, litRegenSize);
        if (e) return e;
        literals = fs->litBuf; in += litCompSize; len -= litCompSize;
        break;
    }
    default: { /* compressed */
        if (len < litCompSize) return ORC_ERR_CORRUPT;
        int64_t used = orc_huf_read_table(&fs->huff, in, litCompSize);
        if (used < 0) { fs->haveHuff = 0; return (int)used; }
        fs->haveHuff = 1;
        int e = fourStreams
                    ? orc_huf_decompress4x(&fs->huff, in + used, litCompSize - (size_t)used, fs->litBuf, litRegenSize)
                    : orc_huf_decompress1x(&fs->huff, in + used, litCompSize - (size_t)used, fs->litBuf, litRegenSize);
        if (e) return e;
        literals = fs->litBuf; in += litCompSize; len -= litCompSize;
        break;
    }
    }
    /* prepareSequences, blockdec.go:505-630 */
    if (len < 1) return ORC_ERR_CORRUPT;
    int nSeqs;
    uint8_t seqHeader = in[0];
    if (seqHeader < 128) { nSeqs = seqHeader; in += 1; len -= 1; }
    else if (seqHeader < 255) {
        if (len < 2) return ORC_ERR_CORRUPT;
        nSeqs = ((int)(seqHeader - 128) << 8) | in[1]; in += 2; len -= 2;
    } else {
        if (len < 3) return ORC_ERR_CORRUPT;
        nSeqs = 0x7f00 + in[1] + ((int)in[2] << 8); in += 3; len -= 3;
    }
    if (nSeqs == 0 && len != 0) return ORC_ERR_CORRUPT; /* ErrUnexpectedBlockSize */
    if (nSeqs == 0) {
        if (*outLen + litRegenSize > outCap) return ORC_ERR_DST_SMALL;
        memcpy(out + *outLen, literals, litRegenSize);
        *outLen += litRegenSize;
        return 0;
    }
    if (len < 1) return ORC_ERR_CORRUPT;
    uint8_t compMode = in[0];
    in += 1; len -= 1;
    if (compMode & 3) return ORC_ERR_CORRUPT; /* "reserved bits not zero" */
    static const unsigned maxSym[3] = {35, 30, 52};
    const base_off *symT[3] = {symLL, symOF, symML};
    const unsigned symTLen[3] = {36, MAX_OFFSET_BITS + 1, 53};
    for (unsigned i = 0; i < 3; i++) {
        unsigned mode = (compMode >> (6 - i * 2)) & 3;
        switch (mode) {
        case 0: fs->cur[i] = &predefDec[i]; break;
        case 1: {
            if (len < 1) return ORC_ERR_CORRUPT;
            uint8_t v = in[0]; in += 1; len -= 1;
            if (v >= symTLen[i]) return ORC_ERR_CORRUPT; /* "rle symbol >= max" */
            fse_dec *f = &fs->tables[i];
            f->actualTableLog = 0;
            f->dt[0].nbBits = 0; f->dt[0].newState = 0;
            f->dt[0].addBits = symT[i][v].addBits; f->dt[0].baseline = symT[i][v].baseLine;
            f->valid = 1;
            fs->cur[i] = f;
            break;
        }
        case 2: {
            int16_t norm[256];
            unsigned symbolLen = 0, tableLog = 0;
            memset(norm, 0, sizeof(norm));
            int64_t used = orc_fse_read_ncount(in, len, maxSym[i], MAX_TABLELOG_DEC, norm, &symbolLen, &tableLog);
            if (used < 0) return (int)used;
            if ((size_t)used > len) return ORC_ERR_CORRUPT;
            in += used; len -= (size_t)used;
            fse_dec *f = &fs->tables[i];
            f->valid = 0;
            fs->cur[i] = NULL;
            int e = build_dtable(f, norm, symbolLen, tableLog, symT[i], symTLen[i]);
            if (e) return e;
            fs->cur[i] = f;
            break;
        }
        default: /* repeat: keep fs->cur[i] */
            break;
        }
    }
    return decode_sequences(fs, in, len, nSeqs, literals, litRegenSize, out, outLen, outCap);
}

/* Header.Decode (zstd/decodeheader.go:94-229): fields of the frame header and of the first block header.
 * out[0..14] = SingleSegment, WindowSize(lo), WindowSize(hi), DictionaryID, HasFCS, FCS(lo), FCS(hi), Skippable,
 * SkippableID, SkippableSize, HeaderSize, FirstBlock.OK, Last, Compressed | HasCheckSum << 1, DecompressedSize,
 * out[15] = CompressedSize.  Returns 0, or ORC_ERR_* (unexpected EOF / magic mismatch / reserved bit). */
ORC_API int orc_zstd_header_decode(const uint8_t *in, size_t n, uint32_t *out) {
    memset(out, 0, 16 * sizeof(uint32_t));
    if (n < 4) return ORC_ERR_CORRUPT;                                  /* io.ErrUnexpectedEOF, :96-98 */
    uint32_t hs = 4;
    const uint8_t *b = in; in += 4; n -= 4;
    if (!(b[0] == 0x28 && b[1] == 0xB5 && b[2] == 0x2F && b[3] == 0xFD)) {
        if (!(b[1] == 0x2A && b[2] == 0x4D && b[3] == 0x18) || (b[0] & 0xf0) != 0x50) return ORC_ERR_MAGIC;  /* :102-104 */
        if (n < 4) return ORC_ERR_CORRUPT;
        out[7] = 1; out[8] = b[0] & 0xf; out[9] = (uint32_t)in[0] | ((uint32_t)in[1] << 8) | ((uint32_t)in[2] << 16) | ((uint32_t)in[3] << 24);
        out[10] = hs + 4;
        return 0;
    }
    if (n < 1) return ORC_ERR_CORRUPT;
    uint8_t fhd = in[0]; in++; n--; hs++;
    const int single = (fhd & (1 << 5)) != 0;
    out[0] = single;
    const uint32_t hasCrc = (fhd & (1 << 2)) != 0;
    if (fhd & (1 << 3)) return ORC_ERR_CORRUPT;                         /* "reserved bit set on frame header", :124-126 */
    if (!single) {                                                      /* :128-142 */
        if (n < 1) return ORC_ERR_CORRUPT;
        uint8_t wd = in[0]; in++; n--; hs++;
        unsigned windowLog = 10 + (wd >> 3);
        uint64_t windowBase = 1ull << windowLog;
        uint64_t ws = windowBase + (windowBase / 8) * (uint64_t)(wd & 7);
        out[1] = (uint32_t)ws; out[2] = (uint32_t)(ws >> 32);
    }
    unsigned size = fhd & 3;                                            /* dictionary id, :144-164 */
    if (size) {
        if (size == 3) size = 4;
        if (n < size) return ORC_ERR_CORRUPT;
        uint32_t id = 0;
        for (unsigned k = 0; k < size; k++) id |= (uint32_t)in[k] << (8 * k);
        out[3] = id; in += size; n -= size; hs += size;
    }
    unsigned fcsSize = 0, v = fhd >> 6;                                 /* :167-176 */
    if (v == 0) { if (single) fcsSize = 1; } else fcsSize = 1u << v;
    if (fcsSize) {
        out[4] = 1;
        if (n < fcsSize) return ORC_ERR_CORRUPT;
        uint64_t fcs = 0;
        for (unsigned k = 0; k < fcsSize; k++) fcs |= (uint64_t)in[k] << (8 * k);
        if (fcsSize == 2) fcs += 256;
        out[5] = (uint32_t)fcs; out[6] = (uint32_t)(fcs >> 32);
        in += fcsSize; n -= fcsSize; hs += fcsSize;
    }
    out[10] = hs;
    out[13] = hasCrc << 1;
    if (n < 3) return 0;                                                /* :203-205: no room for a block header */
    uint32_t bh = (uint32_t)in[0] | ((uint32_t)in[1] << 8) | ((uint32_t)in[2] << 16);
    const uint32_t last = bh & 1, bt = (bh >> 1) & 3, cSize = bh >> 3;
    out[12] = last;                                                     /* FirstBlock.Last is set before the type switch, :208 */
    if (bt == 3) return 0;                                              /* blockTypeReserved */
    if (bt == 1) { out[13] |= 1; out[14] = cSize; out[15] = 1; }        /* RLE */
    else if (bt == 2) { out[13] |= 1; out[15] = cSize; }                /* compressed */
    else { out[14] = cSize; out[15] = cSize; }                          /* raw */
    out[11] = 1;
    return 0;
}

/* The reference's sequence-decoder golden test (zstd/seqdec_test.go:199-302, testdata/seqs.zip <-> seqs-want.zip):
 * three ready-made decoding tables (512 decSymbol words each: LL, ML, OF -- the order readDecoders reads them), the
 * bitstream, nSeqs, the incoming repeat offsets; outputs (mo, ml, ll) per sequence and the final repeat offsets. */
ORC_API int orc_zstd_seqdec_golden(const uint64_t *dtLL, unsigned tlLL, const uint64_t *dtML, unsigned tlML,
                                   const uint64_t *dtOF, unsigned tlOF, const uint8_t *bits, size_t nbytes, int nSeqs,
                                   int64_t *prev3, uint64_t windowSize, size_t nLit, int64_t *out3) {
    init_tables();
    frame_state *fs = (frame_state *)calloc(1, sizeof(frame_state));
    const uint64_t *src[3] = {dtLL, dtOF, dtML};
    const unsigned tl[3] = {tlLL, tlOF, tlML};
    for (int t = 0; t < 3; t++) {
        if (tl[t] > MAX_TABLELOG_DEC) { free(fs); return ORC_ERR_CORRUPT; }
        memcpy(fs->tables[t].dt, src[t], sizeof(fs->tables[t].dt));   /* decSymbol: nbBits | addBits << 8 | newState << 16 | baseline << 32 */
        fs->tables[t].actualTableLog = tl[t]; fs->tables[t].valid = 1;
        fs->cur[t] = &fs->tables[t];
    }
    for (int k = 0; k < 3; k++) fs->recent[k] = prev3[k];
    fs->windowSize = windowSize;
    size_t outLen = 0;
    seqDump = out3;
    int r = decode_sequences(fs, bits, nbytes, nSeqs, NULL, nLit, NULL, &outLen, 0);
    seqDump = NULL;
    for (int k = 0; k < 3; k++) prev3[k] = fs->recent[k];
    free(fs);
    return r;
}

ORC_API int64_t orc_zstd_decode_all(const uint8_t *src, size_t n, uint8_t *dst, size_t cap) {
    init_tables();
    size_t ip = 0, total = 0;
    frame_state *fs = (frame_state *)malloc(sizeof(frame_state));
    if (!fs) return ORC_ERR_INTERNAL;
    fs->litBuf = (uint8_t *)malloc(ORC_ZSTD_MAX_BLOCK + 64);
    int64_t rc = 0;
#define FAIL(code) do { rc = (code); goto out; } while (0)
    for (;;) {
        /* frameDec.reset, framedec.go:65-270 */
        for (;;) {
            if (n - ip == 0) { rc = (int64_t)total; goto out; } /* io.EOF at a frame boundary */
            if (n - ip < 4) FAIL(ORC_ERR_CORRUPT); /* readSmall(3) -> io.ErrUnexpectedEOF */
            const uint8_t *sig = src + ip;
            if (!(sig[1] == 0x2A && sig[2] == 0x4D && sig[3] == 0x18 && (sig[0] & 0xf0) == 0x50)) break;
            ip += 4;
            if (n - ip < 4) FAIL(ORC_ERR_CORRUPT);
            uint32_t skip = orc_ld32(src + ip);
            ip += 4;
            if ((uint64_t)skip > n - ip) FAIL(ORC_ERR_CORRUPT);
            ip += skip;
        }
        if (orc_ld32(src + ip) != 0xFD2FB528u) FAIL(ORC_ERR_MAGIC);
        ip += 4;
        if (n - ip < 1) FAIL(ORC_ERR_CORRUPT);
        uint8_t fhd = src[ip++];
        int singleSegment = (fhd & (1 << 5)) != 0;
        if (fhd & (1 << 3)) FAIL(ORC_ERR_CORRUPT); /* "reserved bit set on frame header" */
        uint64_t windowSize = 0;
        if (!singleSegment) {
            if (n - ip < 1) FAIL(ORC_ERR_CORRUPT);
            uint8_t wd = src[ip++];
            unsigned windowLog = 10 + (wd >> 3);
            uint64_t windowBase = 1ull << windowLog;
            windowSize = windowBase + (windowBase / 8) * (wd & 7);
        }
        if (fhd & 3) {
            unsigned size = fhd & 3; if (size == 3) size = 4;
            if (n - ip < size) FAIL(ORC_ERR_CORRUPT);
            uint32_t id = 0;
            for (unsigned k = 0; k < size; k++) id |= (uint32_t)src[ip + k] << (8 * k);
            ip += size;
            if (id != 0) FAIL(ORC_ERR_UNSUPPORTED); /* ErrUnknownDictionary: no dictionaries registered */
        }
        unsigned fcsSize = 0;
        unsigned v = fhd >> 6;
        if (v == 0) { if (singleSegment) fcsSize = 1; } else fcsSize = 1u << v;
        uint64_t fcs = ~0ull; /* fcsUnknown */
        if (fcsSize) {
            if (n - ip < fcsSize) FAIL(ORC_ERR_CORRUPT);
            fcs = 0;
            for (unsigned k = 0; k < fcsSize; k++) fcs |= (uint64_t)src[ip + k] << (8 * k);
            if (fcsSize == 2) fcs += 256;
            ip += fcsSize;
        }
        int hasCheck = (fhd & (1 << 2)) != 0;
        if (windowSize > MAX_WINDOW_SIZE) FAIL(ORC_ERR_WINDOW);
        if (windowSize == 0 && singleSegment) {
            windowSize = fcs > MIN_WINDOW_SIZE ? fcs : MIN_WINDOW_SIZE;
            if (windowSize > MAX_DECODED_SIZE) FAIL(ORC_ERR_SIZE);
        }
        if (windowSize < MIN_WINDOW_SIZE) FAIL(ORC_ERR_WINDOW);
        if (windowSize > MAX_WINDOW_SIZE && !singleSegment) FAIL(ORC_ERR_WINDOW);
        if (fcs != ~0ull && fcs > MAX_DECODED_SIZE - total) FAIL(ORC_ERR_SIZE);

        /* history.reset, history.go:37-47 */
        fs->haveHuff = 0; fs->huff.loaded = 0;
        fs->cur[0] = fs->cur[1] = fs->cur[2] = NULL;
        fs->recent[0] = 1; fs->recent[1] = 4; fs->recent[2] = 8;
        fs->windowSize = windowSize;

        uint8_t *out = dst + total; /* frame output start */
        size_t outCap = cap - total;
        size_t outLen = 0;
        for (;;) { /* runDecoder, framedec.go:330-412 */
            /* blockDec.reset, blockdec.go:128-211 */
            if (n - ip < 3) FAIL(ORC_ERR_CORRUPT);
            uint32_t bh = (uint32_t)src[ip] | ((uint32_t)src[ip + 1] << 8) | ((uint32_t)src[ip + 2] << 16);
            ip += 3;
            int last = bh & 1;
            unsigned type = (bh >> 1) & 3;
            size_t cSize = bh >> 3;
            switch (type) {
            case 3: FAIL(ORC_ERR_CORRUPT); /* ErrReservedBlockType */
            case 1: /* RLE */
                if (cSize > ORC_ZSTD_MAX_BLOCK || cSize > windowSize) FAIL(ORC_ERR_WINDOW);
                if (n - ip < 1) FAIL(ORC_ERR_CORRUPT);
                if (outLen + cSize > outCap) FAIL(ORC_ERR_DST_SMALL);
                memset(out + outLen, src[ip], cSize);
                outLen += cSize; ip += 1;
                break;
            case 0: /* raw */
                if (cSize > ORC_ZSTD_MAX_BLOCK || cSize > windowSize) FAIL(ORC_ERR_WINDOW);
                if (n - ip < cSize) FAIL(ORC_ERR_CORRUPT);
                if (outLen + cSize > outCap) FAIL(ORC_ERR_DST_SMALL);
                memcpy(out + outLen, src + ip, cSize);
                outLen += cSize; ip += cSize;
                break;
            default: { /* compressed */
                if (cSize > ORC_ZSTD_MAX_BLOCK || (uint64_t)cSize > windowSize) FAIL(ORC_ERR_CORRUPT);
                if (cSize < 2) FAIL(ORC_ERR_CORRUPT); /* ErrBlockTooSmall */
                if (n - ip < cSize) FAIL(ORC_ERR_CORRUPT);
                int e = decode_compressed_block(fs, src + ip, cSize, out, &outLen, outCap);
                if (e) FAIL(e);
                ip += cSize;
                break;
            }
            }
            if ((uint64_t)outLen > MAX_DECODED_SIZE) FAIL(ORC_ERR_SIZE);
            if ((uint64_t)outLen > fcs) FAIL(ORC_ERR_SIZE); /* ErrFrameSizeExceeded */
            if (last) break;
        }
        if (fcs != ~0ull && (uint64_t)outLen != fcs) FAIL(ORC_ERR_SIZE); /* ErrFrameSizeMismatch */
        if (hasCheck) {
            if (n - ip < 4) FAIL(ORC_ERR_CORRUPT);
            uint32_t want = orc_ld32(src + ip);
            ip += 4;
            uint32_t got = (uint32_t)orc_xxh64(out, outLen, 0);
            if (got != want) FAIL(ORC_ERR_CRC);
        }
        total += outLen;
        if (n - ip == 0) break;
    }
    rc = (int64_t)total;
out:
    free(fs->litBuf);
    free(fs);
    return rc;
#undef FAIL
}
