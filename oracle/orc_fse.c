/*
 * oracle/orc_fse.c -- FSE table construction + the standalone 2-state FSE byte
 * codec.  Restates fse/compress.go, fse/decompress.go, fse/fse.go and the
 * identical arithmetic in zstd/fse_encoder.go / zstd/fse_decoder*.go.
 * On the hot path the standalone codec only (de)compresses huff0's <=255
 * Huffman weights (huff0/huff0.go:225, huff0/decompress.go:61).
 * TEST INFRASTRUCTURE ONLY -- see orc_common.h.
 */
#include "orc_fse.h"

static const uint32_t rtbTable[8] = {0, 473195, 504333, 520860, 550000, 700000, 750000, 830000};

/* fse/compress.go:583-683 == zstd/fse_encoder.go:334-427 */
static int normalize2(const uint32_t *count, unsigned symbolLen, uint32_t length, unsigned tableLog,
                      int16_t *norm) {
    const int16_t notYetAssigned = -2;
    uint32_t distributed = 0;
    uint32_t total = length;
    uint32_t lowThreshold = total >> tableLog;
    uint32_t lowOne = (total * 3) >> (tableLog + 1);
    for (unsigned i = 0; i < symbolLen; i++) {
        uint32_t cnt = count[i];
        if (cnt == 0) { norm[i] = 0; continue; }
        if (cnt <= lowThreshold) { norm[i] = -1; distributed++; total -= cnt; continue; }
        if (cnt <= lowOne) { norm[i] = 1; distributed++; total -= cnt; continue; }
        norm[i] = notYetAssigned;
    }
    uint32_t toDistribute = (1u << tableLog) - distributed;
    if ((total / toDistribute) > lowOne) {
        /* risk of rounding to zero */
        lowOne = (total * 3) / (toDistribute * 2);
        for (unsigned i = 0; i < symbolLen; i++) {
            if (norm[i] == notYetAssigned && count[i] <= lowOne) {
                norm[i] = 1; distributed++; total -= count[i];
            }
        }
        toDistribute = (1u << tableLog) - distributed;
    }
    if (distributed == symbolLen + 1) {
        unsigned maxV = 0; uint32_t maxC = 0;
        for (unsigned i = 0; i < symbolLen; i++)
            if (count[i] > maxC) { maxV = i; maxC = count[i]; }
        norm[maxV] = (int16_t)(norm[maxV] + (int16_t)toDistribute);
        return 0;
    }
    if (total == 0) {
        for (uint32_t i = 0; toDistribute > 0; i = (i + 1) % symbolLen) {
            if (norm[i] > 0) { toDistribute--; norm[i]++; }
        }
        return 0;
    }
    uint64_t vStepLog = 62 - (uint64_t)tableLog;
    uint64_t mid = (1ull << (vStepLog - 1)) - 1;
    uint64_t rStep = (((1ull << vStepLog) * (uint64_t)toDistribute) + mid) / (uint64_t)total;
    uint64_t tmpTotal = mid;
    for (unsigned i = 0; i < symbolLen; i++) {
        if (norm[i] == notYetAssigned) {
            uint64_t end = tmpTotal + (uint64_t)count[i] * rStep;
            uint32_t sStart = (uint32_t)(tmpTotal >> vStepLog);
            uint32_t sEnd = (uint32_t)(end >> vStepLog);
            uint32_t weight = sEnd - sStart;
            if (weight < 1) return ORC_ERR_INTERNAL;
            norm[i] = (int16_t)weight;
            tmpTotal = end;
        }
    }
    return 0;
}

/* fse/compress.go:510-581 == zstd/fse_encoder.go:259-330 (after the RLE test) */
int orc_fse_normalize(const uint32_t *count, unsigned symbolLen, uint32_t length, unsigned tableLog,
                      int16_t *norm) {
    uint64_t scale = 62 - (uint64_t)tableLog;
    uint64_t step = (1ull << 62) / (uint64_t)length;
    uint64_t vStep = 1ull << (scale - 20);
    int16_t stillToDistribute = (int16_t)(1 << tableLog);
    unsigned largest = 0;
    int16_t largestP = 0;
    uint32_t lowThreshold = length >> tableLog;

    for (unsigned i = 0; i < symbolLen; i++) {
        uint32_t cnt = count[i];
        if (cnt == 0) { norm[i] = 0; continue; }
        if (cnt <= lowThreshold) {
            norm[i] = -1;
            stillToDistribute--;
        } else {
            int16_t proba = (int16_t)(((uint64_t)cnt * step) >> scale);
            if (proba < 8) {
                uint64_t restToBeat = vStep * (uint64_t)rtbTable[proba];
                uint64_t v = (uint64_t)cnt * step - ((uint64_t)proba << scale);
                if (v > restToBeat) proba++;
            }
            if (proba > largestP) { largestP = proba; largest = i; }
            norm[i] = proba;
            stillToDistribute = (int16_t)(stillToDistribute - proba);
        }
    }
    if ((int16_t)(-stillToDistribute) >= (int16_t)(norm[largest] >> 1)) {
        return normalize2(count, symbolLen, length, tableLog, norm);
    }
    norm[largest] = (int16_t)(norm[largest] + stillToDistribute);
    return 0;
}

int64_t orc_fse_write_ncount(const int16_t *norm, unsigned symbolLen, unsigned tableLog, uint8_t *out,
                             size_t cap) {
    int tableSize = 1 << tableLog;
    int previous0 = 0;
    unsigned charnum = 0;
    uint32_t bitStream = (uint32_t)(tableLog - 5);
    unsigned bitCount = 4;
    int16_t remaining = (int16_t)(tableSize + 1);
    int16_t threshold = (int16_t)tableSize;
    unsigned nbBits = tableLog + 1;
    size_t outP = 0;
#define PUT16()                                                                                        \
    do {                                                                                               \
        if (outP + 2 > cap) return ORC_ERR_DST_SMALL;                                                  \
        out[outP] = (uint8_t)bitStream; out[outP + 1] = (uint8_t)(bitStream >> 8); outP += 2;          \
        bitStream >>= 16;                                                                              \
    } while (0)
    while (remaining > 1) {
        if (previous0) {
            unsigned start = charnum;
            while (norm[charnum] == 0) charnum++;
            while (charnum >= start + 24) {
                start += 24;
                bitStream += 0xFFFFu << bitCount;
                PUT16();
            }
            while (charnum >= start + 3) {
                start += 3;
                bitStream += 3u << bitCount;
                bitCount += 2;
            }
            bitStream += (uint32_t)(charnum - start) << bitCount;
            bitCount += 2;
            if (bitCount > 16) { PUT16(); bitCount -= 16; }
        }
        int16_t count = norm[charnum];
        charnum++;
        int16_t max = (int16_t)((2 * threshold - 1) - remaining);
        if (count < 0) remaining = (int16_t)(remaining + count);
        else remaining = (int16_t)(remaining - count);
        count++;
        if (count >= threshold) count = (int16_t)(count + max);
        bitStream += (uint32_t)count << bitCount;
        bitCount += nbBits;
        if (count < max) bitCount--;
        previous0 = (count == 1);
        if (remaining < 1) return ORC_ERR_INTERNAL;
        while (remaining < threshold) { nbBits--; threshold >>= 1; }
        if (bitCount > 16) { PUT16(); bitCount -= 16; }
    }
    if (outP + 2 > cap) return ORC_ERR_DST_SMALL;
    out[outP] = (uint8_t)bitStream;
    out[outP + 1] = (uint8_t)(bitStream >> 8);
    outP += (bitCount + 7) / 8;
    if (charnum > symbolLen) return ORC_ERR_INTERNAL;
    return (int64_t)outP;
#undef PUT16
}

int orc_fse_build_ctable(const int16_t *norm, unsigned symbolLen, unsigned tableLog, orc_fse_ctable *ct) {
    uint32_t tableSize = 1u << tableLog;
    uint32_t highThreshold = tableSize - 1;
    int16_t cumul[258];
    uint8_t *tableSymbol = ct->tableSymbol;
    cumul[0] = 0;
    for (unsigned u = 0; u < symbolLen; u++) {
        int16_t v = norm[u];
        if (v == -1) {
            cumul[u + 1] = (int16_t)(cumul[u] + 1);
            tableSymbol[highThreshold] = (uint8_t)u;
            highThreshold--;
        } else {
            cumul[u + 1] = (int16_t)(cumul[u] + v);
        }
    }
    if ((uint32_t)cumul[symbolLen] != tableSize) return ORC_ERR_INTERNAL;
    cumul[symbolLen] = (int16_t)(tableSize + 1);

    ct->zeroBits = 0;
    {
        uint32_t step = orc_fse_table_step(tableSize);
        uint32_t tableMask = tableSize - 1;
        uint32_t position = 0;
        int16_t largeLimit = (int16_t)(1 << (tableLog - 1));
        for (unsigned ui = 0; ui < symbolLen; ui++) {
            int16_t v = norm[ui];
            if (v > largeLimit) ct->zeroBits = 1;
            for (int n = 0; n < v; n++) {
                tableSymbol[position] = (uint8_t)ui;
                position = (position + step) & tableMask;
                while (position > highThreshold) position = (position + step) & tableMask;
            }
        }
        if (position != 0) return ORC_ERR_INTERNAL;
    }
    for (uint32_t u = 0; u < tableSize; u++) {
        uint8_t v = tableSymbol[u];
        ct->stateTable[cumul[v]] = (uint16_t)(tableSize + u);
        cumul[v]++;
    }
    {
        int16_t total = 0;
        uint32_t tl = ((uint32_t)tableLog << 16) - (1u << tableLog);
        for (unsigned i = 0; i < symbolLen; i++) {
            int16_t v = norm[i];
            if (v == 0) continue;
            if (v == -1 || v == 1) {
                ct->tt[i].deltaNbBits = tl;
                ct->tt[i].deltaFindState = (int32_t)(total - 1);
                total++;
            } else {
                uint32_t maxBitsOut = (uint32_t)tableLog - orc_highbit32((uint32_t)(v - 1));
                uint32_t minStatePlus = (uint32_t)v << maxBitsOut;
                ct->tt[i].deltaNbBits = (maxBitsOut << 16) - minStatePlus;
                ct->tt[i].deltaFindState = (int32_t)(total - v);
                total = (int16_t)(total + v);
            }
        }
        if (total != (int16_t)tableSize) return ORC_ERR_INTERNAL;
    }
    return 0;
}

/* LSB-first forward bit fetch with zero fill past the end (readNCount helper) */
static inline uint32_t fwd_bits(const uint8_t *in, size_t len, uint64_t bitpos, unsigned n) {
    uint64_t acc = 0;
    size_t byte0 = (size_t)(bitpos >> 3);
    for (unsigned i = 0; i < 8; i++) {
        size_t bi = byte0 + i;
        if (bi < len) acc |= (uint64_t)in[bi] << (8 * i);
    }
    acc >>= (bitpos & 7);
    return (uint32_t)(acc & ((n >= 32) ? 0xffffffffull : ((1ull << n) - 1)));
}

int64_t orc_fse_read_ncount(const uint8_t *in, size_t len, unsigned maxSymbol, unsigned absMaxTableLog,
                            int16_t *norm, unsigned *symbolLenOut, unsigned *tableLogOut) {
    if (len < 4) return ORC_ERR_CORRUPT; /* "input too small" */
    uint64_t bp = 0;
    unsigned nbBits = fwd_bits(in, len, bp, 4) + 5;
    bp += 4;
    if (nbBits > absMaxTableLog) return ORC_ERR_CORRUPT; /* "tableLog too large" */
    unsigned tableLog = nbBits;
    int32_t remaining = (1 << nbBits) + 1;
    int32_t threshold = 1 << nbBits;
    int32_t gotTotal = 0;
    unsigned charnum = 0;
    int previous0 = 0;
    nbBits++;
    while (remaining > 1 && charnum <= maxSymbol) {
        if (previous0) {
            unsigned n0 = charnum;
            while (fwd_bits(in, len, bp, 16) == 0xFFFF) {
                n0 += 24; bp += 16;
                if (bp > 8 * (uint64_t)len + 64) return ORC_ERR_CORRUPT;
            }
            while (fwd_bits(in, len, bp, 2) == 3) { n0 += 3; bp += 2; }
            n0 += fwd_bits(in, len, bp, 2);
            bp += 2;
            if (n0 > 255) return ORC_ERR_CORRUPT; /* "maxSymbolValue too small" */
            while (charnum < n0) { norm[charnum & 0xff] = 0; charnum++; }
        }
        int32_t max = (2 * threshold - 1) - remaining;
        int32_t count;
        uint32_t bitStream = fwd_bits(in, len, bp, 32);
        if ((int32_t)(bitStream & (uint32_t)(threshold - 1)) < max) {
            count = (int32_t)(bitStream & (uint32_t)(threshold - 1));
            bp += nbBits - 1;
        } else {
            count = (int32_t)(bitStream & (uint32_t)(2 * threshold - 1));
            if (count >= threshold) count -= max;
            bp += nbBits;
        }
        count--;
        if (count < 0) { remaining += count; gotTotal -= count; }
        else { remaining -= count; gotTotal += count; }
        norm[charnum & 0xff] = (int16_t)count;
        charnum++;
        previous0 = (count == 0);
        while (remaining < threshold) { nbBits--; threshold >>= 1; }
    }
    if (charnum <= 1) return ORC_ERR_CORRUPT;  /* symbolLen too small */
    if (charnum > 256) return ORC_ERR_CORRUPT; /* symbolLen too big */
    if (remaining != 1) return ORC_ERR_CORRUPT;
    if (bp > 8 * (uint64_t)len) return ORC_ERR_CORRUPT; /* bitCount > 32 */
    if (gotTotal != (1 << tableLog)) return ORC_ERR_CORRUPT;
    *symbolLenOut = charnum;
    *tableLogOut = tableLog;
    return (int64_t)((bp + 7) >> 3);
}

int orc_fse_build_dtable(const int16_t *norm, unsigned symbolLen, unsigned tableLog, orc_fse_dsym *dt) {
    uint32_t tableSize = 1u << tableLog;
    uint32_t highThreshold = tableSize - 1;
    uint16_t symbolNext[256];
    for (unsigned i = 0; i < symbolLen; i++) {
        int16_t v = norm[i];
        if (v == -1) {
            dt[highThreshold].symbol = (uint8_t)i;
            highThreshold--;
            symbolNext[i] = 1;
        } else {
            symbolNext[i] = (uint16_t)v;
        }
    }
    {
        uint32_t tableMask = tableSize - 1;
        uint32_t step = orc_fse_table_step(tableSize);
        uint32_t position = 0;
        for (unsigned ss = 0; ss < symbolLen; ss++) {
            int v = norm[ss];
            for (int i = 0; i < v; i++) {
                dt[position].symbol = (uint8_t)ss;
                position = (position + step) & tableMask;
                while (position > highThreshold) position = (position + step) & tableMask;
            }
        }
        if (position != 0) return ORC_ERR_CORRUPT; /* "corrupted input (position != 0)" */
    }
    for (uint32_t u = 0; u < tableSize; u++) {
        uint8_t symbol = dt[u].symbol;
        uint16_t nextState = symbolNext[symbol];
        symbolNext[symbol] = (uint16_t)(nextState + 1);
        uint8_t nBits = (uint8_t)(tableLog - (uint8_t)orc_highbit32((uint32_t)nextState));
        dt[u].nbBits = nBits;
        uint16_t newState = (uint16_t)(((uint32_t)nextState << nBits) - tableSize);
        if (newState >= tableSize) return ORC_ERR_CORRUPT; /* unreachable for a normalised table */
        if (newState == (uint16_t)u && nBits == 0) return ORC_ERR_CORRUPT; /* "== oldState and no bits" */
        dt[u].newState = newState;
    }
    return 0;
}

/* ---------------- standalone compressor (fse/compress.go:18-78) ---------------- */

static unsigned fse_optimal_tablelog(unsigned tableLogReq, size_t srcLen, unsigned symbolLen) {
    /* fse/compress.go:483-508 (optimalTableLog + minTableLog); uint8 wrap-around kept */
    uint8_t tableLog = (uint8_t)tableLogReq;
    uint32_t minBitsSrc = orc_highbit32((uint32_t)(srcLen - 1)) + 1;
    uint32_t minBitsSymbols = orc_highbit32((uint32_t)(symbolLen - 1)) + 2;
    uint8_t minBits = (uint8_t)(minBitsSrc < minBitsSymbols ? minBitsSrc : minBitsSymbols);
    uint8_t maxBitsSrc = (uint8_t)((uint8_t)orc_highbit32((uint32_t)(srcLen - 1)) - 2);
    if (maxBitsSrc < tableLog) tableLog = maxBitsSrc;
    if (minBits > tableLog) tableLog = minBits;
    if (tableLog < 5) tableLog = 5;
    if (tableLog > ORC_FSE_MAX_TABLELOG) tableLog = ORC_FSE_MAX_TABLELOG;
    return tableLog;
}

ORC_API int64_t orc_fse_compress(const uint8_t *in, size_t n, const uint32_t *countIn, unsigned symbolLen,
                                 unsigned maxCount, unsigned tableLogReq, uint8_t *out, size_t cap) {
    if (n <= 1) return ORC_ERR_INCOMPRESSIBLE;
    uint32_t count[256];
    if (countIn && maxCount) {
        memcpy(count, countIn, sizeof(count));
    } else {
        memset(count, 0, sizeof(count));
        for (size_t i = 0; i < n; i++) count[in[i]]++;
        maxCount = 0; symbolLen = 0;
        for (unsigned i = 0; i < 256; i++) {
            if (!count[i]) continue;
            if (count[i] > maxCount) maxCount = count[i];
            symbolLen = i + 1;
        }
    }
    if (tableLogReq == 0) tableLogReq = 11; /* defaultTablelog, fse/fse.go:31 */
    if (maxCount == n) return ORC_ERR_USE_RLE;
    if (maxCount == 1 || maxCount < (n >> 7)) return ORC_ERR_INCOMPRESSIBLE;
    unsigned tableLog = fse_optimal_tablelog(tableLogReq, n, symbolLen);
    int16_t norm[256];
    memset(norm, 0, sizeof(norm));
    int err = orc_fse_normalize(count, symbolLen, (uint32_t)n, tableLog, norm);
    if (err) return err;
    int64_t hdr = orc_fse_write_ncount(norm, symbolLen, tableLog, out, cap);
    if (hdr < 0) return hdr;
    static __thread orc_fse_ctable ct;
    err = orc_fse_build_ctable(norm, symbolLen, tableLog, &ct);
    if (err) return err;
    if (n <= 2) return ORC_ERR_INTERNAL; /* "compress: src too small" */

    orc_bw bw;
    orc_bw_init(&bw, out + hdr, cap - (size_t)hdr);
    uint16_t c1, c2;
    size_t ip = n;
#define ENC(st, sym)                                                                                   \
    do {                                                                                               \
        orc_symtt t_ = ct.tt[(sym)];                                                                   \
        uint32_t nb_ = ((uint32_t)(st) + t_.deltaNbBits) >> 16;                                        \
        int32_t ds_ = (int32_t)((st) >> (nb_ & 15)) + t_.deltaFindState;                               \
        orc_bw_add(&bw, (st), nb_);                                                                    \
        (st) = ct.stateTable[ds_];                                                                     \
    } while (0)
    if (ip & 1) {
        c1 = orc_fse_cstate_init(&ct, ct.tt[in[ip - 1]]);
        c2 = orc_fse_cstate_init(&ct, ct.tt[in[ip - 2]]);
        ENC(c1, in[ip - 3]);
        ip -= 3;
    } else {
        c2 = orc_fse_cstate_init(&ct, ct.tt[in[ip - 1]]);
        c1 = orc_fse_cstate_init(&ct, ct.tt[in[ip - 2]]);
        ip -= 2;
    }
    if (ip & 2) {
        ENC(c2, in[ip - 1]);
        ENC(c1, in[ip - 2]);
        ip -= 2;
    }
    while (ip >= 4) {
        uint8_t v3 = in[ip - 4], v2 = in[ip - 3], v1 = in[ip - 2], v0 = in[ip - 1];
        ENC(c2, v0);
        ENC(c1, v1);
        ENC(c2, v2);
        ENC(c1, v3);
        ip -= 4;
    }
#undef ENC
    orc_bw_add(&bw, c2, tableLog);
    orc_bw_add(&bw, c1, tableLog);
    orc_bw_close(&bw);
    if (bw.overflow) return ORC_ERR_DST_SMALL;
    size_t total = (size_t)hdr + bw.pos;
    if (total >= n) return ORC_ERR_INCOMPRESSIBLE;
    return (int64_t)total;
}

/* fse.Decompress (fse/decompress.go:18-46, 260-330).  limit = DecompressLimit. */
ORC_API int64_t orc_fse_decompress(const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t limit) {
    int16_t norm[256];
    unsigned symbolLen = 0, tableLog = 0;
    memset(norm, 0, sizeof(norm));
    int64_t hdr = orc_fse_read_ncount(in, n, 255, 15, norm, &symbolLen, &tableLog);
    if (hdr < 0) return hdr;
    if (tableLog > ORC_FSE_MAX_TABLELOG) return ORC_ERR_UNSUPPORTED;
    static __thread orc_fse_dsym dt[ORC_FSE_MAX_TABLESIZE];
    int err = orc_fse_build_dtable(norm, symbolLen, tableLog, dt);
    if (err) return err;
    orc_br br;
    if ((size_t)hdr > n) return ORC_ERR_CORRUPT;
    err = orc_br_init(&br, in + hdr, n - (size_t)hdr);
    if (err) return err;
    /* fse getBits: returns 0 without advancing once the register is exhausted
     * (fse/bitreader.go:47-52) */
#define GETBITS(nb) (((nb) == 0 || orc_br_finished(&br)) ? 0u : orc_br_read(&br, (nb)))
    uint16_t s1 = (uint16_t)GETBITS(tableLog);
    uint16_t s2 = (uint16_t)GETBITS(tableLog);
    size_t o = 0;
#define PUSH(b)                                                                                        \
    do {                                                                                               \
        if (o >= cap) return ORC_ERR_DST_SMALL;                                                        \
        out[o++] = (b);                                                                                \
    } while (0)
#define NEXT(st, dstv)                                                                                 \
    do {                                                                                               \
        orc_fse_dsym e_ = dt[(st)];                                                                    \
        uint16_t lb_ = (uint16_t)GETBITS(e_.nbBits);                                                   \
        (st) = (uint16_t)(e_.newState + lb_);                                                          \
        dstv = e_.symbol;                                                                              \
    } while (0)
    for (;;) {
        uint8_t sym;
        if (orc_br_finished(&br) && dt[s1].nbBits > 0) { PUSH(dt[s1].symbol); PUSH(dt[s2].symbol); break; }
        NEXT(s1, sym); PUSH(sym);
        if (orc_br_finished(&br) && dt[s2].nbBits > 0) { PUSH(dt[s2].symbol); PUSH(dt[s1].symbol); break; }
        NEXT(s2, sym); PUSH(sym);
        if (o >= limit) return ORC_ERR_CORRUPT; /* output size > DecompressLimit */
    }
#undef NEXT
#undef PUSH
#undef GETBITS
    if (orc_br_overread(&br)) return ORC_ERR_CORRUPT;
    return (int64_t)o;
}
