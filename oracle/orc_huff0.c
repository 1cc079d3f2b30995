/*
 * oracle/orc_huff0.c -- huff0 restated: histogram, length-limited canonical
 * Huffman table (bit-exact tie-breaking), table serialisation, 1X/4X encode,
 * table read, 1X/4X decode.
 * Follows huff0/compress.go, huff0/huff0.go, huff0/bitwriter.go,
 * huff0/decompress.go, huff0/decompress_generic.go, huff0/bitreader.go.
 * TEST INFRASTRUCTURE ONLY -- see orc_common.h.
 */
#include "orc_huff0.h"

typedef struct {
    uint32_t count;
    uint16_t parent;
    uint8_t symbol;
    uint8_t nbBits;
} node_t; /* nodeElt, huff0/compress.go:720-742 */

#define HUFF_NODES_LEN 512
#define HUFF_NODES_MASK (HUFF_NODES_LEN - 1)

ORC_API void orc_huf_scratch_init(orc_huf_scratch *s, unsigned wantLogLess, int reuse) {
    memset(s, 0, sizeof(*s));
    s->wantLogLess = wantLogLess;
    s->reuse = reuse;
    s->tableLogReq = 11;
}

/* huffSort, huff0/compress.go:570-607.  nodes = s.nodes[1:] */
static void huff_sort(const uint32_t *count, unsigned symbolLen, node_t *nodes) {
    struct { uint32_t base, current; } rank[32];
    memset(rank, 0, sizeof(rank));
    for (unsigned i = 0; i < symbolLen; i++) {
        uint32_t r = orc_highbit32(count[i] + 1) & 31;
        rank[r].base++;
    }
    const int maxBitLength = 18 + 1;
    for (int n = maxBitLength; n > 0; n--) rank[n - 1].base += rank[n].base;
    for (int n = 0; n < maxBitLength; n++) rank[n].current = rank[n].base;
    for (unsigned n = 0; n < symbolLen; n++) {
        uint32_t c = count[n];
        uint32_t r = (orc_highbit32(c + 1) + 1) & 31;
        uint32_t pos = rank[r].current;
        rank[r].current++;
        node_t prev = nodes[(pos - 1) & HUFF_NODES_MASK];
        while (pos > rank[r].base && c > prev.count) {
            nodes[pos & HUFF_NODES_MASK] = prev;
            pos--;
            prev = nodes[(pos - 1) & HUFF_NODES_MASK];
        }
        node_t e = {c, 0, (uint8_t)n, 0};
        nodes[pos & HUFF_NODES_MASK] = e;
    }
}

/* setMaxHeight, huff0/compress.go:609-718 */
static unsigned set_max_height(node_t *huffNode, int lastNonNull, unsigned maxNbBits) {
    unsigned largestBits = huffNode[lastNonNull].nbBits;
    if (largestBits <= maxNbBits) return largestBits;
    int totalCost = 0;
    int baseCost = 1 << (largestBits - maxNbBits);
    uint32_t n = (uint32_t)lastNonNull;
    while (huffNode[n].nbBits > maxNbBits) {
        totalCost += baseCost - (1 << (largestBits - huffNode[n].nbBits));
        huffNode[n].nbBits = (uint8_t)maxNbBits;
        n--;
    }
    while (huffNode[n].nbBits == maxNbBits) n--;
    totalCost >>= (largestBits - maxNbBits);
    {
        const uint32_t noSymbol = 0xF0F0F0F0u;
        uint32_t rankLast[ORC_HUF_TABLELOG_MAX + 2];
        for (int i = 0; i < ORC_HUF_TABLELOG_MAX + 2; i++) rankLast[i] = noSymbol;
        {
            unsigned currentNbBits = maxNbBits;
            for (int pos = (int)n; pos >= 0; pos--) {
                if (huffNode[pos].nbBits >= currentNbBits) continue;
                currentNbBits = huffNode[pos].nbBits;
                rankLast[maxNbBits - currentNbBits] = (uint32_t)pos;
            }
        }
        while (totalCost > 0) {
            unsigned nBitsToDecrease = orc_highbit32((uint32_t)totalCost) + 1;
            for (; nBitsToDecrease > 1; nBitsToDecrease--) {
                uint32_t highPos = rankLast[nBitsToDecrease];
                uint32_t lowPos = rankLast[nBitsToDecrease - 1];
                if (highPos == noSymbol) continue;
                if (lowPos == noSymbol) break;
                uint32_t highTotal = huffNode[highPos].count;
                uint32_t lowTotal = 2 * huffNode[lowPos].count;
                if (highTotal <= lowTotal) break;
            }
            while (nBitsToDecrease <= ORC_HUF_TABLELOG_MAX && rankLast[nBitsToDecrease] == noSymbol)
                nBitsToDecrease++;
            totalCost -= 1 << (nBitsToDecrease - 1);
            if (rankLast[nBitsToDecrease - 1] == noSymbol)
                rankLast[nBitsToDecrease - 1] = rankLast[nBitsToDecrease];
            huffNode[rankLast[nBitsToDecrease]].nbBits++;
            if (rankLast[nBitsToDecrease] == 0) {
                rankLast[nBitsToDecrease] = noSymbol;
            } else {
                rankLast[nBitsToDecrease]--;
                if (huffNode[rankLast[nBitsToDecrease]].nbBits != maxNbBits - nBitsToDecrease)
                    rankLast[nBitsToDecrease] = noSymbol;
            }
        }
        while (totalCost < 0) {
            if (rankLast[1] == noSymbol) {
                while (huffNode[n].nbBits == maxNbBits) n--;
                huffNode[n + 1].nbBits--;
                rankLast[1] = n + 1;
                totalCost++;
                continue;
            }
            huffNode[rankLast[1] + 1].nbBits--;
            rankLast[1]++;
            totalCost++;
        }
    }
    return maxNbBits;
}

/* optimalTableLog + minTableLog, huff0/compress.go:417-447 */
static unsigned huf_optimal_tablelog(unsigned tableLogReq, size_t srcLen, unsigned symbolLen) {
    uint8_t tableLog = (uint8_t)tableLogReq;
    uint32_t minBitsSrc = orc_highbit32((uint32_t)srcLen) + 1;
    uint32_t minBitsSymbols = orc_highbit32((uint32_t)(symbolLen - 1)) + 2;
    uint8_t minBits = (uint8_t)(minBitsSrc < minBitsSymbols ? minBitsSrc : minBitsSymbols);
    uint8_t maxBitsSrc = (uint8_t)((uint8_t)orc_highbit32((uint32_t)(srcLen - 1)) - 1);
    if (maxBitsSrc < tableLog) tableLog = maxBitsSrc;
    if (minBits > tableLog) tableLog = minBits;
    if (tableLog < 5) tableLog = 5;
    if (tableLog > ORC_HUF_TABLELOG_MAX) tableLog = ORC_HUF_TABLELOG_MAX;
    return tableLog;
}

/* buildCTable, huff0/compress.go:457-567.  Exposed for table-parity tests. */
ORC_API int orc_huf_build_ctable(const uint32_t *count, unsigned symbolLen, size_t srcLen,
                                 unsigned tableLogReq, orc_huf_centry *ctable, unsigned *actualTableLog) {
    node_t nodesStore[HUFF_NODES_LEN + 2];
    memset(nodesStore, 0, sizeof(nodesStore));
    node_t *huffNode0 = nodesStore; /* s.nodes[0:] */
    node_t *huffNode = nodesStore + 1; /* s.nodes[1:] */
    unsigned tl = huf_optimal_tablelog(tableLogReq ? tableLogReq : 11, srcLen, symbolLen);
    huff_sort(count, symbolLen, huffNode);
    for (unsigned i = 0; i < symbolLen; i++) { ctable[i].val = 0; ctable[i].nBits = 0; }

    int16_t startNode = (int16_t)symbolLen;
    unsigned nonNullRank = symbolLen - 1;
    int16_t nodeNb = startNode;
    while (huffNode[nonNullRank].count == 0) nonNullRank--;

    int16_t lowS = (int16_t)nonNullRank;
    int16_t nodeRoot = (int16_t)(nodeNb + lowS - 1);
    int16_t lowN = nodeNb;
    huffNode[nodeNb].count = huffNode[lowS].count + huffNode[lowS - 1].count;
    huffNode[lowS].parent = (uint16_t)nodeNb;
    huffNode[lowS - 1].parent = (uint16_t)nodeNb;
    nodeNb++;
    lowS -= 2;
    for (int16_t n = nodeNb; n <= nodeRoot; n++) huffNode[n].count = 1u << 30;
    huffNode0[0].count = 1u << 31; /* fake entry, strong barrier */

    while (nodeNb <= nodeRoot) {
        int16_t n1, n2;
        if (huffNode0[lowS + 1].count < huffNode0[lowN + 1].count) { n1 = lowS; lowS--; }
        else { n1 = lowN; lowN++; }
        if (huffNode0[lowS + 1].count < huffNode0[lowN + 1].count) { n2 = lowS; lowS--; }
        else { n2 = lowN; lowN++; }
        huffNode[nodeNb].count = huffNode0[n1 + 1].count + huffNode0[n2 + 1].count;
        huffNode0[n1 + 1].parent = (uint16_t)nodeNb;
        huffNode0[n2 + 1].parent = (uint16_t)nodeNb;
        nodeNb++;
    }
    huffNode[nodeRoot].nbBits = 0;
    for (int16_t n = (int16_t)(nodeRoot - 1); n >= startNode; n--)
        huffNode[n].nbBits = (uint8_t)(huffNode[huffNode[n].parent].nbBits + 1);
    for (unsigned n = 0; n <= nonNullRank; n++)
        huffNode[n].nbBits = (uint8_t)(huffNode[huffNode[n].parent].nbBits + 1);
    unsigned maxNbBits = set_max_height(huffNode, (int)nonNullRank, tl);
    *actualTableLog = maxNbBits;
    if (maxNbBits > ORC_HUF_TABLELOG_MAX) return ORC_ERR_INTERNAL;

    uint16_t nbPerRank[ORC_HUF_TABLELOG_MAX + 1];
    uint16_t valPerRank[16];
    memset(nbPerRank, 0, sizeof(nbPerRank));
    memset(valPerRank, 0, sizeof(valPerRank));
    for (unsigned i = 0; i <= nonNullRank; i++) nbPerRank[huffNode[i].nbBits]++;
    {
        uint16_t min = 0;
        for (unsigned n = maxNbBits; n > 0; n--) {
            valPerRank[n] = min;
            min = (uint16_t)(min + nbPerRank[n]);
            min >>= 1;
        }
    }
    for (unsigned i = 0; i <= nonNullRank; i++) ctable[huffNode[i].symbol].nBits = huffNode[i].nbBits;
    for (unsigned n = 0; n < symbolLen; n++) {
        unsigned nbits = ctable[n].nBits & 15;
        uint16_t v = valPerRank[nbits];
        ctable[n].val = v;
        valPerRank[nbits] = (uint16_t)(v + 1);
    }
    return 0;
}

/* cTable.write, huff0/huff0.go:180-247: returns bytes appended or negative */
static int64_t ctable_write(const orc_huf_centry *c, unsigned symbolLen, unsigned huffLog, uint8_t *out,
                            size_t cap) {
    uint8_t bitsToWeight[ORC_HUF_TABLELOG_MAX + 2];
    uint8_t huffWeight[257];
    uint8_t maxSymbolValue = (uint8_t)(symbolLen - 1);
    uint32_t hist[256];
    memset(hist, 0, sizeof(hist));
    memset(huffWeight, 0, sizeof(huffWeight));
    bitsToWeight[0] = 0;
    for (unsigned n = 1; n < huffLog + 1; n++) bitsToWeight[n] = (uint8_t)(huffLog + 1 - n);
    for (unsigned n = 0; n < maxSymbolValue; n++) {
        uint8_t v = bitsToWeight[c[n].nBits] & 15;
        huffWeight[n] = v;
        hist[v]++;
    }
    if (maxSymbolValue >= 2) {
        uint32_t huffMaxCnt = 0;
        uint8_t huffMax = 0;
        for (unsigned i = 0; i < 16; i++) {
            if (hist[i] == 0) continue;
            huffMax = (uint8_t)i;
            if (hist[i] > huffMaxCnt) huffMaxCnt = hist[i];
        }
        uint8_t tmp[512];
        int64_t b = orc_fse_compress(huffWeight, maxSymbolValue, hist, (unsigned)huffMax + 1, huffMaxCnt, 6,
                                     tmp, sizeof(tmp));
        if (b >= 0 && b < (int64_t)(symbolLen >> 1)) {
            if ((size_t)b + 1 > cap) return ORC_ERR_DST_SMALL;
            out[0] = (uint8_t)b;
            memcpy(out + 1, tmp, (size_t)b);
            return b + 1;
        }
    }
    if (maxSymbolValue > (256 - 128)) return ORC_ERR_INCOMPRESSIBLE;
    size_t need = 1 + ((size_t)maxSymbolValue + 1) / 2;
    if (need > cap) return ORC_ERR_DST_SMALL;
    size_t o = 0;
    out[o++] = (uint8_t)(128 | (maxSymbolValue - 1));
    huffWeight[maxSymbolValue] = 0;
    for (unsigned n = 0; n < maxSymbolValue; n += 2) out[o++] = (uint8_t)((huffWeight[n] << 4) | huffWeight[n + 1]);
    return (int64_t)o;
}

/* compress1xDo, huff0/compress.go:233-265 */
static int64_t compress1x_do(const orc_huf_centry *ct, const uint8_t *src, size_t n, uint8_t *dst, size_t cap) {
    orc_bw bw;
    orc_bw_init(&bw, dst, cap);
    for (size_t i = n; i > 0; i--) { /* strictly last symbol first */
        orc_huf_centry e = ct[src[i - 1]];
        orc_bw_add(&bw, e.val, e.nBits);
    }
    orc_bw_close(&bw);
    if (bw.overflow) return ORC_ERR_DST_SMALL;
    return (int64_t)bw.pos;
}

/* compress4X, huff0/compress.go:269-302 */
static int64_t compress4x(const orc_huf_centry *ct, const uint8_t *src, size_t n, uint8_t *dst, size_t cap) {
    if (n < 12) return ORC_ERR_INCOMPRESSIBLE;
    size_t segmentSize = (n + 3) / 4;
    if (cap < 6) return ORC_ERR_DST_SMALL;
    memset(dst, 0, 6);
    size_t o = 6;
    for (int i = 0; i < 4; i++) {
        size_t todo = n > segmentSize ? segmentSize : n;
        int64_t w = compress1x_do(ct, src, todo, dst + o, cap - o);
        if (w < 0) return w;
        src += todo; n -= todo;
        if (w > 65535) return ORC_ERR_INCOMPRESSIBLE;
        if (i < 3) { dst[i * 2] = (uint8_t)w; dst[i * 2 + 1] = (uint8_t)(w >> 8); }
        o += (size_t)w;
    }
    return (int64_t)o;
}

static int estimate_size(const orc_huf_centry *c, const uint32_t *hist, unsigned len) { /* huff0.go:314 */
    uint32_t nbBits = 7;
    for (unsigned i = 0; i < len; i++) nbBits += (uint32_t)c[i].nBits * hist[i];
    return (int)(nbBits >> 3);
}

ORC_API int64_t orc_huf_compress(orc_huf_scratch *s, const uint8_t *in, size_t n, int fourStreams,
                                 uint8_t *out, size_t cap, int *reusedOut) {
    if (reusedOut) *reusedOut = 0;
    s->outTableLen = 0;
    if (n > ORC_HUF_BLOCK_MAX) return ORC_ERR_TOO_BIG; /* prepare(), huff0.go:135 */
    if (s->tableLogReq == 0) s->tableLogReq = 11;
    if (s->reuse == ORC_HUF_REUSE_NONE) s->prevLen = 0;

    /* countSimple, compress.go:351-385 */
    memset(s->count, 0, sizeof(s->count));
    for (size_t i = 0; i < n; i++) s->count[in[i]]++;
    uint32_t m = 0;
    int canReuse = s->prevLen > 0;
    for (unsigned i = 0; i < 256; i++) {
        uint32_t v = s->count[i];
        if (!v) continue;
        if (v > m) m = v;
        s->symbolLen = i + 1;
        if (s->prevLen > 0) {
            if (i >= s->prevLen) canReuse = 0;
            else if (s->prevTable[i].nBits == 0) canReuse = 0;
        }
    }
    size_t maxCount = m;
    size_t wantSize = n;
    if (s->wantLogLess > 0) wantSize -= wantSize >> s->wantLogLess;

    if (maxCount >= n) {
        if (n == 1) return ORC_ERR_INCOMPRESSIBLE;
        return ORC_ERR_USE_RLE;
    }
    if (maxCount == 1 || maxCount < (n >> 7)) return ORC_ERR_INCOMPRESSIBLE;
    if (s->reuse == ORC_HUF_REUSE_MUST && !canReuse) return ORC_ERR_INCOMPRESSIBLE;

#define COMPRESSOR(tbl, dstp, dcap)                                                                    \
    (fourStreams ? compress4x((tbl), in, n, (dstp), (dcap)) : compress1x_do((tbl), in, n, (dstp), (dcap)))

    if ((s->reuse == ORC_HUF_REUSE_PREFER || s->reuse == ORC_HUF_REUSE_MUST) && canReuse) {
        int64_t w = COMPRESSOR(s->prevTable, out, cap);
        if (w >= 0 && (size_t)w < wantSize) {
            if (reusedOut) *reusedOut = 1;
            return w;
        }
        if (s->reuse == ORC_HUF_REUSE_MUST) return ORC_ERR_INCOMPRESSIBLE;
        s->prevLen = 0;
    }

    int err = orc_huf_build_ctable(s->count, s->symbolLen, n, s->tableLogReq, s->ctable, &s->actualTableLog);
    if (err) return err;

    if (s->reuse == ORC_HUF_REUSE_ALLOW && canReuse) {
        int hSize = 0; /* len(s.Out) before the table is written, compress.go:115 */
        int oldSize = estimate_size(s->prevTable, s->count, s->symbolLen);
        int newSize = estimate_size(s->ctable, s->count, s->symbolLen);
        if (oldSize <= hSize + newSize || (size_t)(hSize + 12) >= wantSize) {
            int64_t w = COMPRESSOR(s->prevTable, out, cap);
            if (w < 0) return w;
            if ((size_t)w >= wantSize) return ORC_ERR_INCOMPRESSIBLE;
            if (reusedOut) *reusedOut = 1;
            return w;
        }
    }
    int64_t t = ctable_write(s->ctable, s->symbolLen, s->actualTableLog, out, cap);
    if (t < 0) return t;
    int64_t w = COMPRESSOR(s->ctable, out + t, cap - (size_t)t);
    if (w < 0) return w;
    if ((size_t)(t + w) >= wantSize) return ORC_ERR_INCOMPRESSIBLE;
    memcpy(s->prevTable, s->ctable, sizeof(orc_huf_centry) * s->symbolLen);
    s->prevLen = s->symbolLen;
    s->prevTableLog = s->actualTableLog;
    s->outTableLen = (size_t)t;
    return t + w;
#undef COMPRESSOR
}

/* Convenience one-shot used by the standalone huff0 parity tests (config 4):
 * fresh Scratch{Reuse: ReusePolicyNone}.  status<0 => ORC_ERR_*. */
ORC_API int64_t orc_huf_compress_oneshot(const uint8_t *in, size_t n, int fourStreams, unsigned wantLogLess,
                                         uint8_t *out, size_t cap, uint64_t *tableLen) {
    orc_huf_scratch s;
    orc_huf_scratch_init(&s, wantLogLess, ORC_HUF_REUSE_NONE);
    int reused = 0;
    int64_t r = orc_huf_compress(&s, in, n, fourStreams, out, cap, &reused);
    if (tableLen) *tableLen = s.outTableLen;
    return r;
}

/* ------------------------------- decoder ------------------------------- */

ORC_API int64_t orc_huf_read_table(orc_huf_dtable *d, const uint8_t *in, size_t n) {
    uint8_t huffWeight[258];
    unsigned symbolLen;
    d->loaded = 0;
    if (n <= 1) return ORC_ERR_CORRUPT; /* "input too small for table" */
    unsigned iSize = in[0];
    size_t consumed = 1;
    in++; n--;
    memset(huffWeight, 0, sizeof(huffWeight));
    if (iSize >= 128) {
        unsigned oSize = iSize - 127;
        iSize = (oSize + 1) / 2;
        if (iSize > n) return ORC_ERR_CORRUPT;
        for (unsigned k = 0; k < oSize; k += 2) {
            uint8_t v = in[k / 2];
            huffWeight[k] = v >> 4;
            huffWeight[k + 1] = v & 15;
        }
        symbolLen = oSize;
    } else {
        if (n < iSize) return ORC_ERR_CORRUPT;
        int64_t b = orc_fse_decompress(in, iSize, huffWeight, 256, 255);
        if (b < 0) return ORC_ERR_CORRUPT;
        if (b > 255) return ORC_ERR_CORRUPT; /* "output table too large" */
        symbolLen = (unsigned)b;
    }
    consumed += iSize;

    uint32_t rankStats[16];
    memset(rankStats, 0, sizeof(rankStats));
    uint32_t weightTotal = 0;
    for (unsigned i = 0; i < symbolLen; i++) {
        uint8_t v = huffWeight[i];
        if (v > ORC_HUF_TABLELOG_MAX) return ORC_ERR_CORRUPT; /* "weight too large" */
        rankStats[v & 15]++;
        weightTotal += (1u << (v & 15)) >> 1;
    }
    if (weightTotal == 0) return ORC_ERR_CORRUPT; /* "weights zero" */
    unsigned tableLog = orc_highbit32(weightTotal) + 1;
    if (tableLog > ORC_HUF_TABLELOG_MAX) return ORC_ERR_CORRUPT;
    {
        uint32_t total = 1u << tableLog;
        uint32_t rest = total - weightTotal;
        uint32_t verif = 1u << orc_highbit32(rest);
        uint32_t lastWeight = orc_highbit32(rest) + 1;
        if (verif != rest) return ORC_ERR_CORRUPT; /* "last value not power of two" */
        huffWeight[symbolLen] = (uint8_t)lastWeight;
        symbolLen++;
        rankStats[lastWeight]++;
    }
    if (rankStats[1] < 2 || (rankStats[1] & 1)) return ORC_ERR_CORRUPT; /* "min elt size, even check failed" */
    {
        uint32_t nextRankStart = 0;
        for (unsigned k = 1; k < tableLog + 1; k++) {
            uint32_t current = nextRankStart;
            nextRankStart += rankStats[k] << (k - 1);
            rankStats[k] = current;
        }
    }
    memset(d->dt, 0, sizeof(d->dt));
    for (unsigned k = 0; k < symbolLen; k++) {
        unsigned w = huffWeight[k];
        if (w == 0) continue;
        uint32_t length = (1u << w) >> 1;
        uint16_t entry = (uint16_t)((tableLog + 1 - w) | (k << 8));
        uint32_t r = rankStats[w];
        for (uint32_t i = 0; i < length; i++) d->dt[r + i] = entry;
        rankStats[w] = r + length;
    }
    d->actualTableLog = tableLog;
    d->loaded = 1;
    return (int64_t)consumed;
}

static int decode_stream(const orc_huf_dtable *d, const uint8_t *src, size_t n, uint8_t *dst, size_t count) {
    orc_br br;
    int err = orc_br_init(&br, src, n);
    if (err) return err;
    unsigned tl = d->actualTableLog;
    for (size_t i = 0; i < count; i++) {
        if (orc_br_finished(&br)) return ORC_ERR_CORRUPT; /* short output */
        uint16_t e = d->dt[orc_br_peek(&br, tl)];
        br.pos += (e & 0xff);
        dst[i] = (uint8_t)(e >> 8);
    }
    if (br.pos != br.total) return ORC_ERR_CORRUPT; /* bits remain / over-read */
    return 0;
}

ORC_API int orc_huf_decompress1x(const orc_huf_dtable *d, const uint8_t *src, size_t n, uint8_t *dst,
                                 size_t dstSize) {
    if (!d->loaded) return ORC_ERR_CORRUPT; /* "no table loaded" */
    return decode_stream(d, src, n, dst, dstSize);
}

ORC_API int orc_huf_decompress4x(const orc_huf_dtable *d, const uint8_t *src, size_t n, uint8_t *dst,
                                 size_t dstSize) {
    if (!d->loaded) return ORC_ERR_CORRUPT;
    if (n < 6 + 4) return ORC_ERR_CORRUPT; /* "input too small" */
    size_t dstEvery = (dstSize + 3) / 4;
    size_t start = 6;
    size_t doff = 0;
    for (int i = 0; i < 4; i++) {
        size_t length;
        if (i < 3) {
            length = (size_t)src[i * 2] | ((size_t)src[i * 2 + 1] << 8);
            if (start + length >= n) return ORC_ERR_CORRUPT; /* "truncated input (or invalid offset)" */
        } else {
            length = n - start;
        }
        size_t cnt;
        if (doff >= dstSize) cnt = 0;
        else cnt = (dstSize - doff < dstEvery) ? dstSize - doff : dstEvery;
        if (i < 3 && cnt != dstEvery) return ORC_ERR_CORRUPT;
        int err = decode_stream(d, src + start, length, dst + doff, cnt);
        if (err) return err;
        start += length;
        doff += cnt;
    }
    if (doff != dstSize) return ORC_ERR_CORRUPT;
    return 0;
}
