/*
 * oracle/orc_huff0.h -- huff0 (zstd-flavoured Huffman) oracle interface.
 * TEST INFRASTRUCTURE ONLY -- see orc_common.h.
 */
#ifndef ORC_HUFF0_H
#define ORC_HUFF0_H
#include "orc_common.h"

#define ORC_HUF_BLOCK_MAX ((1 << 18) - 1) /* huff0.BlockSizeMax, huff0/huff0.go:27 */
#define ORC_HUF_TABLELOG_MAX 11

enum { /* huff0.ReusePolicy, huff0/huff0.go:44-61 */
    ORC_HUF_REUSE_ALLOW = 0,
    ORC_HUF_REUSE_PREFER = 1,
    ORC_HUF_REUSE_NONE = 2,
    ORC_HUF_REUSE_MUST = 3
};

typedef struct {
    uint16_t val;
    uint8_t nBits;
} orc_huf_centry; /* cTableEntry, huff0/compress.go:449 */

typedef struct {
    uint32_t count[256];
    unsigned symbolLen;
    unsigned actualTableLog;
    unsigned tableLogReq; /* Scratch.TableLog; 0 -> 11 */
    unsigned wantLogLess; /* Scratch.WantLogLess */
    int reuse;            /* Scratch.Reuse */
    orc_huf_centry ctable[256];
    orc_huf_centry prevTable[256];
    unsigned prevLen; /* len(prevTable) */
    unsigned prevTableLog;
    /* outputs of the last compress call */
    size_t outTableLen; /* len(OutTable); 0 when the table was reused */
} orc_huf_scratch;

void orc_huf_scratch_init(orc_huf_scratch *s, unsigned wantLogLess, int reuse);

/* huff0.Compress1X / Compress4X (huff0/compress.go:14,27).  returns total bytes
 * written (table + data) or a negative ORC_ERR_*; *reused mirrors reUsed. */
int64_t orc_huf_compress(orc_huf_scratch *s, const uint8_t *in, size_t n, int fourStreams, uint8_t *out,
                         size_t cap, int *reused);

/* decoder side */
typedef struct {
    uint16_t dt[1 << ORC_HUF_TABLELOG_MAX]; /* dEntrySingle: nbits | sym<<8 */
    unsigned actualTableLog;
    int loaded;
} orc_huf_dtable;

/* huff0.ReadTable (huff0/decompress.go:29-166): returns bytes consumed or negative */
int64_t orc_huf_read_table(orc_huf_dtable *d, const uint8_t *in, size_t n);
/* Decoder.Decompress1X / 4X: decodes exactly dstSize symbols; stream must be consumed exactly */
int orc_huf_decompress1x(const orc_huf_dtable *d, const uint8_t *src, size_t n, uint8_t *dst, size_t dstSize);
int orc_huf_decompress4x(const orc_huf_dtable *d, const uint8_t *src, size_t n, uint8_t *dst, size_t dstSize);

#endif
