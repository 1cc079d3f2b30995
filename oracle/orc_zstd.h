/*
 * oracle/orc_zstd.h -- zstd block/frame encoder + decoder oracle interface.
 * TEST INFRASTRUCTURE ONLY -- see orc_common.h.
 */
#ifndef ORC_ZSTD_H
#define ORC_ZSTD_H
#include "orc_common.h"
#include "orc_fse.h"
#include "orc_huff0.h"

#define ORC_ZSTD_MAX_BLOCK (128 << 10) /* maxCompressedBlockSize, zstd/blockdec.go:40 */
#define ORC_ZSTD_MINMATCH 3            /* zstdMinMatch, zstd/zstd.go:37 */
#define ORC_ZSTD_MAX_MATCHLEN 131074   /* maxMatchLength, zstd/enc_fast.go:18 */

typedef struct {
    uint32_t litLen;
    uint32_t matchLen; /* match length - 3 */
    uint32_t offset;   /* distance + 3, or 1..3 = repeat code */
    uint8_t llCode, mlCode, ofCode;
} orc_seq; /* seq, zstd/seqdec.go:13-21 */

typedef struct {
    unsigned symbolLen;
    unsigned actualTableLog;
    orc_fse_ctable ct;
    int maxCount;
    int useRLE, preDefined, reUsed;
    uint8_t rleVal, maxBits;
    uint32_t count[256];
    int16_t norm[256];
} orc_fse_enc; /* fseEncoder, zstd/fse_encoder.go:23-40 */

typedef struct {
    size_t size;
    uint8_t *literals;
    size_t nlit, lit_cap;
    orc_seq *seqs;
    size_t nseq, seq_cap;
    orc_fse_enc store[6];
    orc_fse_enc *llEnc, *ofEnc, *mlEnc, *llPrev, *ofPrev, *mlPrev;
    orc_huf_scratch litEnc;
    uint32_t recentOffsets[3], prevRecentOffsets[3];
    int last;
    size_t extraLits;
    uint8_t *tmp; /* huff0 output staging (Scratch.Out) */
} orc_blockenc; /* blockEnc, zstd/blockenc.go:17-33 */

orc_blockenc *orc_blockenc_new(void);
void orc_blockenc_free(orc_blockenc *b);
void orc_blockenc_init_new_encode(orc_blockenc *b); /* initNewEncode, blockenc.go:77-82 */
void orc_blockenc_reset(orc_blockenc *b);           /* reset(nil), blockenc.go:87-98 */
void orc_blockenc_add_literals(orc_blockenc *b, const uint8_t *p, size_t n);
void orc_blockenc_add_seq(orc_blockenc *b, uint32_t litLen, uint32_t matchLenMinus3, uint32_t offset);
/* blockEnc.encode (blockenc.go:481): appends to dst[*pos..cap). returns 0 or negative. */
int orc_blockenc_encode(orc_blockenc *b, const uint8_t *org, size_t orgLen, int raw, int rawAllLits,
                        uint8_t *dst, size_t cap, size_t *pos);

/* match finders (fill b->literals / b->seqs) */
void orc_enc_fast_nohist(orc_blockenc *b, const uint8_t *src, size_t n);   /* enc_fast.go:294 */
void orc_enc_dfast_nohist(orc_blockenc *b, const uint8_t *src, size_t n);  /* enc_dfast.go:372 */

/* Encoder.EncodeAll (zstd/encoder.go:722-839). level: 1 fastest, 2 default, 3 better. returns size or negative. */
int64_t orc_zstd_encode_all(const uint8_t *src, size_t n, int level, int crc, uint8_t *dst, size_t cap);
size_t orc_zstd_max_encoded_size(size_t n, int level, int crc); /* MaxEncodedSize, encoder.go:843 */

/* Decoder.DecodeAll (zstd/decoder.go:319): returns decoded size or negative ORC_ERR_* */
int64_t orc_zstd_decode_all(const uint8_t *src, size_t n, uint8_t *dst, size_t cap);

#endif
