/*
 * oracle/orc_fse.h -- FSE/tANS table construction shared by the standalone FSE
 * codec (fse/compress.go, fse/decompress.go) and zstd's private copy
 * (zstd/fse_encoder.go, zstd/fse_decoder.go).  The two Go copies contain the
 * same arithmetic; the oracle states it once.  TEST INFRASTRUCTURE ONLY.
 */
#ifndef ORC_FSE_H
#define ORC_FSE_H
#include "orc_common.h"

#define ORC_FSE_MAX_TABLELOG 12
#define ORC_FSE_MAX_TABLESIZE (1 << ORC_FSE_MAX_TABLELOG)

typedef struct {
    uint32_t deltaNbBits;
    int32_t deltaFindState;
    uint8_t outBits;
} orc_symtt; /* symbolTransform, zstd/fse_encoder.go:49 / fse/compress.go:329 */

typedef struct {
    uint16_t stateTable[ORC_FSE_MAX_TABLESIZE];
    uint8_t tableSymbol[ORC_FSE_MAX_TABLESIZE];
    orc_symtt tt[256];
    int zeroBits;
} orc_fse_ctable;

/* normalizeCount + normalizeCount2 (fse/compress.go:510-683, zstd/fse_encoder.go:259-427).
 * returns 0 or ORC_ERR_INTERNAL ("weight < 1"). */
int orc_fse_normalize(const uint32_t *count, unsigned symbolLen, uint32_t length, unsigned tableLog,
                      int16_t *norm);
/* writeCount (fse/compress.go:208-327, zstd/fse_encoder.go:488-598). out needs 2 bytes of slack.
 * returns bytes written or negative. */
int64_t orc_fse_write_ncount(const int16_t *norm, unsigned symbolLen, unsigned tableLog, uint8_t *out,
                             size_t cap);
/* buildCTable (fse/compress.go:371-468, zstd/fse_encoder.go:102-204). */
int orc_fse_build_ctable(const int16_t *norm, unsigned symbolLen, unsigned tableLog, orc_fse_ctable *ct);
/* readNCount (fse/decompress.go:48-168, zstd/fse_decoder.go:52-184).
 * maxSymbol: loop guard of the zstd variant (pass 255 for the standalone codec).
 * returns bytes consumed or negative. */
int64_t orc_fse_read_ncount(const uint8_t *in, size_t len, unsigned maxSymbol, unsigned absMaxTableLog,
                            int16_t *norm, unsigned *symbolLen, unsigned *tableLog);

typedef struct {
    uint16_t newState;
    uint8_t symbol;
    uint8_t nbBits;
} orc_fse_dsym;
/* buildDtable (fse/decompress.go:193-258, zstd/fse_decoder_generic.go:11-72) */
int orc_fse_build_dtable(const int16_t *norm, unsigned symbolLen, unsigned tableLog, orc_fse_dsym *dt);

static inline uint32_t orc_fse_table_step(uint32_t tableSize) { return (tableSize >> 1) + (tableSize >> 3) + 3; }

/* cState.init (zstd/fse_encoder.go:676-690, fse/compress.go:86-95) */
static inline uint16_t orc_fse_cstate_init(const orc_fse_ctable *ct, orc_symtt first) {
    uint32_t nbBitsOut = (first.deltaNbBits + (1u << 15)) >> 16;
    int32_t im = (int32_t)((nbBitsOut << 16) - first.deltaNbBits);
    int32_t lu = (im >> nbBitsOut) + first.deltaFindState;
    return ct->stateTable[lu];
}
#endif
