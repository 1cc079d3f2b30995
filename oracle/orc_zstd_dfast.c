/* placeholder until the dfast restatement lands */
#include "orc_zstd.h"
void orc_dfast_encode_all_blocks(orc_blockenc *blk, const uint8_t *src, size_t n, size_t blockSize,
                                 uint8_t *dst, size_t cap, size_t *pos, int *err) {
    (void)blk; (void)src; (void)n; (void)blockSize; (void)dst; (void)cap; (void)pos;
    *err = ORC_ERR_UNSUPPORTED;
}
