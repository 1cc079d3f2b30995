/*
 * oracle/orc_xxh64.c -- XXH64, restating zstd/internal/xxhash/xxhash.go:62-160
 * (Digest.Write / Sum64).  The zstd frame checksum is the low 32 bits, appended
 * little-endian (zstd/enc_base.go:34-38).  TEST INFRASTRUCTURE ONLY.
 */
#include "orc_common.h"

#define P1 11400714785074694791ull
#define P2 14029467366897019727ull
#define P3 1609587929392839161ull
#define P4 9650029242287828579ull
#define P5 2870177450012600261ull

static inline uint64_t rotl(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline uint64_t xround(uint64_t acc, uint64_t in) { /* xxhash.go round() */
    acc += in * P2;
    acc = rotl(acc, 31);
    return acc * P1;
}
static inline uint64_t merge_round(uint64_t acc, uint64_t v) {
    v = xround(0, v);
    acc ^= v;
    return acc * P1 + P4;
}

ORC_API uint64_t orc_xxh64(const void *data, size_t len, uint64_t seed) {
    const uint8_t *p = (const uint8_t *)data;
    const uint8_t *end = p + len;
    uint64_t h;
    if (len >= 32) {
        uint64_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
        const uint8_t *limit = end - 32;
        do {
            v1 = xround(v1, orc_ld64(p));
            v2 = xround(v2, orc_ld64(p + 8));
            v3 = xround(v3, orc_ld64(p + 16));
            v4 = xround(v4, orc_ld64(p + 24));
            p += 32;
        } while (p <= limit);
        h = rotl(v1, 1) + rotl(v2, 7) + rotl(v3, 12) + rotl(v4, 18);
        h = merge_round(h, v1);
        h = merge_round(h, v2);
        h = merge_round(h, v3);
        h = merge_round(h, v4);
    } else {
        h = seed + P5;
    }
    h += (uint64_t)len;
    while (p + 8 <= end) {
        uint64_t k1 = xround(0, orc_ld64(p));
        h ^= k1;
        h = rotl(h, 27) * P1 + P4;
        p += 8;
    }
    if (p + 4 <= end) {
        h ^= (uint64_t)orc_ld32(p) * P1;
        h = rotl(h, 23) * P2 + P3;
        p += 4;
    }
    while (p < end) {
        h ^= (uint64_t)(*p) * P5;
        h = rotl(h, 11) * P1;
        p++;
    }
    h ^= h >> 33;
    h *= P2;
    h ^= h >> 29;
    h *= P3;
    h ^= h >> 32;
    return h;
}
