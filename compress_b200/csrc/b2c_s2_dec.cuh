// compress_b200/csrc/b2c_s2_dec.cuh -- S2 / Snappy block decoder for sm_100a.
//
// One warp decodes one block (what one s2.Decode call gets, s2/decode.go:58-86): uvarint length, then the tag
// stream of s2Decode (s2/decode_other.go:22-287; asm twins s2/decode_amd64.s, decode_arm64.s) including the S2
// repeat-offset extension (tagCopy1 with offset 0, decode_other.go:74-101).  Every lane walks the tags
// redundantly (uniform control flow, broadcast loads); literal runs and copies are moved by all 32 lanes, an
// overlapping copy (offset < length) with a modulo gather.  Errors follow the reference: any malformed stream is
// ErrCorrupt, a block longer than the destination is reported as too large.
#pragma once
#include "b2c_common.cuh"

namespace b2c {

constexpr int S2DEC_WARPS = 4;
enum { S2_ERR_DST = -4, S2_ERR_CORRUPT = -5 };

struct S2DecParams {
    const uint8_t *src_base; uint64_t src_stride; const uint64_t *src_offsets; const uint32_t *src_sizes;
    uint8_t *dst_base; uint64_t dst_stride; const uint64_t *dst_offsets; const uint32_t *dst_caps; uint32_t dst_cap;
    int64_t *out_sizes;      // decoded bytes or negative error
    uint32_t nchunks;
};

B2C_DEV int64_t s2_decode_block(const uint8_t *src, uint32_t slen, uint8_t *dst, uint32_t cap, unsigned lane) {
    // decodedLen (s2/decode.go:36-49): binary.Uvarint, at most 5 bytes, value <= 2^32-1
    uint64_t v = 0;
    uint32_t s = 0, shift = 0;
    for (;;) {
        if (s >= slen || s >= 10) return S2_ERR_CORRUPT;
        const uint8_t b = src[s++];
        if (b < 0x80) {
            if (s == 10 && b > 1) return S2_ERR_CORRUPT;
            v |= (uint64_t)b << shift;
            break;
        }
        v |= (uint64_t)(b & 0x7f) << shift;
        shift += 7;
    }
    if (s > 5 || v > 0xffffffffull) return S2_ERR_CORRUPT;
    if (v > cap) return S2_ERR_DST;
    const uint32_t dlen = (uint32_t)v;
    uint32_t d = 0, offset = 0;
    while (s < slen) {
        const uint32_t tag = src[s];
        uint32_t length;
        if ((tag & 3) == 0) {
            uint32_t x = tag >> 2;
            if (x < 60) s += 1;
            else if (x == 60) { s += 2; if (s > slen) return S2_ERR_CORRUPT; x = src[s - 1]; }
            else if (x == 61) { s += 3; if (s > slen) return S2_ERR_CORRUPT; x = (uint32_t)src[s - 2] | (uint32_t)src[s - 1] << 8; }
            else if (x == 62) {
                s += 4; if (s > slen) return S2_ERR_CORRUPT;
                x = (uint32_t)src[s - 3] | (uint32_t)src[s - 2] << 8 | (uint32_t)src[s - 1] << 16;
            } else {
                s += 5; if (s > slen) return S2_ERR_CORRUPT;
                x = (uint32_t)src[s - 4] | (uint32_t)src[s - 3] << 8 | (uint32_t)src[s - 2] << 16 | (uint32_t)src[s - 1] << 24;
            }
            const uint64_t l64 = (uint64_t)x + 1;
            if (l64 > dlen - d || l64 > slen - s) return S2_ERR_CORRUPT;
            length = (uint32_t)l64;
            for (uint32_t k = lane; k < length; k += 32) dst[d + k] = src[s + k];
            d += length; s += length;
            __syncwarp();
            continue;
        }
        if ((tag & 3) == 1) {
            s += 2;
            if (s > slen) return S2_ERR_CORRUPT;
            length = (tag >> 2) & 7;
            const uint32_t toffset = ((tag & 0xe0) << 3) | src[s - 1];
            if (toffset == 0) {            // repeat: keep the last offset, extended length codes
                if (length == 5) { s += 1; if (s > slen) return S2_ERR_CORRUPT; length = (uint32_t)src[s - 1] + 4; }
                else if (length == 6) {
                    s += 2; if (s > slen) return S2_ERR_CORRUPT;
                    length = ((uint32_t)src[s - 2] | (uint32_t)src[s - 1] << 8) + (1 << 8);
                } else if (length == 7) {
                    s += 3; if (s > slen) return S2_ERR_CORRUPT;
                    length = ((uint32_t)src[s - 3] | (uint32_t)src[s - 2] << 8 | (uint32_t)src[s - 1] << 16) + (1 << 16);
                }
            } else offset = toffset;
            length += 4;
        } else if ((tag & 3) == 2) {
            s += 3;
            if (s > slen) return S2_ERR_CORRUPT;
            length = 1 + (tag >> 2);
            offset = (uint32_t)src[s - 2] | (uint32_t)src[s - 1] << 8;
        } else {
            s += 5;
            if (s > slen) return S2_ERR_CORRUPT;
            length = 1 + (tag >> 2);
            offset = (uint32_t)src[s - 4] | (uint32_t)src[s - 3] << 8 | (uint32_t)src[s - 2] << 16 | (uint32_t)src[s - 1] << 24;
        }
        // offset > 2^31 is "offset <= 0" for the reference's int on 32-bit and simply > d on 64-bit: corrupt either way
        if (offset == 0 || d < offset || length > dlen - d) return S2_ERR_CORRUPT;
        const uint8_t *from = dst + d - offset;
        if (offset >= length || offset >= 32) {
            for (uint32_t k0 = 0; k0 < length; k0 += 32) {
                const uint32_t k = k0 + lane;
                if (k < length) dst[d + k] = from[k];
                if (offset < length) __syncwarp();
            }
        } else {
            for (uint32_t k = lane; k < length; k += 32) dst[d + k] = from[k % offset];
        }
        d += length;
        __syncwarp();
    }
    if (d != dlen) return S2_ERR_CORRUPT;
    return (int64_t)dlen;
}

B2C_DEV void s2_decode_warp(const S2DecParams &P, uint32_t warpGlobal, uint32_t totalWarps) {
    const unsigned lane = threadIdx.x & 31;
    for (uint32_t c = warpGlobal; c < P.nchunks; c += totalWarps) {
        const uint8_t *src = P.src_base + (P.src_offsets ? P.src_offsets[c] : (uint64_t)c * P.src_stride);
        uint8_t *dst = P.dst_base + (P.dst_offsets ? P.dst_offsets[c] : (uint64_t)c * P.dst_stride);
        const uint32_t cap = P.dst_caps ? P.dst_caps[c] : P.dst_cap;
        __syncwarp();
        const int64_t r = s2_decode_block(src, P.src_sizes[c], dst, cap, lane);
        __syncwarp();
        if (lane == 0) P.out_sizes[c] = r;
    }
}

#ifndef B2C_EMU
extern "C" __global__ void __launch_bounds__(S2DEC_WARPS * 32) b2c_s2_decode_kernel(S2DecParams P) {
    s2_decode_warp(P, blockIdx.x * S2DEC_WARPS + (threadIdx.x >> 5), gridDim.x * S2DEC_WARPS);
}
#endif

}  // namespace b2c
