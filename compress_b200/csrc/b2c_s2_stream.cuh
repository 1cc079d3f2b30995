// compress_b200/csrc/b2c_s2_stream.cuh -- the S2 / Snappy framing format on the device.
//
// Reference: s2.Writer's block goroutine (s2/writer.go:395-470) writes, per block, a 4-byte chunk header (type, 24-bit
// length), the masked CRC32-C of the UNCOMPRESSED block (crc(), s2/s2.go:118-126) and either the encoded block (uvarint
// length + tags) or -- when the block encoder reports "not compressible" -- the raw bytes; a stream starts with the
// identifier chunk (magicChunk / magicChunkSnappy, s2/s2.go:78-82).  s2.Reader (s2/reader.go:249-420) walks the chunks,
// decodes, and compares the checksum.
//
// Here: the block encoders are the parse kernels of b2c_lz.cuh; this file adds the checksum (one warp per block: every
// lane runs the table-driven CRC over its own contiguous piece, lane 0 folds the 32 pieces with the GF(2) identity
// crc(A || B) = crc(A) * x^(8 |B|) mod P  xor  crc(B)), and the placement of headers and payloads in the output stream.
#pragma once
#include "b2c_common.cuh"

namespace b2c {

constexpr uint32_t CRC32C_POLY = 0x82F63B78u;      // Castagnoli, reflected (crc32.MakeTable(crc32.Castagnoli), s2/s2.go:112)

#ifdef B2C_EMU
#define B2C_HDS static inline
#else
#define B2C_HDS static __host__ __device__ inline
#endif

// a(x) * b(x) mod P, operands and result in the reflected representation (bit 31 = x^0)
B2C_HDS uint32_t crc_multmodp(uint32_t a, uint32_t b) {
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) { p ^= b; if ((a & (m - 1)) == 0) break; }
        m >>= 1;
        b = (b & 1) ? (b >> 1) ^ CRC32C_POLY : b >> 1;
    }
    return p;
}
// x^(8 * nbytes) mod P (square-and-multiply; nbytes < 2^28)
B2C_HDS uint32_t crc_xpow8(uint32_t nbytes) {
    uint32_t sq = 1u << 30;                       // x^1
    sq = crc_multmodp(sq, sq); sq = crc_multmodp(sq, sq); sq = crc_multmodp(sq, sq);     // x^8
    uint32_t p = 1u << 31;                        // x^0
    for (uint32_t n = nbytes; n; n >>= 1) {
        if (n & 1) p = crc_multmodp(sq, p);
        sq = crc_multmodp(sq, sq);
    }
    return p;
}
B2C_HDS uint32_t crc_mask(uint32_t c) { return ((c >> 15) | (c << 17)) + 0xa282ead8u; }   // s2/s2.go:124

// the byte table (256 words of shared memory), filled by the calling threads
B2C_DEV void crc32c_fill_table(uint32_t *tab, unsigned tid, unsigned nthreads) {
    for (uint32_t i = tid; i < 256; i += nthreads) {
        uint32_t c = i;
#pragma unroll
        for (int k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ CRC32C_POLY : c >> 1;
        tab[i] = c;
    }
}
// CRC32-C (standard pre / post conditioning) of p[0, n) by one warp; result on every lane
B2C_DEV uint32_t crc32c_warp(const uint8_t *p, uint32_t n, const uint32_t *tab, unsigned lane) {
    const uint32_t seg = (((n + 31) / 32) + 3) & ~3u;          // bytes per lane (a multiple of 4)
    const uint32_t lo = lane * seg < n ? lane * seg : n, hi = lo + seg < n ? lo + seg : n;
    uint32_t c = 0xffffffffu;
    uint32_t i = lo;
    // head to a 4-byte boundary, words, tail
    for (; i < hi && ((reinterpret_cast<uintptr_t>(p) + i) & 3); i++) c = tab[(c ^ p[i]) & 0xff] ^ (c >> 8);
    for (; i + 4 <= hi; i += 4) {
        const uint32_t w = *reinterpret_cast<const uint32_t *>(p + i);
        c = tab[(c ^ w) & 0xff] ^ (c >> 8);
        c = tab[(c ^ (w >> 8)) & 0xff] ^ (c >> 8);
        c = tab[(c ^ (w >> 16)) & 0xff] ^ (c >> 8);
        c = tab[(c ^ (w >> 24)) & 0xff] ^ (c >> 8);
    }
    for (; i < hi; i++) c = tab[(c ^ p[i]) & 0xff] ^ (c >> 8);
    c ^= 0xffffffffu;
    if (hi == lo) c = 0;                                         // an empty piece: crc("") = 0
    // fold: acc = crc(piece 0); acc = acc * x^(8 |piece i|) ^ crc(piece i)
    const uint32_t pfull = crc_xpow8(seg);
    uint32_t acc = 0;
    for (int l = 0; l < 32; l++) {
        const uint32_t cl = __shfl_sync(FULLMASK, c, l), ll = __shfl_sync(FULLMASK, hi - lo, l);
        if (l == 0) acc = cl;
        else if (ll == seg) acc = crc_multmodp(pfull, acc) ^ cl;
        else if (ll) acc = crc_multmodp(crc_xpow8(ll), acc) ^ cl;
    }
    return acc;
}

// One block of a stream: where its uncompressed bytes are, what the block encoder made of it.
struct S2StreamParams {
    const uint8_t *src; uint64_t total;       // the whole input
    uint32_t block;                           // block size of the stream (<= 65536)
    const uint8_t *slots; uint64_t slot_stride; const int64_t *enc_sizes;     // s2.Encode-style output per block of the sub-batch
    uint32_t *crc;                            // masked checksum per block of the sub-batch
    int64_t *piece;                           // bytes the block occupies in the stream (header + checksum + payload)
    const uint64_t *offsets;                  // exclusive scan of piece
    const uint64_t *base; uint32_t k;         // bytes of the earlier sub-batches (base[k])
    uint8_t *dst; uint64_t cap;
    uint32_t c0, m;                           // blocks [c0, c0 + m) of the stream
    uint32_t snappy;
    int32_t *err;                             // set when the stream does not fit dst
};
B2C_DEV uint32_t s2s_block_len(const S2StreamParams &P, uint32_t c) {
    const uint64_t off = (uint64_t)(P.c0 + c) * P.block;
    return (uint32_t)(P.total - off < P.block ? P.total - off : P.block);
}
// s2.Encode's "store as one literal" form (s2/encode.go:60-70 via the kernels' store path): uvarint + literal header + n bytes.
// The framing format wants those blocks as uncompressed chunks (encodeBlock returned 0, s2/writer.go:428-436).
B2C_DEV bool s2s_is_stored(uint32_t n, int64_t enc) {
    const uint32_t hdr = n < 128 ? 1u : (n < 16384 ? 2u : 3u);
    const uint32_t lh = n == 0 ? 0u : (n <= 60 ? 1u : (n <= 256 ? 2u : 3u));
    return enc == (int64_t)(hdr + lh + n);
}
// warp per block: checksum + the size of its chunk
B2C_DEV void s2s_crc_block(const S2StreamParams &P, uint32_t c, const uint32_t *tab, unsigned lane) {
    const uint32_t n = s2s_block_len(P, c);
    const uint32_t crc = crc32c_warp(P.src + (uint64_t)(P.c0 + c) * P.block, n, tab, lane);
    if (lane == 0) {
        P.crc[c] = crc_mask(crc);
        const int64_t e = P.enc_sizes[c];
        P.piece[c] = e < 0 ? e : (int64_t)(8 + (s2s_is_stored(n, e) ? n : (uint32_t)e));
    }
}
// CTA per block: chunk header, checksum, payload (+ the stream identifier in front of block 0)
B2C_DEV void s2s_place_block(const S2StreamParams &P, uint32_t c, unsigned tid, unsigned nthreads) {
    const int64_t pc = P.piece[c];
    if (pc < 0) { if (tid == 0) *P.err = (int32_t)pc; return; }
    const uint64_t at = 10 + P.base[P.k] + P.offsets[c];
    if (at + (uint64_t)pc > P.cap) { if (tid == 0) *P.err = -4; return; }
    const uint32_t n = s2s_block_len(P, c);
    const bool stored = s2s_is_stored(n, P.enc_sizes[c]);
    uint8_t *d = P.dst + at;
    if (tid == 0) {
        if (P.c0 + c == 0) {
            const char *magic = P.snappy ? "\xff\x06\x00\x00sNaPpY" : "\xff\x06\x00\x00S2sTwO";
            for (int i = 0; i < 10; i++) P.dst[i] = (uint8_t)magic[i];
        }
        const uint32_t ln = (uint32_t)pc - 4;              // checksum + payload
        d[0] = stored ? 0x01 : 0x00; d[1] = (uint8_t)ln; d[2] = (uint8_t)(ln >> 8); d[3] = (uint8_t)(ln >> 16);
        const uint32_t crc = P.crc[c];
        d[4] = (uint8_t)crc; d[5] = (uint8_t)(crc >> 8); d[6] = (uint8_t)(crc >> 16); d[7] = (uint8_t)(crc >> 24);
    }
    const uint8_t *s = stored ? P.src + (uint64_t)(P.c0 + c) * P.block : P.slots + (uint64_t)c * P.slot_stride;
    coop_copy(d + 8, s, (uint32_t)pc - 8, tid, nthreads);
}

// ---- reading a stream: the chunk table is made on the host (a serial walk over 4-byte headers); the device decodes the
// compressed blocks (b2c_s2_dec.cuh), copies the uncompressed ones and checks every block's checksum
struct S2StreamBlock { uint64_t src_off, dst_off; uint32_t src_len, dst_len, type, crc; };
B2C_DEV void s2s_verify_block(const S2StreamBlock *blk, const uint8_t *src, uint8_t *dst, const int64_t *dec_sizes, int32_t *status,
                              uint32_t c, const uint32_t *tab, unsigned lane) {
    const S2StreamBlock b = blk[c];
    if (b.type == 1) {      // uncompressed chunk: copy
        for (uint32_t i = lane; i < b.dst_len; i += 32) dst[b.dst_off + i] = src[b.src_off + i];
        __syncwarp();
    } else if (dec_sizes[c] != (int64_t)b.dst_len) {
        if (lane == 0) status[c] = dec_sizes[c] < 0 ? (int32_t)dec_sizes[c] : -5;
        return;
    }
    const uint32_t crc = crc_mask(crc32c_warp(dst + b.dst_off, b.dst_len, tab, lane));
    if (lane == 0) status[c] = crc == b.crc ? 0 : -9;        // ErrCRC
}

#ifndef B2C_EMU
constexpr int S2S_WARPS = 4;
extern "C" __global__ void __launch_bounds__(S2S_WARPS * 32) b2c_s2_stream_crc_kernel(S2StreamParams P) {
    __shared__ uint32_t tab[256];
    crc32c_fill_table(tab, threadIdx.x, blockDim.x);
    __syncthreads();
    const unsigned lane = threadIdx.x & 31;
    for (uint32_t c = blockIdx.x * S2S_WARPS + (threadIdx.x >> 5); c < P.m; c += gridDim.x * S2S_WARPS) s2s_crc_block(P, c, tab, lane);
}
extern "C" __global__ void __launch_bounds__(256) b2c_s2_stream_place_kernel(S2StreamParams P) {
    for (uint32_t c = blockIdx.x; c < P.m; c += gridDim.x) s2s_place_block(P, c, threadIdx.x, blockDim.x);
}
extern "C" __global__ void __launch_bounds__(S2S_WARPS * 32) b2c_s2_stream_verify_kernel(const S2StreamBlock *blk, const uint8_t *src, uint8_t *dst,
                                                                                     const int64_t *dec_sizes, int32_t *status, uint32_t n) {
    __shared__ uint32_t tab[256];
    crc32c_fill_table(tab, threadIdx.x, blockDim.x);
    __syncthreads();
    const unsigned lane = threadIdx.x & 31;
    for (uint32_t c = blockIdx.x * S2S_WARPS + (threadIdx.x >> 5); c < n; c += gridDim.x * S2S_WARPS)
        s2s_verify_block(blk, src, dst, dec_sizes, status, c, tab, lane);
}
#endif

}  // namespace b2c
