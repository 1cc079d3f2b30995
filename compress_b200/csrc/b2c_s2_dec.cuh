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

struct S2Head { uint32_t state, nrec, firstLit, dlen; };     // staged form: per block; state 0 = the staged kernels' job, 1 = the one-warp kernel's
struct S2DecParams {
    const uint8_t *src_base; uint64_t src_stride; const uint64_t *src_offsets; const uint32_t *src_sizes;
    uint8_t *dst_base; uint64_t dst_stride; const uint64_t *dst_offsets; const uint32_t *dst_caps; uint32_t dst_cap;
    int64_t *out_sizes;      // decoded bytes or negative error
    uint32_t nchunks;
    // staged form (nullptr: every block goes through the one-warp kernel)
    S2Head *heads;           // [nchunks]
    uint64_t *recs;          // block c: s2_rec_off(P, c), capacity src_sizes[c] / 3 + 1 records
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

// ------------------------------------------------------------------------------------------------ staged form
// The one-warp decoder above walks the tags with all 32 lanes in lockstep: the walk is serial, so 31 lanes idle through
// it.  The staged form splits the two halves of s2Decode (s2/decode_other.go:22-287) the way the staged zstd decoder does:
//   walk  one LANE per block: the tag stream -> one 8-byte record per element: literal length | copy length << 17 |
//         offset << 34 | gap << 51, where an element is "a literal run (possibly empty) followed by a copy (possibly none)"
//         and gap = bytes between the end of the previous element's literal bytes and the start of this element's (the
//         previous copy tag and this literal header), so a prefix sum places every literal run in the input
//   exec  one WARP per block: 32 elements per step, places by warp scans, literal bytes gathered from the step's window of
//         the input (staged in shared memory), copies in dependency waves (lz_exec_match_waves)
// Blocks of at most 64 KiB decoded / 128 KiB encoded with at most size/3 + 1 elements; anything else, and anything a stage
// does not like, is left to the one-warp kernel (state 1), which also produces the error values.
constexpr uint32_t S2S_MAX_DLEN = 65536, S2S_MAX_SLEN = 1u << 17, S2S_WIN = 1024;
B2C_DEV const uint8_t *s2_src(const S2DecParams &P, uint32_t c) { return P.src_base + (P.src_offsets ? P.src_offsets[c] : (uint64_t)c * P.src_stride); }
B2C_DEV uint64_t s2_rec_off(const S2DecParams &P, uint32_t c) {
    return (P.src_offsets ? P.src_offsets[c] : (uint64_t)c * P.src_stride) / 3 + c;
}
B2C_DEV uint64_t s2_rec_pack(uint32_t ll, uint32_t ml, uint32_t off, uint32_t gap) {
    return (uint64_t)ll | ((uint64_t)ml << 17) | ((uint64_t)off << 34) | ((uint64_t)gap << 51);
}
B2C_DEV void s2s_walk_lane(const S2DecParams &P, uint32_t c) {
    const uint8_t *src = s2_src(P, c);
    const uint32_t slen = P.src_sizes[c], cap = P.dst_caps ? P.dst_caps[c] : P.dst_cap;
    S2Head *hd = P.heads + c;
    uint64_t *recs = P.recs + s2_rec_off(P, c);
    const uint32_t recCap = slen / 3 + 1;         // (an element is at least a two-byte copy; three bytes on average holds for all but degenerate blocks)
#define S2LEG() do { hd->state = 1; return; } while (0)
    if (slen == 0 || slen > S2S_MAX_SLEN) S2LEG();
    uint64_t v = 0;
    uint32_t s = 0, shift = 0;
    for (;;) {
        if (s >= slen || s >= 5) S2LEG();
        const uint8_t b = src[s++];
        v |= (uint64_t)(b & 0x7f) << shift;
        if (b < 0x80) break;
        shift += 7;
    }
    if (v > S2S_MAX_DLEN || v > cap) S2LEG();
    const uint32_t dlen = (uint32_t)v;
    // Element i = [literal run of ll_i bytes][copy]; its record carries gap_i = the bytes between the end of the previous
    // element's literal data and the start of its own (the previous copy tag + its own literal header), so the execution
    // kernel finds every literal run with one prefix sum: L_i = C_(i-1) + gap_i, C_i = L_i + ll_i, C_(-1) = end of the varint.
    // One step of the loop takes one element: a literal tag (if there is one) and the copy tag behind it.  Tags and the four
    // bytes behind them arrive as aligned words through the read-only path and are decoded with selects: the 32 lanes of a
    // warp sit on 32 different tags, and a branch per tag kind would make the warp run every kind's path at every step.
    uint32_t nrec = 0, offset = 0, prevTb = 0;
    const uint32_t firstLit = s;
    const uint32_t smis = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3);
    const uint32_t *sw = reinterpret_cast<const uint32_t *>(src - smis);
    const uint32_t nsw = (slen + smis + 3) >> 2;                  // aligned words that hold block bytes
#define S2_FETCH(pos_, tag_, ext_)                                                                                          \
    do {                                                                                                                    \
        const uint32_t wi_ = ((pos_) + smis) >> 2, sh_ = (((pos_) + smis) & 3) * 8;                                          \
        const uint32_t w0_ = B2C_LDG(sw + wi_), w1_ = wi_ + 1 < nsw ? B2C_LDG(sw + wi_ + 1) : 0u,                            \
                       w2_ = wi_ + 2 < nsw ? B2C_LDG(sw + wi_ + 2) : 0u;                                                     \
        const uint32_t lo_ = __funnelshift_r(w0_, w1_, sh_), hi_ = __funnelshift_r(w1_, w2_, sh_);                          \
        (tag_) = lo_ & 0xff; (ext_) = (lo_ >> 8) | (hi_ << 24);          /* the tag byte; bytes pos+1 .. pos+4 */           \
    } while (0)
    while (s < slen) {
        uint32_t tag, ext;
        S2_FETCH(s, tag, ext);
        uint32_t ll = 0, hb = 0;
        bool copyFollows = true;
        if ((tag & 3) == 0) {
            const uint32_t x6 = tag >> 2;
            const uint32_t nb = x6 < 60 ? 0u : x6 - 59;                                      // extra length bytes (0 .. 4)
            hb = 1 + nb;
            if (s + hb > slen) S2LEG();
            const uint32_t x = nb == 0 ? x6 : (nb == 4 ? ext : (ext & ((1u << (8 * nb)) - 1)));
            if (x >= S2S_MAX_DLEN) S2LEG();
            ll = x + 1;
            if (ll > slen - (s + hb)) S2LEG();
            s += hb + ll;
            if (s < slen) { S2_FETCH(s, tag, ext); copyFollows = (tag & 3) != 0; }           // (another literal: no copy in this element)
            else copyFollows = false;                                                         // trailing literal run
        }
        uint32_t length = 0, off = 0, tb = 0;
        if (copyFollows) {
            // copy1 (2 bytes; offset 0 = repeat with 0 .. 3 extra length bytes), copy2 (3 bytes), copy4 (5 bytes)
            const uint32_t kind = tag & 3, x6 = tag >> 2, l3 = x6 & 7;
            const uint32_t toff1 = ((tag & 0xe0) << 3) | (ext & 0xff);
            const bool rep = kind == 1 && toff1 == 0;
            const uint32_t rb = (rep && l3 >= 5) ? l3 - 4 : 0u;                               // repeat: extra length bytes
            tb = kind == 1 ? 2 + rb : (kind == 2 ? 3u : 5u);
            if (s + tb > slen) S2LEG();
            const uint32_t rext = (ext >> 8) & (rb == 0 ? 0u : ((1u << (8 * rb)) - 1));       // bytes s+2 .. of a repeat
            const uint32_t len1 = (rb == 0 ? l3 : rext + (rb == 1 ? 4u : (rb == 2 ? 256u : 65536u))) + 4;
            length = kind == 1 ? len1 : 1 + x6;
            if (!rep) offset = kind == 1 ? toff1 : (kind == 2 ? (ext & 0xffff) : ext);
            if (offset == 0 || offset > S2S_MAX_DLEN || length > S2S_MAX_DLEN) S2LEG();
            off = offset;
            s += tb;
        }
        if (nrec >= recCap) S2LEG();
        recs[nrec++] = s2_rec_pack(ll, length, off, prevTb + hb);
        prevTb = tb;
    }
#undef S2_FETCH
    hd->state = 0; hd->nrec = nrec; hd->firstLit = firstLit; hd->dlen = dlen;
#undef S2LEG
}

// stg: S2S_WIN + 8 bytes of shared memory for this warp
B2C_DEV void s2s_exec_warp(const S2DecParams &P, uint32_t c, uint8_t *stg, unsigned lane) {
    S2Head *hd = P.heads + c;
    if (hd->state != 0) return;
    const uint8_t *src = s2_src(P, c);
    uint8_t *out = P.dst_base + (P.dst_offsets ? P.dst_offsets[c] : (uint64_t)c * P.dst_stride);
    const uint32_t slen = P.src_sizes[c], nrec = hd->nrec, dlen = hd->dlen;
    const uint64_t *recs = P.recs + s2_rec_off(P, c);
    uint32_t d = 0, litSrc = hd->firstLit;
    bool bad = false;
    uint64_t nextRec = lane < nrec ? recs[lane] : 0;
    for (uint32_t base = 0; base < nrec; base += 32) {
        const uint32_t cnt = nrec - base < 32 ? nrec - base : 32;
        const bool mine = lane < cnt;
        const uint64_t rec = mine ? nextRec : 0;
        if (base + 32 < nrec) nextRec = (base + 32 + lane < nrec) ? recs[base + 32 + lane] : 0;
        const uint32_t myLL = (uint32_t)(rec & 0x1ffff), myML = (uint32_t)((rec >> 17) & 0x1ffff), myMO = (uint32_t)((rec >> 34) & 0x1ffff);
        const uint32_t myGap = (uint32_t)((rec >> 51) & 15);
        const uint32_t lenIncl = warp_scan_incl(myLL + myML), llIncl = warp_scan_incl(myLL), srcIncl = warp_scan_incl(myLL + myGap);
        const uint32_t myOut = d + (lenIncl - (myLL + myML)), myDst = myOut + myLL;
        const uint32_t myLit = litSrc + (srcIncl - myLL);                // L_i = C_start + sum_(j <= i)(gap_j + ll_j) - ll_i
        bool err = false;
        if (mine) {
            if (lenIncl > dlen - d) err = true;                          // (d <= dlen always)
            else if (myLL > slen || myLit > slen - myLL) err = true;
            else if (myML && (myMO == 0 || myMO > myDst)) err = true;
        }
        if (__any_sync(FULLMASK, err)) { bad = true; break; }
        // ---- literal runs: every literal byte of the step finds its element by a binary search over the lanes' counts; the
        // bytes come from the step's window of the input [litSrc, litSrc + sum(ll + gap)), staged in shared memory when it is
        // small enough (else straight from global memory)
        {
            const uint32_t totLit = __shfl_sync(FULLMASK, llIncl, 31);
            const uint32_t win = __shfl_sync(FULLMASK, srcIncl, 31);
            const bool useS = win <= S2S_WIN && totLit > 0;
            if (useS) {
                const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(src + litSrc) & 3);
                const uint32_t *gw = reinterpret_cast<const uint32_t *>(src + litSrc - mis);
                const uint32_t nw = (win + mis + 3) >> 2;                 // aligned words holding window bytes
                uint32_t *sw = reinterpret_cast<uint32_t *>(stg);
                for (uint32_t i = lane; i < nw; i += 32) sw[i] = gw[i];
                __syncwarp();
                const uint8_t *wb = stg + mis;                             // window byte j at wb[j]
                for (uint32_t k0 = 0; k0 < totLit; k0 += 32) {
                    const uint32_t k = k0 + lane;
                    uint32_t owner = 0;
#pragma unroll
                    for (int st = 16; st > 0; st >>= 1) {
                        const uint32_t cc = owner + st - 1;
                        const uint32_t v = __shfl_sync(FULLMASK, llIncl, (int)(cc & 31));
                        if (cc < 32 && v <= k) owner += st;
                    }
                    const uint32_t oIncl = __shfl_sync(FULLMASK, llIncl, (int)(owner & 31)), oLL = __shfl_sync(FULLMASK, myLL, (int)(owner & 31));
                    const uint32_t oOut = __shfl_sync(FULLMASK, myOut, (int)(owner & 31)), oLit = __shfl_sync(FULLMASK, myLit, (int)(owner & 31));
                    if (k < totLit) { const uint32_t j = k - (oIncl - oLL); out[oOut + j] = wb[oLit - litSrc + j]; }
                }
                __syncwarp();
            } else if (totLit) {
                // long runs: each by the whole warp
                for (int l = 0; l < 32; l++) {
                    const uint32_t n = __shfl_sync(FULLMASK, myLL, l), from = __shfl_sync(FULLMASK, myLit, l), to = __shfl_sync(FULLMASK, myOut, l);
                    for (uint32_t i = lane; i < n; i += 32) out[to + i] = src[from + i];
                }
                __syncwarp();
            }
        }
        __syncwarp();
        lz_exec_match_waves(out, mine && myML != 0, myDst, myMO, myML, lane);
        d += __shfl_sync(FULLMASK, lenIncl, 31);
        litSrc += __shfl_sync(FULLMASK, srcIncl, 31);
    }
    if (!bad && d != dlen) bad = true;
    __syncwarp();
    if (lane == 0) {
        if (bad) hd->state = 1;
        else P.out_sizes[c] = (int64_t)dlen;
    }
}

B2C_DEV void s2_decode_warp(const S2DecParams &P, uint32_t warpGlobal, uint32_t totalWarps) {
    const unsigned lane = threadIdx.x & 31;
    for (uint32_t c = warpGlobal; c < P.nchunks; c += totalWarps) {
        const uint8_t *src = P.src_base + (P.src_offsets ? P.src_offsets[c] : (uint64_t)c * P.src_stride);
        uint8_t *dst = P.dst_base + (P.dst_offsets ? P.dst_offsets[c] : (uint64_t)c * P.dst_stride);
        const uint32_t cap = P.dst_caps ? P.dst_caps[c] : P.dst_cap;
        __syncwarp();
        if (P.heads && P.heads[c].state == 0) continue;       // decoded by the staged kernels
        const int64_t r = s2_decode_block(src, P.src_sizes[c], dst, cap, lane);
        __syncwarp();
        if (lane == 0) P.out_sizes[c] = r;
    }
}

#ifndef B2C_EMU
extern "C" __global__ void __launch_bounds__(32) b2c_s2_walk_kernel(S2DecParams P) {
    const uint32_t c = blockIdx.x * 32 + threadIdx.x;
    if (c < P.nchunks) s2s_walk_lane(P, c);
}
extern "C" __global__ void __launch_bounds__(S2DEC_WARPS * 32) b2c_s2_exec_kernel(S2DecParams P) {
    __shared__ __align__(16) uint8_t stg[S2DEC_WARPS][S2S_WIN + 16];
    const uint32_t c = blockIdx.x * S2DEC_WARPS + (threadIdx.x >> 5);
    if (c < P.nchunks) s2s_exec_warp(P, c, stg[threadIdx.x >> 5], threadIdx.x & 31);
}
extern "C" __global__ void __launch_bounds__(S2DEC_WARPS * 32) b2c_s2_decode_kernel(S2DecParams P) {
    s2_decode_warp(P, blockIdx.x * S2DEC_WARPS + (threadIdx.x >> 5), gridDim.x * S2DEC_WARPS);
}
#endif

}  // namespace b2c
