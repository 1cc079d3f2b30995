// compress_b200/csrc/b2c_zstd_enc.cuh -- fused zstd "SpeedFastest" chunk encoder for sm_100a.
//
// One persistent CTA (1024 threads, one per SM) turns one independent <= 64 KiB chunk into one
// complete zstd frame (what zstd.Encoder.EncodeAll does per call, zstd/encoder.go:722-839) or a
// bare block.  It replaces, for the GPU path, the reference's
//   zstd/enc_fast.go:294-531  fastEncoder.EncodeNoHist      (match finding)
//   zstd/blockenc.go:481-826  blockEnc.encode               (entropy stage, bit-exact here)
//   zstd/frameenc.go:25-92    frameHeader.appendTo
//   zstd/internal/xxhash      XXH64 frame checksum
//
// B200-first design (not a port of the serial Go loop):
//   * the chunk is staged into shared memory with one TMA bulk copy (cp.async.bulk + mbarrier);
//   * match finding is split over 31 warps, each parsing its own 2 KiB sub-range 32 positions at a
//     time: every lane hashes its position (hash6, zstd/hash.go), probes a private per-warp table
//     (most recent occurrence in the sub-range) and a CTA-wide table holding the EARLIEST
//     occurrence of each hash in the whole chunk (built by a race-free min-reduction pre-pass, so
//     every candidate lies before the probing position and the output is deterministic), verifies
//     4 bytes in shared memory, and the warp picks matches greedily with ballot/ffs and extends
//     them 128 bytes per step with ballot.  The 32nd warp computes XXH64 meanwhile;
//   * literals are gathered into shared memory, Huffman/FSE tables are built with the
//     reference's exact tie-breaking, the three tANS chains are walked speculatively in parallel
//     (b2c_seq.cuh) and all bitstreams are packed at prefix-summed bit offsets into a staging
//     buffer that is written back with one coalesced copy.
// Sequences differ from the reference's greedy parse (the parse is position-parallel), the
// entropy stage is byte-identical to blockEnc.encode for the same (literals, sequences).
#pragma once
#include "b2c_common.cuh"
#include "b2c_fse.cuh"
#include "b2c_huff.cuh"
#include "b2c_seq.cuh"

namespace b2c {

constexpr int ENC_NT = 1024;              // threads per CTA
constexpr int ENC_NW = ENC_NT / 32;       // 32 warps
constexpr int ENC_NPARSE = ENC_NW - 1;    // 31 parsing warps, warp 31 hashes (XXH64)
constexpr int ENC_EBITS = 15;             // earliest-occurrence table: 32 Ki x u16
constexpr int ENC_LBITS = 10;             // per-warp recent table: 1 Ki x u16
constexpr uint32_t ENC_MAX_CHUNK = 1u << 16;
constexpr uint32_t ENC_SRC_BYTES = ENC_MAX_CHUNK + 128;   // chunk + zero padding
constexpr uint32_t ENC_E_BYTES = (1u << ENC_EBITS) * 2;
constexpr uint32_t ENC_L_BYTES = ENC_NPARSE * (1u << ENC_LBITS) * 2;
constexpr uint32_t ENC_SEGCAP = 560;      // max records per parse warp: ceil(2144 / 4) + slack
constexpr uint32_t ENC_MAXSEQ = 16384 + 64;
constexpr uint32_t ENC_CODES_SMEM_CAP = 8192;

// per-CTA global scratch layout (bytes)
constexpr uint32_t SCR_REC = 0;                                   // parse records: NPARSE x SEGCAP x 8
constexpr uint32_t SCR_LL = SCR_REC + ENC_NPARSE * ENC_SEGCAP * 8;  // u16[MAXSEQ]
constexpr uint32_t SCR_ML = SCR_LL + ENC_MAXSEQ * 2;              // u16[MAXSEQ]
constexpr uint32_t SCR_OF = SCR_ML + ENC_MAXSEQ * 2;              // u32[MAXSEQ]
constexpr uint32_t SCR_STB = SCR_OF + ENC_MAXSEQ * 4;             // u16[3][MAXSEQ]
constexpr uint32_t SCR_CODES = SCR_STB + ENC_MAXSEQ * 6;          // u8[3][MAXSEQ] (only if nseq > smem cap)
constexpr uint32_t ENC_SCRATCH_BYTES = ((SCR_CODES + ENC_MAXSEQ * 3 + 255) / 256) * 256;

enum { ENC_FLAG_CRC = 1, ENC_FLAG_FRAME = 2 };

struct EncShared {
    uint32_t cnt[32];       // records per parse warp
    uint32_t tail[32];      // trailing literal bytes of each sub-range
    uint32_t sumLL[32];     // literal bytes inside each sub-range (incl. tail)
    uint32_t seqBase[32];
    uint32_t litBase[32];
    uint32_t carry[32];     // literal bytes carried into the warp's first sequence
    uint32_t nseq, nlit, n, sub;
    uint32_t kind;          // 0 compressed, 1 raw block, 2 RLE block
    uint32_t rleLen;
    uint32_t litMode;       // 0 raw, 1 RLE, 2 compressed
    uint32_t litPayload;    // compressed literal payload bytes
    uint32_t lhSize;
    uint32_t pos;           // running output byte offset inside the staging buffer
    uint32_t seqBitsBase;   // bit offset where the sequence bitstream starts (inside staging)
    uint32_t outBytes;
    uint32_t fhSize;
    uint64_t xxh;
    uint64_t mbar;
    HufWork hw;
    SeqWork sw;
};

constexpr uint32_t ENC_SMEM_SRC = 0;
constexpr uint32_t ENC_SMEM_E = ENC_SMEM_SRC + ENC_SRC_BYTES;
constexpr uint32_t ENC_SMEM_L = ENC_SMEM_E + ENC_E_BYTES;
constexpr uint32_t ENC_SMEM_SH = ENC_SMEM_L + ENC_L_BYTES;
constexpr uint32_t ENC_SMEM_BYTES = ENC_SMEM_SH + ((sizeof(EncShared) + 15) / 16) * 16;

struct ZstdEncParams {
    const uint8_t *const *srcs;   // per-chunk source pointers (device) or nullptr
    const uint8_t *src_base;      // used when srcs == nullptr: chunk i at src_base + i * src_stride
    uint64_t src_stride;
    const uint32_t *src_sizes;    // per-chunk sizes (<= 65536); nullptr => all chunks are src_size_all
    uint32_t src_size_all;
    uint8_t *const *dsts;         // per-chunk destination pointers or nullptr
    uint8_t *dst_base;            // used when dsts == nullptr
    uint64_t dst_stride;
    uint32_t dst_cap;             // capacity of every destination slot
    int64_t *out_sizes;           // bytes written per chunk, negative = error
    uint32_t nchunks;
    uint32_t flags;
    uint8_t *scratch;             // gridDim.x * ENC_SCRATCH_BYTES
    // optional debug dump (tests): per chunk {nseq, nlit, kind, litMode} + seq triples + literals
    uint32_t *dbg_hdr;            // [nchunks][4]
    uint32_t *dbg_seqs;           // [nchunks][dbg_seq_cap][3]
    uint8_t *dbg_lits;            // [nchunks][65536]
    uint32_t dbg_seq_cap;
    unsigned long long *dbg_cycles;  // optional [nchunks][16][32] per-warp arrival stamps (clock64)
};

#ifdef B2C_EMU
#define B2C_PHASE(k) do { } while (0)
#else
// Every warp's lane 0 stamps clock64 right after each barrier.  BAR.SYNC does not block at issue (the wait is
// deferred to the next access of barrier-protected state), so the stamp captures the warp's ARRIVAL time at the
// preceding barrier; the barrier's release time is the maximum over warps (tools/phase_times.py).
#define B2C_PHASE(k)                                                                                   \
    do {                                                                                               \
        if (P.dbg_cycles && (threadIdx.x & 31) == 0)                                                   \
            P.dbg_cycles[((uint64_t)chunk * 16 + (k)) * 32 + (threadIdx.x >> 5)] = (unsigned long long)clock64(); \
    } while (0)
#endif

// zstd/hash.go:27 hashLen(u, 32, 6): top 32 bits of ((u << 16) * prime6bytes)
B2C_DEV uint32_t enc_hash6(uint64_t u) {
    const uint64_t prime6 = 227718039650203ull;
    return (uint32_t)(((u << 16) * prime6) >> 32);
}

// length of the common prefix of src[a..limitA) and src[b..], cooperative over the warp
B2C_DEV uint32_t warp_match_len(const uint8_t *src, uint32_t a, uint32_t b, uint32_t limitA) {
    unsigned lane = lane_id();
    uint32_t rem = limitA - a;
    uint32_t k = 0;
    for (;;) {
        uint32_t pos = k + 4 * lane;
        uint32_t eq = 0;
        if (pos < rem) {
            uint32_t x = ld32u(src, a + pos) ^ ld32u(src, b + pos);
            eq = x ? (uint32_t)(__ffs((int)x) - 1) >> 3 : 4u;
            uint32_t nv = rem - pos;
            if (eq > nv) eq = nv;
        }
        unsigned stop = __ballot_sync(FULLMASK, eq < 4);
        if (stop) {
            int fl = __ffs((int)stop) - 1;
            uint32_t e = __shfl_sync(FULLMASK, eq, fl);
            return k + 4 * (uint32_t)fl + e;
        }
        k += 128;
    }
}
// number of equal bytes going backwards: src[a-1-i] == src[b-1-i], i < maxBack
B2C_DEV uint32_t warp_match_back(const uint8_t *src, uint32_t a, uint32_t b, uint32_t maxBack) {
    unsigned lane = lane_id();
    uint32_t k = 0;
    for (;;) {
        uint32_t i = k + lane;
        bool ok = (i < maxBack) && (src[a - 1 - i] == src[b - 1 - i]);
        unsigned stop = __ballot_sync(FULLMASK, !ok);
        if (stop) return k + (uint32_t)(__ffs((int)stop) - 1);
        k += 32;
    }
}

// XXH64 of src[0..n) (8-byte aligned shared memory), computed by lanes 0..3 of one warp
B2C_DEV uint64_t warp_xxh64(const uint8_t *src, uint32_t n) {
    const uint64_t P1 = 11400714785074694791ull, P2 = 14029467366897019727ull, P3 = 1609587929392839161ull,
                   P4 = 9650029242287828579ull, P5 = 2870177450012600261ull;
    unsigned lane = lane_id();
    uint64_t h = 0;
    uint32_t p = 0;
    if (n >= 32) {
        uint64_t v = (lane == 0) ? P1 + P2 : (lane == 1) ? P2 : (lane == 2) ? 0ull : (0ull - P1);
        uint32_t stripes = n / 32;
        if (lane < 4) {
            const uint64_t *q = reinterpret_cast<const uint64_t *>(src) + lane;
            for (uint32_t i = 0; i < stripes; i++) {
                uint64_t in = q[4 * i];
                v += in * P2;
                v = (v << 31) | (v >> 33);
                v *= P1;
            }
        }
        p = stripes * 32;
        uint64_t v1 = __shfl_sync(FULLMASK, v, 0), v2 = __shfl_sync(FULLMASK, v, 1), v3 = __shfl_sync(FULLMASK, v, 2),
                 v4 = __shfl_sync(FULLMASK, v, 3);
        h = ((v1 << 1) | (v1 >> 63)) + ((v2 << 7) | (v2 >> 57)) + ((v3 << 12) | (v3 >> 52)) + ((v4 << 18) | (v4 >> 46));
#define XMERGE(vv)                                                                                     \
    do {                                                                                               \
        uint64_t t_ = (vv) * P2; t_ = (t_ << 31) | (t_ >> 33); t_ *= P1;                               \
        h ^= t_; h = h * P1 + P4;                                                                      \
    } while (0)
        XMERGE(v1); XMERGE(v2); XMERGE(v3); XMERGE(v4);
#undef XMERGE
    } else {
        h = P5;
    }
    h += (uint64_t)n;
    while (p + 8 <= n) {
        uint64_t k1 = ld64u(src, p) * P2; k1 = (k1 << 31) | (k1 >> 33); k1 *= P1;
        h ^= k1; h = ((h << 27) | (h >> 37)) * P1 + P4;
        p += 8;
    }
    if (p + 4 <= n) {
        h ^= (uint64_t)ld32u(src, p) * P1;
        h = ((h << 23) | (h >> 41)) * P2 + P3;
        p += 4;
    }
    while (p < n) {
        h ^= (uint64_t)src[p] * P5;
        h = ((h << 11) | (h >> 53)) * P1;
        p++;
    }
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return h;
}

// ------------------------------------------------------------------------------------------------
// The chunk encoder.  All 1024 threads call.  smem = dynamic shared memory base (16-byte aligned).
B2C_DEV void zstd_encode_chunk(uint8_t *smem, const ZstdEncParams &P, uint32_t chunk, uint8_t *scratch) {
    const unsigned tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    uint8_t *src = smem + ENC_SMEM_SRC;
    uint16_t *E = reinterpret_cast<uint16_t *>(smem + ENC_SMEM_E);
    uint16_t *Lall = reinterpret_cast<uint16_t *>(smem + ENC_SMEM_L);
    EncShared *sh = reinterpret_cast<EncShared *>(smem + ENC_SMEM_SH);
    uint8_t *lit = smem + ENC_SMEM_E;                                      // after the parse
    uint8_t *stage = smem + ENC_SMEM_SRC;                                  // after the literal gather
    uint32_t *whist = reinterpret_cast<uint32_t *>(smem + ENC_SMEM_L);     // [32][256] literal histograms
    uint8_t *codesS = smem + ENC_SMEM_L + 32 * 256 * 4;                    // u8[3][CODES_SMEM_CAP]

    const uint8_t *gsrc = P.srcs ? P.srcs[chunk] : P.src_base + (uint64_t)chunk * P.src_stride;
    uint8_t *gdst = P.dsts ? P.dsts[chunk] : P.dst_base + (uint64_t)chunk * P.dst_stride;
    const uint32_t n = P.src_sizes ? P.src_sizes[chunk] : P.src_size_all;
    const bool frame = (P.flags & ENC_FLAG_FRAME) != 0;
    const bool crc = frame && (P.flags & ENC_FLAG_CRC) != 0;

    if (n > ENC_MAX_CHUNK) {
        if (tid == 0) P.out_sizes[chunk] = -3;  // too big for this kernel
        return;
    }

    B2C_PHASE(0);
    // ---------------------------------------------------------------- P0: stage the chunk
    {
#ifndef B2C_EMU
        bool bulk = ((reinterpret_cast<uintptr_t>(gsrc) & 15) == 0) && ((n & 15) == 0) && n > 0;
        if (bulk) {
            if (tid == 0) {
                mbar_init(&sh->mbar, 1);
                mbar_fence_init();
            }
            __syncthreads();
            if (tid == 0) {
                // order the previous chunk's generic-proxy accesses to this buffer before the async-proxy write
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&sh->mbar, n);
                tma_load_1d(src, gsrc, n, &sh->mbar);
            }
        } else
#endif
        {
            for (uint32_t i = tid; i < n; i += ENC_NT) src[i] = gsrc[i];
        }
        // zero padding behind the chunk (unaligned 8-byte loads may look at up to n+11)
        for (uint32_t i = n + tid; i < ((n + 128 + 15) & ~15u) && i < ENC_SRC_BYTES; i += ENC_NT) src[i] = 0;
        for (uint32_t i = tid; i < (1u << ENC_EBITS) / 2; i += ENC_NT) reinterpret_cast<uint32_t *>(E)[i] = 0xffffffffu;
#ifndef B2C_EMU
        if (bulk) mbar_wait(&sh->mbar, 0);
#endif
        if (tid == 0) { sh->n = n; sh->kind = 0; sh->sw.err = 0; }
        __syncthreads();
#ifndef B2C_EMU
        if (bulk && tid == 0) { asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&sh->mbar))); }
#endif
    }
    const uint32_t sub = (((n + ENC_NPARSE - 1) / ENC_NPARSE) + 31) & ~31u;  // sub-range size
    B2C_PHASE(1);

    // ---------------------------------------------------------------- P1: earliest-occurrence table
    const uint32_t npos = (n >= 8) ? n - 7 : 0;  // positions with 8 readable bytes
    {
        // round 1: plain stores, highest positions first so low positions tend to land last
        for (int32_t k = (int32_t)((npos + ENC_NT - 1) / ENC_NT) - 1; k >= 0; k--) {
            uint32_t p = (uint32_t)k * ENC_NT + tid;
            if (p < npos) E[enc_hash6(ld64u(src, p)) >> (32 - ENC_EBITS)] = (uint16_t)p;
        }
        __syncthreads();
        // fix-up rounds: a slot only ever decreases, so this converges to the exact minimum
        for (;;) {
            int changed = 0;
            for (uint32_t p = tid; p < npos; p += ENC_NT) {
                uint32_t h = enc_hash6(ld64u(src, p)) >> (32 - ENC_EBITS);
                if (E[h] > p) { E[h] = (uint16_t)p; changed = 1; }
            }
            if (!__syncthreads_or(changed)) break;
        }
    }

    B2C_PHASE(2);
    // ---------------------------------------------------------------- P2: parse (warps 0..30), XXH64 (warp 31)
    if (w == ENC_NPARSE) {
        if (crc) {
            uint64_t h = warp_xxh64(src, n);
            if (lane == 0) sh->xxh = h;
        }
    } else {
        uint16_t *L = Lall + w * (1u << ENC_LBITS);
        for (uint32_t i = lane; i < (1u << ENC_LBITS) / 2; i += 32) reinterpret_cast<uint32_t *>(L)[i] = 0xffffffffu;
        __syncwarp();
        const uint32_t b0 = w * sub;
        const uint32_t e0 = (b0 + sub < n) ? b0 + sub : n;
        uint32_t nrec = 0, sumML = 0;
        uint32_t nextEmit = b0;
        if (b0 < n) {
            uint2 *rec = reinterpret_cast<uint2 *>(scratch + SCR_REC) + w * ENC_SEGCAP;
            uint32_t cur = b0, ownNew = 0, rep0 = 0;
            while (cur < e0) {
                uint32_t p = cur + lane;
                bool valid = (p < e0) && (p < npos);
                uint64_t cv = valid ? ld64u(src, p) : 0;
                uint32_t h32 = enc_hash6(cv);
                uint32_t hL = h32 >> (32 - ENC_LBITS), hE = h32 >> (32 - ENC_EBITS);
                uint32_t candL = 0xffff, candE = 0xffff;
                if (valid) { candL = L[hL]; candE = E[hE]; }
                __syncwarp();
                // insert the window's positions; the highest position wins a slot (deterministic)
                uint32_t rel = p - b0;
                if (valid) L[hL] = (uint16_t)rel;
                __syncwarp();
                for (;;) {
                    bool lose = valid && (L[hL] < rel);
                    if (!__any_sync(FULLMASK, lose)) break;
                    if (lose) L[hL] = (uint16_t)rel;
                    __syncwarp();
                }
                // verify candidates: recent-in-sub-range, earliest-in-chunk, then the repeat offset
                int32_t q = -1;
                if (valid) {
                    uint32_t c32 = (uint32_t)cv;
                    if (candL != 0xffff && ld32u(src, b0 + candL) == c32) q = (int32_t)(b0 + candL);
                    else if (candE < p && ld32u(src, candE) == c32) q = (int32_t)candE;
                    else if (ownNew >= 1 && p >= rep0 && ld32u(src, p - rep0) == c32) q = (int32_t)(p - rep0);
                }
                unsigned mask = __ballot_sync(FULLMASK, q >= 0);
                uint32_t next = cur + 32;
                uint32_t from = 0;
                while (true) {
                    unsigned m = (from < 32) ? (mask & (0xffffffffu << from)) : 0u;
                    if (m == 0) break;
                    int f = __ffs((int)m) - 1;
                    uint32_t pf = cur + (uint32_t)f;
                    uint32_t qf = (uint32_t)__shfl_sync(FULLMASK, q, f);
                    uint32_t len = warp_match_len(src, pf, qf, e0);
                    if (len < 4) { mask &= ~(1u << f); continue; }
                    uint32_t off = pf - qf;
                    uint32_t maxBack = pf - nextEmit; if (qf < maxBack) maxBack = qf;
                    uint32_t back = warp_match_back(src, pf, qf, maxBack);
                    uint32_t s = pf - back;
                    len += back;
                    bool isrep = (ownNew >= 1) && (off == rep0) && (s > nextEmit);
                    if (lane == 0) {
                        // record: x = litLen | (matchLen-3) << 16 ; y = dist (0 = repeat) | matchStart << 16
                        rec[nrec] = make_uint2((s - nextEmit) | ((len - 3) << 16), (isrep ? 0u : off) | (s << 16));
                    }
                    nrec++;
                    sumML += len;
                    if (!isrep) { rep0 = off; ownNew++; }
                    nextEmit = s + len;
                    if (nextEmit >= cur + 32) { next = nextEmit; break; }
                    from = nextEmit - cur;
                }
                cur = next;
            }
        }
        if (lane == 0) {
            sh->cnt[w] = nrec;
            sh->tail[w] = (b0 < n) ? e0 - nextEmit : 0;
            sh->sumLL[w] = (b0 < n) ? (e0 - b0) - sumML : 0;
        }
    }
    __syncthreads();
    B2C_PHASE(3);

    // ---------------------------------------------------------------- P3: global sequence/literal layout
    if (tid == 0) {
        uint32_t nseq = 0, nlit = 0, carry = 0;
        for (int v = 0; v < ENC_NPARSE; v++) {
            sh->seqBase[v] = nseq; sh->litBase[v] = nlit; sh->carry[v] = carry;
            nseq += sh->cnt[v]; nlit += sh->sumLL[v];
            if (sh->cnt[v]) carry = sh->tail[v]; else carry += sh->tail[v];
        }
        sh->nseq = nseq; sh->nlit = nlit;
        // blockEnc.encode early decisions (blockenc.go:481-503)
        uint32_t kind = 0;
        if (nseq == 0) kind = 1;  // encodeLits(..., rawAllLits=true) => raw block
        else {
            int saved = (int)n - (int)nlit - (int)(n >> 6);
            if (saved < 16) kind = 1;
        }
        sh->kind = kind;
        // frame header size (frameenc.go:25-92); single segment when 1024 < n <= window
        uint32_t fh = 0;
        if (frame) {
            bool single = n > 1024;
            fh = 4 + 1 + (single ? 0 : 1);
            if (n >= 256) fh += (n >= 65536 + 256) ? 4 : 2; else if (single) fh += 1;
        }
        sh->fhSize = fh;
    }
    __syncthreads();
    const uint32_t nseq = sh->nseq, nlit = sh->nlit, fh = sh->fhSize;
    uint16_t *seqLL = reinterpret_cast<uint16_t *>(scratch + SCR_LL);
    uint16_t *seqML = reinterpret_cast<uint16_t *>(scratch + SCR_ML);
    uint32_t *seqOF = reinterpret_cast<uint32_t *>(scratch + SCR_OF);
    uint8_t *codes = (nseq <= ENC_CODES_SMEM_CAP) ? codesS : scratch + SCR_CODES;
    const uint32_t codeStride = (nseq <= ENC_CODES_SMEM_CAP) ? ENC_CODES_SMEM_CAP : ENC_MAXSEQ;
    uint32_t kind = sh->kind;

    // ---------------------------------------------------------------- P4: gather literals, compact sequences, codes
    if (kind == 0) {
        if (w < ENC_NPARSE) {
            const uint2 *rec = reinterpret_cast<const uint2 *>(scratch + SCR_REC) + w * ENC_SEGCAP;
            const uint32_t cntw = sh->cnt[w], sbase = sh->seqBase[w], carry = sh->carry[w];
            uint32_t lbase = sh->litBase[w];
            const uint32_t b0 = w * sub;
            const uint32_t e0 = (b0 + sub < n) ? b0 + sub : n;
            for (uint32_t i0 = 0; i0 < cntw; i0 += 32) {
                uint32_t i = i0 + lane;
                uint32_t ll = 0, ml3 = 0, dist = 0, ms = 0;
                if (i < cntw) {
                    uint2 r = rec[i];
                    ll = r.x & 0xffff; ml3 = r.x >> 16; dist = r.y & 0xffff; ms = r.y >> 16;
                }
                uint32_t incl = warp_scan_incl(ll);
                uint32_t lpos = lbase + incl - ll;  // literal index of my run
                // copy my literal run src[ms-ll .. ms) -> lit[lpos ..)
                for (uint32_t k = 0; k < ll; k++) lit[lpos + k] = src[ms - ll + k];
                if (i < cntw) {
                    uint32_t llt = ll + ((i == 0) ? carry : 0u);
                    uint32_t ofv = dist ? dist + 3 : 1u;
                    uint32_t gi = sbase + i;
                    seqLL[gi] = (uint16_t)llt; seqML[gi] = (uint16_t)ml3; seqOF[gi] = ofv;
                    codes[gi] = (uint8_t)seq_ll_code(llt);
                    codes[codeStride + gi] = (uint8_t)highbit32(ofv);
                    codes[2 * codeStride + gi] = (uint8_t)seq_ml_code(ml3);
                }
                lbase += __shfl_sync(FULLMASK, incl, 31);
            }
            // trailing literals of the sub-range
            uint32_t tl = sh->tail[w];
            for (uint32_t k = lane; k < tl; k += 32) lit[lbase + k] = src[e0 - tl + k];
        }
    }
    __syncthreads();
    // single-sequence RLE block test (blockenc.go:484-493) needs org[0]; nlit <= 1
    if (kind == 0 && nseq == 1 && nlit <= 1 && tid == 0) {
        uint32_t ll0 = seqLL[0], of0 = seqOF[0];
        if (ll0 == nlit && of0 - 3 == 1) { sh->kind = 2; sh->rleLen = (uint32_t)seqML[0] + 3 + ll0; }
    }
    if (P.dbg_hdr && kind == 0) {
        for (uint32_t i = tid; i < nseq && i < P.dbg_seq_cap; i += ENC_NT) {
            uint32_t *d = P.dbg_seqs + ((uint64_t)chunk * P.dbg_seq_cap + i) * 3;
            d[0] = seqLL[i]; d[1] = seqML[i]; d[2] = seqOF[i];
        }
        for (uint32_t i = tid; i < nlit; i += ENC_NT) P.dbg_lits[(uint64_t)chunk * 65536 + i] = lit[i];
    }
    __syncthreads();
    kind = sh->kind;
    B2C_PHASE(4);

    if (kind == 0) {
        // ------------------------------------------------------------ P5: histograms
        // sequence code histograms: per-warp match.any merge into sh->sw.hist
        for (uint32_t i = tid; i < 3 * 64; i += ENC_NT) (&sh->sw.hist[0][0])[i] = 0;
        if (tid < 3) sh->sw.maxSym[tid] = 0;
        for (uint32_t i = tid; i < 32 * 256; i += ENC_NT) whist[i] = 0;
        __syncthreads();
        {
            // reuse whist rows [w][0..191] as this warp's private (3 x 64) code histogram
            uint32_t *hw3 = whist + w * 256;
            for (uint32_t base = w * 32; base < nseq; base += ENC_NT) {
                uint32_t i = base + lane;
                bool valid = i < nseq;
                unsigned act = __ballot_sync(FULLMASK, valid);
                if (valid) {
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        unsigned code = codes[c * codeStride + i];
                        unsigned peers = __match_any_sync(act, code);
                        if (lane == (unsigned)(__ffs((int)peers) - 1)) hw3[c * 64 + code] += (uint32_t)__popc(peers);
                    }
                }
                __syncwarp();
            }
        }
        __syncthreads();
        if (tid < 192) {
            uint32_t c = 0;
            for (int k = 0; k < ENC_NW; k++) c += whist[k * 256 + tid];
            (&sh->sw.hist[0][0])[tid] = c;
        }
        __syncthreads();
        if (tid < 3) {
            uint32_t mx = 0;
            for (uint32_t s = 0; s < 64; s++) if (sh->sw.hist[tid][s]) mx = s;
            sh->sw.maxSym[tid] = mx;
        }
        // literal histogram (only needed when Huffman is attempted: nlit > 16)
        if (nlit > 16) huf_histogram(lit, nlit, whist, &sh->hw, tid, ENC_NT, 0);
        else __syncthreads();

        B2C_PHASE(5);
        // ------------------------------------------------------------ P6: tables
        if (nlit > 16) { huf_bt_stats(&sh->hw, nlit, tid); } else if (tid == 0) sh->hw.status = HUF_INCOMPRESSIBLE;
        __syncthreads();
        const bool hufTry = sh->hw.status == HUF_OK;
        if (hufTry) huf_bt_sort(&sh->hw, tid, ENC_NT);
        __syncthreads();
        B2C_PHASE(6);
        // serial table builders side by side: warp 0 Huffman tree, warps 1..3 the FSE tables
        if (tid == 0 && hufTry) huf_bt_tree(&sh->hw, nlit);
        if (lane == 0 && w >= 1 && w <= 3) {
            int which = (int)w - 1;
            seq_build_table(&sh->sw, which, nseq, codes[which * codeStride + 0]);
        }
        __syncthreads();
        B2C_PHASE(7);
        if (hufTry) huf_bt_bits(&sh->hw, tid, ENC_NT);
        __syncthreads();
        if (hufTry) huf_bt_vals(&sh->hw, tid, ENC_NT);
        __syncthreads();
        if (tid == 0 && hufTry) huf_bt_write(&sh->hw);
        // meanwhile: the three state chains only need codes + FSE tables
        if (w >= 1 && w <= 3) {
            int which = (int)w - 1;
            uint16_t *stb = reinterpret_cast<uint16_t *>(scratch + SCR_STB) + which * ENC_MAXSEQ;
            seq_chain(&sh->sw, which, codes + which * codeStride, nseq, stb);
        }
        __syncthreads();

        B2C_PHASE(8);
        // ------------------------------------------------------------ P7: literals section
        const bool four = nlit >= 1024;
        HufEncState hst;
        uint32_t payload = 0;
        const bool hufOK = sh->hw.status == HUF_OK;   // may have turned INCOMPRESSIBLE in huf_bt_write
        if (hufOK) payload = huf_enc_sizes(&sh->hw, lit, nlit, four ? 1 : 0, tid, ENC_NT, 0, &hst);
        if (tid == 0) {
            // huff0 compress(): out >= wantSize => ErrIncompressible (compress.go:155-158, WantLogLess 4)
            uint32_t mode = 2;
            if (!hufOK) mode = (sh->hw.status == HUF_USE_RLE) ? 1u : 0u;
            else {
                uint32_t wantSize = nlit - (nlit >> 4);
                if (payload >= wantSize) mode = 0;
                else if (payload + 5 > nlit) {
                    // blockenc.go:534-544: compare with the raw representation
                    uint32_t inBits = 32 - (uint32_t)__clz((int)nlit);
                    uint32_t szRaw = inBits < 5 ? 1 : (inBits < 12 ? 2 : 3);
                    uint32_t compBits = payload ? 32 - (uint32_t)__clz((int)payload) : 0;
                    uint32_t szComp = (compBits <= 10 && inBits <= 10) ? 3 : ((compBits <= 14 && inBits <= 14) ? 4 : 5);
                    if (payload + szComp >= nlit + szRaw) mode = 0;
                }
            }
            uint32_t lh;
            if (mode == 2) {
                uint32_t inBits = 32 - (uint32_t)__clz((int)nlit);
                uint32_t compBits = payload ? 32 - (uint32_t)__clz((int)payload) : 0;
                lh = (compBits <= 10 && inBits <= 10) ? 3 : ((compBits <= 14 && inBits <= 14) ? 4 : 5);
            } else {
                uint32_t inBits = nlit ? 32 - (uint32_t)__clz((int)nlit) : 0;
                lh = inBits < 5 ? 1 : (inBits < 12 ? 2 : 3);
            }
            sh->litMode = mode; sh->lhSize = lh; sh->litPayload = payload;
        }
        __syncthreads();
        const uint32_t litMode = sh->litMode, lhSize = sh->lhSize;
        const uint32_t litOff = fh + 3 + lhSize;  // staging offset of the literal payload
        const uint32_t litBytes = (litMode == 2) ? sh->litPayload : (litMode == 1 ? 1u : nlit);
        // sequence section header: nSeq (1..3 bytes) + modes byte + NCount tables (LL, OF, ML)
        const uint32_t nsHdr = (nseq < 128) ? 1u : (nseq < 0x7f00 ? 2u : 3u);
        const uint32_t seqOff = litOff + litBytes;
        const uint32_t tblOff = seqOff + nsHdr + 1;
        const uint32_t bsOff = tblOff + sh->sw.ncountLen[0] + sh->sw.ncountLen[1] + sh->sw.ncountLen[2];

        B2C_PHASE(9);
        // P8 sizes of the sequence bitstream (every thread owns a run of consecutive t = nseq-1-idx)
        const uint16_t *stbLL = reinterpret_cast<const uint16_t *>(scratch + SCR_STB);
        const uint16_t *stbOF = stbLL + ENC_MAXSEQ;
        const uint16_t *stbML = stbOF + ENC_MAXSEQ;
        const uint32_t per = (nseq + ENC_NT - 1) / ENC_NT;
        uint32_t tA = tid * per, tB = tA + per;
        if (tA > nseq) tA = nseq;
        if (tB > nseq) tB = nseq;
        uint32_t mybits = 0;
        for (uint32_t t = tA; t < tB; t++) {
            uint32_t idx = nseq - 1 - t;
            uint32_t cl = codes[idx], co = codes[codeStride + idx], cm = codes[2 * codeStride + idx];
            mybits += seq_ll_bits(cl) + seq_ml_bits(cm) + co;
            if (t) mybits += (stbLL[idx] >> 12) + (stbOF[idx] >> 12) + (stbML[idx] >> 12);
        }
        uint32_t totalBits;
        uint32_t exBits = group_scan_excl(mybits, sh->sw.scan, 0, ENC_NT, tid, &totalBits);
        const FseCTable *tLL = seq_table(&sh->sw, TBL_LL), *tOF = seq_table(&sh->sw, TBL_OF), *tML = seq_table(&sh->sw, TBL_ML);
        const uint32_t flushBits = tML->tableLog + tOF->tableLog + tLL->tableLog;
        const uint32_t bsBytes = (totalBits + flushBits + 1 + 7) >> 3;
        const uint32_t blockBytes = (bsOff - fh - 3) + bsBytes;  // block content size
        const uint32_t total = fh + 3 + blockBytes + (crc ? 4u : 0u);
        // blockenc.go:811-817: not smaller than the input => raw block.  Also covers staging overflow.
        const bool useRaw = (blockBytes >= n) || (total + 8 > ENC_SRC_BYTES) || sh->sw.err;
        __syncthreads();  // everyone is done reading src/stage-overlapping data? (src no longer needed)
        B2C_PHASE(10);
        if (!useRaw) {
            // zero the staging words that receive bit-granular output
            uint32_t zw = (total + 8 + 3) / 4;
            for (uint32_t i = tid; i < zw; i += ENC_NT) reinterpret_cast<uint32_t *>(stage)[i] = 0;
            __syncthreads();
            // literals
            if (litMode == 2) {
                huf_enc_pack(&sh->hw, lit, four ? 1 : 0, stage, litOff, tid, ENC_NT, 0, &hst);
            } else if (litMode == 0) {
                for (uint32_t i = tid; i < nlit; i += ENC_NT) stage[litOff + i] = lit[i];
            } else if (tid == 0) {
                stage[litOff] = lit[0];
            }
            __syncthreads();  // byte stores above must not race the word atomics below
            B2C_PHASE(11);
            // sequence bitstream
            {
                BitRun br;
                br.init(reinterpret_cast<uint32_t *>(stage), bsOff * 8 + exBits);
                for (uint32_t t = tA; t < tB; t++) {
                    uint32_t idx = nseq - 1 - t;
                    uint32_t cl = codes[idx], co = codes[codeStride + idx], cm = codes[2 * codeStride + idx];
                    if (t) {
                        uint32_t so = stbOF[idx], sm = stbML[idx], sl = stbLL[idx];
                        br.add(so & 0xfff, so >> 12);
                        br.add(sm & 0xfff, sm >> 12);
                        br.add(sl & 0xfff, sl >> 12);
                    }
                    uint32_t lb = seq_ll_bits(cl), mb = seq_ml_bits(cm);
                    br.add((uint32_t)seqLL[idx] & ((1u << lb) - 1), lb);
                    br.add((uint32_t)seqML[idx] & ((1u << mb) - 1), mb);
                    br.add(seqOF[idx] & ((1u << co) - 1), co);
                }
                if (tB == nseq && tA < tB) {
                    // final states: ml, of, ll (blockenc.go:804-806) + end mark
                    br.add(sh->sw.finalState[TBL_ML] & ((1u << tML->tableLog) - 1), tML->tableLog);
                    br.add(sh->sw.finalState[TBL_OF] & ((1u << tOF->tableLog) - 1), tOF->tableLog);
                    br.add(sh->sw.finalState[TBL_LL] & ((1u << tLL->tableLog) - 1), tLL->tableLog);
                    br.add(1u, 1);
                }
                br.finish();
            }
            __syncthreads();
            B2C_PHASE(12);
            // byte-granular headers (after all word-granular atomics)
            if (tid == 0) {
                uint32_t o = 0;
                if (frame) {
                    stage[o++] = 0x28; stage[o++] = 0xB5; stage[o++] = 0x2F; stage[o++] = 0xFD;
                    bool single = n > 1024;
                    uint32_t fcs = (n >= 256) ? ((n >= 65536 + 256) ? 2u : 1u) : 0u;
                    stage[o++] = (uint8_t)((crc ? 4u : 0u) | (single ? 32u : 0u) | (fcs << 6));
                    if (!single) {
                        // WindowSize(n) (enc_base.go:42-50): max(1 << bits.Len(n), 1024)
                        uint32_t ws = 1u << (32 - (uint32_t)__clz((int)n));
                        if (ws < 1024) ws = 1024;
                        stage[o++] = (uint8_t)(((32 - (uint32_t)__clz((int)(ws - 1))) - 10) << 3);
                    }
                    if (fcs == 0) { if (single) stage[o++] = (uint8_t)n; }
                    else if (fcs == 1) { uint32_t v = n - 256; stage[o++] = (uint8_t)v; stage[o++] = (uint8_t)(v >> 8); }
                    else { stage[o++] = (uint8_t)n; stage[o++] = (uint8_t)(n >> 8); stage[o++] = (uint8_t)(n >> 16); stage[o++] = (uint8_t)(n >> 24); }
                }
                uint32_t bh = 1u | (2u << 1) | (blockBytes << 3);  // last block, compressed
                stage[o++] = (uint8_t)bh; stage[o++] = (uint8_t)(bh >> 8); stage[o++] = (uint8_t)(bh >> 16);
                // literals header (blockenc.go:153-238)
                uint64_t lh;
                if (litMode == 2) {
                    uint64_t comp = sh->litPayload;
                    if (lhSize == 3) lh = 2u | ((four ? 1u : 0u) << 2) | ((uint64_t)nlit << 4) | (comp << 14);
                    else if (lhSize == 4) lh = 2u | (2u << 2) | ((uint64_t)nlit << 4) | (comp << 18);
                    else lh = 2u | (3u << 2) | ((uint64_t)nlit << 4) | (comp << 22);
                } else {
                    uint64_t ty = (litMode == 1) ? 1u : 0u;
                    if (lhSize == 1) lh = ty | ((uint64_t)nlit << 3);
                    else if (lhSize == 2) lh = ty | (1u << 2) | ((uint64_t)nlit << 4);
                    else lh = ty | (3u << 2) | ((uint64_t)nlit << 4);
                }
                for (uint32_t k = 0; k < lhSize; k++) stage[o++] = (uint8_t)(lh >> (8 * k));
                // sequences header
                o = seqOff;
                if (nseq < 128) stage[o++] = (uint8_t)nseq;
                else if (nseq < 0x7f00) { stage[o++] = (uint8_t)(128 + (nseq >> 8)); stage[o++] = (uint8_t)nseq; }
                else { uint32_t v = nseq - 0x7f00; stage[o++] = 255; stage[o++] = (uint8_t)v; stage[o++] = (uint8_t)(v >> 8); }
                stage[o++] = (uint8_t)((sh->sw.mode[TBL_LL] << 6) | (sh->sw.mode[TBL_OF] << 4) | (sh->sw.mode[TBL_ML] << 2));
                for (int c = 0; c < 3; c++)
                    for (uint32_t k = 0; k < sh->sw.ncountLen[c]; k++) stage[o++] = sh->sw.ncount[c][k];
                if (crc) {
                    uint32_t c32 = (uint32_t)sh->xxh;
                    uint32_t e = total - 4;
                    stage[e] = (uint8_t)c32; stage[e + 1] = (uint8_t)(c32 >> 8); stage[e + 2] = (uint8_t)(c32 >> 16); stage[e + 3] = (uint8_t)(c32 >> 24);
                }
                sh->outBytes = total;
            }
            __syncthreads();
            B2C_PHASE(13);
            // one coalesced write-back
            if (total <= P.dst_cap) {
                if ((reinterpret_cast<uintptr_t>(gdst) & 15) == 0) {
                    const uint4 *s4 = reinterpret_cast<const uint4 *>(stage);
                    uint4 *d4 = reinterpret_cast<uint4 *>(gdst);
                    uint32_t n16 = total / 16;
                    for (uint32_t i = tid; i < n16; i += ENC_NT) d4[i] = s4[i];
                    for (uint32_t i = n16 * 16 + tid; i < total; i += ENC_NT) gdst[i] = stage[i];
                } else {
                    for (uint32_t i = tid; i < total; i += ENC_NT) gdst[i] = stage[i];
                }
                if (tid == 0) P.out_sizes[chunk] = (int64_t)total;
            } else if (tid == 0) P.out_sizes[chunk] = -4;  // destination too small
            if (P.dbg_hdr && tid == 0) {
                uint32_t *d = P.dbg_hdr + (uint64_t)chunk * 4;
                d[0] = nseq; d[1] = nlit; d[2] = 0; d[3] = litMode;
            }
            __syncthreads();
            B2C_PHASE(14);
            return;
        }
        kind = 1;  // fall through to the raw block
    }

    // ---------------------------------------------------------------- raw / RLE block (+ frame)
    {
        __syncthreads();
        // header assembled in the (now free) L region; payload streamed from global memory
        uint8_t *hdr = smem + ENC_SMEM_L;
        if (tid == 0) {
            uint32_t o = 0;
            if (frame) {
                hdr[o++] = 0x28; hdr[o++] = 0xB5; hdr[o++] = 0x2F; hdr[o++] = 0xFD;
                if (n == 0) {
                    // WithZeroFrames: single segment, no checksum, FCS byte 0 (encoder.go:732-751)
                    hdr[o++] = 32; hdr[o++] = 0;
                } else {
                    bool single = n > 1024;
                    uint32_t fcs = (n >= 256) ? ((n >= 65536 + 256) ? 2u : 1u) : 0u;
                    hdr[o++] = (uint8_t)((crc ? 4u : 0u) | (single ? 32u : 0u) | (fcs << 6));
                    if (!single) {
                        uint32_t ws = 1u << (32 - (uint32_t)__clz((int)n));
                        if (ws < 1024) ws = 1024;
                        hdr[o++] = (uint8_t)(((32 - (uint32_t)__clz((int)(ws - 1))) - 10) << 3);
                    }
                    if (fcs == 0) { if (single) hdr[o++] = (uint8_t)n; }
                    else if (fcs == 1) { uint32_t v = n - 256; hdr[o++] = (uint8_t)v; hdr[o++] = (uint8_t)(v >> 8); }
                    else { hdr[o++] = (uint8_t)n; hdr[o++] = (uint8_t)(n >> 8); hdr[o++] = (uint8_t)(n >> 16); hdr[o++] = (uint8_t)(n >> 24); }
                }
            }
            uint32_t bh = (kind == 2) ? (1u | (1u << 1) | (sh->rleLen << 3)) : (1u | (0u << 1) | (n << 3));
            hdr[o++] = (uint8_t)bh; hdr[o++] = (uint8_t)(bh >> 8); hdr[o++] = (uint8_t)(bh >> 16);
            sh->pos = o;
        }
        __syncthreads();
        const uint32_t hlen = sh->pos;
        const uint32_t body = (kind == 2) ? 1u : n;
        const bool crcHere = crc && n > 0;
        const uint32_t total = hlen + body + (crcHere ? 4u : 0u);
        if (total <= P.dst_cap) {
            for (uint32_t i = tid; i < hlen; i += ENC_NT) gdst[i] = hdr[i];
            for (uint32_t i = tid; i < body; i += ENC_NT) gdst[hlen + i] = gsrc[i];
            if (crcHere && tid < 4) gdst[hlen + body + tid] = (uint8_t)((uint32_t)sh->xxh >> (8 * tid));
            if (tid == 0) P.out_sizes[chunk] = (int64_t)total;
        } else if (tid == 0) P.out_sizes[chunk] = -4;
        if (P.dbg_hdr && tid == 0) {
            uint32_t *d = P.dbg_hdr + (uint64_t)chunk * 4;
            d[0] = nseq; d[1] = nlit; d[2] = kind; d[3] = 0;
        }
        __syncthreads();
    }
}

#ifndef B2C_EMU
extern "C" __global__ void __launch_bounds__(ENC_NT, 1) b2c_zstd_encode_kernel(ZstdEncParams P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *scratch = P.scratch + (uint64_t)blockIdx.x * ENC_SCRATCH_BYTES;
    // predefined FSE tables: once per CTA
    EncShared *sh = reinterpret_cast<EncShared *>(smem + ENC_SMEM_SH);
    if (threadIdx.x < 3) seq_build_predef(&sh->sw, (int)threadIdx.x);
    __syncthreads();
    for (uint32_t c = blockIdx.x; c < P.nchunks; c += gridDim.x) zstd_encode_chunk(smem, P, c, scratch);
}
#endif

}  // namespace b2c
