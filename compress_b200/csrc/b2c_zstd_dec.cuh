// compress_b200/csrc/b2c_zstd_dec.cuh -- zstd frame/block decoder for sm_100a.
//
// One warp decodes one independent input (a frame, or several concatenated / skippable frames: the unit of
// zstd.Decoder.DecodeAll, zstd/decoder.go:319).  Replaces, for the GPU path, the reference's
//   zstd/framedec.go:65-412   frameDec.reset / runDecoder / checkCRC
//   zstd/blockdec.go:128-630  blockDec.reset / decodeLiterals / prepareSequences
//   zstd/seqdec.go:221-500    sequenceDecs.decodeSync / adjustOffset   (+ seqdec_amd64.s)
//   zstd/fse_decoder*.go      readNCount / buildDtable / transform     (+ fse_decoder_amd64.s)
//   huff0/decompress.go:29-166 ReadTable, huff0/decompress_*.go Decompress1X/4X (+ decompress_amd64.s)
// with the reference's validity rules (exact bit consumption, per-block output limit, window checks, ...).
// Known deviation: Huffman-weight FSE tables with tableLog > 9 are reported as unsupported (the zstd format
// caps them at 6; the reference's standalone fse package would take up to 12).  No dictionaries.
//
// Mapping: decoding one block is a serial bit-stream walk, so parallelism comes from the number of inputs (one
// warp each, thousands resident) and, inside a warp, from (a) the 4 Huffman streams on 4 lanes, (b) literal
// and match copies done by all 32 lanes (matches with offset < length are expanded with a modulo gather, so
// they need no byte-serial loop), (c) table fills strided over lanes.  Every lane executes the header /
// sequence state machine redundantly (uniform control flow, broadcast loads), so no shuffles are needed.
#pragma once
#include "b2c_common.cuh"

namespace b2c {

constexpr int DEC_WARPS = 2;                       // warps per CTA
constexpr uint32_t DEC_MAX_BLOCK = 128u << 10;     // maxCompressedBlockSize (zstd/blockdec.go:40)
constexpr uint32_t DEC_LIT_SCRATCH = DEC_MAX_BLOCK + 64;
constexpr uint32_t DEC_TLOG_MAX = 9;               // tablelogAbsoluteMax (zstd/fse_decoder.go:15)
constexpr uint32_t DEC_MAX_MATCHLEN = 131074;      // maxMatchLen (zstd/blockdec.go:48)

enum {
    DEC_ERR_DST = -4, DEC_ERR_CORRUPT = -5, DEC_ERR_MAGIC = -7, DEC_ERR_WINDOW = -8, DEC_ERR_CRC = -9,
    DEC_ERR_SIZE = -10, DEC_ERR_UNSUPPORTED = -11
};

struct DecSym { uint32_t bits; uint32_t baseline; };   // in registers: bits = nbBits | addBits << 8 | newState << 16
// in shared memory a table entry is one word: nbBits | symbol << 8 | newState << 16 (what buildDtable produces);
// the symbol's (baseline, extra bits) come from a 64-entry per-CTA table (DecCta) at lookup time.
struct DecCta { uint2 codeTab[3][64]; };   // x = baseline, y = extra bits

struct DecWarp {
    uint32_t fse[3][1u << DEC_TLOG_MAX];   // per-block tables (LL, OF, ML)
    uint32_t pre[3][64];                   // predefined tables (built once per warp)
    uint32_t dtb[1u << DEC_TLOG_MAX];      // dtable before the transform: nbBits | symbol << 8 | newState << 16
    uint16_t hufDt[2048];
    int16_t norm[256];
    uint16_t symbolNext[256];
    uint8_t weight[260];
    uint8_t symTmp[1u << DEC_TLOG_MAX];
    uint32_t rank[16];
    uint32_t flag;
};
constexpr uint32_t DEC_WARP_BYTES = ((sizeof(DecWarp) + 15) / 16) * 16;
constexpr uint32_t DEC_CTA_BYTES = ((sizeof(DecCta) + 15) / 16) * 16;
constexpr uint32_t DEC_SMEM_BYTES = DEC_WARPS * DEC_WARP_BYTES + DEC_CTA_BYTES;

struct ZstdDecParams {
    // input i: src_base + (src_offsets ? src_offsets[i] : i * src_stride), src_sizes[i] bytes
    const uint8_t *src_base; uint64_t src_stride; const uint64_t *src_offsets; const uint32_t *src_sizes;
    // output i: dst_base + (dst_offsets ? dst_offsets[i] : i * dst_stride), capacity dst_caps ? dst_caps[i] : dst_cap
    uint8_t *dst_base; uint64_t dst_stride; const uint64_t *dst_offsets; const uint32_t *dst_caps; uint32_t dst_cap;
    int64_t *out_sizes;                                                        // decoded bytes or negative error
    uint32_t nchunks;
    uint8_t *lit_scratch;                                                      // [gridDim.x * DEC_WARPS][DEC_LIT_SCRATCH]
    // staged path (b2c_zstd_dec_staged.cuh); fd == nullptr: every input goes through the one-warp decoder, else only the
    // inputs the staged stages marked
    struct FdChunk *fd;                                                        // [nchunks]
    struct FdBlock *fd_blk;                                                    // [nchunks][fd_maxb] block records
    uint32_t fd_maxb;                                                          // blocks per input the staged path takes (FD_MAXB, or up to FD_MAXB_LONG)
    uint32_t fd_per_block;                                                     // 1: the literal and sequence kernels run one unit per (input, block)
    uint32_t *fd_tabs;                                                         // [nchunks][fd_maxb][FD_TAB_ENTRIES] packed entries
    uint16_t *fd_huf;                                                          // [nchunks][fd_maxb][2048] Huffman decoding tables
    const uint32_t *fd_const;                                                   // code maps + predefined tables (FD_CONST_*)
    uint64_t *fd_seqs;                                                         // sequence records, fd_seq_off(c)
    uint8_t *fd_lits;                                                          // decoded literals, fd_lit_off(c)
    uint64_t fd_lit_stride;                                                    // literal area of input c at c * stride (0: dst_offsets[c])
};
B2C_DEV bool dec_staged_done(const ZstdDecParams &P, uint32_t c);               // b2c_zstd_dec_staged.cuh

// ---- backward bit reader over global memory (zstd/bitreader.go semantics: exact consumption required).
// The stream is a little-endian integer; `total` payload bits sit below the end mark; reads take bits from the top
// down; bits below bit 0 read as zero (pos > total is the over-read condition the callers test).
// `bits` holds the next `avail` bits left-aligned (next bit = MSB); 32 more are appended whenever a read would run
// dry, from two aligned word loads (byte loads with zero fill at the two ends of the buffer).  All 32-bit arithmetic:
// a stream is at most one compressed block (128 KiB).
struct BrB {
    const uint8_t *in; uint32_t len; uint32_t total; uint32_t fed;   // fed = bits moved into the window so far
    uint64_t bits; uint32_t avail;
    // look-ahead queue: the four aligned words below the window, requested four refills early (an L2 round trip lasts
    // several refills of a Huffman / sequence walk).  The queue holds the words as loaded; the funnel shift that turns a
    // word into stream bits happens when it is consumed, so nothing waits on a load at request time.
    uint32_t ahead, ahead1, ahead2, ahead3;
    int32_t fb;                                          // stream bytes below fb have not been requested yet
    // word-aligned view of the stream: in + fb keeps its alignment (fb moves by 4), so every refill needs one new
    // aligned word and a fixed funnel shift
    const uint32_t *wp; uint32_t wsh, whi;               // wp = aligned word holding byte fb; whi = the word above `ahead`
    B2C_DEV uint32_t load32_slow(int32_t idx) const {    // bytes outside [0, len) read as zero
        uint32_t v = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int32_t bi = idx + i;
            if (bi >= 0 && (uint32_t)bi < len) v |= (uint32_t)in[bi] << (8 * i);
        }
        return v;
    }
    // the aligned word below wp, bytes that lie before the start of the stream cleared; moves the word view down
    B2C_DEV uint32_t fetch_below() {
        const int32_t sidx = fb - (int32_t)(wsh >> 3) - 4;    // stream index of that word's first byte
        uint32_t v;
        if (sidx >= 0) v = wp[-1];
        else if (sidx <= -4) v = 0u;
        else v = wp[-1] & (0xffffffffu << (8 * (uint32_t)(-sidx)));   // the word holding in[0]: safe to read
        wp -= 1; fb -= 4;
        return v;
    }
    B2C_DEV int init(const uint8_t *p, uint32_t n) {
        in = p; len = n; total = 0; fed = 0; bits = 0; avail = 0; fb = 0; ahead = ahead1 = ahead2 = ahead3 = 0; wp = nullptr; wsh = 0; whi = 0;
        if (n < 1) return -1;
        const uint8_t v = p[n - 1];
        if (v == 0) return -1;
        total = 8 * (n - 1) + highbit32(v);
        const int32_t cbyte = (int32_t)((total + 7) >> 3) - 8;        // window = the 8 bytes ending at the top payload byte
        const uint64_t win = (uint64_t)load32_slow(cbyte) | ((uint64_t)load32_slow(cbyte + 4) << 32);
        const uint32_t k = (uint32_t)((int32_t)total - 8 * cbyte);    // payload bits inside the window: 57..64
        bits = win << (64 - k);
        avail = k; fed = k; fb = cbyte;
        const uintptr_t a = reinterpret_cast<uintptr_t>(p) + (uintptr_t)(intptr_t)cbyte;   // may lie below p for tiny streams
        wp = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
        wsh = (uint32_t)(a & 3) * 8;
        // the word holding byte cbyte: only its bytes below cbyte are ever used, and only when they are stream bytes
        whi = (cbyte >= 1 && wsh != 0) ? *wp : 0u;
        if (cbyte >= 1 && wsh != 0 && (int32_t)(wsh >> 3) > cbyte) whi &= 0xffffffffu << (8 * ((wsh >> 3) - (uint32_t)cbyte));
        ahead = fetch_below(); ahead1 = fetch_below(); ahead2 = fetch_below(); ahead3 = fetch_below();
        return 0;
    }
    B2C_DEV void refill32() {     // requires avail <= 32
        const uint32_t lo = ahead;
        const uint32_t v = wsh ? __funnelshift_r(lo, whi, wsh) : lo;
        whi = lo;
        bits |= (uint64_t)v << (32 - avail);
        avail += 32; fed += 32;
        ahead = ahead1; ahead1 = ahead2; ahead2 = ahead3;
        ahead3 = fetch_below();
    }
    B2C_DEV uint32_t pos() const { return fed - avail; }          // bits consumed
    // next n bits, 0 <= n <= 32, without consuming them
    B2C_DEV uint32_t peek(uint32_t n) {
        if (avail < n) refill32();
        return (uint32_t)((bits >> 1) >> (63 - n));
    }
    // consume n bits; n <= the n of the preceding peek
    B2C_DEV void skip(uint32_t n) { bits <<= n; avail -= n; }
    B2C_DEV uint32_t read(uint32_t n) {
        const uint32_t v = peek(n);
        skip(n);
        return v;
    }
    B2C_DEV bool finished() const { return pos() >= total; }
};

// forward LSB-first bit fetch with zero fill (readNCount)
B2C_DEV uint32_t dec_fwd_bits(const uint8_t *in, uint32_t len, uint32_t bitpos, uint32_t n) {
    uint64_t acc = 0;
    uint32_t byte0 = bitpos >> 3;
#pragma unroll
    for (uint32_t i = 0; i < 8; i++) { uint32_t bi = byte0 + i; if (bi < len) acc |= (uint64_t)in[bi] << (8 * i); }
    acc >>= (bitpos & 7);
    return (uint32_t)(acc & ((n >= 32) ? 0xffffffffull : ((1ull << n) - 1)));
}

// readNCount (zstd/fse_decoder.go:52-184, fse/decompress.go:48-168).  Uniform across the warp; lane 0 stores.
// returns bytes consumed or -1
B2C_DEV int dec_read_ncount(const uint8_t *in, uint32_t len, uint32_t maxSymbol, uint32_t absMaxLog, int16_t *norm,
                            uint32_t *symbolLenOut, uint32_t *tableLogOut, unsigned lane) {
    if (len < 4) return -1;
    uint32_t bp = 0;
    uint32_t nbBits = dec_fwd_bits(in, len, bp, 4) + 5;
    bp += 4;
    if (nbBits > absMaxLog) return -1;
    uint32_t tableLog = nbBits;
    int32_t remaining = (1 << nbBits) + 1, threshold = 1 << nbBits, gotTotal = 0;
    uint32_t charnum = 0;
    bool previous0 = false;
    nbBits++;
    while (remaining > 1 && charnum <= maxSymbol) {
        if (previous0) {
            uint32_t n0 = charnum;
            while (dec_fwd_bits(in, len, bp, 16) == 0xFFFF) { n0 += 24; bp += 16; if (bp > 8 * len + 64) return -1; }
            while (dec_fwd_bits(in, len, bp, 2) == 3) { n0 += 3; bp += 2; }
            n0 += dec_fwd_bits(in, len, bp, 2);
            bp += 2;
            if (n0 > 255) return -1;
            while (charnum < n0) { if (lane == 0) norm[charnum & 0xff] = 0; charnum++; }
        }
        int32_t max = (2 * threshold - 1) - remaining;
        int32_t count;
        uint32_t bitStream = dec_fwd_bits(in, len, bp, 32);
        if ((int32_t)(bitStream & (uint32_t)(threshold - 1)) < max) {
            count = (int32_t)(bitStream & (uint32_t)(threshold - 1));
            bp += nbBits - 1;
        } else {
            count = (int32_t)(bitStream & (uint32_t)(2 * threshold - 1));
            if (count >= threshold) count -= max;
            bp += nbBits;
        }
        count--;
        if (count < 0) { remaining += count; gotTotal -= count; }
        else { remaining -= count; gotTotal += count; }
        if (lane == 0) norm[charnum & 0xff] = (int16_t)count;
        charnum++;
        previous0 = (count == 0);
        while (remaining < threshold) { nbBits--; threshold >>= 1; }
    }
    __syncwarp();
    if (charnum <= 1 || charnum > 256) return -1;
    if (remaining != 1) return -1;
    if (bp > 8 * len) return -1;
    if (gotTotal != (1 << tableLog)) return -1;
    *symbolLenOut = charnum; *tableLogOut = tableLog;
    return (int)((bp + 7) >> 3);
}

// buildDtable (zstd/fse_decoder_generic.go:11-72, fse/decompress.go:193-258): serial spread by lane 0.
// dtb[u] = nbBits | symbol << 8 | newState << 16
B2C_DEV int dec_build_dtable(DecWarp *dw, uint32_t *outTab, uint32_t symbolLen, uint32_t tableLog, unsigned lane) {
    uint32_t tableSize = 1u << tableLog;
    if (lane == 0) {
        int err = 0;
        const int16_t *norm = dw->norm;
        uint8_t *symOut = dw->symTmp;
        uint32_t highThreshold = tableSize - 1;
        for (uint32_t i = 0; i < symbolLen; i++) {
            int16_t v = norm[i];
            if (v == -1) { symOut[highThreshold] = (uint8_t)i; highThreshold--; dw->symbolNext[i] = 1; }
            else dw->symbolNext[i] = (uint16_t)v;
        }
        uint32_t tableMask = tableSize - 1, step = (tableSize >> 1) + (tableSize >> 3) + 3, position = 0;
        for (uint32_t ss = 0; ss < symbolLen; ss++) {
            int v = norm[ss];
            for (int i = 0; i < v; i++) {
                symOut[position] = (uint8_t)ss;
                position = (position + step) & tableMask;
                while (position > highThreshold) position = (position + step) & tableMask;
            }
        }
        if (position != 0) err = 1;
        if (!err) {
            for (uint32_t u = 0; u < tableSize; u++) {
                uint8_t symbol = symOut[u];
                uint32_t nextState = dw->symbolNext[symbol];
                dw->symbolNext[symbol] = (uint16_t)(nextState + 1);
                uint32_t nBits = tableLog - highbit32(nextState);
                uint32_t newState = ((nextState << nBits) - tableSize) & 0xffff;
                if (newState >= tableSize) err = 1;
                if (newState == u && nBits == 0) err = 1;   // "newState == oldState and no bits"
                outTab[u] = nBits | ((uint32_t)symbol << 8) | (newState << 16);
            }
        }
        dw->flag = (uint32_t)err;
    }
    __syncwarp();
    return dw->flag ? -1 : 0;
}

// code -> (baseline, extra bits): zstd/fse_predefined.go:78-107 (symbolTableX)
B2C_DEV void dec_code_base(int which, uint32_t c, uint32_t *base, uint32_t *bits) {
    if (which == 0) {       // literal lengths
        if (c < 16) { *base = c; *bits = 0; return; }
        const uint8_t b[20] = {1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
        uint32_t v = 16;
        for (uint32_t i = 16; i < c; i++) v += 1u << b[i - 16];
        *base = v; *bits = b[c - 16];
    } else if (which == 2) {  // match lengths
        if (c < 32) { *base = c + 3; *bits = 0; return; }
        const uint8_t b[21] = {1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
        uint32_t v = 35;
        for (uint32_t i = 32; i < c; i++) v += 1u << b[i - 32];
        *base = v; *bits = b[c - 32];
    } else {                 // offsets: code 0 -> (0,0), 1 -> (1,1), c >= 2 -> ((1 << c) - 3, c)
        if (c == 0) { *base = 0; *bits = 0; }
        else if (c == 1) { *base = 1; *bits = 1; }
        else { *base = (1u << c) - 3; *bits = c; }
    }
}
B2C_DEV uint32_t dec_code_limit(int which) { return which == 0 ? 36u : (which == 1 ? 31u : 53u); }

// transform (zstd/fse_decoder.go:282-299) only has to reject symbols without a (baseline, bits) entry: the table
// keeps the symbol, the mapping is applied at lookup (dec_lookup).
B2C_DEV int dec_check_symbols(const uint32_t *tab, uint32_t tableSize, int which, unsigned lane) {
    bool bad = false;
    for (uint32_t i = lane; i < tableSize; i += 32) bad = bad || (((tab[i] >> 8) & 0xff) >= dec_code_limit(which));
    const bool anyBad = __any_sync(FULLMASK, bad);
    __syncwarp();
    return anyBad ? -1 : 0;
}
B2C_DEV DecSym dec_lookup(const uint32_t *tab, const uint2 *codeTab, uint32_t state) {
    const uint32_t e = tab[state];
    const uint2 c = codeTab[(e >> 8) & 63];
    DecSym r;
    r.bits = (e & 0xffff00ffu) | (c.y << 8);
    r.baseline = c.x;
    return r;
}

B2C_DEV void dec_build_predef(DecWarp *dw, unsigned lane) {
    const int8_t llN[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
    const int8_t ofN[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};
    const int8_t mlN[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                            1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
    for (int which = 0; which < 3; which++) {
        uint32_t sl = which == 0 ? 36u : (which == 1 ? 29u : 53u), tl = which == 1 ? 5u : 6u;
        if (lane == 0)
            for (uint32_t i = 0; i < sl; i++) dw->norm[i] = which == 0 ? llN[i] : (which == 1 ? ofN[i] : mlN[i]);
        __syncwarp();
        dec_build_dtable(dw, dw->pre[which], sl, tl, lane);
    }
}

// ---- Huffman ------------------------------------------------------------------------------------------------
// standalone FSE decode of the weights (fse/decompress.go:260-330); uniform, lane 0 stores. returns count or -1/-2
B2C_DEV int dec_fse_weights(DecWarp *dw, const uint8_t *in, uint32_t n, unsigned lane) {
    uint32_t symbolLen = 0, tableLog = 0;
    int hdr = dec_read_ncount(in, n, 255, 15, dw->norm, &symbolLen, &tableLog, lane);
    if (hdr < 0) return -1;
    if (tableLog > DEC_TLOG_MAX) return -2;
    if (dec_build_dtable(dw, dw->dtb, symbolLen, tableLog, lane)) return -1;
    if ((uint32_t)hdr > n) return -1;
    BrB br;
    if (br.init(in + hdr, n - (uint32_t)hdr)) return -1;
    const uint32_t *dt = dw->dtb;
#define WGET(nb) (((nb) == 0 || br.finished()) ? 0u : br.read(nb))
    uint32_t s1 = WGET(tableLog), s2 = WGET(tableLog);
    uint32_t o = 0;
    int rc = 0;
    for (;;) {
        uint32_t e1 = dt[s1], e2 = dt[s2];
        if (br.finished() && (e1 & 0xff) > 0) {
            if (o + 2 > 256) { rc = -1; break; }
            if (lane == 0) { dw->weight[o] = (uint8_t)(e1 >> 8); dw->weight[o + 1] = (uint8_t)(e2 >> 8); }
            o += 2; break;
        }
        { uint32_t lb = WGET(e1 & 0xff); s1 = (e1 >> 16) + lb; if (o >= 256) { rc = -1; break; } if (lane == 0) dw->weight[o] = (uint8_t)(e1 >> 8); o++; }
        e2 = dt[s2]; e1 = dt[s1];
        if (br.finished() && (e2 & 0xff) > 0) {
            if (o + 2 > 256) { rc = -1; break; }
            if (lane == 0) { dw->weight[o] = (uint8_t)(e2 >> 8); dw->weight[o + 1] = (uint8_t)(e1 >> 8); }
            o += 2; break;
        }
        { uint32_t lb = WGET(e2 & 0xff); s2 = (e2 >> 16) + lb; if (o >= 256) { rc = -1; break; } if (lane == 0) dw->weight[o] = (uint8_t)(e2 >> 8); o++; }
        if (o >= 255) { rc = -1; break; }   // DecompressLimit = 255
    }
#undef WGET
    __syncwarp();
    if (rc) return -1;
    if (br.pos() > br.total) return -1;
    return (int)o;
}

// huff0.ReadTable (huff0/decompress.go:29-166). returns bytes consumed, -1 corrupt, -2 unsupported
B2C_DEV int dec_huf_read_table(DecWarp *dw, const uint8_t *in, uint32_t n, uint32_t *tableLogOut, unsigned lane) {
    if (n <= 1) return -1;
    uint32_t iSize = in[0];
    in++; n--;
    uint32_t symbolLen;
    if (iSize >= 128) {
        uint32_t oSize = iSize - 127;
        iSize = (oSize + 1) / 2;
        if (iSize > n) return -1;
        for (uint32_t k = lane * 2; k < oSize; k += 64) {
            uint8_t v = in[k / 2];
            dw->weight[k] = v >> 4; dw->weight[k + 1] = v & 15;
        }
        __syncwarp();
        symbolLen = oSize;
    } else {
        if (n < iSize) return -1;
        int b = dec_fse_weights(dw, in, iSize, lane);
        if (b == -2) return -2;
        if (b < 0 || b > 255) return -1;
        symbolLen = (uint32_t)b;
    }
    // rank statistics (uniform)
    uint32_t rankStats[16];
#pragma unroll
    for (int i = 0; i < 16; i++) rankStats[i] = 0;
    uint32_t weightTotal = 0;
    for (uint32_t i = 0; i < symbolLen; i++) {
        uint32_t v = dw->weight[i];
        if (v > 11) return -1;
        rankStats[v]++;
        weightTotal += (1u << v) >> 1;
    }
    if (weightTotal == 0) return -1;
    uint32_t tableLog = highbit32(weightTotal) + 1;
    if (tableLog > 11) return -1;
    {
        uint32_t total = 1u << tableLog, rest = total - weightTotal;
        uint32_t verif = 1u << highbit32(rest), lastWeight = highbit32(rest) + 1;
        if (verif != rest) return -1;
        if (lane == 0) dw->weight[symbolLen] = (uint8_t)lastWeight;
        symbolLen++;
        rankStats[lastWeight]++;
    }
    __syncwarp();
    if (rankStats[1] < 2 || (rankStats[1] & 1)) return -1;
    {
        uint32_t nextRankStart = 0;
        for (uint32_t k = 1; k < tableLog + 1; k++) { uint32_t c = nextRankStart; nextRankStart += rankStats[k] << (k - 1); rankStats[k] = c; }
    }
    for (uint32_t i = lane; i < 2048; i += 32) dw->hufDt[i] = 0;
    __syncwarp();
    for (uint32_t k = 0; k < symbolLen; k++) {
        uint32_t w = dw->weight[k];
        if (w == 0) continue;
        uint32_t length = (1u << w) >> 1;
        uint16_t entry = (uint16_t)((tableLog + 1 - w) | (k << 8));
        uint32_t r = rankStats[w];
        for (uint32_t i = lane; i < length; i += 32) dw->hufDt[r + i] = entry;
        rankStats[w] = r + length;
    }
    __syncwarp();
    *tableLogOut = tableLog;
    return (int)(1 + iSize);
}

// one Huffman stream, one lane: exactly `count` symbols, exact consumption (huff0/decompress_generic.go).
// WORDS: symbols are produced four at a time and leave as one aligned 32-bit store (an over-read shows up as
// pos > total at the end -- bits below the start of the stream read as zero -- so the loop needs no per-symbol end
// test).  Measured on B200: the word form is 18 % faster for standalone huff0 blocks (64 K symbols per stream) and
// 17 % slower inside the zstd block decoder (8 K symbols per stream, scattered alignment), so each uses its own.
template <bool WORDS>
__device__ __noinline__ int dec_huf_stream(const uint16_t *dt, uint32_t tl, const uint8_t *src, uint32_t n, uint8_t *dst, uint32_t count) {
    BrB br;
    if (br.init(src, n)) return -1;
    uint32_t i = 0;
#define HUF_ONE(out)                                                                                   \
    do {                                                                                               \
        const uint16_t e_ = dt[br.peek(tl)];                                                           \
        br.skip(e_ & 0xff);                                                                            \
        (out) = (uint32_t)(e_ >> 8);                                                                   \
    } while (0)
    if (!WORDS) {
        // two symbols per end test and refill test (2 * tl <= 22 bits; a refill leaves at least 32)
        for (; i + 2 <= count; i += 2) {
            if (br.finished()) return -1;
            if (br.avail < 2 * tl) br.refill32();
            const uint16_t e0 = dt[(uint32_t)(br.bits >> (64 - tl))];
            br.skip(e0 & 0xff);
            const uint16_t e1 = dt[(uint32_t)(br.bits >> (64 - tl))];
            br.skip(e1 & 0xff);
            dst[i] = (uint8_t)(e0 >> 8);
            dst[i + 1] = (uint8_t)(e1 >> 8);
        }
        for (; i < count; i++) {
            if (br.finished()) return -1;
            uint32_t s0; HUF_ONE(s0);
            dst[i] = (uint8_t)s0;
        }
    }
    while (i < count && ((reinterpret_cast<uintptr_t>(dst + i) & 3) != 0)) {
        if (br.finished()) return -1;
        uint32_t s0; HUF_ONE(s0);
        dst[i++] = (uint8_t)s0;
    }
    for (; i + 4 <= count; i += 4) {
        if (br.finished()) return -1;
        uint32_t s0, s1, s2, s3;
        HUF_ONE(s0); HUF_ONE(s1); HUF_ONE(s2); HUF_ONE(s3);
        *reinterpret_cast<uint32_t *>(dst + i) = s0 | (s1 << 8) | (s2 << 16) | (s3 << 24);
    }
    for (; i < count; i++) {
        if (br.finished()) return -1;
        uint32_t s0; HUF_ONE(s0);
        dst[i] = (uint8_t)s0;
    }
#undef HUF_ONE
    return br.pos() == br.total ? 0 : -1;
}

// Decompress1X / Decompress4X body after the table (huff0/decompress.go:234-, :622-): 4 streams on 4 lanes.
// Returns 0 or -1 per lane; callers vote.  Every lane must call.
template <bool WORDS>
B2C_DEV int dec_huf_streams(const uint16_t *dt, uint32_t tl, const uint8_t *hs, uint32_t hl, uint8_t *dst, uint32_t dstSize,
                            bool four, unsigned lane) {
    int e = 0;
    if (four) {
        if (hl < 6 + 4) return -1;
        const uint32_t dstEvery = (dstSize + 3) / 4;
        const uint32_t l0 = hs[0] | ((uint32_t)hs[1] << 8), l1 = hs[2] | ((uint32_t)hs[3] << 8), l2 = hs[4] | ((uint32_t)hs[5] << 8);
        const uint32_t s0 = 6, s1 = s0 + l0, s2 = s1 + l1, s3 = s2 + l2;
        if (s1 >= hl || s2 >= hl || s3 >= hl) return -1;   // "truncated input (or invalid offset)"
        if (3 * dstEvery > dstSize) return -1;
        const uint32_t cnt3 = dstSize - 3 * dstEvery;
        if (lane < 4) {
            const uint32_t st = lane == 0 ? s0 : (lane == 1 ? s1 : (lane == 2 ? s2 : s3));
            const uint32_t ln = lane == 0 ? l0 : (lane == 1 ? l1 : (lane == 2 ? l2 : hl - s3));
            const uint32_t cnt = lane < 3 ? dstEvery : cnt3;
            e = dec_huf_stream<WORDS>(dt, tl, hs + st, ln, dst + lane * dstEvery, cnt);
        }
    } else {
        if (lane == 0) e = dec_huf_stream<WORDS>(dt, tl, hs, hl, dst, dstSize);
    }
    return e;
}

// ---- the decoder: one warp, one input ---------------------------------------------------------------
B2C_DEV int64_t zstd_decode_input(DecWarp *dw, const DecCta *dc, const uint8_t *src, uint32_t n, uint8_t *dst, uint32_t cap,
                                  uint8_t *litbuf, unsigned lane) {
    uint32_t ip = 0;
    uint64_t total = 0;
#define DFAIL(code) return (int64_t)(code)
    for (;;) {
        // ---- frame header (framedec.go:65-270), skippable frames first
        for (;;) {
            if (n - ip == 0) return (int64_t)total;
            if (n - ip < 4) DFAIL(DEC_ERR_CORRUPT);
            if (!(src[ip + 1] == 0x2A && src[ip + 2] == 0x4D && src[ip + 3] == 0x18 && (src[ip] & 0xf0) == 0x50)) break;
            ip += 4;
            if (n - ip < 4) DFAIL(DEC_ERR_CORRUPT);
            uint32_t skip = (uint32_t)src[ip] | ((uint32_t)src[ip + 1] << 8) | ((uint32_t)src[ip + 2] << 16) | ((uint32_t)src[ip + 3] << 24);
            ip += 4;
            if (skip > n - ip) DFAIL(DEC_ERR_CORRUPT);
            ip += skip;
        }
        if (!(src[ip] == 0x28 && src[ip + 1] == 0xB5 && src[ip + 2] == 0x2F && src[ip + 3] == 0xFD)) DFAIL(DEC_ERR_MAGIC);
        ip += 4;
        if (n - ip < 1) DFAIL(DEC_ERR_CORRUPT);
        uint32_t fhd = src[ip++];
        bool singleSegment = (fhd & 32) != 0;
        if (fhd & 8) DFAIL(DEC_ERR_CORRUPT);
        uint64_t windowSize = 0;
        if (!singleSegment) {
            if (n - ip < 1) DFAIL(DEC_ERR_CORRUPT);
            uint32_t wd = src[ip++];
            uint32_t windowLog = 10 + (wd >> 3);
            uint64_t windowBase = 1ull << windowLog;
            windowSize = windowBase + (windowBase / 8) * (wd & 7);
        }
        if (fhd & 3) {
            uint32_t size = fhd & 3; if (size == 3) size = 4;
            if (n - ip < size) DFAIL(DEC_ERR_CORRUPT);
            uint32_t id = 0;
            for (uint32_t k = 0; k < size; k++) id |= (uint32_t)src[ip + k] << (8 * k);
            ip += size;
            if (id != 0) DFAIL(DEC_ERR_UNSUPPORTED);
        }
        uint32_t fcsSize = 0, v6 = fhd >> 6;
        if (v6 == 0) { if (singleSegment) fcsSize = 1; } else fcsSize = 1u << v6;
        uint64_t fcs = ~0ull;
        if (fcsSize) {
            if (n - ip < fcsSize) DFAIL(DEC_ERR_CORRUPT);
            fcs = 0;
            for (uint32_t k = 0; k < fcsSize; k++) fcs |= (uint64_t)src[ip + k] << (8 * k);
            if (fcsSize == 2) fcs += 256;
            ip += fcsSize;
        }
        bool hasCheck = (fhd & 4) != 0;
        if (windowSize > (1ull << 29)) DFAIL(DEC_ERR_WINDOW);
        if (windowSize == 0 && singleSegment) {
            windowSize = fcs > 1024 ? fcs : 1024;
            if (windowSize > (64ull << 30)) DFAIL(DEC_ERR_SIZE);   // ErrDecoderSizeExceeded (maxDecodedSize default)
        }
        if (windowSize < 1024) DFAIL(DEC_ERR_WINDOW);
        if (fcs != ~0ull && fcs > (64ull << 30) - total) DFAIL(DEC_ERR_SIZE);

        // history.reset
        bool haveHuff = false;
        uint32_t hufLog = 0;
        const uint32_t *cur[3] = {nullptr, nullptr, nullptr};
        uint32_t tlog[3] = {0, 0, 0};
        int64_t rep0 = 1, rep1 = 4, rep2 = 8;
        uint8_t *out = dst + total;
        uint64_t outCap = (uint64_t)cap - total;
        uint64_t o = 0;

        for (;;) {   // blocks
            if (n - ip < 3) DFAIL(DEC_ERR_CORRUPT);
            uint32_t bh = (uint32_t)src[ip] | ((uint32_t)src[ip + 1] << 8) | ((uint32_t)src[ip + 2] << 16);
            ip += 3;
            bool last = bh & 1;
            uint32_t type = (bh >> 1) & 3, cSize = bh >> 3;
            if (type == 3) DFAIL(DEC_ERR_CORRUPT);
            if (type == 1) {
                if (cSize > DEC_MAX_BLOCK || cSize > windowSize) DFAIL(DEC_ERR_WINDOW);
                if (n - ip < 1) DFAIL(DEC_ERR_CORRUPT);
                if (o + cSize > outCap) DFAIL(DEC_ERR_DST);
                uint8_t b = src[ip];
                for (uint32_t i = lane; i < cSize; i += 32) out[o + i] = b;
                o += cSize; ip += 1;
            } else if (type == 0) {
                if (cSize > DEC_MAX_BLOCK || cSize > windowSize) DFAIL(DEC_ERR_WINDOW);
                if (n - ip < cSize) DFAIL(DEC_ERR_CORRUPT);
                if (o + cSize > outCap) DFAIL(DEC_ERR_DST);
                for (uint32_t i = lane; i < cSize; i += 32) out[o + i] = src[ip + i];
                o += cSize; ip += cSize;
            } else {
                if (cSize > DEC_MAX_BLOCK || (uint64_t)cSize > windowSize) DFAIL(DEC_ERR_CORRUPT);
                if (cSize < 2) DFAIL(DEC_ERR_CORRUPT);
                if (n - ip < cSize) DFAIL(DEC_ERR_CORRUPT);
                const uint8_t *in = src + ip;
                uint32_t len = cSize;
                ip += cSize;
                // ---- literals section (blockdec.go:275-474)
                uint32_t litType = in[0] & 3, sizeFormat = (in[0] >> 2) & 3;
                uint32_t litRegen = 0, litComp = 0;
                bool four = false;
                if (litType < 2) {
                    if (sizeFormat == 0 || sizeFormat == 2) { litRegen = in[0] >> 3; in += 1; len -= 1; }
                    else if (sizeFormat == 1) { litRegen = (in[0] >> 4) + ((uint32_t)in[1] << 4); in += 2; len -= 2; }
                    else { if (len < 3) DFAIL(DEC_ERR_CORRUPT); litRegen = (in[0] >> 4) + ((uint32_t)in[1] << 4) + ((uint32_t)in[2] << 12); in += 3; len -= 3; }
                } else {
                    if (sizeFormat <= 1) {
                        if (len < 3) DFAIL(DEC_ERR_CORRUPT);
                        uint32_t v = (in[0] >> 4) + ((uint32_t)in[1] << 4) + ((uint32_t)in[2] << 12);
                        litRegen = v & 1023; litComp = v >> 10; four = sizeFormat == 1; in += 3; len -= 3;
                    } else if (sizeFormat == 2) {
                        if (len < 4) DFAIL(DEC_ERR_CORRUPT);
                        uint32_t v = (in[0] >> 4) + ((uint32_t)in[1] << 4) + ((uint32_t)in[2] << 12) + ((uint32_t)in[3] << 20);
                        litRegen = v & 16383; litComp = v >> 14; four = true; in += 4; len -= 4;
                    } else {
                        if (len < 5) DFAIL(DEC_ERR_CORRUPT);
                        uint64_t v = (uint64_t)(in[0] >> 4) + ((uint64_t)in[1] << 4) + ((uint64_t)in[2] << 12) + ((uint64_t)in[3] << 20) + ((uint64_t)in[4] << 28);
                        litRegen = (uint32_t)(v & 262143); litComp = (uint32_t)(v >> 18); four = true; in += 5; len -= 5;
                    }
                }
                if (litRegen > windowSize || litRegen > DEC_MAX_BLOCK) DFAIL(DEC_ERR_WINDOW);
                const uint8_t *literals = litbuf;
                if (litType == 0) {
                    if (len < litRegen) DFAIL(DEC_ERR_CORRUPT);
                    literals = in; in += litRegen; len -= litRegen;
                } else if (litType == 1) {
                    if (len < 1) DFAIL(DEC_ERR_CORRUPT);
                    uint8_t b = in[0];
                    for (uint32_t i = lane; i < litRegen; i += 32) litbuf[i] = b;
                    in += 1; len -= 1;
                } else {
                    if (len < litComp) DFAIL(DEC_ERR_CORRUPT);
                    const uint8_t *hs = in;
                    uint32_t hl = litComp;
                    if (litType == 2) {
                        int used = dec_huf_read_table(dw, hs, hl, &hufLog, lane);
                        if (used == -2) DFAIL(DEC_ERR_UNSUPPORTED);
                        if (used < 0) DFAIL(DEC_ERR_CORRUPT);
                        haveHuff = true;
                        hs += used; hl -= (uint32_t)used;
                    } else if (!haveHuff) DFAIL(DEC_ERR_CORRUPT);   // treeless without history
                    int e = dec_huf_streams<false>(dw->hufDt, hufLog, hs, hl, litbuf, litRegen, four, lane);
                    if (__any_sync(FULLMASK, e != 0)) DFAIL(DEC_ERR_CORRUPT);
                    in += litComp; len -= litComp;
                }
                __syncwarp();
                // ---- sequences section header (blockdec.go:505-630)
                if (len < 1) DFAIL(DEC_ERR_CORRUPT);
                uint32_t nSeqs, sh0 = in[0];
                if (sh0 < 128) { nSeqs = sh0; in += 1; len -= 1; }
                else if (sh0 < 255) { if (len < 2) DFAIL(DEC_ERR_CORRUPT); nSeqs = ((sh0 - 128) << 8) | in[1]; in += 2; len -= 2; }
                else { if (len < 3) DFAIL(DEC_ERR_CORRUPT); nSeqs = 0x7f00 + in[1] + ((uint32_t)in[2] << 8); in += 3; len -= 3; }
                if (nSeqs == 0) {
                    if (len != 0) DFAIL(DEC_ERR_CORRUPT);
                    if (o + litRegen > outCap) DFAIL(DEC_ERR_DST);
                    for (uint32_t i = lane; i < litRegen; i += 32) out[o + i] = literals[i];
                    o += litRegen;
                } else {
                    if (len < 1) DFAIL(DEC_ERR_CORRUPT);
                    uint32_t compMode = in[0];
                    in += 1; len -= 1;
                    if (compMode & 3) DFAIL(DEC_ERR_CORRUPT);
                    const uint32_t maxSym[3] = {35, 30, 52};
                    for (int i = 0; i < 3; i++) {
                        uint32_t mode = (compMode >> (6 - i * 2)) & 3;
                        if (mode == 0) { cur[i] = dw->pre[i]; tlog[i] = (i == 1) ? 5 : 6; }
                        else if (mode == 1) {
                            if (len < 1) DFAIL(DEC_ERR_CORRUPT);
                            uint32_t v = in[0]; in += 1; len -= 1;
                            if (v >= dec_code_limit(i)) DFAIL(DEC_ERR_CORRUPT);
                            __syncwarp();
                            if (lane == 0) dw->fse[i][0] = v << 8;
                            __syncwarp();
                            cur[i] = dw->fse[i]; tlog[i] = 0;
                        } else if (mode == 2) {
                            uint32_t symbolLen = 0, tl = 0;
                            __syncwarp();
                            int used = dec_read_ncount(in, len, maxSym[i], DEC_TLOG_MAX, dw->norm, &symbolLen, &tl, lane);
                            if (used < 0 || (uint32_t)used > len) DFAIL(DEC_ERR_CORRUPT);
                            in += used; len -= (uint32_t)used;
                            cur[i] = nullptr;
                            if (dec_build_dtable(dw, dw->fse[i], symbolLen, tl, lane)) DFAIL(DEC_ERR_CORRUPT);
                            if (dec_check_symbols(dw->fse[i], 1u << tl, i, lane)) DFAIL(DEC_ERR_CORRUPT);
                            cur[i] = dw->fse[i]; tlog[i] = tl;
                        }  // mode 3: repeat, keep cur[i]
                    }
                    if (!cur[0] || !cur[1] || !cur[2]) DFAIL(DEC_ERR_CORRUPT);   // "sequence decoder not defined"
                    // ---- decodeSync (seqdec.go:221-445)
                    BrB br;
                    if (br.init(in, len)) DFAIL(DEC_ERR_CORRUPT);
                    DecSym llS = dec_lookup(cur[0], dc->codeTab[0], br.read(tlog[0]));
                    DecSym ofS = dec_lookup(cur[1], dc->codeTab[1], br.read(tlog[1]));
                    DecSym mlS = dec_lookup(cur[2], dc->codeTab[2], br.read(tlog[2]));
                    const uint64_t startSize = o;
                    const uint64_t maxBlockSize = windowSize < DEC_MAX_BLOCK ? windowSize : DEC_MAX_BLOCK;
                    uint32_t litPos = 0;
                    // Sequences are handled 32 at a time.  Step 1 (every lane, redundantly): the serial bitstream walk;
                    // lane j keeps sequence j.  Step 2: a warp scan places them; all literal runs are copied at once;
                    // matches are copied in waves -- a match runs as soon as its source ends before the destination of
                    // every match of the batch that is still pending (seqdec.go:221-445 executes them one by one).
                    for (uint32_t base = 0; base < nSeqs; base += 32) {
                        const uint32_t want = (nSeqs - base < 32) ? nSeqs - base : 32;
                        uint32_t cnt = want;
                        uint32_t myLL = 0, myML = 0, myMO = 0;
                        bool myOver = false;
                        for (uint32_t j = 0; j < want; j++) {
                            if (br.pos() > br.total) { if (lane == j) myOver = true; cnt = j + 1; break; }
                            int64_t ll = llS.baseline, ml = mlS.baseline, mo = ofS.baseline;
                            const uint32_t moB = (ofS.bits >> 8) & 0xff;
                            mo += br.read(moB);
                            {   // match-length and literal-length extra bits in one read (<= 16 + 16)
                                const uint32_t nML = (mlS.bits >> 8) & 0xff, nLL = (llS.bits >> 8) & 0xff;
                                const uint32_t both = br.read(nML + nLL);
                                ml += both >> nLL;
                                ll += both & ((1u << nLL) - 1);
                            }
                            if (moB > 1) { rep2 = rep1; rep1 = rep0; rep0 = mo; }
                            else {
                                if (ll == 0) mo++;
                                if (mo == 0) mo = rep0;
                                else {
                                    int64_t temp = (mo == 3) ? rep0 - 1 : (mo == 1 ? rep1 : rep2);
                                    if (temp == 0) temp = 1;
                                    if (mo != 1) rep2 = rep1;
                                    rep1 = rep0; rep0 = temp; mo = temp;
                                }
                            }
                            if (lane == j) { myLL = (uint32_t)ll; myML = (uint32_t)ml; myMO = (uint32_t)mo; }
                            if (base + j + 1 == nSeqs) break;
                            const uint32_t nl = llS.bits & 0xff, nm = mlS.bits & 0xff, no = ofS.bits & 0xff;
                            const uint32_t all = br.read(nl + nm + no);      // <= 9 + 9 + 8 state bits in one read
                            const uint32_t bl = all >> (nm + no), bm = (all >> no) & ((1u << nm) - 1), bo = all & ((1u << no) - 1);
                            llS = dec_lookup(cur[0], dc->codeTab[0], ((llS.bits >> 16) + bl) & ((1u << DEC_TLOG_MAX) - 1));
                            mlS = dec_lookup(cur[2], dc->codeTab[2], ((mlS.bits >> 16) + bm) & ((1u << DEC_TLOG_MAX) - 1));
                            ofS = dec_lookup(cur[1], dc->codeTab[1], ((ofS.bits >> 16) + bo) & ((1u << DEC_TLOG_MAX) - 1));
                        }
                        // ---- placement (ll, ml < 2^18 each, 32 of them: no overflow in 32 bits)
                        const bool mine = lane < cnt;
                        const uint32_t lenIncl = warp_scan_incl(mine ? myLL + myML : 0u);
                        const uint32_t llIncl = warp_scan_incl(mine ? myLL : 0u);
                        const uint64_t myOut = o + (lenIncl - (myLL + myML));        // where my literals go
                        const uint32_t myLit = litPos + (llIncl - myLL);
                        const uint64_t myDst = myOut + myLL;                          // where my match goes
                        // ---- the reference's per-sequence checks, in its order; the first failing sequence decides
                        int err = 0;
                        if (mine) {
                            if (myOver) err = DEC_ERR_CORRUPT;
                            else if (myLit > litRegen || myLL > litRegen - myLit) err = DEC_ERR_CORRUPT;
                            else if (myDst + myML - startSize > maxBlockSize) err = DEC_ERR_CORRUPT;
                            else if (myML > DEC_MAX_MATCHLEN) err = DEC_ERR_CORRUPT;
                            else if (myDst + myML > outCap) err = DEC_ERR_DST;
                            else if (myMO == 0 && myML > 0) err = DEC_ERR_CORRUPT;
                            else if ((uint64_t)myMO > myDst || (uint64_t)myMO > windowSize) err = DEC_ERR_CORRUPT;
                        }
                        const unsigned bad = __ballot_sync(FULLMASK, err != 0);
                        if (bad) DFAIL(__shfl_sync(FULLMASK, err, __ffs((int)bad) - 1));
                        // ---- literal runs of the batch: their source bytes are one contiguous piece of the literal buffer, so
                        // the warp copies it with coalesced loads; every byte finds its sequence by a binary search over the
                        // lanes' inclusive literal counts (a per-lane byte loop left 31 lanes waiting on one L2 round trip
                        // per byte)
                        {
                            const uint32_t totLit = __shfl_sync(FULLMASK, llIncl, 31);
                            const uint32_t myOutLo = (uint32_t)myOut, myOutHi = (uint32_t)(myOut >> 32);
                            for (uint32_t k0 = 0; k0 < totLit; k0 += 32) {
                                const uint32_t k = k0 + lane;
                                uint32_t owner = 0;                       // first lane whose inclusive count exceeds k
#pragma unroll
                                for (int st = 16; st > 0; st >>= 1) {
                                    const uint32_t c = owner + st - 1;
                                    const uint32_t v = __shfl_sync(FULLMASK, llIncl, (int)(c & 31));
                                    if (c < 32 && v <= k) owner += st;
                                }
                                const uint32_t oIncl = __shfl_sync(FULLMASK, llIncl, (int)(owner & 31));
                                const uint32_t oLL = __shfl_sync(FULLMASK, myLL, (int)(owner & 31));
                                const uint32_t oLo = __shfl_sync(FULLMASK, myOutLo, (int)(owner & 31));
                                const uint32_t oHi = __shfl_sync(FULLMASK, myOutHi, (int)(owner & 31));
                                if (k < totLit) {
                                    const uint64_t ob = ((uint64_t)oHi << 32) | oLo;
                                    out[ob + (k - (oIncl - oLL))] = literals[litPos + k];
                                }
                            }
                        }
                        __syncwarp();
                        // ---- matches in waves
                        bool pending = mine && myML > 0;
                        const uint64_t srcEnd = (myMO >= myML) ? myDst - myMO + myML : myDst;   // self-overlap: source ends at dst
                        for (;;) {
                            uint64_t key = pending ? myDst : ~0ull;
                            uint32_t khi = (uint32_t)(key >> 32), klo = (uint32_t)key;
                            // 64-bit minimum over the warp
#pragma unroll
                            for (int d = 16; d > 0; d >>= 1) {
                                const uint32_t ohi = __shfl_xor_sync(FULLMASK, khi, d), olo = __shfl_xor_sync(FULLMASK, klo, d);
                                if (ohi < khi || (ohi == khi && olo < klo)) { khi = ohi; klo = olo; }
                            }
                            const uint64_t minDst = ((uint64_t)khi << 32) | klo;
                            if (minDst == ~0ull) break;
                            const bool ready = pending && (srcEnd <= minDst || myDst == minDst);
                            const bool longM = ready && myML >= 96;
                            if (ready && !longM) {
                                const uint8_t *from = out + myDst - myMO;
                                uint32_t k = 0;
                                if (myMO >= 8) {        // the next eight source bytes are final: eight loads in flight per round trip
                                    for (; k + 8 <= myML; k += 8) {
                                        uint8_t t8[8];
#pragma unroll
                                        for (int q = 0; q < 8; q++) t8[q] = from[k + q];
#pragma unroll
                                        for (int q = 0; q < 8; q++) out[myDst + k + q] = t8[q];
                                    }
                                }
                                for (; k < myML; k++) out[myDst + k] = from[k];
                            }
                            for (unsigned m = __ballot_sync(FULLMASK, longM); m; m &= m - 1) {
                                const int f = __ffs((int)m) - 1;
                                const uint64_t fd = __shfl_sync(FULLMASK, myDst, f);
                                const uint32_t fo = __shfl_sync(FULLMASK, myMO, f), fn = __shfl_sync(FULLMASK, myML, f);
                                const uint8_t *from = out + fd - fo;
                                if (fo >= fn || fo >= 32) {
                                    for (uint32_t k0 = 0; k0 < fn; k0 += 32) {
                                        const uint32_t k = k0 + lane;
                                        if (k < fn) out[fd + k] = from[k];
                                        if (fo < fn) __syncwarp();
                                    }
                                } else {
                                    for (uint32_t k = lane; k < fn; k += 32) out[fd + k] = from[k % fo];
                                }
                                __syncwarp();
                            }
                            if (ready) pending = false;
                            __syncwarp();
                        }
                        o += __shfl_sync(FULLMASK, lenIncl, 31);
                        litPos += __shfl_sync(FULLMASK, llIncl, 31);
                    }
                    uint32_t rest = litRegen - litPos;
                    if ((uint64_t)rest + o - startSize > maxBlockSize) DFAIL(DEC_ERR_CORRUPT);
                    if (o + rest > outCap) DFAIL(DEC_ERR_DST);
                    for (uint32_t k = lane; k < rest; k += 32) out[o + k] = literals[litPos + k];
                    o += rest;
                    if (br.pos() != br.total) DFAIL(DEC_ERR_CORRUPT);
                }
                __syncwarp();
            }
            if (o > (64ull << 30)) DFAIL(DEC_ERR_SIZE);
            if (o > fcs) DFAIL(DEC_ERR_SIZE);
            if (last) break;
        }
        __syncwarp();
        if (fcs != ~0ull && o != fcs) DFAIL(DEC_ERR_SIZE);
        if (hasCheck) {
            if (n - ip < 4) DFAIL(DEC_ERR_CORRUPT);
            uint32_t want = (uint32_t)src[ip] | ((uint32_t)src[ip + 1] << 8) | ((uint32_t)src[ip + 2] << 16) | ((uint32_t)src[ip + 3] << 24);
            ip += 4;
            // XXH64 of the frame content: lanes 0..3 hold the accumulators
            const uint64_t P1 = 11400714785074694791ull, P2 = 14029467366897019727ull, P3 = 1609587929392839161ull,
                           P4 = 9650029242287828579ull, P5 = 2870177450012600261ull;
            uint64_t h;
            uint64_t p = 0;
            if (o >= 32) {
                uint64_t v = (lane == 0) ? P1 + P2 : (lane == 1) ? P2 : (lane == 2) ? 0ull : (0ull - P1);
                uint64_t stripes = o / 32;
                if (lane < 4) {
                    if ((reinterpret_cast<uintptr_t>(out) & 7) == 0) {
                        const uint64_t *q8 = reinterpret_cast<const uint64_t *>(out) + lane;
                        for (uint64_t i = 0; i < stripes; i++) {
                            uint64_t inw = q8[4 * i];
                            v += inw * P2; v = (v << 31) | (v >> 33); v *= P1;
                        }
                    } else {
                        for (uint64_t i = 0; i < stripes; i++) {
                            const uint8_t *q = out + 32 * i + 8 * lane;
                            uint64_t inw = 0;
#pragma unroll
                            for (int b = 0; b < 8; b++) inw |= (uint64_t)q[b] << (8 * b);
                            v += inw * P2; v = (v << 31) | (v >> 33); v *= P1;
                        }
                    }
                }
                p = stripes * 32;
                uint64_t v1 = __shfl_sync(FULLMASK, v, 0), v2 = __shfl_sync(FULLMASK, v, 1), v3 = __shfl_sync(FULLMASK, v, 2), v4 = __shfl_sync(FULLMASK, v, 3);
                h = ((v1 << 1) | (v1 >> 63)) + ((v2 << 7) | (v2 >> 57)) + ((v3 << 12) | (v3 >> 52)) + ((v4 << 18) | (v4 >> 46));
#define XMERGE(vv) do { uint64_t t_ = (vv) * P2; t_ = (t_ << 31) | (t_ >> 33); t_ *= P1; h ^= t_; h = h * P1 + P4; } while (0)
                XMERGE(v1); XMERGE(v2); XMERGE(v3); XMERGE(v4);
#undef XMERGE
            } else h = P5;
            h += o;
            while (p + 8 <= o) {
                uint64_t k1 = 0;
                for (int b = 0; b < 8; b++) k1 |= (uint64_t)out[p + b] << (8 * b);
                k1 *= P2; k1 = (k1 << 31) | (k1 >> 33); k1 *= P1;
                h ^= k1; h = ((h << 27) | (h >> 37)) * P1 + P4; p += 8;
            }
            if (p + 4 <= o) {
                uint32_t k4 = 0;
                for (int b = 0; b < 4; b++) k4 |= (uint32_t)out[p + b] << (8 * b);
                h ^= (uint64_t)k4 * P1; h = ((h << 23) | (h >> 41)) * P2 + P3; p += 4;
            }
            while (p < o) { h ^= (uint64_t)out[p] * P5; h = ((h << 11) | (h >> 53)) * P1; p++; }
            h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
            if ((uint32_t)h != want) DFAIL(DEC_ERR_CRC);
        }
        total += o;
        if (n - ip == 0) break;
    }
#undef DFAIL
    return (int64_t)total;
}

B2C_DEV void zstd_decode_warp(uint8_t *smem, const ZstdDecParams &P, uint32_t warpGlobal, uint32_t totalWarps) {
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    DecWarp *dw = reinterpret_cast<DecWarp *>(smem + w * DEC_WARP_BYTES);
    DecCta *dc = reinterpret_cast<DecCta *>(smem + DEC_WARPS * DEC_WARP_BYTES);
    uint8_t *litbuf = P.lit_scratch + (uint64_t)warpGlobal * DEC_LIT_SCRATCH;
    // code -> (baseline, extra bits) maps, one copy per CTA
    for (uint32_t i = threadIdx.x; i < 3 * 64; i += blockDim.x) {
        const int which = (int)(i / 64);
        const uint32_t c = i % 64;
        uint32_t base = 0, bits = 0;
        if (c < dec_code_limit(which)) dec_code_base(which, c, &base, &bits);
        dc->codeTab[which][c] = make_uint2(base, bits);
    }
    __syncthreads();
    dec_build_predef(dw, lane);
    for (uint32_t c = warpGlobal; c < P.nchunks; c += totalWarps) {
        if (P.fd && dec_staged_done(P, c)) continue;
        const uint8_t *src = P.src_base + (P.src_offsets ? P.src_offsets[c] : (uint64_t)c * P.src_stride);
        uint8_t *dst = P.dst_base + (P.dst_offsets ? P.dst_offsets[c] : (uint64_t)c * P.dst_stride);
        uint32_t cap = P.dst_caps ? P.dst_caps[c] : P.dst_cap;
        __syncwarp();
        int64_t r = zstd_decode_input(dw, dc, src, P.src_sizes[c], dst, cap, litbuf, lane);
        __syncwarp();
        if (lane == 0) P.out_sizes[c] = r;
    }
}

#ifndef B2C_EMU
extern "C" __global__ void __launch_bounds__(DEC_WARPS * 32) b2c_zstd_decode_kernel(ZstdDecParams P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    zstd_decode_warp(smem, P, blockIdx.x * DEC_WARPS + (threadIdx.x >> 5), gridDim.x * DEC_WARPS);
}
#endif

}  // namespace b2c
