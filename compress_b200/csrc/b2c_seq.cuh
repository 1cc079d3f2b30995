// compress_b200/csrc/b2c_seq.cuh -- zstd sequence section on the device.
//
// B200-native replacement for the reference's
//   zstd/blockenc.go:601-610 (nSeq header), :611-724 (modes, chooseComp, NCount tables),
//   :725-808 (3-state backward FSE bitstream), :831-893 (genCodes)
//   zstd/seqenc.go:48-112 (llCode/mlCode/ofCode + extra-bit tables)
//   zstd/fse_encoder.go (normalizeCount/buildCTable/writeCount/approxSize, via b2c_fse.cuh)
//   zstd/fse_predefined.go:118-156 (default distributions)
// The three tANS state chains are serial in the reference.  Here each chain is
// cut into 32 segments walked by the 32 lanes of one warp: a lane first runs a
// short warm-up over the symbols preceding its segment (tANS encoder states
// forget their past at ~nbBits per step), then its segment; lanes whose assumed
// start state disagrees with the predecessor's verified final state re-run, so
// the result is always the exact serial chain.  Bits are then packed by all
// threads at prefix-summed offsets.  Output bytes equal the oracle's.
#pragma once
#include "b2c_common.cuh"
#include "b2c_fse.cuh"

namespace b2c {

enum { TBL_LL = 0, TBL_OF = 1, TBL_ML = 2 };
constexpr int SEQ_WARMUP = 24;
constexpr uint32_t SEQ_TABLE_ERR = 0xffffffffu;  // ncountLen marker: table construction failed

struct SeqWork {
    FseCTable cur[3];
    FseCTable predef[3];
    uint32_t hist[3][64];
    uint32_t maxSym[3];
    uint32_t mode[3];          // 0 predefined, 1 RLE, 2 FSE
    uint32_t used[3];          // 0 -> predef[i], 1 -> cur[i]
    uint8_t ncount[3][96];
    uint32_t ncountLen[3];
    uint32_t finalState[3];
    uint32_t scan[40];
    uint32_t totalBits;
    int32_t err;
    uint16_t fseScratch[3][200];   // fse_build_ctable_warp
};

B2C_DEV uint32_t seq_ll_code(uint32_t litLength) {
    // seqenc.go:48-75
    if (litLength <= 15) return litLength;
    if (litLength <= 63) {
        if (litLength < 24) return 16 + ((litLength - 16) >> 1);
        if (litLength < 32) return 20 + ((litLength - 24) >> 2);
        if (litLength < 40) return 22;
        if (litLength < 48) return 23;
        return 24;
    }
    return highbit32(litLength) + 19;
}
B2C_DEV uint32_t seq_ml_code(uint32_t mlBase) {
    // seqenc.go:77-107
    if (mlBase <= 31) return mlBase;
    if (mlBase <= 127) {
        if (mlBase < 40) return 32 + ((mlBase - 32) >> 1);
        if (mlBase < 48) return 36 + ((mlBase - 40) >> 2);
        if (mlBase < 56) return 38;
        if (mlBase < 64) return 39;
        if (mlBase < 80) return 40;
        if (mlBase < 96) return 41;
        return 42;
    }
    return highbit32(mlBase) + 36;
}
B2C_DEV uint32_t seq_ll_bits(uint32_t code) {  // llBitsTable, seqenc.go:61-66
    if (code < 16) return 0;
    if (code < 20) return 1;
    if (code < 22) return 2;
    if (code < 24) return 3;
    if (code == 24) return 4;
    return code - 19;  // 25 -> 6, 26 -> 7 ... 35 -> 16
}
B2C_DEV uint32_t seq_ml_bits(uint32_t code) {  // mlBitsTable, seqenc.go:90-97
    if (code < 32) return 0;
    if (code < 36) return 1;
    if (code < 38) return 2;
    if (code < 40) return 3;
    if (code < 42) return 4;
    if (code == 42) return 5;
    return code - 36;  // 43 -> 7 ... 52 -> 16
}

// Build the three predefined encoder tables (one thread each; tid 0..2 of the caller's choice).
B2C_DEV void seq_build_predef(SeqWork *sw, int which) {
    const int8_t llN[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2,
                            2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
    const int8_t ofN[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1,
                            1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};
    const int8_t mlN[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                            1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                            1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
    FseCTable *ct = &sw->predef[which];
    for (int i = 0; i < FSE_MAX_SYM; i++) ct->norm[i] = 0;
    if (which == TBL_LL) { for (int i = 0; i < 36; i++) ct->norm[i] = llN[i]; ct->symbolLen = 36; ct->tableLog = 6; }
    else if (which == TBL_OF) { for (int i = 0; i < 29; i++) ct->norm[i] = ofN[i]; ct->symbolLen = 29; ct->tableLog = 5; }
    else { for (int i = 0; i < 53; i++) ct->norm[i] = mlN[i]; ct->symbolLen = 53; ct->tableLog = 6; }
    ct->useRLE = 0; ct->rleVal = 0;
    fse_build_ctable(ct);
}

// fseEncoder.optimalTableLog (fse_encoder.go:429-455)
B2C_DEV uint32_t seq_optimal_tablelog(uint32_t length, uint32_t symbolLen) {
    uint8_t tableLog = 8;
    uint32_t minBitsSrc = fse_hb(length) + 1;
    uint32_t minBitsSymbols = fse_hb(symbolLen - 1) + 2;
    uint8_t minBits = (uint8_t)minBitsSymbols;
    if (minBitsSrc < minBitsSymbols) minBits = (uint8_t)minBitsSrc;
    uint8_t maxBitsSrc = (uint8_t)((uint8_t)fse_hb(length - 1) - 2);
    if (maxBitsSrc < tableLog) tableLog = maxBitsSrc;
    if (minBits > tableLog) tableLog = minBits;
    if (tableLog < 5) tableLog = 5;
    if (tableLog > 8) tableLog = 8;
    return tableLog;
}

// One warp builds table `which` from sw->hist[which] (fresh block: no previous tables); every lane calls.
// firstCode = code of sequence 0 (setRLE uses b.sequences[0]).  Lane 0 runs the serial steps (normalisation, size
// estimates, NCount), the table itself is filled by all lanes.
B2C_DEV void seq_build_table(SeqWork *sw, int which, uint32_t nseq, uint32_t firstCode, unsigned lane) {
    FseCTable *ct = &sw->cur[which];
    const uint32_t *hist = sw->hist[which];
    const uint32_t symbolLen = sw->maxSym[which] + 1;
    uint32_t maxCount = warp_max(hist[lane] > hist[lane + 32] ? hist[lane] : hist[lane + 32]);   // bins >= symbolLen are zero
    if (lane == 0) {
        ct->symbolLen = symbolLen;
        ct->tableLog = seq_optimal_tablelog(nseq, symbolLen);
        ct->rleVal = 0;
    }
    if (maxCount == nseq) {
        // useRLE: setRLE(b.sequences[0].code), fse_encoder.go:208-221
        if (lane == 0) {
            ct->useRLE = 1; ct->rleVal = firstCode; ct->tableLog = 0;
            ct->stateTable[0] = 0; ct->deltaNbBits[firstCode] = 0; ct->deltaFindState[firstCode] = 0;
            sw->mode[which] = 1; sw->used[which] = 1;
            sw->ncount[which][0] = (uint8_t)firstCode; sw->ncountLen[which] = 1;
        }
        __syncwarp();
        return;
    }
    ct->norm[lane] = 0; ct->norm[lane + 32] = 0;
    __syncwarp();
    int bad = 0;
    if (lane == 0) {
        ct->useRLE = 0;
        bad = fse_normalize(hist, symbolLen, nseq, ct->tableLog, ct->norm);
    }
    bad = __shfl_sync(FULLMASK, bad, 0);
    __syncwarp();
    if (!bad) bad = fse_build_ctable_warp(ct, sw->fseScratch[which], lane);
    if (bad) {
        if (lane == 0) { sw->mode[which] = 0; sw->used[which] = 0; sw->ncountLen[which] = SEQ_TABLE_ERR; }
        __syncwarp();
        return;
    }
    if (lane == 0) {
        // chooseComp, blockenc.go:633-661 (prev == never valid for an independent block)
        uint32_t nSize = fse_approx_size(ct, hist, symbolLen) + (((symbolLen * ct->tableLog) >> 3) + 3) * 8;
        uint32_t predefSize = fse_approx_size(&sw->predef[which], hist, symbolLen);
        nSize = nSize + ((nSize + 2 * 8 * 16) >> 4);
        if (predefSize <= nSize) { sw->mode[which] = 0; sw->used[which] = 0; sw->ncountLen[which] = 0; }
        else {
            sw->mode[which] = 2; sw->used[which] = 1;
            int w = fse_write_ncount(ct->norm, symbolLen, ct->tableLog, sw->ncount[which]);
            if (w < 0) { sw->mode[which] = 0; sw->used[which] = 0; sw->ncountLen[which] = SEQ_TABLE_ERR; }
            else sw->ncountLen[which] = (uint32_t)w;
        }
    }
    __syncwarp();
}

B2C_DEV const FseCTable *seq_table(const SeqWork *sw, int which) {
    return sw->used[which] ? &sw->cur[which] : &sw->predef[which];
}

// One warp walks chain `which` over t = 1..nseq-1 (t = 0 is the last sequence) and stores
// stb[idx] = (state & mask(nb)) | nb << 12 for idx = nseq-1-t.  codes[] is indexed by sequence.
B2C_DEV void seq_chain(SeqWork *sw, int which, const uint8_t *codes, uint32_t nseq, uint16_t *stb) {
    const FseCTable *ct = seq_table(sw, which);
    unsigned lane = lane_id();
    uint32_t m = nseq - 1;  // number of steps
    if (ct->useRLE) {
        for (uint32_t t = 1 + lane; t <= m; t += 32) stb[nseq - 1 - t] = 0;
        if (lane == 0) sw->finalState[which] = 0;
        __syncwarp();
        return;
    }
    uint32_t seg = (m + 31) / 32;
    uint32_t t0 = 1 + lane * seg;
    uint32_t t1 = t0 + seg;  // exclusive
    if (t0 > m + 1) t0 = m + 1;
    if (t1 > m + 1) t1 = m + 1;
    // assumed start state (state after step t0-1)
    uint32_t start;
    {
        uint32_t tw = (t0 > (uint32_t)SEQ_WARMUP) ? t0 - SEQ_WARMUP : 1;  // first warm-up step
        uint32_t st = fse_init_state(ct, codes[nseq - 1 - (tw - 1)]);
        for (uint32_t t = tw; t < t0; t++) {
            uint32_t sym = codes[nseq - 1 - t];
            uint32_t nb = (st + ct->deltaNbBits[sym]) >> 16;
            st = ct->stateTable[(int32_t)(st >> nb) + (int32_t)ct->deltaFindState[sym]];
        }
        start = st;
    }
    bool need = true;   // segment must be (re)computed
    uint32_t fin = start;
    for (;;) {
        if (need) {
            uint32_t st = start;
            for (uint32_t t = t0; t < t1; t++) {
                uint32_t sym = codes[nseq - 1 - t];
                uint32_t nb = (st + ct->deltaNbBits[sym]) >> 16;
                stb[nseq - 1 - t] = (uint16_t)((st & ((1u << nb) - 1)) | (nb << 12));
                st = ct->stateTable[(int32_t)(st >> nb) + (int32_t)ct->deltaFindState[sym]];
            }
            fin = st;
            need = false;
        }
        // verify against the predecessor's final state; lane 0 is exact by construction
        uint32_t prevFin = __shfl_up_sync(FULLMASK, fin, 1);
        bool bad = (lane > 0) && (t0 <= m) && (prevFin != start);
        unsigned badMask = __ballot_sync(FULLMASK, bad);
        if (badMask == 0) break;
        // only the lowest mismatching lane is guaranteed to see a verified predecessor
        if (lane == (unsigned)(__ffs((int)badMask) - 1)) { start = prevFin; need = true; }
    }
    // final state = state after step m: held by the last lane that has work (or lane 0 when m == 0)
    uint32_t lastLane = (m == 0) ? 0 : (m - 1) / seg;
    uint32_t f = __shfl_sync(FULLMASK, fin, (int)lastLane);
    if (lane == 0) sw->finalState[which] = f;
    __syncwarp();
}

}  // namespace b2c
