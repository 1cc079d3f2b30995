// compress_b200/csrc/b2c_huf0.cuh -- standalone huff0 block coder for sm_100a.
//
// Compress: one CTA per block (<= 262143 bytes, huff0.BlockSizeMax), what huff0.Compress4X / Compress1X do for a fresh
// Scratch (huff0/compress.go:14-141): histogram, table (huffSort + two-queue tree + setMaxHeight, identical
// tie-breaking), table description (FSE-compressed or 4-bit weights), 1 or 4 streams.  The output bytes equal the
// reference's for the same input (tests compare with the oracle), errors map to ErrIncompressible / ErrUseRLE /
// ErrTooBig.  Every thread encodes a contiguous run of symbols straight into the (zeroed) destination; bit runs of
// neighbouring threads are merged with atomicOr on the boundary words.
// Decompress: one warp per block: ReadTable (huff0/decompress.go:29-166) + Decompress4X / 1X
// (huff0/decompress.go:234,622; exact bit consumption as decompress_generic.go), 4 streams on 4 lanes.
#pragma once
#include "b2c_huff.cuh"
#include "b2c_zstd_dec.cuh"
#include "b2c_zstd_dec_staged.cuh"

namespace b2c {

constexpr int HUF0_NT = 1024;
constexpr uint32_t HUF0_BLOCK_MAX = (1u << 18) - 1;   // huff0.BlockSizeMax (huff0/huff0.go:27)
enum { HUF0_FLAG_4X = 1 };
enum { HUF0_ERR_INCOMPRESSIBLE = -1, HUF0_ERR_USE_RLE = -2, HUF0_ERR_TOO_BIG = -3, HUF0_ERR_DST = -4, HUF0_ERR_CORRUPT = -5, HUF0_ERR_UNSUPPORTED = -11 };

struct Huf0Shared {
    HufWork hw;
    uint32_t whist[(HUF0_NT / 32) * 256];
    uint32_t flag;
};
constexpr uint32_t HUF0_SMEM_BYTES = ((sizeof(Huf0Shared) + 15) / 16) * 16;

struct Huf0Params {
    const uint8_t *src_base; uint64_t src_stride; const uint32_t *src_sizes; uint32_t src_size_all;
    uint8_t *dst_base; uint64_t dst_stride; uint32_t dst_cap;
    const uint32_t *dst_sizes;   // decompress: exact decoded size of every block (Decompress4X(in, dstSize))
    int64_t *out_sizes;
    uint32_t nchunks;
    uint32_t flags;
    // staged decompress: records of the preparation pass; state 0 = decoded by the stream kernel, FD_LEGACY = this kernel's
    // job, HUF0_STATE_ERR = the error in pad[0]
    FdChunk *fd;
    FdBlock *fd_blk;             // one block record per input (fd_maxb = 1)
    uint16_t *fd_huf;
};
constexpr uint32_t HUF0_STATE_ERR = 2;

B2C_DEV void huf0_compress_block(Huf0Shared *sh, const Huf0Params &P, uint32_t chunk) {
    const unsigned tid = threadIdx.x;
    HufWork *hw = &sh->hw;
    const uint8_t *in = P.src_base + (uint64_t)chunk * P.src_stride;
    uint8_t *out = P.dst_base + (uint64_t)chunk * P.dst_stride;
    const uint32_t n = P.src_sizes ? P.src_sizes[chunk] : P.src_size_all;
    const bool four = (P.flags & HUF0_FLAG_4X) != 0;
    int64_t result = 0;
    if (n > HUF0_BLOCK_MAX) {                        // prepare(): ErrTooBig (huff0/huff0.go:135)
        if (tid == 0) P.out_sizes[chunk] = HUF0_ERR_TOO_BIG;
        return;
    }
    huf_histogram(in, n, sh->whist, hw, tid, HUF0_NT, 0);   // countSimple (compress.go:351), all 32 warps
    if (tid == 0) { hw->status = HUF_INCOMPRESSIBLE; hw->tableDescLen = 0; hw->tableLog = 0; }
    __syncthreads();
    huf_build_table(hw, n, tid, HUF0_NT, 0);         // sets hw->status (compress.go:66-80 early outs included)
    __syncthreads();
    const int status = hw->status;
    if (status != HUF_OK) result = (status == HUF_USE_RLE) ? HUF0_ERR_USE_RLE : HUF0_ERR_INCOMPRESSIBLE;
    else if (four && n < 12) result = HUF0_ERR_INCOMPRESSIBLE;          // compress4X (compress.go:270)
    if (result == 0) {
        HufEncState st;
        const uint32_t total = huf_enc_sizes(hw, sh->whist, in, n, four ? 1 : 0, tid, HUF0_NT, 0, &st);   // whist is free again: reused as the packed code table
        bool bad = total >= n;                                          // wantSize = len(in) (WantLogLess 0)
        if (four) for (int k = 0; k < 4; k++) bad = bad || hw->streamBytes[k] > 65535;   // jump table limit (:288)
        if (bad) result = HUF0_ERR_INCOMPRESSIBLE;
        else if (total > P.dst_cap) result = HUF0_ERR_DST;
        else {
            // zero the destination words the bit runs are OR-ed into
            uint32_t *o32 = reinterpret_cast<uint32_t *>(out);
            for (uint32_t i = tid; i < (total + 3) / 4; i += HUF0_NT) o32[i] = 0;
            __syncthreads();
            huf_enc_pack(hw, sh->whist, in, four ? 1 : 0, out, 0, tid, HUF0_NT, 0, &st);
            result = (int64_t)total;
        }
    }
    __syncthreads();
    if (tid == 0) P.out_sizes[chunk] = result;
}

// one warp per block; dw = this warp's DecWarp scratch
B2C_DEV int64_t huf0_decompress_block(DecWarp *dw, const uint8_t *src, uint32_t n, uint8_t *dst, uint32_t dstSize, bool four,
                                      unsigned lane) {
    if (dstSize > HUF0_BLOCK_MAX) return HUF0_ERR_TOO_BIG;             // MaxDecodedSize default = BlockSizeMax
    uint32_t tl = 0;
    const int used = dec_huf_read_table(dw, src, n, &tl, lane);
    if (used == -2) return HUF0_ERR_UNSUPPORTED;     // weight table description beyond this decoder's limits (fse tableLog > 9)
    if (used < 0) return HUF0_ERR_CORRUPT;
    const int e = dec_huf_streams<true>(dw->hufDt, tl, src + used, n - (uint32_t)used, dst, dstSize, four, lane);
    if (__any_sync(FULLMASK, e != 0)) return HUF0_ERR_CORRUPT;
    return (int64_t)dstSize;
}

// Staged decompress, pass 1 (one warp per block): ReadTable, the decoding table to global memory and a one-block record for
// the literal-stream kernel of the staged zstd decoder (fd_lit_warp: four lanes per block, eight blocks per warp, two-level
// tables in shared memory, 16-byte stores) -- the same Huffman streams, so the same kernel.  Whatever is unusual keeps the
// one-warp path (state FD_LEGACY), which also produces the error values.
B2C_DEV void huf0_prep_block(DecWarp *dw, const Huf0Params &P, uint32_t c, unsigned lane) {
    const uint8_t *src = P.src_base + (uint64_t)c * P.src_stride;
    const uint32_t n = P.src_sizes[c], want = P.dst_sizes ? P.dst_sizes[c] : P.dst_cap;
    FdChunk *ck = P.fd + c;
    uint32_t state = FD_LEGACY;
    int64_t err = 0;
    uint32_t tl = 0;
    int used = 0;
    if (P.dst_stride && want > P.dst_stride && want <= HUF0_BLOCK_MAX) { state = HUF0_STATE_ERR; err = HUF0_ERR_DST; }
    else if (want > HUF0_BLOCK_MAX) { state = HUF0_STATE_ERR; err = HUF0_ERR_TOO_BIG; }
    else {
        used = dec_huf_read_table(dw, src, n, &tl, lane);
        __syncwarp();
        if (used == -2) { state = HUF0_STATE_ERR; err = HUF0_ERR_UNSUPPORTED; }
        else if (used < 0) { state = HUF0_STATE_ERR; err = HUF0_ERR_CORRUPT; }
        else if (n - (uint32_t)used < (1u << 18) && want > 0 && (reinterpret_cast<uintptr_t>(P.dst_base + (uint64_t)c * P.dst_stride) & 3) == 0) {
            uint16_t *dt = P.fd_huf + (uint64_t)c * 2048;
            for (uint32_t i = lane; i < (1u << tl); i += 32) dt[i] = dw->hufDt[i];
            state = 0;
        }
    }
    __syncwarp();
    if (lane == 0) {
        ck->state = state; ck->nBlocks = 1; ck->hasCheck = 0; ck->check = 0; ck->fcs = want; ck->windowSize = 0;
        ck->pad[0] = (uint32_t)(int32_t)err;
        FdBlock *bk = P.fd_blk + c;
        bk->type = 2; bk->size = 0; bk->srcOff = 0; bk->litKind = 2; bk->litOff = 0; bk->litRegen = want;
        bk->nSeqs = 0; bk->bitsOff = 0; bk->bitsLen = 0; bk->tab[0] = bk->tab[1] = bk->tab[2] = 0; bk->tlog = 0; bk->seqOff = 0;
        bk->hufOff = (uint32_t)(used > 0 ? used : 0);
        bk->hufInfo = ((n - (uint32_t)(used > 0 ? used : 0)) & 0x3ffffu) | (((P.flags & HUF0_FLAG_4X) ? 1u : 0u) << 18) | (tl << 19);
    }
}

// huff0.ReadTable (huff0/decompress.go:29-166) as a service: one warp per input.  Row i of `tables` (260 bytes) receives
// [0] actualTableLog, [1] 0, [2..3] bytes the table description occupies (LE), [4..259] the code length of every symbol
// (0 = not present); out_sizes[i] = the same byte count, or a negative error.
B2C_DEV void huf0_read_table_block(DecWarp *dw, const uint8_t *src, uint32_t n, uint8_t *row, int64_t *res, unsigned lane) {
    for (uint32_t i = lane; i < 260; i += 32) dw->weight[i] = 0;
    __syncwarp();
    uint32_t tl = 0;
    const int used = dec_huf_read_table(dw, src, n, &tl, lane);
    __syncwarp();
    if (used < 0) { if (lane == 0) *res = (used == -2) ? HUF0_ERR_UNSUPPORTED : HUF0_ERR_CORRUPT; return; }
    for (uint32_t sym = lane; sym < 256; sym += 32) {
        const uint32_t wgt = dw->weight[sym];
        row[4 + sym] = (uint8_t)(wgt ? tl + 1 - wgt : 0);
    }
    if (lane == 0) { row[0] = (uint8_t)tl; row[1] = 0; row[2] = (uint8_t)used; row[3] = (uint8_t)(used >> 8); *res = used; }
}

#ifndef B2C_EMU
extern "C" __global__ void __launch_bounds__(DEC_WARPS * 32) b2c_huf_read_table_kernel(Huf0Params P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    DecWarp *dw = reinterpret_cast<DecWarp *>(smem + w * DEC_WARP_BYTES);
    const uint32_t totalWarps = gridDim.x * DEC_WARPS;
    for (uint32_t c = blockIdx.x * DEC_WARPS + w; c < P.nchunks; c += totalWarps) {
        __syncwarp();
        huf0_read_table_block(dw, P.src_base + (uint64_t)c * P.src_stride, P.src_sizes[c], P.dst_base + (uint64_t)c * P.dst_stride,
                              P.out_sizes + c, lane);
        __syncwarp();
    }
}
extern "C" __global__ void __launch_bounds__(DEC_WARPS * 32) b2c_huf_dec_prep_kernel(Huf0Params P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    DecWarp *dw = reinterpret_cast<DecWarp *>(smem + w * DEC_WARP_BYTES);
    const uint32_t totalWarps = gridDim.x * DEC_WARPS;
    for (uint32_t c = blockIdx.x * DEC_WARPS + w; c < P.nchunks; c += totalWarps) {
        __syncwarp();
        huf0_prep_block(dw, P, c, lane);
        __syncwarp();
    }
}
extern "C" __global__ void __launch_bounds__(HUF0_NT, 1) b2c_huf_compress_kernel(Huf0Params P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    Huf0Shared *sh = reinterpret_cast<Huf0Shared *>(smem);
    for (uint32_t c = blockIdx.x; c < P.nchunks; c += gridDim.x) { huf0_compress_block(sh, P, c); __syncthreads(); }
}
extern "C" __global__ void __launch_bounds__(DEC_WARPS * 32) b2c_huf_decompress_kernel(Huf0Params P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    DecWarp *dw = reinterpret_cast<DecWarp *>(smem + w * DEC_WARP_BYTES);
    const uint32_t totalWarps = gridDim.x * DEC_WARPS;
    for (uint32_t c = blockIdx.x * DEC_WARPS + w; c < P.nchunks; c += totalWarps) {
        __syncwarp();
        const uint32_t want = P.dst_sizes ? P.dst_sizes[c] : P.dst_cap;
        if (P.fd) {             // staged call: only what the stream kernel left
            const uint32_t st = P.fd[c].state;
            if (st == 0) { if (lane == 0) P.out_sizes[c] = (int64_t)want; continue; }
            if (st == HUF0_STATE_ERR) { if (lane == 0) P.out_sizes[c] = (int64_t)(int32_t)P.fd[c].pad[0]; continue; }
        }
        // the exact decoded size must fit the block's slot (it is the caller's number, not the stream's)
        const int64_t r = (P.dst_stride && want > P.dst_stride && want <= HUF0_BLOCK_MAX) ? (int64_t)HUF0_ERR_DST
                          : huf0_decompress_block(dw, P.src_base + (uint64_t)c * P.src_stride, P.src_sizes[c],
                                                  P.dst_base + (uint64_t)c * P.dst_stride, want, (P.flags & HUF0_FLAG_4X) != 0, lane);
        __syncwarp();
        if (lane == 0) P.out_sizes[c] = r;
    }
}
#endif

}  // namespace b2c
