// tests/emu/emu_kernels.cpp -- compiles the CUDA kernel sources with g++ on top of the fiber
// SIMT emulator and exposes C entry points for pytest.  TEST INFRASTRUCTURE ONLY (see simt_emu.h).
#include "simt_emu.h"
#include "../../compress_b200/csrc/b2c_zstd_enc.cuh"
#include "../../compress_b200/csrc/b2c_lz.cuh"
#include "../../compress_b200/csrc/b2c_frame.cuh"
#include "../../compress_b200/csrc/b2c_zstd_dec.cuh"
#include "../../compress_b200/csrc/b2c_zstd_dec_staged.cuh"
#include "../../compress_b200/csrc/b2c_s2_dec.cuh"
#include "../../compress_b200/csrc/b2c_s2_stream.cuh"
#include "../../compress_b200/csrc/b2c_huf0.cuh"
#include <vector>
#include <algorithm>
#include <cstdlib>

using namespace b2c;

extern "C" {

void emu_set_lane_order(int desc) { emu::lane_order_desc = desc; }
static int emu_s2_staged = 1, emu_s2_staged_count = 0;
void emu_set_s2_staged(int on) { emu_s2_staged = on; }
int emu_get_s2_staged_count(void) { return emu_s2_staged_count; }     // blocks of the last emu_s2_decode the staged kernels finished
static int emu_dec_maxb = 0;      // 0: the device's policy; else the staged decoder's blocks-per-input (4: per input, > 4: per block)
void emu_set_dec_maxb(int maxb) { emu_dec_maxb = maxb; }

// Encode nchunks chunks laid out contiguously (chunk i = src + i*stride, size sizes[i]) with the same kernels the
// device runs.  level 1 / 2 / 3; parse must be 0 (kept in the signature for the test helpers).  dst slots of dst_stride bytes.  Optional debug dumps (may be null; dbg_lits rows of blockmax bytes).
int emu_zstd_encode_lv(const uint8_t *src, uint64_t stride, const uint32_t *sizes, uint32_t nchunks, uint8_t *dst,
                       uint64_t dst_stride, int64_t *out_sizes, uint32_t flags, uint32_t *dbg_hdr, uint32_t *dbg_seqs,
                       uint8_t *dbg_lits, uint32_t dbg_seq_cap, int level, int parse) {
    const uint32_t blockmax = level >= 2 ? 131072u : 65536u;
    if (parse != 0) return -1;
    std::vector<uint8_t> scratch(std::max<size_t>(16, std::max(LzLayout<1>::SCRATCH_BYTES, std::max(LzLayout<2>::SCRATCH_BYTES, LzLayout<5>::SCRATCH_BYTES))), 0xCD);
    ChunkWork *work = (ChunkWork *)aligned_alloc(16, sizeof(ChunkWork) * (size_t)nchunks);
    memset(work, 0xCD, sizeof(ChunkWork) * (size_t)nchunks);
    const uint64_t pstride = wk_pool_stride(blockmax);
    uint8_t *pool = (uint8_t *)aligned_alloc(16, pstride * (size_t)nchunks);
    memset(pool, 0xCD, pstride * (size_t)nchunks);
    ZstdEncParams P;
    memset(&P, 0, sizeof(P));
    P.src_base = src; P.src_stride = stride; P.src_sizes = sizes; P.src_size_all = 0;
    P.dst_base = dst; P.dst_stride = dst_stride; P.dst_cap = (uint32_t)dst_stride;
    P.out_sizes = out_sizes; P.nchunks = nchunks; P.flags = flags; P.scratch = scratch.data(); P.work = work;
    P.pool = pool; P.pool_stride = pstride; P.maxseq = wk_maxseq(blockmax); P.blockmax = blockmax;
    P.big = blockmax > 65536; P.level = (uint32_t)level;
    P.dbg_hdr = dbg_hdr; P.dbg_seqs = dbg_seqs; P.dbg_lits = dbg_lits; P.dbg_seq_cap = dbg_seq_cap;
    // K1 parse
    {
        if (level >= 3)
            emu::launch(1, LzCfg<5>::NT, LzLayout<5>::SMEM_BYTES, [&]() {
                for (uint32_t c = 0; c < P.nchunks; c++) lz_parse_chunk<5, LZ_MODE_ZSTD>(emu::dyn_smem, P, c, P.scratch);
            });
        else if (level >= 2)
            emu::launch(1, LzCfg<2>::NT, LzLayout<2>::SMEM_BYTES, [&]() {
                for (uint32_t c = 0; c < P.nchunks; c++) lz_parse_chunk<2, LZ_MODE_ZSTD>(emu::dyn_smem, P, c, P.scratch);
            });
        else
            emu::launch(1, LzCfg<1>::NT, LzLayout<1>::SMEM_BYTES, [&]() {
                for (uint32_t c = 0; c < P.nchunks; c++) lz_parse_chunk<1, LZ_MODE_ZSTD>(emu::dyn_smem, P, c, P.scratch);
            });
        emu::launch(1, HIST_NT, HIST_SMEM_BYTES, [&]() {
            for (uint32_t c = 0; c < P.nchunks; c++) zstd_hist_chunk(emu::dyn_smem, P, c);
        });
    }
    // K2 tables
    static TablesShared ts;
    emu::launch(1, TABLES_NT, 0, [&]() {
        if (threadIdx.x < 3) seq_build_predef(&ts.sw, (int)threadIdx.x);
        __syncthreads();
        zstd_tables_loop(&ts, P, 0, 1);
    });
    // K3 chains + K5 xxh64 (four more warps per CTA)
    emu::launch((nchunks + 31) / 32, CHAIN_NT + CHAIN_XXH_NT, CHAIN_SMEM_BYTES, [&]() {
        if (threadIdx.x < CHAIN_NT) zstd_chains_block(reinterpret_cast<uint32_t *>(emu::dyn_smem), P, blockIdx.x * 32);
        else {
            const unsigned t = threadIdx.x - CHAIN_NT;
            zstd_xxh_quad(P, blockIdx.x * 32 + (t >> 2), t & 3, (t & 31) & ~3u);
        }
    });
    // K4 pack
    if (blockmax > 65536)
        emu::launch(nchunks, PACK_NT, PackCfg<131072>::SMEM_BYTES, [&]() { zstd_pack_chunk<131072>(emu::dyn_smem, P, blockIdx.x); });
    else
        emu::launch(nchunks, PACK_NT, PACK_SMEM_BYTES, [&]() { zstd_pack_chunk<65536>(emu::dyn_smem, P, blockIdx.x); });
    free(work);
    free(pool);
    return 0;
}
// Frame mode (b2c_zstd_encode_frames_device) with the device's kernels and launch order: frame f = sizes[f] bytes at
// src + offs[f]; frames are written back to back into dst; out_offsets / out_sizes per frame.  Optional debug dumps per BLOCK
// (rows of the level's virtual block size), n_blocks_out = number of blocks planned, blk_desc_out (optional, 4 u64 per
// block: off, len, hist, flags).
int emu_zstd_encode_frames(const uint8_t *src, const uint64_t *offs, const uint64_t *sizes, uint32_t nframes, uint8_t *dst,
                           uint64_t dst_cap, uint64_t *out_offsets, int64_t *out_sizes, int crc, int level, uint32_t *dbg_hdr,
                           uint32_t *dbg_seqs, uint8_t *dbg_lits, uint32_t dbg_seq_cap, uint32_t dbg_block_cap,
                           uint32_t *n_blocks_out, uint64_t *blk_desc_out) {
    std::vector<EncBlockDesc> blocks;
    std::vector<FrameDesc> frames;
    if (!frame_plan(level, crc != 0, offs, sizes, nframes, blocks, frames)) return -1;
    const uint32_t nblocks = (uint32_t)blocks.size();
    if (n_blocks_out) *n_blocks_out = nblocks;
    if (blk_desc_out)
        for (uint32_t i = 0; i < nblocks && i < dbg_block_cap; i++) {
            blk_desc_out[4 * i] = blocks[i].off; blk_desc_out[4 * i + 1] = blocks[i].len;
            blk_desc_out[4 * i + 2] = blocks[i].hist; blk_desc_out[4 * i + 3] = blocks[i].flags;
        }
    if (dbg_hdr && nblocks > dbg_block_cap) return -2;
    const FrameGeom g = frame_geom(level);
    const uint32_t blockmax = level >= 2 ? 131072u : 65536u;
    const uint64_t slotB = (uint64_t)g.block + 512;
    std::vector<uint8_t> slots((size_t)slotB * nblocks + 64, 0xCD);
    std::vector<int64_t> bsizes(nblocks, 0);
    std::vector<uint8_t> scratch(std::max<size_t>(16, std::max(LzLayout<1>::SCRATCH_BYTES, std::max(LzLayout<2>::SCRATCH_BYTES, LzLayout<5>::SCRATCH_BYTES))), 0xCD);
    ChunkWork *work = (ChunkWork *)aligned_alloc(16, sizeof(ChunkWork) * (size_t)nblocks);
    memset(work, 0xCD, sizeof(ChunkWork) * (size_t)nblocks);
    const uint64_t pstride = wk_pool_stride(blockmax);
    uint8_t *pool = (uint8_t *)aligned_alloc(16, pstride * (size_t)nblocks);
    memset(pool, 0xCD, pstride * (size_t)nblocks);
    ZstdEncParams P;
    memset(&P, 0, sizeof(P));
    P.desc = blocks.data(); P.src_base = src;
    P.dst_base = slots.data(); P.dst_stride = slotB; P.dst_cap = (uint32_t)slotB;
    P.out_sizes = bsizes.data(); P.nchunks = nblocks; P.flags = 0; P.scratch = scratch.data(); P.work = work;
    P.pool = pool; P.pool_stride = pstride; P.maxseq = wk_maxseq(blockmax); P.blockmax = blockmax;
    P.big = blockmax > 65536; P.level = (uint32_t)level;
    P.dbg_hdr = dbg_hdr; P.dbg_seqs = dbg_seqs; P.dbg_lits = dbg_lits; P.dbg_seq_cap = dbg_seq_cap;
    std::vector<uint64_t> xxh(nframes, 0);
    emu::launch((nframes + 3) / 4, 128, 4 * 2 * XXH_TILE, [&]() {
        const unsigned w = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const uint32_t f = blockIdx.x * 4 + w;
        if (f >= nframes || !frames[f].crc) return;
        const uint64_t h = xxh64_warp(src + frames[f].off, frames[f].size, emu::dyn_smem + w * 2 * XXH_TILE, lane);
        if (lane == 0) xxh[f] = h;
    });
    if (level >= 3)
        emu::launch(1, LzCfg<5>::NT, LzLayout<5>::SMEM_BYTES, [&]() {
            for (uint32_t c = 0; c < P.nchunks; c++) lz_parse_chunk<5, LZ_MODE_ZSTD>(emu::dyn_smem, P, c, P.scratch);
        });
    else if (level >= 2)
        emu::launch(1, LzCfg<2>::NT, LzLayout<2>::SMEM_BYTES, [&]() {
            for (uint32_t c = 0; c < P.nchunks; c++) lz_parse_chunk<2, LZ_MODE_ZSTD>(emu::dyn_smem, P, c, P.scratch);
        });
    else
        emu::launch(1, LzCfg<1>::NT, LzLayout<1>::SMEM_BYTES, [&]() {
            for (uint32_t c = 0; c < P.nchunks; c++) lz_parse_chunk<1, LZ_MODE_ZSTD>(emu::dyn_smem, P, c, P.scratch);
        });
    emu::launch(1, HIST_NT, HIST_SMEM_BYTES, [&]() {
        for (uint32_t c = 0; c < P.nchunks; c++) zstd_hist_chunk(emu::dyn_smem, P, c);
    });
    static TablesShared ts;
    emu::launch(1, TABLES_NT, 0, [&]() {
        if (threadIdx.x < 3) seq_build_predef(&ts.sw, (int)threadIdx.x);
        __syncthreads();
        zstd_tables_loop(&ts, P, 0, 1);
    });
    emu::launch((nblocks + 31) / 32, CHAIN_NT, CHAIN_SMEM_BYTES, [&]() {
        zstd_chains_block(reinterpret_cast<uint32_t *>(emu::dyn_smem), P, blockIdx.x * 32);
    });
    if (blockmax > 65536)
        emu::launch(nblocks, PACK_NT, PackCfg<131072>::SMEM_BYTES, [&]() { zstd_pack_chunk<131072>(emu::dyn_smem, P, blockIdx.x); });
    else
        emu::launch(nblocks, PACK_NT, PACK_SMEM_BYTES, [&]() { zstd_pack_chunk<65536>(emu::dyn_smem, P, blockIdx.x); });
    // scan of the block sizes (b2c_scan_sizes_kernel), placement, completion
    std::vector<uint64_t> scan(nblocks + 1, 0), pos(nblocks, 0);
    for (uint32_t i = 0; i < nblocks; i++) scan[i + 1] = scan[i] + (bsizes[i] > 0 ? (uint64_t)bsizes[i] : 0);
    emu::launch(nblocks, 256, 0, [&]() {
        frame_place_block(slots.data(), slotB, bsizes.data(), scan.data(), 0, blocks.data(), frames.data(), dst, dst_cap,
                          pos.data(), 0, blockIdx.x, threadIdx.x, blockDim.x);
    });
    emu::launch((nframes + 127) / 128, 128, 0, [&]() {
        const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
        if (f < nframes) frame_finish_one(frames.data(), pos.data(), bsizes.data(), xxh.data(), dst, dst_cap, out_offsets, out_sizes, f);
    });
    free(work);
    free(pool);
    return 0;
}

// masked CRC32-C of p[0, n) by the device's one-warp routine (the S2 stream checksum)
uint32_t emu_s2_stream_crc(const uint8_t *p, uint32_t n) {
    uint32_t out = 0;
    emu::launch(1, 32, 1024, [&]() {
        uint32_t *tab = reinterpret_cast<uint32_t *>(emu::dyn_smem);
        crc32c_fill_table(tab, threadIdx.x, blockDim.x);
        __syncthreads();
        const uint32_t c = crc32c_warp(p, n, tab, threadIdx.x);
        if (threadIdx.x == 0) out = crc_mask(c);
    });
    return out;
}

int emu_zstd_encode(const uint8_t *src, uint64_t stride, const uint32_t *sizes, uint32_t nchunks, uint8_t *dst,
                    uint64_t dst_stride, int64_t *out_sizes, uint32_t flags, uint32_t *dbg_hdr, uint32_t *dbg_seqs,
                    uint8_t *dbg_lits, uint32_t dbg_seq_cap) {
    return emu_zstd_encode_lv(src, stride, sizes, nchunks, dst, dst_stride, out_sizes, flags, dbg_hdr, dbg_seqs, dbg_lits,
                              dbg_seq_cap, 1, 0);
}

// Decode n inputs (input i = src + src_off[i], src_sizes[i] bytes) into dst + dst_off[i] (capacity dst_caps[i]).
// mode 0: the device's launch sequence (staged kernels, then the one-warp decoder over the inputs they marked; dst_off
// must be increasing with non-overlapping [dst_off[i], dst_off[i] + dst_caps[i]) ranges); mode 1: one-warp decoder only.
// staged_out (optional, n entries): 1 where the staged kernels produced the result.
int emu_zstd_decode_mode(const uint8_t *src, const uint64_t *src_off, const uint32_t *src_sizes, uint32_t n, uint8_t *dst,
                         const uint64_t *dst_off, const uint32_t *dst_caps, int64_t *out_sizes, int mode, uint8_t *staged_out) {
    uint32_t grid = (n + DEC_WARPS - 1) / DEC_WARPS;
    if (grid > 2) grid = 2;
    std::vector<uint8_t> lit((size_t)grid * DEC_WARPS * DEC_LIT_SCRATCH, 0xCD);
    ZstdDecParams P;
    memset(&P, 0, sizeof(P));
    P.src_base = src; P.src_offsets = src_off; P.src_sizes = src_sizes;
    P.dst_base = dst; P.dst_offsets = dst_off; P.dst_caps = dst_caps;
    P.out_sizes = out_sizes; P.nchunks = n; P.lit_scratch = lit.data();
    std::vector<FdChunk> fd;
    std::vector<FdBlock> fdb;
    std::vector<uint32_t> tabs;
    std::vector<uint64_t> seqs;
    std::vector<uint8_t> lits;
    std::vector<uint16_t> hufs;
    if (mode == 0 && n > 0) {
        const uint64_t span = dst_off[n - 1] + dst_caps[n - 1];
        // the device's choice of the staged form (b2c_api.cu launch_decode); emu_dec_maxb overrides it (tests)
        uint32_t maxb = FD_MAXB;
        if (n <= 4096) { maxb = 65536 / n; if (maxb > FD_MAXB_LONG) maxb = FD_MAXB_LONG; }
        if (emu_dec_maxb) maxb = (uint32_t)emu_dec_maxb;
        const bool perBlock = maxb > FD_MAXB;
        fd.resize(n);
        memset(fd.data(), 0xCD, sizeof(FdChunk) * (size_t)n);
        fdb.resize((size_t)n * maxb);
        memset(fdb.data(), 0xCD, sizeof(FdBlock) * (size_t)n * maxb);
        tabs.assign((size_t)n * maxb * FD_TAB_ENTRIES, 0xCDCDCDCDu);
        seqs.assign((size_t)(span / 3) + 2 * (size_t)n + 8, 0xCDCDCDCDCDCDCDCDull);
        lits.assign((size_t)span + 64, 0xCD);
        P.fd = fd.data(); P.fd_blk = fdb.data(); P.fd_maxb = maxb; P.fd_per_block = perBlock ? 1u : 0u;
        P.fd_tabs = tabs.data(); P.fd_seqs = seqs.data(); P.fd_lits = lits.data(); P.fd_lit_stride = 0;
        std::vector<uint32_t> fdc(FD_CONST_ENTRIES);
        hufs.assign((size_t)n * maxb * 2048, 0xCDCD);
        emu::launch(1, 32, DEC_WARP_BYTES, [&]() { fd_init_warp(emu::dyn_smem, fdc.data(), threadIdx.x); });
        P.fd_huf = hufs.data(); P.fd_const = fdc.data();
        const uint32_t units = perBlock ? n * maxb : n;
        emu::launch((n + 31) / 32, 32, 0, [&]() {
            const uint32_t c = blockIdx.x * 32 + threadIdx.x;
            if (c < P.nchunks) fd_scan_lane(P, c);
        });
        if (perBlock) {
            emu::launch((units + 31) / 32, 32, 0, [&]() {
                const uint32_t u = blockIdx.x * 32 + threadIdx.x;
                const uint32_t c = u / P.fd_maxb;
                if (c < P.nchunks) fd_scan_block_lane(P, c, u % P.fd_maxb);
            });
            emu::launch((n + 31) / 32, 32, 0, [&]() {
                const uint32_t c = blockIdx.x * 32 + threadIdx.x;
                if (c < P.nchunks) fd_link_input(P, c);
            });
        }
        {
            const unsigned groups = (units + FD_LIT_GROUP - 1) / FD_LIT_GROUP;
            emu::launch((groups + FD_LIT_WARPS - 1) / FD_LIT_WARPS, FD_LIT_WARPS * 32, FD_LIT_WARPS * FD_LIT_WARP_BYTES, [&]() {
                const unsigned w = threadIdx.x >> 5;
                fd_lit_warp(emu::dyn_smem + w * FD_LIT_WARP_BYTES, P, blockIdx.x * FD_LIT_WARPS + w, threadIdx.x & 31);
            });
        }
        emu::launch((units + 31) / 32, 32, 0, [&]() {
            const uint32_t u = blockIdx.x * 32 + threadIdx.x;
            const uint32_t *bl = P.fd_const + FD_CONST_BASE, *bm = P.fd_const + FD_CONST_BASE + 128;
            if (P.fd_per_block) {
                const uint32_t c = u / P.fd_maxb;
                if (c < P.nchunks) fd_seq_lane(P, c, bl, bm, (int)(u % P.fd_maxb));
            } else if (u < P.nchunks) fd_seq_lane(P, u, bl, bm);
        });
        emu::launch((n + FD_EXEC_WARPS - 1) / FD_EXEC_WARPS, FD_EXEC_WARPS * 32, 0, [&]() {
            const uint32_t c = blockIdx.x * FD_EXEC_WARPS + (threadIdx.x >> 5);
            if (c < P.nchunks) fd_exec_input(P, c, threadIdx.x & 31);
        });
        if (perBlock)
            emu::launch((n + 3) / 4, 128, 4 * 2 * XXH_TILE, [&]() {
                const uint32_t c = blockIdx.x * 4 + (threadIdx.x >> 5);
                if (c < P.nchunks) fd_xxh_warp(P, c, emu::dyn_smem + (threadIdx.x >> 5) * 2 * XXH_TILE, threadIdx.x & 31);
            });
        else
            emu::launch((unsigned)(((uint64_t)n * 4 + 127) / 128), 128, 0, [&]() {
                const unsigned gt = blockIdx.x * blockDim.x + threadIdx.x;
                fd_xxh_quad(P, gt >> 2, gt & 3, (threadIdx.x & 31) & ~3u);
            });
        if (staged_out) for (uint32_t i = 0; i < n; i++) staged_out[i] = fd[i].state == 0;
    }
    emu::launch(grid, DEC_WARPS * 32, DEC_SMEM_BYTES, [&]() {
        zstd_decode_warp(emu::dyn_smem, P, blockIdx.x * DEC_WARPS + (threadIdx.x >> 5), gridDim.x * DEC_WARPS);
    });
    return 0;
}
int emu_zstd_decode(const uint8_t *src, const uint64_t *src_off, const uint32_t *src_sizes, uint32_t n, uint8_t *dst,
                    const uint64_t *dst_off, const uint32_t *dst_caps, int64_t *out_sizes) {
    return emu_zstd_decode_mode(src, src_off, src_sizes, n, dst, dst_off, dst_caps, out_sizes, 0, nullptr);
}

// S2 (snappy = 0) / Snappy-compatible (snappy = 1) block encode of nchunks chunks (chunk i = src + i*stride).
// better: 0 = s2.Encode's class, 1 = s2.EncodeBetter's; parse must be 0
int emu_s2_encode_lv(const uint8_t *src, uint64_t stride, const uint32_t *sizes, uint32_t nchunks, uint8_t *dst,
                     uint64_t dst_stride, int64_t *out_sizes, int snappy, int better, int parse) {
    if (parse != 0) return -1;
    std::vector<uint8_t> scratch(std::max<size_t>(16, std::max(LzLayout<3>::SCRATCH_BYTES, LzLayout<4>::SCRATCH_BYTES)), 0xCD);
    ZstdEncParams P;
    memset(&P, 0, sizeof(P));
    P.src_base = src; P.src_stride = stride; P.src_sizes = sizes;
    P.dst_base = dst; P.dst_stride = dst_stride; P.dst_cap = (uint32_t)dst_stride;
    P.out_sizes = out_sizes; P.nchunks = nchunks; P.scratch = scratch.data(); P.blockmax = 65536;
    if (better) {
        emu::launch(1, LzCfg<4>::NT, LzLayout<4>::SMEM_BYTES, [&]() {
            for (uint32_t c = 0; c < P.nchunks; c++) {
                if (snappy) lz_parse_chunk<4, LZ_MODE_SNAPPY>(emu::dyn_smem, P, c, P.scratch);
                else lz_parse_chunk<4, LZ_MODE_S2>(emu::dyn_smem, P, c, P.scratch);
            }
        });
    } else {
        emu::launch(1, LzCfg<3>::NT, LzLayout<3>::SMEM_BYTES, [&]() {
            for (uint32_t c = 0; c < P.nchunks; c++) {
                if (snappy) lz_parse_chunk<3, LZ_MODE_SNAPPY>(emu::dyn_smem, P, c, P.scratch);
                else lz_parse_chunk<3, LZ_MODE_S2>(emu::dyn_smem, P, c, P.scratch);
            }
        });
    }
    return 0;
}
int emu_s2_encode(const uint8_t *src, uint64_t stride, const uint32_t *sizes, uint32_t nchunks, uint8_t *dst,
                  uint64_t dst_stride, int64_t *out_sizes, int snappy) {
    return emu_s2_encode_lv(src, stride, sizes, nchunks, dst, dst_stride, out_sizes, snappy, 0, 0);
}

// b2c_s2_encode_stream_device with the device's kernels (one sub-batch): blocks, checksums, scan, placement.
// Returns the stream length, or a negative error.
int64_t emu_s2_encode_stream(const uint8_t *src, uint64_t n, uint32_t block, uint8_t *dst, uint64_t cap, int snappy, int better) {
    const uint32_t nblocks = (uint32_t)((n + block - 1) / block);
    const char *magic = snappy ? "\xff\x06\x00\x00sNaPpY" : "\xff\x06\x00\x00S2sTwO";
    if (cap < 10) return -4;
    if (nblocks == 0) { memcpy(dst, magic, 10); return 10; }
    uint64_t slotB = block + (block / 6) + 64;           // >= s2.MaxEncodedLen(block)
    slotB = (slotB + 15) & ~15ull;
    std::vector<uint8_t> slots((size_t)slotB * nblocks + 64, 0xCD), scratch(std::max<size_t>(16, std::max(LzLayout<3>::SCRATCH_BYTES, LzLayout<4>::SCRATCH_BYTES)), 0xCD);
    std::vector<int64_t> enc(nblocks, 0), piece(nblocks, 0);
    std::vector<uint32_t> crc(nblocks, 0);
    std::vector<uint64_t> scan(nblocks + 1, 0), base(2, 0);
    ZstdEncParams E;
    memset(&E, 0, sizeof(E));
    E.src_base = src; E.src_stride = block; E.src_size_all = block; E.src_total = n;
    E.dst_base = slots.data(); E.dst_stride = slotB; E.dst_cap = (uint32_t)slotB;
    E.out_sizes = enc.data(); E.nchunks = nblocks; E.scratch = scratch.data(); E.blockmax = 65536;
    if (better)
        emu::launch(1, LzCfg<4>::NT, LzLayout<4>::SMEM_BYTES, [&]() {
            for (uint32_t c = 0; c < E.nchunks; c++) {
                if (snappy) lz_parse_chunk<4, LZ_MODE_SNAPPY>(emu::dyn_smem, E, c, E.scratch);
                else lz_parse_chunk<4, LZ_MODE_S2>(emu::dyn_smem, E, c, E.scratch);
            }
        });
    else
        emu::launch(1, LzCfg<3>::NT, LzLayout<3>::SMEM_BYTES, [&]() {
            for (uint32_t c = 0; c < E.nchunks; c++) {
                if (snappy) lz_parse_chunk<3, LZ_MODE_SNAPPY>(emu::dyn_smem, E, c, E.scratch);
                else lz_parse_chunk<3, LZ_MODE_S2>(emu::dyn_smem, E, c, E.scratch);
            }
        });
    int32_t err = 0;
    S2StreamParams P;
    memset(&P, 0, sizeof(P));
    P.src = src; P.total = n; P.block = block; P.slots = slots.data(); P.slot_stride = slotB; P.enc_sizes = enc.data();
    P.crc = crc.data(); P.piece = piece.data(); P.offsets = scan.data(); P.base = base.data(); P.k = 0;
    P.dst = dst; P.cap = cap; P.c0 = 0; P.m = nblocks; P.snappy = snappy ? 1u : 0u; P.err = &err;
    emu::launch((nblocks + 3) / 4, 128, 1024, [&]() {
        uint32_t *tab = reinterpret_cast<uint32_t *>(emu::dyn_smem);
        crc32c_fill_table(tab, threadIdx.x, blockDim.x);
        __syncthreads();
        const uint32_t c = blockIdx.x * 4 + (threadIdx.x >> 5);
        if (c < P.m) s2s_crc_block(P, c, tab, threadIdx.x & 31);
    });
    for (uint32_t i = 0; i < nblocks; i++) scan[i + 1] = scan[i] + (piece[i] > 0 ? (uint64_t)piece[i] : 0);
    emu::launch(nblocks, 256, 0, [&]() { s2s_place_block(P, blockIdx.x, threadIdx.x, blockDim.x); });
    if (err) return err;
    return (int64_t)(10 + scan[nblocks]);
}

int emu_s2_decode(const uint8_t *src, const uint64_t *src_off, const uint32_t *src_sizes, uint32_t n, uint8_t *dst,
                  const uint64_t *dst_off, const uint32_t *dst_caps, int64_t *out_sizes) {
    S2DecParams P;
    memset(&P, 0, sizeof(P));
    P.src_base = src; P.src_offsets = src_off; P.src_sizes = src_sizes;
    P.dst_base = dst; P.dst_offsets = dst_off; P.dst_caps = dst_caps;
    P.out_sizes = out_sizes; P.nchunks = n;
    // the device's launch sequence: tag walk (one lane per block), execution (one warp per block), then the one-warp kernel over
    // what they left.  emu_s2_staged = 0 runs the one-warp kernel alone.  staged_out (optional): 1 where the staged kernels did it.
    std::vector<S2Head> heads;
    std::vector<uint64_t> recs;
    if (emu_s2_staged && n > 0) {
        uint64_t span = 0;
        for (uint32_t i = 0; i < n; i++) span = std::max<uint64_t>(span, src_off[i] + src_sizes[i]);
        heads.resize(n);
        memset(heads.data(), 0xCD, sizeof(S2Head) * (size_t)n);
        recs.assign((size_t)(span / 3) + n + 16, 0xCDCDCDCDCDCDCDCDull);
        P.heads = heads.data(); P.recs = recs.data();
        emu::launch((n + 31) / 32, 32, 0, [&]() {
            const uint32_t c = blockIdx.x * 32 + threadIdx.x;
            if (c < P.nchunks) s2s_walk_lane(P, c);
        });
        emu::launch((n + S2DEC_WARPS - 1) / S2DEC_WARPS, S2DEC_WARPS * 32, S2DEC_WARPS * (S2S_WIN + 16), [&]() {
            const uint32_t c = blockIdx.x * S2DEC_WARPS + (threadIdx.x >> 5);
            if (c < P.nchunks) s2s_exec_warp(P, c, emu::dyn_smem + (threadIdx.x >> 5) * (S2S_WIN + 16), threadIdx.x & 31);
        });
        emu_s2_staged_count = 0;
        for (uint32_t i = 0; i < n; i++) emu_s2_staged_count += heads[i].state == 0;
    }
    emu::launch(2, S2DEC_WARPS * 32, 0, [&]() {
        s2_decode_warp(P, blockIdx.x * S2DEC_WARPS + (threadIdx.x >> 5), gridDim.x * S2DEC_WARPS);
    });
    return 0;
}

// standalone huff0: blocks at src + i*stride -> dst + i*dst_stride
int emu_huf_compress(const uint8_t *src, uint64_t stride, const uint32_t *sizes, uint32_t n, uint8_t *dst,
                     uint64_t dst_stride, int64_t *out_sizes, int four) {
    Huf0Params P;
    memset(&P, 0, sizeof(P));
    P.src_base = src; P.src_stride = stride; P.src_sizes = sizes;
    P.dst_base = dst; P.dst_stride = dst_stride; P.dst_cap = (uint32_t)dst_stride;
    P.out_sizes = out_sizes; P.nchunks = n; P.flags = four ? HUF0_FLAG_4X : 0;
    emu::launch(1, HUF0_NT, HUF0_SMEM_BYTES, [&]() {
        Huf0Shared *sh = reinterpret_cast<Huf0Shared *>(emu::dyn_smem);
        for (uint32_t c = 0; c < P.nchunks; c++) { huf0_compress_block(sh, P, c); __syncthreads(); }
    });
    return 0;
}
int emu_huf_decompress(const uint8_t *src, uint64_t stride, const uint32_t *sizes, uint32_t n, uint8_t *dst,
                       uint64_t dst_stride, const uint32_t *dst_sizes, int64_t *out_sizes, int four) {
    // the device's launch sequence: table pass, the staged decoder's literal-stream kernel, the one-warp kernel over the rest
    Huf0Params P;
    memset(&P, 0, sizeof(P));
    P.src_base = src; P.src_stride = stride; P.src_sizes = sizes; P.dst_base = dst; P.dst_stride = dst_stride;
    P.dst_sizes = dst_sizes; P.out_sizes = out_sizes; P.nchunks = n; P.flags = four ? HUF0_FLAG_4X : 0;
    std::vector<FdChunk> fd(n);
    memset(fd.data(), 0xCD, sizeof(FdChunk) * (size_t)n);
    std::vector<FdBlock> fdb(n);
    memset(fdb.data(), 0xCD, sizeof(FdBlock) * (size_t)n);
    std::vector<uint16_t> hufs((size_t)n * 2048, 0xCDCD);
    const bool staged = (dst_stride & 3) == 0 && n > 0;
    if (staged) {
        P.fd = fd.data(); P.fd_blk = fdb.data(); P.fd_huf = hufs.data();
        emu::launch(1, DEC_WARPS * 32, DEC_SMEM_BYTES, [&]() {
            const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
            DecWarp *dw = reinterpret_cast<DecWarp *>(emu::dyn_smem + w * DEC_WARP_BYTES);
            for (uint32_t c = w; c < n; c += DEC_WARPS) { __syncwarp(); huf0_prep_block(dw, P, c, lane); __syncwarp(); }
        });
        ZstdDecParams Z;
        memset(&Z, 0, sizeof(Z));
        Z.src_base = src; Z.src_stride = stride; Z.src_sizes = sizes; Z.dst_base = dst; Z.dst_stride = dst_stride;
        Z.out_sizes = out_sizes; Z.nchunks = n; Z.fd = fd.data(); Z.fd_blk = fdb.data(); Z.fd_maxb = 1; Z.fd_huf = hufs.data(); Z.fd_lits = dst; Z.fd_lit_stride = dst_stride;
        const unsigned groups = (n + FD_LIT_GROUP - 1) / FD_LIT_GROUP;
        emu::launch((groups + FD_LIT_WARPS - 1) / FD_LIT_WARPS, FD_LIT_WARPS * 32, FD_LIT_WARPS * FD_LIT_WARP_BYTES, [&]() {
            const unsigned w = threadIdx.x >> 5;
            fd_lit_warp(emu::dyn_smem + w * FD_LIT_WARP_BYTES, Z, blockIdx.x * FD_LIT_WARPS + w, threadIdx.x & 31);
        });
    }
    emu::launch(1, DEC_WARPS * 32, DEC_SMEM_BYTES, [&]() {
        const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        DecWarp *dw = reinterpret_cast<DecWarp *>(emu::dyn_smem + w * DEC_WARP_BYTES);
        for (uint32_t c = w; c < n; c += DEC_WARPS) {
            __syncwarp();
            if (P.fd) {
                const uint32_t st = P.fd[c].state;
                if (st == 0) { if (lane == 0) out_sizes[c] = (int64_t)dst_sizes[c]; continue; }
                if (st == HUF0_STATE_ERR) { if (lane == 0) out_sizes[c] = (int64_t)(int32_t)P.fd[c].pad[0]; continue; }
            }
            const int64_t r = (dst_stride && dst_sizes[c] > dst_stride && dst_sizes[c] <= HUF0_BLOCK_MAX) ? (int64_t)HUF0_ERR_DST
                              : huf0_decompress_block(dw, src + (uint64_t)c * stride, sizes[c], dst + (uint64_t)c * dst_stride,
                                                      dst_sizes[c], four != 0, lane);
            __syncwarp();
            if (lane == 0) out_sizes[c] = r;
        }
    });
    return 0;
}

// huff0.ReadTable rows (260 bytes each, layout of include/b2c.h) for n inputs at src + i*stride
int emu_huf_read_table(const uint8_t *src, uint64_t stride, const uint32_t *sizes, uint32_t n, uint8_t *rows, int64_t *out_sizes) {
    emu::launch(1, DEC_WARPS * 32, DEC_SMEM_BYTES, [&]() {
        const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        DecWarp *dw = reinterpret_cast<DecWarp *>(emu::dyn_smem + w * DEC_WARP_BYTES);
        for (uint32_t c = w; c < n; c += DEC_WARPS) {
            __syncwarp();
            huf0_read_table_block(dw, src + (uint64_t)c * stride, sizes[c], rows + (uint64_t)c * 260, out_sizes + c, lane);
            __syncwarp();
        }
    });
    return 0;
}

uint32_t emu_lz_smem_bytes(int level) { return level >= 2 ? LzLayout<2>::SMEM_BYTES : LzLayout<1>::SMEM_BYTES; }
uint32_t emu_pack_smem_bytes() { return PACK_SMEM_BYTES; }
uint64_t emu_chunkwork_bytes() { return sizeof(ChunkWork); }
}
