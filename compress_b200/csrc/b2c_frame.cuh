// compress_b200/csrc/b2c_frame.cuh -- frame mode of the zstd encoder: planning (host), frame header, placement of the
// encoded blocks and completion of the frames (device).  Shared by the C ABI (b2c_api.cu) and the CPU emulator.
//
// zstd.Encoder.EncodeAll of an input larger than one block (zstd/encoder.go:796-830): ONE frame per input -- header with
// the content size, the blocks, the XXH64 of the whole content.  Every block is encoded by the same six-kernel pipeline
// as an independent chunk, with two differences: the match finder sees the `hist` bytes before the block (previous
// blocks of the same frame, already in memory), and the blocks are bare (no frame header / checksum of their own).
// Blocks are entropy-coded independently (no repeat-mode tables between blocks, which would serialise a frame's
// blocks; the reference's seqCoders.setPrev / compModeRepeat, zstd/seqenc.go:19-42, is a size optimisation only).
#pragma once
#include <vector>
#include "b2c_zstd_enc.cuh"

#ifdef B2C_EMU
#define B2C_HD static inline
#else
#define B2C_HD static __host__ __device__ inline
#endif

namespace b2c {

struct FrameGeom { uint32_t block, hist; };
// block + history must fit the parse kernel's staged chunk: 64 KiB at level 1, 128 KiB at levels 2 and 3.  Measured under the
// emulator (tools/emu_frame_ratio.py): halving the history from 32 to 16 KiB (64 to 32 KiB) costs 0.7 % (0.2 %) of output
// and saves a third of the redundant table insertions and a third of the blocks.
#ifndef FRAME_BLOCK1
#define FRAME_BLOCK1 49152u   // level 1: 48 KiB blocks that see the 16 KiB before them
#endif
#ifndef FRAME_BLOCK2
#define FRAME_BLOCK2 98304u   // levels 2, 3: 96 KiB blocks, 32 KiB of history
#endif
static inline FrameGeom frame_geom(int level) {
    return level == 1 ? FrameGeom{FRAME_BLOCK1, 65536u - FRAME_BLOCK1} : FrameGeom{FRAME_BLOCK2, 131072u - FRAME_BLOCK2};
}
static inline uint32_t frame_window(int level) { return level == 1 ? (4u << 20) : (8u << 20); }   // encoder_options.go:246-263

// frameHeader.appendTo (zstd/frameenc.go:25-92) for a frame of `size` bytes, no dictionary; SingleSegment and
// WindowSize as encodeAll sets them (zstd/encoder.go:755-766, fastBase.WindowSize zstd/enc_base.go:42-50)
B2C_HD uint32_t frame_header_put(uint8_t *o8, uint64_t size, uint32_t window, bool crc) {
    uint32_t o = 0;
    o8[o++] = 0x28; o8[o++] = 0xB5; o8[o++] = 0x2F; o8[o++] = 0xFD;
    if (size == 0) { o8[o++] = 32; o8[o++] = 0; return o; }       // WithZeroFrames: single segment, no checksum (encoder.go:732-751)
    const bool single = size <= window && size > 1024;
    uint32_t fcs = 0;
    if (size >= 256) fcs++;
    if (size >= 65536 + 256) fcs++;
    if (size >= 0xffffffffull) fcs++;
    o8[o++] = (uint8_t)((crc ? 4u : 0u) | (single ? 32u : 0u) | (fcs << 6));
    if (!single) {
        uint32_t ws = window;
        if (size < window) {
            uint32_t bl = 0;
            for (uint64_t v = size; v; v >>= 1) bl++;
            ws = 1u << bl;
            if (ws < 1024) ws = 1024;
        }
        uint32_t wl = 0;
        for (uint32_t v = ws - 1; v; v >>= 1) wl++;
        o8[o++] = (uint8_t)((wl - 10) << 3);
    }
    if (fcs == 0) { if (single) o8[o++] = (uint8_t)size; }
    else if (fcs == 1) { const uint64_t v = size - 256; o8[o++] = (uint8_t)v; o8[o++] = (uint8_t)(v >> 8); }
    else if (fcs == 2) { for (int k = 0; k < 4; k++) o8[o++] = (uint8_t)(size >> (8 * k)); }
    else { for (int k = 0; k < 8; k++) o8[o++] = (uint8_t)(size >> (8 * k)); }
    return o;
}

struct FrameDesc {
    uint64_t off, size;          // input bytes of the frame
    uint64_t extra;              // header + checksum bytes of all earlier frames of the call
    uint32_t first, nblk;        // its blocks in the call's block list
    uint32_t hdr, crc;           // header bytes, checksum bytes (0 or 4)
    uint32_t window, pad;
};

// Host: the block list and the frame table of a call.  Frame f = sizes[f] bytes at offset offs[f] of the source buffer.
static inline bool frame_plan(int level, bool crc, const uint64_t *offs, const uint64_t *sizes, uint32_t nframes,
                              std::vector<EncBlockDesc> &blocks, std::vector<FrameDesc> &frames) {
    const FrameGeom g = frame_geom(level);
    blocks.clear();
    frames.resize(nframes);
    uint64_t extra = 0;
    for (uint32_t f = 0; f < nframes; f++) {
        FrameDesc &F = frames[f];
        F.off = offs[f]; F.size = sizes[f]; F.extra = extra; F.first = (uint32_t)blocks.size();
        F.window = frame_window(level); F.pad = 0;
        F.crc = (crc && F.size) ? 4u : 0u;
        uint8_t tmp[16];
        F.hdr = frame_header_put(tmp, F.size, F.window, F.crc != 0);
        uint64_t done = 0;
        do {
            EncBlockDesc b;
            const uint64_t left = F.size - done;
            b.off = F.off + done; b.len = (uint32_t)(left < g.block ? left : g.block);
            b.hist = (uint32_t)(done < g.hist ? done : g.hist);
            b.frame = f; b.flags = (done + b.len == F.size) ? 1u : 0u;
            blocks.push_back(b);
            done += b.len;
        } while (done < F.size);
        F.nblk = (uint32_t)blocks.size() - F.first;
        extra += F.hdr + F.crc;
        if (blocks.size() > 0x7fffffffull) return false;
    }
    return true;
}

// Device: block c of a sub-batch (global block index c0 + c) moves from its slot to its place in the packed output
// (all threads of a CTA call; base = bytes of the earlier sub-batches, offsets = exclusive scan of this sub-batch's sizes)
B2C_DEV void frame_place_block(const uint8_t *slots, uint64_t slot_stride, const int64_t *sizes, const uint64_t *offsets,
                               uint64_t base, const EncBlockDesc *desc, const FrameDesc *fr, uint8_t *packed, uint64_t cap,
                               uint64_t *pos, uint32_t c0, uint32_t c, unsigned tid, unsigned nthreads) {
    const int64_t sz = sizes[c];
    const FrameDesc &F = fr[desc[c0 + c].frame];
    const uint64_t at = base + offsets[c] + F.extra + F.hdr;
    const bool ok = sz > 0 && at + (uint64_t)sz + 4 <= cap;
    if (tid == 0) pos[c0 + c] = ok ? at : ~0ull;        // ~0: block failed or does not fit
    if (!ok) return;
    coop_copy(packed + at, slots + (uint64_t)c * slot_stride, (uint32_t)sz, tid, nthreads);
}
// Device, one thread per frame: header in front of its first block, checksum behind its last, offset and size for the
// caller.  sizes_all: encoded size of every block of the call (<= 0 = error)
B2C_DEV void frame_finish_one(const FrameDesc *fr, const uint64_t *pos, const int64_t *sizes_all, const uint64_t *xxh,
                              uint8_t *packed, uint64_t cap, uint64_t *out_offsets, int64_t *out_sizes, uint32_t f) {
    const FrameDesc F = fr[f];
    int64_t err = 0;
    uint64_t end = 0;
    for (uint32_t b = 0; b < F.nblk; b++) {
        const int64_t sz = sizes_all[F.first + b];
        if (sz <= 0) { err = sz < 0 ? sz : -101; break; }
        if (pos[F.first + b] == ~0ull) { err = -4; break; }      // B2C_ERR_DST_SMALL
        end = pos[F.first + b] + (uint64_t)sz;
    }
    if (F.nblk == 0) err = -102;
    const uint64_t start = err ? 0 : pos[F.first] - F.hdr;
    if (!err && end + F.crc > cap) err = -4;
    if (err) { out_offsets[f] = 0; out_sizes[f] = err; return; }
    uint8_t h[16];
    const uint32_t hl = frame_header_put(h, F.size, F.window, F.crc != 0);
    for (uint32_t i = 0; i < hl; i++) packed[start + i] = h[i];
    if (F.crc) { const uint32_t c32 = (uint32_t)xxh[f]; for (int i = 0; i < 4; i++) packed[end + i] = (uint8_t)(c32 >> (8 * i)); }
    out_offsets[f] = start;
    out_sizes[f] = (int64_t)(end + F.crc - start);
}

}  // namespace b2c
