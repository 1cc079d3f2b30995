"""Driver for the SIMT-emulated kernels (tests/emu).  Test infrastructure only."""
import numpy as np


def emu_encode(E, chunks, flags=3, desc=0, seq_cap=None, level=1, parse=0):
    """level 1 / 2; parse 0 = the tile-ordered parse (product path), 1 = the round-1 parse (level 1 only)."""
    E.emu_set_lane_order(desc)
    n = len(chunks)
    stride = 65536 if level == 1 else 131072
    if seq_cap is None:
        seq_cap = stride // 4 + 64
    src = np.zeros(n * stride + 64, dtype=np.uint8)
    sizes = np.zeros(n, dtype=np.uint32)
    for i, c in enumerate(chunks):
        m = min(len(c), stride)
        src[i * stride:i * stride + m] = np.frombuffer(c, dtype=np.uint8)[:m]
        sizes[i] = len(c)
    dstride = stride + 512
    dst = np.zeros(n * dstride, dtype=np.uint8)
    outs = np.zeros(n, dtype=np.int64)
    hdr = np.zeros((n, 4), dtype=np.uint32)
    seqs = np.zeros((n, seq_cap, 3), dtype=np.uint32)
    lits = np.zeros((n, stride), dtype=np.uint8)
    rc = E.emu_zstd_encode_lv(src.ctypes.data, stride, sizes.ctypes.data, n, dst.ctypes.data, dstride, outs.ctypes.data,
                              flags, hdr.ctypes.data, seqs.ctypes.data, lits.ctypes.data, seq_cap, level, parse)
    assert rc == 0
    frames = [bytes(dst[i * dstride:i * dstride + max(int(outs[i]), 0)]) for i in range(n)]
    return frames, outs, hdr, seqs, lits


def frame_header_len(n):
    """frameHeader.appendTo for a single chunk of n bytes, no dict (zstd/frameenc.go:25-92)."""
    single = n > 1024
    fh = 4 + 1 + (0 if single else 1)
    if n >= 256:
        fh += 2 if n < 65536 + 256 else 4
    elif single:
        fh += 1
    return fh


def emu_decode(E, inputs, caps, desc=0, mode=0, staged=None):
    """Run the emulated decode kernels on a list of compressed byte strings; returns (sizes, outputs).
    mode 0: the device's launch sequence (staged kernels, then the one-warp decoder over what they marked); mode 1: the
    one-warp decoder alone.  staged: optional list, filled with 1 / 0 per input (result produced by the staged kernels)."""
    E.emu_set_lane_order(desc)
    n = len(inputs)
    src_off = np.zeros(n, dtype=np.uint64)
    dst_off = np.zeros(n, dtype=np.uint64)
    sizes = np.array([len(b) for b in inputs], dtype=np.uint32)
    capv = np.array(caps, dtype=np.uint32)
    so = do = 0
    for i in range(n):
        src_off[i] = so
        dst_off[i] = do
        so += (len(inputs[i]) + 15) & ~15
        do += (int(capv[i]) + 15) & ~15
    src = np.full(so + 64, 0xA5, dtype=np.uint8)
    for i, b in enumerate(inputs):
        src[int(src_off[i]):int(src_off[i]) + len(b)] = np.frombuffer(b, dtype=np.uint8)
    dst = np.zeros(do + 64, dtype=np.uint8)
    outs = np.zeros(n, dtype=np.int64)
    flags = np.zeros(max(n, 1), dtype=np.uint8)
    E.emu_zstd_decode_mode(src.ctypes.data, src_off.ctypes.data, sizes.ctypes.data, n, dst.ctypes.data, dst_off.ctypes.data,
                           capv.ctypes.data, outs.ctypes.data, mode, flags.ctypes.data)
    if staged is not None:
        staged[:] = [int(x) for x in flags[:n]]
    res = [bytes(dst[int(dst_off[i]):int(dst_off[i]) + max(int(outs[i]), 0)]) for i in range(n)]
    return outs, res


def emu_s2_encode(E, blocks, snappy=False, desc=0, better=False, parse=0):
    E.emu_set_lane_order(desc)
    n = len(blocks)
    stride = 65536
    src = np.zeros(n * stride + 64, dtype=np.uint8)
    sizes = np.zeros(n, dtype=np.uint32)
    for i, c in enumerate(blocks):
        src[i * stride:i * stride + len(c)] = np.frombuffer(c, dtype=np.uint8)
        sizes[i] = len(c)
    dstride = 65536 + 512
    dst = np.zeros(n * dstride, dtype=np.uint8)
    outs = np.zeros(n, dtype=np.int64)
    rc = E.emu_s2_encode_lv(src.ctypes.data, stride, sizes.ctypes.data, n, dst.ctypes.data, dstride, outs.ctypes.data,
                            1 if snappy else 0, 1 if better else 0, parse)
    assert rc == 0
    return [bytes(dst[i * dstride:i * dstride + max(int(outs[i]), 0)]) for i in range(n)], outs


def emu_s2_decode(E, inputs, caps, desc=0):
    E.emu_set_lane_order(desc)
    n = len(inputs)
    src_off = np.zeros(n, dtype=np.uint64)
    dst_off = np.zeros(n, dtype=np.uint64)
    sizes = np.array([len(b) for b in inputs], dtype=np.uint32)
    capv = np.array(caps, dtype=np.uint32)
    so = do = 0
    for i in range(n):
        src_off[i] = so
        dst_off[i] = do
        so += (len(inputs[i]) + 15) & ~15
        do += (int(capv[i]) + 15) & ~15
    src = np.full(so + 64, 0xA5, dtype=np.uint8)
    for i, b in enumerate(inputs):
        src[int(src_off[i]):int(src_off[i]) + len(b)] = np.frombuffer(b, dtype=np.uint8)
    dst = np.full(do + 64, 0x5A, dtype=np.uint8)
    outs = np.zeros(n, dtype=np.int64)
    E.emu_s2_decode(src.ctypes.data, src_off.ctypes.data, sizes.ctypes.data, n, dst.ctypes.data, dst_off.ctypes.data,
                    capv.ctypes.data, outs.ctypes.data)
    res = [bytes(dst[int(dst_off[i]):int(dst_off[i]) + max(int(outs[i]), 0)]) for i in range(n)]
    return outs, res, dst, dst_off


def emu_huf_compress(E, blocks, four=True, desc=0):
    E.emu_set_lane_order(desc)
    n = len(blocks)
    stride = max(16, (max(len(b) for b in blocks) + 15) // 16 * 16)
    src = np.zeros((n, stride), dtype=np.uint8)
    for i, b in enumerate(blocks):
        src[i, :len(b)] = np.frombuffer(b, dtype=np.uint8)
    sizes = np.array([len(b) for b in blocks], dtype=np.uint32)
    dst = np.full((n, stride), 0xEE, dtype=np.uint8)
    outs = np.zeros(n, dtype=np.int64)
    E.emu_huf_compress(src.ctypes.data, stride, sizes.ctypes.data, n, dst.ctypes.data, stride, outs.ctypes.data, 1 if four else 0)
    return [(bytes(dst[i, :outs[i]]) if outs[i] >= 0 else None, int(outs[i])) for i in range(n)]


def emu_huf_decompress(E, blocks, dst_sizes, four=True, desc=0):
    E.emu_set_lane_order(desc)
    n = len(blocks)
    stride = max(16, (max(len(b) for b in blocks) + 15) // 16 * 16)
    src = np.zeros((n, stride), dtype=np.uint8)
    for i, b in enumerate(blocks):
        src[i, :len(b)] = np.frombuffer(b, dtype=np.uint8)
    sizes = np.array([len(b) for b in blocks], dtype=np.uint32)
    ds = np.array(dst_sizes, dtype=np.uint32)
    dstride = max(16, (int(ds.max()) + 15) // 16 * 16)
    dst = np.zeros((n, dstride), dtype=np.uint8)
    outs = np.zeros(n, dtype=np.int64)
    E.emu_huf_decompress(src.ctypes.data, stride, sizes.ctypes.data, n, dst.ctypes.data, dstride, ds.ctypes.data,
                         outs.ctypes.data, 1 if four else 0)
    return [(bytes(dst[i, :outs[i]]) if outs[i] >= 0 else None, int(outs[i])) for i in range(n)]


def emu_encode_frames(E, inputs, level=1, crc=True, desc=0, dump=True, misalign=0):
    """Frame mode under the emulator (b2c_zstd_encode_frames_device's launch sequence): one frame per input of any size.
    Returns (frames, blocks): blocks = list of dicts per planned block (frame-relative order) with off / len / hist / last and,
    when dump, the parse: nseq, nlit, kind, tri (n x 3), lits."""
    import ctypes as c
    E.emu_set_lane_order(desc)
    E.emu_zstd_encode_frames.argtypes = [c.c_void_p] * 3 + [c.c_uint32, c.c_void_p, c.c_uint64, c.c_void_p, c.c_void_p, c.c_int,
                                                            c.c_int, c.c_void_p, c.c_void_p, c.c_void_p, c.c_uint32, c.c_uint32,
                                                            c.c_void_p, c.c_void_p]
    n = len(inputs)
    offs = np.zeros(n, dtype=np.uint64)
    sizes = np.array([len(x) for x in inputs], dtype=np.uint64)
    tot = misalign          # (frames normally start 16-byte aligned; misalign shifts all of them)
    for i, x in enumerate(inputs):
        offs[i] = tot
        tot += (len(x) + 15) & ~15
    src = np.zeros(tot + 256, dtype=np.uint8)
    for i, x in enumerate(inputs):
        src[int(offs[i]):int(offs[i]) + len(x)] = np.frombuffer(x, dtype=np.uint8)
    fblock = 49152 if level == 1 else 98304
    vblock = 65536 if level == 1 else 131072
    nblk_max = sum(max(1, (len(x) + fblock - 1) // fblock) for x in inputs)
    cap = sum(len(x) + 3 * max(1, (len(x) + fblock - 1) // fblock) + 32 for x in inputs) + 64
    dst = np.zeros(cap, dtype=np.uint8)
    foff = np.zeros(n, dtype=np.uint64)
    fsz = np.zeros(n, dtype=np.int64)
    seq_cap = vblock // 4 + 64
    hdr = np.zeros((nblk_max, 4), dtype=np.uint32)
    seqs = np.zeros((nblk_max, seq_cap, 3), dtype=np.uint32) if dump else None
    lits = np.zeros((nblk_max, vblock), dtype=np.uint8) if dump else None
    nb = c.c_uint32(0)
    bd = np.zeros((nblk_max, 4), dtype=np.uint64)
    rc = E.emu_zstd_encode_frames(src.ctypes.data, offs.ctypes.data, sizes.ctypes.data, n, dst.ctypes.data, cap, foff.ctypes.data,
                                  fsz.ctypes.data, 1 if crc else 0, level, hdr.ctypes.data if dump else None,
                                  seqs.ctypes.data if dump else None, lits.ctypes.data if dump else None, seq_cap, nblk_max,
                                  c.byref(nb), bd.ctypes.data)
    assert rc == 0, rc
    assert nb.value == nblk_max
    frames = []
    for i in range(n):
        assert fsz[i] > 0, (i, fsz[i])
        frames.append(bytes(dst[int(foff[i]):int(foff[i]) + int(fsz[i])]))
    blocks = []
    for b in range(nblk_max):
        d = {"off": int(bd[b, 0]), "len": int(bd[b, 1]), "hist": int(bd[b, 2]), "last": int(bd[b, 3]) & 1}
        if dump:
            nseq, nlit, kind, _ = [int(v) for v in hdr[b]]
            d.update(nseq=nseq, nlit=nlit, kind=kind, tri=seqs[b][:nseq].copy(), lits=bytes(lits[b][:nlit]) if kind == 0 else None)
        blocks.append(d)
    return frames, blocks, offs
