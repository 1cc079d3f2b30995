"""Driver for the SIMT-emulated kernels (tests/emu).  Test infrastructure only."""
import numpy as np


def emu_encode(E, chunks, flags=3, desc=0, seq_cap=20000):
    E.emu_set_lane_order(desc)
    n = len(chunks)
    stride = 65536
    src = np.zeros(n * stride + 64, dtype=np.uint8)
    sizes = np.zeros(n, dtype=np.uint32)
    for i, c in enumerate(chunks):
        src[i * stride:i * stride + len(c)] = np.frombuffer(c, dtype=np.uint8)
        sizes[i] = len(c)
    dstride = 65536 + 512
    dst = np.zeros(n * dstride, dtype=np.uint8)
    outs = np.zeros(n, dtype=np.int64)
    hdr = np.zeros((n, 4), dtype=np.uint32)
    seqs = np.zeros((n, seq_cap, 3), dtype=np.uint32)
    lits = np.zeros((n, 65536), dtype=np.uint8)
    E.emu_zstd_encode(src.ctypes.data, stride, sizes.ctypes.data, n, dst.ctypes.data, dstride, outs.ctypes.data, flags,
                      hdr.ctypes.data, seqs.ctypes.data, lits.ctypes.data, seq_cap)
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
