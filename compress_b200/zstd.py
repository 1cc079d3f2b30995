"""Host-side mirror of the reference's zstd encoder interface for the accelerated path.

Names follow klauspost/compress/zstd: ``Encoder.EncodeAll`` (zstd/encoder.go:722),
``Encoder.MaxEncodedSize`` (:843), levels ``SpeedFastest``/``SpeedDefault``
(zstd/encoder_options.go), ``WithEncoderCRC``.  The work is done by libb200comp.so
(hand-written sm_100a kernels) through the C ABI in include/b2c.h; PyTorch is only the owner of
device buffers and streams.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import lib, check, B2CError

SpeedFastest = 1
SpeedDefault = 2
CHUNK = 1 << 16           # SpeedFastest block size, zstd/encoder_options.go:248-252
SLOT = CHUNK + 512        # per-chunk output slot (>= MaxEncodedSize(CHUNK))
FLAG_CRC = 1
FLAG_FRAME = 2


class Encoder:
    """zstd.Encoder for independent chunks on one B200.

    EncodeAll(src) splits src into 64 KiB chunks, each encoded as one complete zstd frame (the
    reference's EncodeAll emits a single multi-block frame with cross-block history; concatenated
    frames decode to the same bytes, zstd/encoder.go:719).
    """

    def __init__(self, level=SpeedFastest, crc=True, device=0, max_chunks=4096):
        if not torch.cuda.is_available() or lib.b2c_device_count() == 0:
            raise B2CError("no CUDA device: compress_b200 has no CPU fallback")
        if level != SpeedFastest:
            raise B2CError("only SpeedFastest is implemented on the GPU path so far")
        self.level = level
        self.flags = (FLAG_CRC if crc else 0) | FLAG_FRAME
        self.device = device
        self.max_chunks = max_chunks
        self._ctx = lib.b2c_ctx_create(device, max_chunks)
        if not self._ctx:
            raise B2CError("b2c_ctx_create failed")

    def close(self):
        if self._ctx:
            lib.b2c_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self):
        return int(lib.b2c_launch_count(self._ctx))

    @property
    def sm_count(self):
        return int(lib.b2c_sm_count(self._ctx))

    def MaxEncodedSize(self, size):
        return int(lib.b2c_zstd_bound(size, self.level))

    # ---- device-resident batch -------------------------------------------------------------
    def encode_device(self, src, sizes=None, chunk=CHUNK, dst=None, out_sizes=None, flags=None):
        """src: uint8 CUDA tensor holding nchunks chunks at stride `chunk` bytes.
        sizes: optional uint32 CUDA tensor (per-chunk sizes); default all `chunk` bytes.
        Returns (dst [nchunks, SLOT] uint8, out_sizes [nchunks] int64), both on the device. Async."""
        assert src.is_cuda and src.dtype == torch.uint8
        nchunks = src.numel() // chunk if sizes is None else sizes.numel()
        if dst is None:
            dst = torch.empty((nchunks, SLOT), dtype=torch.uint8, device=src.device)
        if out_sizes is None:
            out_sizes = torch.empty((nchunks,), dtype=torch.int64, device=src.device)
        stream = torch.cuda.current_stream(src.device).cuda_stream
        rc = lib.b2c_zstd_encode_device(
            self._ctx, self.level, self.flags if flags is None else flags, src.data_ptr(), chunk,
            None if sizes is None else sizes.data_ptr(), chunk, dst.data_ptr(), SLOT, out_sizes.data_ptr(),
            nchunks, ctypes.c_void_p(stream))
        check(rc, self._ctx)
        return dst, out_sizes

    def encode_device_debug(self, src, sizes=None, chunk=CHUNK, flags=None, seq_cap=20000):
        """Like encode_device but also returns the parse (sequence triples + literals) per chunk."""
        nchunks = src.numel() // chunk if sizes is None else sizes.numel()
        dev = src.device
        dst = torch.empty((nchunks, SLOT), dtype=torch.uint8, device=dev)
        out_sizes = torch.empty((nchunks,), dtype=torch.int64, device=dev)
        hdr = torch.zeros((nchunks, 4), dtype=torch.int32, device=dev)
        seqs = torch.zeros((nchunks, seq_cap, 3), dtype=torch.int32, device=dev)
        lits = torch.zeros((nchunks, 65536), dtype=torch.uint8, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        rc = lib.b2c_zstd_encode_device_debug(
            self._ctx, self.flags if flags is None else flags, src.data_ptr(), chunk,
            None if sizes is None else sizes.data_ptr(), chunk, dst.data_ptr(), SLOT, out_sizes.data_ptr(), nchunks,
            hdr.data_ptr(), seqs.data_ptr(), lits.data_ptr(), seq_cap, ctypes.c_void_p(stream))
        check(rc, self._ctx)
        return dst, out_sizes, hdr, seqs, lits

    # ---- host buffers (what the cgo shim calls) -----------------------------------------------
    def encode_chunks(self, chunks):
        """chunks: list of bytes-like (each <= 64 KiB).  Returns list of encoded frames (bytes)."""
        n = len(chunks)
        if n == 0:
            return []
        bufs = [np.frombuffer(c, dtype=np.uint8) if len(c) else np.zeros(0, dtype=np.uint8) for c in chunks]
        outs = [np.empty(self.MaxEncodedSize(len(c)) + 16, dtype=np.uint8) for c in chunks]
        srcs = (ctypes.c_void_p * n)(*[b.ctypes.data for b in bufs])
        ssz = (ctypes.c_size_t * n)(*[len(c) for c in chunks])
        dsts = (ctypes.c_void_p * n)(*[o.ctypes.data for o in outs])
        dcap = (ctypes.c_size_t * n)(*[o.size for o in outs])
        res = (ctypes.c_int64 * n)()
        rc = lib.b2c_zstd_encode_chunks(self._ctx, self.level, self.flags, srcs, ssz, dsts, dcap, res, n)
        check(rc, self._ctx)
        out = []
        for i in range(n):
            if res[i] < 0:
                raise B2CError(f"chunk {i}: {lib.b2c_strerror(int(res[i])).decode()}")
            out.append(outs[i][: res[i]].tobytes())
        return out

    def encode_packed(self, src, dst=None, chunk=CHUNK):
        """src: contiguous host buffer (bytes / numpy / CPU torch tensor, ideally pinned).  Returns
        (dst uint8 tensor (pinned), total, sizes int64 ndarray, offsets uint64 ndarray): dst[:total] is the
        concatenation of one frame per `chunk` bytes of src."""
        if isinstance(src, torch.Tensor):
            assert not src.is_cuda and src.dtype == torch.uint8 and src.is_contiguous()
            sptr, nbytes = src.data_ptr(), src.numel()
        else:
            arr = np.frombuffer(src, dtype=np.uint8)
            sptr, nbytes = arr.ctypes.data, arr.size
        nchunks = max(1, (nbytes + chunk - 1) // chunk)
        cap = nbytes + nchunks * 32 + 64
        if dst is None:
            dst = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
        sizes = np.empty(nchunks, dtype=np.int64)
        offs = np.empty(nchunks, dtype=np.uint64)
        total = ctypes.c_size_t(0)
        rc = lib.b2c_zstd_encode_packed(self._ctx, self.level, self.flags, sptr, nbytes, chunk, dst.data_ptr(),
                                        dst.numel(), sizes.ctypes.data, offs.ctypes.data, ctypes.byref(total))
        check(rc, self._ctx)
        return dst, int(total.value), sizes, offs

    def EncodeAll(self, src, dst=None):
        """EncodeAll will encode all input in src and append it to dst (zstd/encoder.go:715-729)."""
        src = bytes(src)
        buf, total, _, _ = self.encode_packed(src)
        out = bytes(buf[:total].numpy())
        if dst is not None:
            dst += out
            return dst
        return out
