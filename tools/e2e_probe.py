#!/usr/bin/env python
"""Where does the end-to-end time go?  PCIe copy rates on this box next to b2c_zstd_encode_packed at several batch sizes."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
from compress_b200 import zstd

CH = 65536
n = 8288
dev = torch.device("cuda", 0)
src = H.synth_text_torch(n * CH, dev)
host_in = src.cpu().pin_memory()
host_out = torch.empty(n * CH + n * 32 + 64, dtype=torch.uint8, pin_memory=True)
dbuf = torch.empty(n * CH, dtype=torch.uint8, device=dev)


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


t = timed(lambda: dbuf.copy_(host_in, non_blocking=True))
print("H2D %.1f GB/s (%.2f ms for %d MB)" % (n * CH / t / 1e9, t * 1e3, n * CH >> 20))
half = host_out[: 240 << 20]
dhalf = dbuf[: 240 << 20]
t = timed(lambda: half.copy_(dhalf, non_blocking=True))
print("D2H %.1f GB/s (%.2f ms for 240 MB)" % ((240 << 20) / t / 1e9, t * 1e3))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def both():
    with torch.cuda.stream(s1):
        dbuf.copy_(host_in, non_blocking=True)
    with torch.cuda.stream(s2):
        half.copy_(dhalf, non_blocking=True)
t = timed(both)
print("H2D + D2H concurrently: %.2f ms" % (t * 1e3))
for mc in (296, 592, 1184, 2368, 4144):
    enc = zstd.Encoder(max_chunks=mc)
    t = timed(lambda: enc.encode_packed(host_in, dst=host_out), reps=4)
    print("encode_packed max_chunks=%d: %.2f ms  %.1f GB/s" % (mc, t * 1e3, n * CH / t / 1e9))
    enc.close()
enc = zstd.Encoder(max_chunks=64)
dst = torch.empty((n, zstd.SLOT), dtype=torch.uint8, device=dev)
outs = torch.empty((n,), dtype=torch.int64, device=dev)
t = timed(lambda: enc.encode_device(src, dst=dst, out_sizes=outs))
print("device-resident encode of the same %d chunks: %.2f ms" % (n, t * 1e3))
