#!/usr/bin/env python
"""Decode throughput of the zstd decode kernel on the bench workload (tuning aid; honours B2C_LIB)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import helpers as H
from compress_b200 import zstd
n = 16384
src = H.synth_text_torch(n * 65536, "cuda", seed=1000)
enc = zstd.Encoder(max_chunks=64)
frames, sizes = enc.encode_device(src)
torch.cuda.synchronize()
dec = zstd.Decoder()
sz = sizes.to(torch.int32)
out = torch.empty((n, 65536), dtype=torch.uint8, device="cuda")
res = torch.empty((n,), dtype=torch.int64, device="cuda")
for _ in range(2):
    dec.decode_device(frames, sz, src_stride=zstd.SLOT, dst=out, dst_cap=65536, out_sizes=res)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    dec.decode_device(frames, sz, src_stride=zstd.SLOT, dst=out, dst_cap=65536, out_sizes=res)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 3
assert bool((res == 65536).all()) and torch.equal(out.view(-1), src)
print("decode %.2f ms  %.1f GB/s" % (ms, n * 65536 / ms / 1e6))
