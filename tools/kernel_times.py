#!/usr/bin/env python
"""Per-kernel device times of one encode pass over N chunks (profile events of the library), plus huff0
compress and S2 encode rates.  Used to compare tuning builds: B2C_LIB=<variant.so> python tools/kernel_times.py"""
import hashlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
from compress_b200 import zstd, huff0

CH = 65536
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
steps = 5
dev = torch.device("cuda", 0)
src = H.synth_text_torch(n * CH, dev)
enc = zstd.Encoder()
dst = torch.empty((n, zstd.SLOT), dtype=torch.uint8, device=dev)
outs = torch.empty(n, dtype=torch.int64, device=dev)
for _ in range(3):
    enc.encode_device(src, dst=dst, out_sizes=outs)
torch.cuda.synchronize()
enc.profile(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    enc.encode_device(src, dst=dst, out_sizes=outs)
e1.record()
torch.cuda.synchronize()
ms, calls = enc.profile_read()
tot = e0.elapsed_time(e1) / steps
o = outs.cpu()
digest = hashlib.sha1(o.numpy().tobytes()).hexdigest()[:12]
first = bytes(dst[0, : int(o[0])].cpu().numpy())
print("lib", os.environ.get("B2C_LIB", "default"))
print("  step %.3f ms = %.1f GB/s; ratio %.4f; sizes sha1 %s; frame0 sha1 %s" % (
    tot, n * CH / tot / 1e6, float(o.sum()) / (n * CH), digest, hashlib.sha1(first).hexdigest()[:12]))
print("  " + "  ".join("%s %.3f" % (k.replace("b2c_zstd_", "").replace("_kernel", ""), v / calls) for k, v in ms.items()))

# huff0 4X compress on 262143-byte blocks of the same text (BASELINE config 4)
hc = huff0.Codec()
hb, hstride = 262143, 262144
hn = (n * CH) // hstride
hsz = torch.full((hn,), hb, dtype=torch.int32, device=dev)
hdst = torch.empty((hn, hstride), dtype=torch.uint8, device=dev)
hout = torch.empty((hn,), dtype=torch.int64, device=dev)
for _ in range(2):
    hc.compress_device(src, hstride, hsz, True, dst=hdst, out_sizes=hout)
torch.cuda.synchronize()
e0.record()
for _ in range(3):
    hc.compress_device(src, hstride, hsz, True, dst=hdst, out_sizes=hout)
e1.record()
torch.cuda.synchronize()
hms = e0.elapsed_time(e1) / 3
print("  huff0 compress4x %.3f ms = %.1f GB/s; sizes sha1 %s" % (
    hms, hn * hb / hms / 1e6, hashlib.sha1(hout.cpu().numpy().tobytes()).hexdigest()[:12]))
