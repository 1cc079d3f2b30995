#!/usr/bin/env python
"""Per-kernel device times (ms per GiB of input) of the zstd encode pipeline at level 1 or 2, from the library's own
CUDA-event profile.  usage: enc_times.py [level] [GiB]   (B2C_PARSE=r1 selects the round-1 parse kernel at level 1)"""
import hashlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
from compress_b200 import zstd

level = int(sys.argv[1]) if len(sys.argv) > 1 else 1
gib = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
steps = 5
dev = torch.device("cuda", 0)
enc = zstd.Encoder(level=level)
CH = enc.block
n = int(gib * (1 << 30)) // CH
src = H.synth_text_torch(n * CH, dev)
dst = torch.empty((n, enc.slot), dtype=torch.uint8, device=dev)
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
assert int(o.min()) > 0
print("level %d  parse %s  %d x %d KiB" % (level, os.environ.get("B2C_PARSE", "lz"), n, CH >> 10))
print("  step %.3f ms = %.1f GB/s; ratio %.4f; sizes sha1 %s" % (
    tot, n * CH / tot / 1e6, float(o.sum()) / (n * CH), hashlib.sha1(o.numpy().tobytes()).hexdigest()[:12]))
print("  per step: " + "  ".join("%s %.3f" % (k.replace("b2c_zstd_", "").replace("b2c_", "").replace("_kernel", ""), v / steps)
                                 for k, v in ms.items()) + "   (%d launches of the pipeline per step)" % (calls // steps))
