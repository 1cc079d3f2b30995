#!/usr/bin/env python
"""S2 block decode on 1 GiB of synthetic text: time, and how many blocks the staged kernels finished."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
from compress_b200 import s2 as s2mod
dev = torch.device("cuda", 0)
n, CH = 16384, 65536
src = H.synth_text_torch(n * CH, dev)
c = s2mod.Codec()
for better in (False, True):
    sd, ss = c.encode_device(src, better=better)
    si = ss.to(torch.int32)
    dout = torch.empty((n, CH), dtype=torch.uint8, device=dev)
    dres = torch.empty((n,), dtype=torch.int64, device=dev)
    for _ in range(2):
        c.decode_device(sd, si, src_stride=s2mod.SLOT, dst=dout, dst_cap=CH, out_sizes=dres)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        c.decode_device(sd, si, src_stride=s2mod.SLOT, dst=dout, dst_cap=CH, out_sizes=dres)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    assert bool((dres == CH).all()) and torch.equal(dout.view(-1), src)
    print("s2 %s decode %.3f ms = %.1f GB/s; staged blocks %d of %d" % ("better" if better else "fast", ms, n * CH / ms / 1e6, c.staged_count(n), n))
