#!/usr/bin/env python
"""Standalone huff0 Compress4X / Decompress4X throughput on 262143-byte blocks of synthetic text (BASELINE config 4)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
from compress_b200 import huff0
dev = torch.device("cuda", 0)
hb, hstride, hn = 262143, 262144, 4096
src = H.synth_text_torch(hn * hstride, dev)
hc = huff0.Codec()
hsz = torch.full((hn,), hb, dtype=torch.int32, device=dev)
hdst = torch.empty((hn, hstride), dtype=torch.uint8, device=dev)
hout = torch.empty((hn,), dtype=torch.int64, device=dev)
def timed(fn, reps=3):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
cms = timed(lambda: hc.compress_device(src, hstride, hsz, True, dst=hdst, out_sizes=hout))
assert int(hout.min()) > 0
hcsz = hout.to(torch.int32)
hdec = torch.empty((hn, hstride), dtype=torch.uint8, device=dev)
hres = torch.empty((hn,), dtype=torch.int64, device=dev)
dms = timed(lambda: hc.decompress_device(hdst.view(-1), hstride, hcsz, hsz, hstride, True, dst=hdec, out_sizes=hres))
assert bool((hres == hb).all()) and torch.equal(hdec[:, :hb], src.view(hn, hstride)[:, :hb])
print("huff0 %d x %d B: Compress4X %.3f ms = %.1f GB/s (ratio %.4f); Decompress4X %.3f ms = %.1f GB/s" % (
    hn, hb, cms, hn * hb / cms / 1e6, float(hout.sum()) / (hn * hb), dms, hn * hb / dms / 1e6))
