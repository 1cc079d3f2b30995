#!/usr/bin/env python
"""Per-kernel device times (ms per GiB of input) of the zstd encode pipeline at level 1 or 2, from the library's own
CUDA-event profile.  usage: enc_times.py [level] [GiB]"""
import hashlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
from compress_b200 import zstd

# --stream: run everything on a torch side stream instead of the (legacy) default stream
if "--stream" in sys.argv:
    _side = torch.cuda.Stream()
    torch.cuda.set_stream(_side)
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
from compress_b200 import _lib
print("level %d  %d x %d KiB   lib md5 %s" % (level, n, CH >> 10,
                                                    hashlib.md5(open(_lib.LIB_PATH, "rb").read()).hexdigest()[:10]))
print("  step %.3f ms = %.1f GB/s; ratio %.4f; sizes sha1 %s" % (
    tot, n * CH / tot / 1e6, float(o.sum()) / (n * CH), hashlib.sha1(o.numpy().tobytes()).hexdigest()[:12]))
print("  per step: " + "  ".join("%s %.3f" % (k.replace("b2c_zstd_", "").replace("b2c_", "").replace("_kernel", ""), v / steps)
                                 for k, v in ms.items()) + "   (%d launches of the pipeline per step)" % (calls // steps))

if "--decode" in sys.argv:
    dec = zstd.Decoder()
    dsz = outs.to(torch.int32)
    dout = torch.empty((n, CH), dtype=torch.uint8, device=dev)
    dres = torch.empty((n,), dtype=torch.int64, device=dev)
    for _ in range(2):
        dec.decode_device(dst, dsz, src_stride=enc.slot, dst=dout, dst_cap=CH, out_sizes=dres)
    torch.cuda.synchronize()
    d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d0.record()
    for _ in range(3):
        dec.decode_device(dst, dsz, src_stride=enc.slot, dst=dout, dst_cap=CH, out_sizes=dres)
    d1.record()
    torch.cuda.synchronize()
    ms = d0.elapsed_time(d1) / 3
    assert bool((dres == CH).all()) and torch.equal(dout.view(-1), src)
    print("  decode %.3f ms = %.1f GB/s (output bytes)" % (ms, n * CH / ms / 1e6))
    dec.profile(True)
    for _ in range(3):
        dec.decode_device(dst, dsz, src_stride=enc.slot, dst=dout, dst_cap=CH, out_sizes=dres)
    pm = dec.profile_read()
    dec.profile(False)
    print("  decode per step: " + "  ".join("%s %.3f" % (k.replace("b2c_zstd_", "").replace("_kernel", ""), v / 3) for k, v in pm.items()))
if "--s2" in sys.argv and level == 1:
    from compress_b200 import s2 as s2mod
    c = s2mod.Codec()
    sd = torch.empty((n, s2mod.SLOT), dtype=torch.uint8, device=dev)
    ss = torch.empty((n,), dtype=torch.int64, device=dev)
    for name, snappy, better in (("s2", False, False), ("s2-better", False, True), ("snappy", True, False)):
        for _ in range(2):
            c.encode_device(src, snappy=snappy, better=better, dst=sd, out_sizes=ss)
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(3):
            c.encode_device(src, snappy=snappy, better=better, dst=sd, out_sizes=ss)
        a1.record()
        torch.cuda.synchronize()
        ms = a0.elapsed_time(a1) / 3
        print("  %-10s %.3f ms = %.1f GB/s; ratio %.4f" % (name, ms, n * CH / ms / 1e6, float(ss.sum()) / (n * CH)))
if "--s2stream" in sys.argv and level == 1:
    from compress_b200 import s2 as s2mod
    c = s2mod.Codec()
    sdst = torch.empty(int(_lib.lib.b2c_s2_stream_bound(n * CH, 65536)) + 16, dtype=torch.uint8, device=dev)
    for name, better in (("s2 stream", False), ("s2 stream (better)", True)):
        for _ in range(2):
            _, tot, err = c.encode_stream_device(src, better=better, dst=sdst)
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(3):
            _, tot, err = c.encode_stream_device(src, better=better, dst=sdst)
        a1.record()
        torch.cuda.synchronize()
        ms = a0.elapsed_time(a1) / 3
        assert int(err.item()) == 0
        print("  %-18s %.3f ms = %.1f GB/s; ratio %.4f" % (name, ms, n * CH / ms / 1e6, int(tot.cpu().numpy()[0]) / (n * CH)))
if "--frames" in sys.argv:
    # frame mode: the same bytes as 1 MiB inputs, one frame each (blocks see their history); --frame-bytes=N: other frame size
    fs = 1 << 20
    for a in sys.argv:
        if a.startswith("--frame-bytes="):
            fs = int(a.split("=")[1])
    nf = (n * CH) // fs
    offs, lens = [i * fs for i in range(nf)], [fs] * nf
    fdst = torch.empty(nf * (fs + 4096), dtype=torch.uint8, device=dev)
    for _ in range(2):
        _, foff, fsz = enc.encode_frames_device(src, offs, lens, dst=fdst)
    torch.cuda.synchronize()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(3):
        _, foff, fsz = enc.encode_frames_device(src, offs, lens, dst=fdst)
    f1.record()
    torch.cuda.synchronize()
    fms = f0.elapsed_time(f1) / 3
    fz = fsz.cpu()
    assert int(fz.min()) > 0
    print("  frame mode (%d x %d KiB frames) %.3f ms = %.1f GB/s; ratio %.4f" % (nf, fs >> 10, fms, nf * fs / fms / 1e6, float(fz.sum()) / (nf * fs)))
    enc.profile(True)
    for _ in range(2):
        enc.encode_frames_device(src, offs, lens, dst=fdst)
    pm, pc = enc.profile_read()
    enc.profile(False)
    print("  frame mode per step: " + "  ".join("%s %.3f" % (k.replace("b2c_zstd_", "").replace("b2c_", "").replace("_kernel", ""), v / 2)
                                                for k, v in pm.items()))
if "--frames-decode" in sys.argv:
    # GPU decode of the multi-block frames (one-warp decoder: more than four blocks per frame)
    dec = zstd.Decoder()
    fo = foff.to(torch.int64)
    fsz32 = fsz.to(torch.int32)
    dout = torch.empty((nf, fs), dtype=torch.uint8, device=dev)
    dres = torch.empty((nf,), dtype=torch.int64, device=dev)
    for _ in range(2):
        dec.decode_device(fdst, fsz32, src_offsets=foff, dst=dout, dst_cap=fs, out_sizes=dres)
    torch.cuda.synchronize()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(2):
        dec.decode_device(fdst, fsz32, src_offsets=foff, dst=dout, dst_cap=fs, out_sizes=dres)
    g1.record()
    torch.cuda.synchronize()
    gms = g0.elapsed_time(g1) / 2
    assert bool((dres == fs).all()) and torch.equal(dout.view(-1), src[: nf * fs])
    print("  decode of the %d multi-block frames (%d KiB each) %.3f ms = %.1f GB/s (output bytes)" % (nf, fs >> 10, gms, nf * fs / gms / 1e6))
    dec.profile(True)
    for _ in range(2):
        dec.decode_device(fdst, fsz32, src_offsets=foff, dst=dout, dst_cap=fs, out_sizes=dres)
    pm = dec.profile_read()
    dec.profile(False)
    print("  frame decode per step: " + "  ".join("%s %.3f" % (k.replace("b2c_zstd_", "").replace("_kernel", ""), v / 2) for k, v in pm.items())
          + "   staged inputs: %d of %d" % (dec.staged_count(nf), nf))
