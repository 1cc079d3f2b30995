#!/usr/bin/env python
"""Output size of the emulated zstd encoder (levels 1-3) per corpus for a build of the kernels with extra macros, next
to the oracle (reference algorithm).  Design tool: scores parse variants on the CPU before any GPU time is spent.
usage: emu_ratio.py [-DLZ_X=1 ...] [--levels 1,2]"""
import ctypes, os, subprocess, sys, hashlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
from emu_util import emu_encode

defs = [a for a in sys.argv[1:] if a.startswith("-D")]
levels = [1]
for a in sys.argv[1:]:
    if a.startswith("--levels"):
        levels = [int(x) for x in a.split("=")[1].split(",")]
tag = hashlib.md5(" ".join(defs).encode()).hexdigest()[:8]
so = "/tmp/libb2c_emu_%s.so" % tag
emu_dir = os.path.join(ROOT, "tests", "emu")
subprocess.run(["g++", "-O2", "-fPIC", "-std=c++17", "-w", "-I" + emu_dir] + defs + ["-shared", "-o", so,
                os.path.join(emu_dir, "simt_emu.cpp"), os.path.join(emu_dir, "emu_kernels.cpp")], check=True)
H.EMU_SO = so
H.build_emu = lambda: None
E = H.emu()
tw = H.golden("twain.txt")
for level in levels:
    block = 65536 if level == 1 else 131072
    corp = {"twain": [tw[i:i + block] for i in range(0, len(tw) - block, block)], "html": [H.golden("html.txt")],
            "e": [H.golden("e.txt")[:block]], "synth": [H.synth_text(block, s) for s in (3, 4, 5, 6)]}
    row = []
    for name, chunks in corp.items():
        frames = emu_encode(E, chunks, level=level)[0]
        for c, f in zip(chunks[:2], frames[:2]):
            assert H.libzstd_decode(f, len(c)) == c if hasattr(H, "libzstd_decode") else True
        got = sum(len(f) for f in frames)
        ref = sum(H.oracle_encode(c, level=level)[0] for c in chunks)
        row.append("%s %d (%+.2f%%)" % (name, got, 100.0 * (got - ref) / ref))
    print("L%d %s: %s" % (level, " ".join(defs) or "(default)", "  ".join(row)))
