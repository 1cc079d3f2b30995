#!/usr/bin/env python
"""Frame-mode output size under the emulator for a build with extra macros (e.g. -DFRAME_BLOCK1=49152u), next to
independent chunks and the reference's single frame (oracle).  usage: emu_frame_ratio.py [-D...] [--levels=1,2]"""
import ctypes, os, subprocess, sys, hashlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
import emu_util
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
syn = H.synth_text(1 << 19, 9)
import numpy as np
# emu_encode_frames asserts the planned block count for the default geometry: bypass by computing from the macro
for level in levels:
    for name, data in (("twain", tw[:6 * 65536]), ("synth", syn)):
        fb = None
        for d in defs:
            if d.startswith("-DFRAME_BLOCK%d=" % (1 if level == 1 else 2)):
                fb = int(d.split("=")[1].rstrip("u"))
        import types
        src = emu_util.emu_encode_frames
        if fb:
            code = open(os.path.join(ROOT, "tests", "emu_util.py")).read().replace("fblock = 32768 if level == 1 else 65536", "fblock = %d" % fb)
            ns = {}
            exec(compile(code, "emu_util_mod", "exec"), ns)
            src = ns["emu_encode_frames"]
        frames, _, _ = src(E, [data], level=level, dump=False)
        assert H.libzstd_decode(frames[0], len(data)) == data
        block = 65536 if level == 1 else 131072
        indep = sum(len(f) for f in emu_util.emu_encode(E, [data[i:i + block] for i in range(0, len(data), block)], level=level)[0])
        ref = H.oracle_encode(data, level=level)[0]
        print("L%d %s %s: frame %d  indep %d (%+.2f%%)  ref %d (%+.2f%%)" % (level, " ".join(defs), name, len(frames[0]), indep,
              100.0 * (len(frames[0]) - indep) / indep, ref, 100.0 * (len(frames[0]) - ref) / ref))
