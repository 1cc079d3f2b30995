#!/usr/bin/env python
"""Mutation campaign of the emulated decode kernels against the oracle (a longer run of tests/test_emu_decoder_mutations.py:
the reference's FuzzDecodeAll / FuzzDecoder idea, zstd/fuzz_test.go:30-152).  Sources: oracle frames at levels 1-3, libzstd
frames at several levels (multi-block, treeless literals, repeat-mode tables), frame-mode output of the emulated encoder,
S2 blocks in all modes.  Every mutated stream must get the oracle's verdict: the same error / success class and, when
accepted, the same bytes -- in both staged forms of the zstd decoder.  Test infrastructure, CPU only.

  python tools/fuzz_emu.py [--seed S] [--rounds R] [--per P]
Under AddressSanitizer: make -C tests/emu asan CXX=/usr/bin/g++, LD_PRELOAD the libasan, run, then make clean all."""
import argparse
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

import helpers as H                                                    # noqa: E402
from emu_util import emu_decode, emu_s2_decode, emu_encode_frames     # noqa: E402
from test_oracle_s2 import s2_decode as orc_s2_decode, s2_encode as orc_s2_encode   # noqa: E402


def mutations(blob, rng, count):
    out = []
    n = len(blob)
    for k in range(count):
        b = bytearray(blob)
        kind = int(rng.integers(0, 8))
        if kind == 0:
            i = int(rng.integers(0, n)); b[i] ^= 1 << int(rng.integers(0, 8))
        elif kind == 1:
            i = int(rng.integers(0, min(n, 64))); b[i] = int(rng.integers(0, 256))
        elif kind == 2:
            b = b[: int(rng.integers(1, n))]
        elif kind == 3:
            i = int(rng.integers(0, n)); del b[i]
        elif kind == 4:
            i = int(rng.integers(0, n)); L = int(rng.integers(1, 5))
            b[i:i + L] = bytes(rng.integers(0, 256, L, dtype=np.uint8))[: max(0, min(L, n - i))]
        elif kind == 5:    # several bit flips
            for _ in range(int(rng.integers(2, 6))):
                i = int(rng.integers(0, n)); b[i] ^= 1 << int(rng.integers(0, 8))
        elif kind == 6:    # copy a slice of the stream over another place (keeps local structure valid)
            L = int(rng.integers(1, min(n, 40) + 1)); s = int(rng.integers(0, n - L + 1)); d = int(rng.integers(0, n - L + 1))
            b[d:d + L] = blob[s:s + L]
        else:              # insert a byte
            i = int(rng.integers(0, n)); b.insert(i, int(rng.integers(0, 256)))
        out.append(bytes(b))
    return out


def libzstd_encode(data, level):
    Z = H.libzstd()
    cap = Z.ZSTD_compressBound(len(data))
    out = ctypes.create_string_buffer(cap)
    r = Z.ZSTD_compress(out, cap, bytes(data), len(data), level)
    assert not Z.ZSTD_isError(r)
    return out.raw[:r]


def sources(rng):
    tw, html = H.golden("twain.txt"), H.golden("html.txt")
    r = lambda n, hi: bytes(rng.integers(0, hi, n, dtype=np.uint8))
    return [tw[:3000], tw[5000:5000 + 66000], tw[:200000], html[:30000], bytes(7000), b"abcd" * 3000, r(9000, 5),
            r(3000, 256), tw[:40000] + r(3000, 256) + tw[:40000], b"x"]


def run_zstd(E, rng, per, stats):
    srcs = sources(rng)
    frames = []
    for s in srcs:
        for lv in (1, 2, 3):
            frames.append((H.oracle_encode(s, lv)[1], len(s)))
        for lv in (-3, 1, 3, 9, 19):
            frames.append((libzstd_encode(s, lv), len(s)))
    big = [s for s in srcs if len(s) > 20000]
    for lv in (1, 2):
        fr = emu_encode_frames(E, big, level=lv, dump=False)[0]
        frames += [(bytes(f), len(s)) for f, s in zip(fr, big)]
    cases = []
    for f, n in frames:
        cases += [(m, n + 4096) for m in mutations(f, rng, per)]
    # two frames back to back, the second one damaged
    for k in range(0, len(frames) - 1, 5):
        (f0, n0), (f1, n1) = frames[k], frames[k + 1]
        cases += [(f0 + m, n0 + n1 + 4096) for m in mutations(f1, rng, 2)]
    want = [H.oracle_decode(c, cap) for c, cap in cases]
    for form, maxb in (("per-block", 0), ("per-input", 4)):
        E.emu_set_dec_maxb(maxb)
        staged = []
        sizes, outs = emu_decode(E, [c for c, _ in cases], [cap for _, cap in cases], staged=staged)
        for i, ((c, cap), (ro, wb)) in enumerate(zip(cases, want)):
            if int(sizes[i]) != ro or (ro >= 0 and outs[i] != wb):
                path = "/tmp/fuzz_fail_%s_%d.zst" % (form, i)
                open(path, "wb").write(c)
                raise SystemExit("MISMATCH zstd %s: oracle %d emu %d cap %d -> %s" % (form, ro, int(sizes[i]), cap, path))
        stats["zstd_" + form] = stats.get("zstd_" + form, 0) + len(cases)
        stats["zstd_staged_" + form] = stats.get("zstd_staged_" + form, 0) + sum(staged)
    E.emu_set_dec_maxb(0)
    stats["zstd_ok"] = stats.get("zstd_ok", 0) + sum(1 for r, _ in want if r >= 0)


def run_s2(E, rng, per, stats):
    cases = []
    for s in sources(rng):
        s = s[:65536]
        for mode in (0, 1, 2):
            blk = orc_s2_encode(s, mode)
            cases += [(m, len(s) + (0 if k & 1 else 64)) for k, m in enumerate(mutations(blk, rng, per))]
    sizes, outs, _, _ = emu_s2_decode(E, [c for c, _ in cases], [cap for _, cap in cases])
    for i, (c, cap) in enumerate(cases):
        ro, wb = orc_s2_decode(c, cap)
        if int(sizes[i]) != ro or (ro >= 0 and outs[i] != wb):
            path = "/tmp/fuzz_fail_s2_%d.bin" % i
            open(path, "wb").write(c)
            raise SystemExit("MISMATCH s2: oracle %d emu %d cap %d -> %s" % (ro, int(sizes[i]), cap, path))
    stats["s2"] = stats.get("s2", 0) + len(cases)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--rounds", type=int, default=3)
    ap.add_argument("--per", type=int, default=12)
    a = ap.parse_args()
    H.build_oracle()
    E = H.emu()
    stats = {}
    t0 = time.time()
    for r in range(a.rounds):
        rng = np.random.default_rng(a.seed * 1000 + r)
        run_zstd(E, rng, a.per, stats)
        run_s2(E, rng, a.per, stats)
        print("round %d  %.0f s  %s" % (r, time.time() - t0, stats), flush=True)
    print("clean: every mutated stream got the oracle's verdict")


if __name__ == "__main__":
    main()
