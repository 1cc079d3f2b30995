#!/usr/bin/env python
"""Round-trip campaign of the emulated encode kernels (the shape of the reference's FuzzEncoding, zstd/fuzz_test.go:154-322;
a longer run of tests/test_emu_encoder_random.py): random structured inputs at ragged sizes through
  * the chunk encoders at levels 1-3 (both lane orders),
  * frame mode at levels 1-3 (inputs of several blocks, with history),
  * S2 fast / better and the Snappy-compatible variants,
  * standalone huff0 4X / 1X (output bytes equal to the oracle's),
each decoded by the oracle, by libzstd (zstd) and by the emulated decode kernels, which must return the input.
Test infrastructure, CPU only.   python tools/fuzz_emu_enc.py [--seed S] [--rounds R]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

import helpers as H                                                                  # noqa: E402
from emu_util import (emu_encode, emu_decode, emu_s2_encode, emu_s2_decode, emu_encode_frames,          # noqa: E402
                      emu_huf_compress, emu_huf_decompress)
from test_emu_huf0 import orc_compress as orc_huf_compress, orc_decompress as orc_huf_decompress   # noqa: E402
from test_emu_encoder_random import _structured                                     # noqa: E402
from test_oracle_s2 import s2_decode as orc_s2_decode                               # noqa: E402


def fail(what, i, data):
    path = "/tmp/fuzz_enc_fail_%s_%d.bin" % (what, i)
    open(path, "wb").write(data)
    raise SystemExit("MISMATCH %s input %d (%d bytes) -> %s" % (what, i, len(data), path))


def check_zstd(E, what, inputs, frames):
    for i, (c, f) in enumerate(zip(inputs, frames)):
        f = bytes(f)
        n, got = H.oracle_decode(f, len(c) + 16)
        if n != len(c) or got != c:
            fail(what + "_oracle", i, c)
        if H.libzstd_decode(f, max(len(c), 1)) != c:
            fail(what + "_libzstd", i, c)
    sizes, outs = emu_decode(E, [bytes(f) for f in frames], [len(c) + 16 for c in inputs])
    for i, c in enumerate(inputs):
        if int(sizes[i]) != len(c) or outs[i] != c:
            fail(what + "_emudec", i, c)


def one_round(E, rng, stats):
    edge = [65536, 65535, 65521, 8191, 4097, 1000, 333, 69, 68, 67, 33, 9, 8, 7, 1]
    sizes = [int(rng.integers(1, 65537)) for _ in range(10)] + list(rng.choice(edge, 5))
    for level, blk in ((1, 65536), (2, 131072), (3, 131072)):
        chunks = [_structured(rng, min(int(n) * (blk // 65536), blk)) for n in sizes]
        frames = emu_encode(E, chunks, level=level, desc=int(rng.integers(0, 2)))[0]
        check_zstd(E, "chunks_L%d" % level, chunks, frames)
        stats["chunks"] = stats.get("chunks", 0) + len(chunks)
    for level in (1, 2, 3):
        inputs = [_structured(rng, int(rng.integers(1, 400000))) for _ in range(4)] + [_structured(rng, 49152 * 2), b""]
        inputs = [x for x in inputs if x]
        frames = emu_encode_frames(E, inputs, level=level, dump=False, desc=int(rng.integers(0, 2)))[0]
        check_zstd(E, "frames_L%d" % level, inputs, frames)
        stats["frames"] = stats.get("frames", 0) + len(inputs)
    blocks = [_structured(rng, int(n)) for n in sizes]
    for snappy in (False, True):
        for better in (False, True):
            enc = emu_s2_encode(E, blocks, snappy=snappy, better=better, desc=int(rng.integers(0, 2)))[0]
            sizes2, outs, _, _ = emu_s2_decode(E, [bytes(e) for e in enc], [len(b) for b in blocks])
            for i, (b, e) in enumerate(zip(blocks, enc)):
                n, got = orc_s2_decode(bytes(e), len(b))
                if n != len(b) or got != b:
                    fail("s2_oracle_%d%d" % (snappy, better), i, b)
                if int(sizes2[i]) != len(b) or outs[i] != b:
                    fail("s2_emudec_%d%d" % (snappy, better), i, b)
            stats["s2"] = stats.get("s2", 0) + len(blocks)
    # standalone huff0: bytes and result class equal the oracle's, and the emulated decompressor returns the input
    hb = [_structured(rng, int(rng.integers(1, 200000))) for _ in range(6)] + [bytes(rng.integers(0, int(rng.integers(2, 200)), int(rng.integers(100, 100000)), dtype=np.uint8)) for _ in range(6)]
    for four in (True, False):
        got = emu_huf_compress(E, hb, four)
        ok = []
        for i, (b, (comp, code)) in enumerate(zip(hb, got)):
            wcomp, wcode = orc_huf_compress(b, four)
            if (code if code < 0 else 0) != (wcode if wcode < 0 else 0) or comp != wcomp:
                fail("huf0_compress_%d" % four, i, b)
            if code > 0:
                ok.append((comp, b))
        if ok:
            back = emu_huf_decompress(E, [c for c, _ in ok], [len(b) for _, b in ok], four)
            for i, ((c, b), (out, code)) in enumerate(zip(ok, back)):
                if out != b or orc_huf_decompress(c, len(b), four)[1] != b:
                    fail("huf0_decompress_%d" % four, i, b)
        stats["huf0"] = stats.get("huf0", 0) + len(hb)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--rounds", type=int, default=5)
    a = ap.parse_args()
    H.build_oracle()
    E = H.emu()
    stats = {}
    t0 = time.time()
    for r in range(a.rounds):
        one_round(E, np.random.default_rng(a.seed * 7919 + r), stats)
        print("round %d  %.0f s  %s" % (r, time.time() - t0, stats), flush=True)
    print("clean: every input came back through every decoder")


if __name__ == "__main__":
    main()
