"""Randomised structure tests for the parse (zstd frames, S2 and Snappy blocks) under the SIMT emulator: mixtures of
periodic runs, repeated segments at many distances, noise and text at ragged sizes must always round-trip through
the oracle decoders (the shape of the reference's FuzzEncoding, zstd/fuzz_test.go:154-322).  CPU only."""
import numpy as np

import helpers as H
from emu_util import emu_encode, emu_s2_encode
from test_oracle_s2 import s2_decode as orc_s2_decode


def _structured(rng, n):
    tw = H.golden("twain.txt")
    out = bytearray()
    while len(out) < n:
        kind = int(rng.integers(0, 7))
        L = int(rng.integers(1, 3000))
        if kind == 0:
            out += bytes(rng.integers(0, 256, L, dtype=np.uint8))
        elif kind == 1:
            out += bytes([int(rng.integers(0, 256))]) * L
        elif kind == 2:
            p = bytes(rng.integers(0, 256, int(rng.integers(2, 40)), dtype=np.uint8))
            out += (p * (L // len(p) + 1))[:L]
        elif kind == 3 and len(out) > 8:
            d = int(rng.integers(1, len(out)))
            for _ in range(L):
                out.append(out[-d])
        elif kind == 4:
            o = int(rng.integers(0, len(tw) - L))
            out += tw[o:o + L]
        elif kind == 5:
            out += bytes(rng.integers(0, 4, L, dtype=np.uint8))
        else:
            out += bytes(rng.integers(97, 123, L, dtype=np.uint8))
    return bytes(out[:n])


def test_emu_random_structures(emu_lib, oracle_lib):
    rng = np.random.default_rng(31337)
    sizes = [65536, 65535, 65521, 40000, 8191, 4097, 1000, 333, 68, 69, 67, 33, 9, 8, 7]
    chunks = [_structured(rng, n) for n in sizes] + [_structured(rng, 65536) for _ in range(5)]
    frames, outs, hdr, _, _ = emu_encode(emu_lib, chunks)
    for c, f, r in zip(chunks, frames, outs):
        assert r == len(f) > 0
        n, got = H.oracle_decode(f, len(c) + 16)
        assert n == len(c) and got == c
        assert H.libzstd_decode(f, max(len(c), 1)) == c
    for snappy in (False, True):
        enc, outs = emu_s2_encode(emu_lib, chunks, snappy=snappy)
        for c, e in zip(chunks, enc):
            n, got = orc_s2_decode(e, len(c))
            assert n == len(c) and got == c
