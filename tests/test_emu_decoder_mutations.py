"""Decoders versus the oracle on corrupted input (the reference's FuzzDecodeAll / FuzzDecoder / s2 FuzzEncodingBlocks
idea, zstd/fuzz_test.go:30-152): every mutated stream must get the oracle's verdict -- same error class, or the same
bytes when it is still accepted.  Runs the kernel sources under the SIMT emulator.  CPU only."""
import numpy as np

import helpers as H
from emu_util import emu_decode, emu_s2_decode, emu_huf_decompress
from test_oracle_s2 import s2_decode as orc_s2_decode, s2_encode as orc_s2_encode
from test_emu_huf0 import orc_compress as orc_huf_compress, orc_decompress as orc_huf_decompress


def _mutations(blob, rng, count):
    out = []
    n = len(blob)
    for k in range(count):
        b = bytearray(blob)
        kind = k % 5
        if kind == 0:      # single bit flip
            i = int(rng.integers(0, n)); b[i] ^= 1 << int(rng.integers(0, 8))
        elif kind == 1:    # byte overwrite, biased to the headers
            i = int(rng.integers(0, min(n, 64))); b[i] = int(rng.integers(0, 256))
        elif kind == 2:    # truncate
            b = b[: int(rng.integers(1, n))]
        elif kind == 3:    # delete a byte
            i = int(rng.integers(0, n)); del b[i]
        else:              # overwrite a short run anywhere
            i = int(rng.integers(0, n)); L = int(rng.integers(1, 5))
            b[i:i + L] = bytes(rng.integers(0, 256, L, dtype=np.uint8))[: max(0, min(L, n - i))]
        out.append(bytes(b))
    return out


def test_zstd_decoder_mutations(emu_lib, oracle_lib, staged_form):
    rng = np.random.default_rng(2024)
    tw = H.golden("twain.txt")
    sources = [tw[:3000], tw[1000:1000 + 70000], H.golden("html.txt")[:20000], bytes(5000), b"ab" * 4000,
               bytes(rng.integers(0, 5, 9000, dtype=np.uint8))]
    cases = []
    for s in sources:
        frame = H.oracle_encode(s)[1]
        cases += [(m, len(s) + 4096) for m in _mutations(frame, rng, 45)]
    outs, res = emu_decode(emu_lib, [c for c, _ in cases], [cap for _, cap in cases])
    agree_ok = 0
    for (c, cap), r, got in zip(cases, outs, res):
        ro, want = H.oracle_decode(c, cap)
        assert ro == r, (ro, int(r), len(c))
        if ro >= 0:
            assert got == want
            agree_ok += 1
    assert 0 < agree_ok < len(cases)       # both outcomes occur


def test_s2_decoder_mutations(emu_lib, oracle_lib):
    rng = np.random.default_rng(7)
    tw = H.golden("twain.txt")
    sources = [tw[:2000], tw[:65536], bytes(3000), b"abc" * 3000, bytes(rng.integers(0, 4, 8000, dtype=np.uint8))]
    cases = []
    for s in sources:
        for mode in (0, 1, 2):
            blk = orc_s2_encode(s, mode)
            cases += [(m, len(s)) for m in _mutations(blk, rng, 25)]
    outs, res, _, _ = emu_s2_decode(emu_lib, [c for c, _ in cases], [cap for _, cap in cases])
    both = set()
    for (c, cap), r, got in zip(cases, outs, res):
        ro, want = orc_s2_decode(c, cap)
        assert ro == r, (ro, int(r))
        if ro >= 0:
            assert got == want
        both.add(ro >= 0)
    assert both == {True, False}


def test_huf0_decoder_mutations(emu_lib, oracle_lib):
    rng = np.random.default_rng(11)
    tw = H.golden("twain.txt")
    sources = [tw[:5000], tw[:100000], bytes(rng.integers(0, 6, 20000, dtype=np.uint8))]
    for four in (True, False):
        cases = []
        for s in sources:
            comp, code = orc_huf_compress(s, four)
            assert code > 0
            cases += [(m, len(s)) for m in _mutations(comp, rng, 30)]
        got = emu_huf_decompress(emu_lib, [c for c, _ in cases], [d for _, d in cases], four)
        for (g, code), (c, d) in zip(got, cases):
            wcode, wout = orc_huf_decompress(c, d, four)
            assert (code < 0) == (wcode < 0), (code, wcode)
            if code >= 0:
                assert g == wout
