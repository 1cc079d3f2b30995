"""Standalone huff0 kernels (b2c_huf0.cuh) under the SIMT emulator: output bytes must equal the oracle's
huff0.Compress4X / Compress1X (fresh Scratch), errors must match, decompress must invert.  CPU only."""
import ctypes

import numpy as np

import helpers as H
from emu_util import emu_huf_compress, emu_huf_decompress


def orc_compress(data, four):
    L = H.oracle()
    L.orc_huf_compress_oneshot.restype = ctypes.c_int64
    L.orc_huf_compress_oneshot.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint, ctypes.c_char_p,
                                           ctypes.c_size_t, ctypes.c_void_p]
    cap = len(data) + 1024
    out = ctypes.create_string_buffer(cap)
    r = L.orc_huf_compress_oneshot(bytes(data), len(data), 1 if four else 0, 0, out, cap, None)
    return (out.raw[:r] if r >= 0 else None), int(r)


def orc_decompress(comp, dst_size, four):
    """huff0.ReadTable + Decompress4X/1X through the oracle: (code, bytes)."""
    L = H.oracle()

    class DT(ctypes.Structure):
        _fields_ = [("dt", ctypes.c_uint16 * 2048), ("actualTableLog", ctypes.c_uint), ("loaded", ctypes.c_int)]
    d = DT()
    L.orc_huf_read_table.restype = ctypes.c_int64
    L.orc_huf_read_table.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t]
    used = L.orc_huf_read_table(ctypes.byref(d), bytes(comp), len(comp))
    if used < 0:
        return -5, None
    out = ctypes.create_string_buffer(max(dst_size, 1))
    fn = L.orc_huf_decompress4x if four else L.orc_huf_decompress1x
    fn.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t]
    rest = bytes(comp)[used:]
    e = fn(ctypes.byref(d), rest, len(rest), out, dst_size)
    return (-5, None) if e else (dst_size, out.raw[:dst_size])


def _blocks():
    rng = np.random.default_rng(3)
    tw = H.golden("twain.txt")
    skew = bytes(np.minimum(rng.geometric(0.3, 262143), 200).astype(np.uint8))
    return [b"", b"a", b"ab" * 3, b"abcabcabcabc", tw[:11], tw[:12], tw[:13], tw[:100], tw[:1000], tw[:65536], tw[:262143],
            skew, skew[:70001], bytes(1000), b"a" * 999 + b"b", bytes(rng.integers(0, 256, 20000, dtype=np.uint8)),
            bytes(rng.integers(0, 2, 5000, dtype=np.uint8)), bytes(rng.integers(0, 130, 3000, dtype=np.uint8)),
            H.synth_text(262143), bytes(range(256)) * 4, bytes([i % 7 for i in range(4000)]) + tw[:3000]]


def test_emu_huf_compress_parity(emu_lib, oracle_lib):
    blocks = _blocks()
    for four in (True, False):
        want = [orc_compress(b, four) for b in blocks]
        for desc in (0, 1):
            got = emu_huf_compress(emu_lib, blocks, four, desc)
            for i, (g, w) in enumerate(zip(got, want)):
                assert g[1] == w[1], (four, i, g[1], w[1])
                assert g[0] == w[0], (four, i)
    # too big: 262144 bytes is rejected like huff0.prepare (huff0.go:135)
    big = emu_huf_compress(emu_lib, [bytes(np.random.default_rng(1).integers(0, 4, 262144, dtype=np.uint8))])
    assert big[0][1] == -3


def test_emu_huf_decompress(emu_lib, oracle_lib):
    blocks = [b for b in _blocks()]
    for four in (True, False):
        comp = [(orc_compress(b, four), b) for b in blocks]
        ok = [(c[0], b) for c, b in comp if c[1] > 0]
        assert len(ok) >= 10
        got = emu_huf_decompress(emu_lib, [c for c, _ in ok], [len(b) for _, b in ok], four)
        for (g, code), (_, b) in zip(got, ok):
            assert code == len(b) and g == b
        # wrong size, truncated stream and garbage: same verdict as the oracle's ReadTable + Decompress (exact bit consumption)
        c0, b0 = ok[-1]
        cases = [(c0, len(b0) - 1), (c0[:-1], len(b0)), (c0[:40], len(b0)), (bytes([200]) * 50, 100), (c0[:-2] + b"\x00\x00", len(b0))]
        bad = emu_huf_decompress(emu_lib, [c for c, _ in cases], [d for _, d in cases], four)
        for (g, code), (c, d) in zip(bad, cases):
            wcode, wout = orc_decompress(c, d, four)
            assert (code < 0) == (wcode < 0), (code, wcode)
            if code >= 0:
                assert g == wout
        assert bad[0][1] == -5 and bad[3][1] == -5 and bad[4][1] == -5


def test_emu_read_table_matches_oracle(emu_lib, oracle_lib):
    """huff0.ReadTable (huff0/decompress.go:29-166): bytes consumed, table log and the code length of every symbol equal
    the oracle's for table descriptions produced by the compressor (raw 4-bit weights and FSE-compressed weights)."""
    import ctypes
    import numpy as np
    from emu_util import emu_huf_compress
    rng = np.random.default_rng(3)
    blocks = [H.golden("twain.txt")[:50000], H.golden("e.txt")[:30000], bytes(rng.integers(0, 7, 20000, dtype=np.uint8)),
              bytes(rng.integers(97, 123, 4000, dtype=np.uint8)), H.golden("html.txt")]
    comp = [c for c, code in emu_huf_compress(emu_lib, blocks, four=True)]
    assert all(c is not None for c in comp)
    n = len(comp)
    stride = max(len(c) for c in comp) + 16
    src = np.zeros((n, stride), dtype=np.uint8)
    for i, c in enumerate(comp):
        src[i, :len(c)] = np.frombuffer(c, dtype=np.uint8)
    sizes = np.array([len(c) for c in comp], dtype=np.uint32)
    rows = np.zeros((n, 260), dtype=np.uint8)
    outs = np.zeros(n, dtype=np.int64)
    emu_lib.emu_huf_read_table(src.ctypes.data, stride, sizes.ctypes.data, n, rows.ctypes.data, outs.ctypes.data)

    class DT(ctypes.Structure):
        _fields_ = [("dt", ctypes.c_uint16 * 2048), ("actualTableLog", ctypes.c_uint), ("loaded", ctypes.c_int)]
    L = oracle_lib
    L.orc_huf_read_table.restype = ctypes.c_int64
    L.orc_huf_read_table.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t]
    for i, c in enumerate(comp):
        d = DT()
        used = L.orc_huf_read_table(ctypes.byref(d), c, len(c))
        assert used > 0 and outs[i] == used == int(rows[i, 2]) | (int(rows[i, 3]) << 8)
        assert rows[i, 0] == d.actualTableLog
        want = [0] * 256
        for e in d.dt[: 1 << d.actualTableLog]:
            want[e >> 8] = e & 0xff
        assert list(rows[i, 4:260]) == want, i
    # truncated description -> corrupt
    bad = comp[0][:3]
    src[0, :] = 0
    src[0, :3] = np.frombuffer(bad, dtype=np.uint8)
    sizes[0] = 3
    emu_lib.emu_huf_read_table(src.ctypes.data, stride, sizes.ctypes.data, 1, rows.ctypes.data, outs.ctypes.data)
    assert outs[0] == -5
