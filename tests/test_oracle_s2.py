"""The S2 oracle (oracle/orc_s2.c) pinned against the reference's byte-exact KATs, its decode table, the golden
Snappy block and pyarrow's independent Snappy codec.  CPU only."""
import ctypes

import numpy as np
import pytest

import helpers as H


def _L():
    L = H.oracle()
    c = ctypes
    for nm in ("orc_s2_emit_literal", "orc_s2_emit_repeat", "orc_s2_emit_copy", "orc_s2_emit_copy_norepeat",
               "orc_s2_max_encoded_len", "orc_s2_encode", "orc_s2_decode", "orc_s2_encode_block"):
        getattr(L, nm).restype = c.c_int64
    L.orc_s2_emit_literal.argtypes = [c.c_char_p, c.c_char_p, c.c_size_t]
    L.orc_s2_emit_repeat.argtypes = [c.c_char_p, c.c_int64, c.c_int64]
    L.orc_s2_emit_copy.argtypes = [c.c_char_p, c.c_int64, c.c_int64]
    L.orc_s2_emit_copy_norepeat.argtypes = [c.c_char_p, c.c_int64, c.c_int64]
    L.orc_s2_max_encoded_len.argtypes = [c.c_int64]
    L.orc_s2_encode.argtypes = [c.c_char_p, c.c_size_t, c.c_char_p, c.c_int64, c.c_int]
    L.orc_s2_decode.argtypes = [c.c_char_p, c.c_size_t, c.c_char_p, c.c_size_t]
    return L


def s2_encode(data, mode=0):
    L = _L()
    cap = L.orc_s2_max_encoded_len(len(data))
    out = ctypes.create_string_buffer(cap + 16)
    r = L.orc_s2_encode(out, cap, bytes(data), len(data), mode)
    assert r > 0 or (r == 1 and len(data) == 0), r
    return out.raw[:r]


def s2_decode(comp, cap):
    L = _L()
    out = ctypes.create_string_buffer(max(cap, 1))
    r = L.orc_s2_decode(out, cap, bytes(comp), len(comp))
    return r, out.raw[:max(r, 0)]


def test_emit_literal_kat(oracle_lib):
    # s2/s2_test.go:827-862 TestEmitLiteral
    cases = [(1, b"\x00"), (2, b"\x04"), (59, b"\xe8"), (60, b"\xec"), (61, b"\xf0\x3c"), (62, b"\xf0\x3d"),
             (254, b"\xf0\xfd"), (255, b"\xf0\xfe"), (256, b"\xf0\xff"), (257, b"\xf4\x00\x01"), (65534, b"\xf4\xfd\xff"),
             (65535, b"\xf4\xfe\xff"), (65536, b"\xf4\xff\xff")]
    L = _L()
    nines = b"\x99" * 65536
    for length, want in cases:
        dst = ctypes.create_string_buffer(70000)
        n = L.orc_s2_emit_literal(dst, nines[:length], length)
        assert dst.raw[:n].endswith(nines[:length])
        assert dst.raw[:n - length] == want, length


def test_emit_copy_kat(oracle_lib):
    # s2/s2_test.go:864-942 TestEmitCopy
    cases = [
        (8, 4, b"\x01\x08"), (8, 11, b"\x1d\x08"), (8, 12, b"\x2e\x08\x00"), (8, 13, b"\x32\x08\x00"), (8, 59, b"\xea\x08\x00"),
        (8, 60, b"\xee\x08\x00"), (8, 61, b"\xf2\x08\x00"), (8, 62, b"\xf6\x08\x00"), (8, 63, b"\xfa\x08\x00"),
        (8, 64, b"\xfe\x08\x00"), (8, 65, b"\x11\x08\x15\x001"), (8, 66, b"\x11\x08\x15\x002"), (8, 67, b"\x11\x08\x15\x003"),
        (8, 68, b"\x11\x08\x15\x004"), (8, 69, b"\x11\x08\x15\x005"), (8, 80, b"\x11\x08\x15\x00@"),
        (8, 800, b"\x11\x08\x19\x00\x14\x02"), (8, 800000, b"\x11\x08\x1d\x00\xf44\x0b"),
        (256, 4, b"\x21\x00"), (256, 11, b"\x3d\x00"), (256, 12, b"\x2e\x00\x01"), (256, 13, b"\x32\x00\x01"),
        (256, 59, b"\xea\x00\x01"), (256, 60, b"\xee\x00\x01"), (256, 61, b"\xf2\x00\x01"), (256, 62, b"\xf6\x00\x01"),
        (256, 63, b"\xfa\x00\x01"), (256, 64, b"\xfe\x00\x01"), (256, 65, b"1\x00\x15\x001"), (256, 66, b"1\x00\x15\x002"),
        (256, 67, b"1\x00\x15\x003"), (256, 68, b"1\x00\x15\x004"), (256, 69, b"1\x00\x15\x005"), (256, 80, b"1\x00\x15\x00@"),
        (256, 800, b"1\x00\x19\x00\x14\x02"), (256, 80000, b"1\x00\x1d\x00t8\x00"),
        (2048, 4, b"\x0e\x00\x08"), (2048, 11, b"\x2a\x00\x08"), (2048, 12, b"\x2e\x00\x08"), (2048, 13, b"\x32\x00\x08"),
        (2048, 59, b"\xea\x00\x08"), (2048, 60, b"\xee\x00\x08"), (2048, 61, b"\xf2\x00\x08"), (2048, 62, b"\xf6\x00\x08"),
        (2048, 63, b"\xfa\x00\x08"), (2048, 64, b"\xfe\x00\x08"), (2048, 65, b"\xee\x00\x08\x05\x00"),
        (2048, 66, b"\xee\x00\x08\x09\x00"), (2048, 67, b"\xee\x00\x08\x0d\x00"), (2048, 68, b"\xee\x00\x08\x11\x00"),
        (2048, 69, b"\xee\x00\x08\x15\x00\x01"), (2048, 80, b"\xee\x00\x08\x15\x00\x0c"),
        (2048, 800, b"\xee\x00\x08\x19\x00\xe0\x01"), (2048, 80000, b"\xee\x00\x08\x1d\x00\x40\x38\x00"),
        (204800, 4, b"\x0f\x00\x20\x03\x00"), (204800, 65, b"\xff\x00\x20\x03\x00\x03\x00\x20\x03\x00"),
        (204800, 69, b"\xff\x00\x20\x03\x00\x05\x00"), (204800, 800, b"\xff\x00\x20\x03\x00\x19\x00\xdc\x01"),
        (204800, 80000, b"\xff\x00\x20\x03\x00\x1d\x00\x3c\x38\x00"),
    ]
    L = _L()
    for off, length, want in cases:
        dst = ctypes.create_string_buffer(1024)
        n = L.orc_s2_emit_copy(dst, off, length)
        assert dst.raw[:n] == want, (off, length, dst.raw[:n])


def test_max_encoded_len(oracle_lib):
    # s2/s2_test.go:37-60 TestMaxEncodedLen (64-bit rows)
    L = _L()
    assert L.orc_s2_max_encoded_len(0) == 1
    assert L.orc_s2_max_encoded_len(1 << 24) == (1 << 24) + 4 + 5
    m32 = 0xffffffff
    assert L.orc_s2_max_encoded_len(m32 - 5 - 5) == m32          # MaxBlockSize
    for k in range(0, 10):
        assert L.orc_s2_max_encoded_len(m32 - k) == -1
    assert L.orc_s2_max_encoded_len(-1) == -1 and L.orc_s2_max_encoded_len(-2) == -1


def test_invalid_varint(oracle_lib):
    # s2/s2_test.go:214-250 TestInvalidVarint
    from s2_vectors import INVALID_VARINT
    for inp in INVALID_VARINT:
        r, _ = s2_decode(inp, 100)
        assert r == -5, inp


def test_decode_table(oracle_lib):
    # s2/s2_test.go:252-470 TestDecode
    from s2_vectors import DECODE_TABLE
    for i, (inp, want, ok) in enumerate(DECODE_TABLE):
        r, got = s2_decode(inp, 100)
        if ok:
            assert r == len(want) and got == want, i
        else:
            assert r == -5, i


def test_golden_snappy_block(oracle_lib):
    # s2/s2_test.go:599-618 TestDecodeGoldenInput: the .rawsnappy block decodes to the text
    comp = H.golden("s2_twain.txt.rawsnappy")
    want = H.golden("s2_twain.txt")
    r, got = s2_decode(comp, len(want))
    assert r == len(want) and got == want


def _corpus():
    rng = np.random.default_rng(5)
    tw = H.golden("twain.txt")
    return [b"", b"a", b"abc" * 5, bytes(31), bytes(32), bytes(100), tw[:33], tw[:1000], tw[:65536], tw[:65537], tw[:200000],
            H.golden("html.txt")[:100000], H.golden("e.txt")[:70000], H.synth_text(300000), bytes(70000), b"ab" * 40000,
            bytes(rng.integers(0, 256, 5000, dtype=np.uint8)), bytes(rng.integers(0, 256, 70000, dtype=np.uint8)),
            bytes(rng.integers(0, 3, 70000, dtype=np.uint8)), (tw[:3000] + bytes(rng.integers(0, 256, 70000, dtype=np.uint8))) * 2]


def test_roundtrip_all_modes(oracle_lib):
    # s2/s2_test.go:93-153 roundtrip(): Encode / EncodeBetter / EncodeSnappy decode back, sizes within MaxEncodedLen
    L = _L()
    for i, s in enumerate(_corpus()):
        for mode in (0, 1, 2):
            comp = s2_encode(s, mode)
            assert len(comp) <= L.orc_s2_max_encoded_len(len(s))
            r, got = s2_decode(comp, len(s))
            assert r == len(s) and got == s, (i, mode)


def test_snappy_mode_is_snappy(oracle_lib):
    # EncodeSnappy output must decode with an independent Snappy decoder (pyarrow); S2 text compresses
    pa = pytest.importorskip("pyarrow")
    codec = pa.Codec("snappy")
    for i, s in enumerate(_corpus()):
        if len(s) == 0:
            continue
        comp = s2_encode(s, 2)
        out = codec.decompress(comp, decompressed_size=len(s))
        assert out.to_pybytes() == s, i
        # and pyarrow's own Snappy blocks decode through the oracle's s2Decode
        pc = codec.compress(s).to_pybytes()
        r, got = s2_decode(pc, len(s))
        assert r == len(s) and got == s, i
    tw = H.golden("twain.txt")[:65536]
    assert len(s2_encode(tw, 0)) < 0.72 * len(tw) and len(s2_encode(tw, 1)) < len(s2_encode(tw, 0))
