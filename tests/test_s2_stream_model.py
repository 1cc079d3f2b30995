"""The S2 framing model used by the GPU stream tests, pinned on the CPU: CRC32-C known answers, the GF(2) combination rule
the device's one-warp checksum relies on, the device routine itself under the emulator, and a model round trip with the
oracle's block codecs (s2/s2.go:118-126 crc(); s2/writer.go:395-470; s2/reader.go:249-420)."""
import ctypes
import os

import numpy as np
import pytest

import helpers as H
import s2_stream_ref as R


def test_crc32c_known_answers():
    assert R.crc32c(b"") == 0
    assert R.crc32c(b"123456789") == 0xE3069283                 # the CRC-32C check value
    assert R.crc32c(bytes(32)) == 0x8A9136AA                     # RFC 3720 B.4: 32 bytes of zeros
    assert R.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43            # RFC 3720 B.4: 32 bytes of ones
    assert R.crc32c(bytes(range(32))) == 0x46DD794E              # RFC 3720 B.4: incrementing
    c = R.crc32c(b"abc")
    assert R.masked_crc(b"abc") == ((((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF)


def test_crc_combination_rule():
    rng = np.random.default_rng(1)
    for la, lb in ((0, 0), (1, 0), (0, 5), (1000, 777), (2048, 2048), (65536, 3)):
        a, b = rng.integers(0, 256, la, dtype=np.uint8).tobytes(), rng.integers(0, 256, lb, dtype=np.uint8).tobytes()
        assert R.crc_combine(R.crc32c(a), R.crc32c(b), lb) == R.crc32c(a + b)


def test_device_checksum_routine_under_the_emulator(emu_lib):
    E = emu_lib
    E.emu_s2_stream_crc.restype = ctypes.c_uint32
    E.emu_s2_stream_crc.argtypes = [ctypes.c_void_p, ctypes.c_uint32]
    rng = np.random.default_rng(2)
    for n in (0, 1, 3, 4, 31, 32, 33, 127, 128, 129, 1000, 4096, 65535, 65536):
        for mis in (0, 1, 3):
            buf = np.zeros(n + 16, dtype=np.uint8)
            buf[mis:mis + n] = rng.integers(0, 256, n, dtype=np.uint8)
            got = E.emu_s2_stream_crc(buf.ctypes.data + mis, n)
            assert got == R.masked_crc(bytes(buf[mis:mis + n])), (n, mis)


def test_model_round_trip_with_oracle_blocks():
    from test_oracle_s2 import s2_decode, _L
    L = _L()
    L.orc_s2_encode_block.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int64, ctypes.c_int]

    def enc(blk):
        out = ctypes.create_string_buffer(L.orc_s2_max_encoded_len(len(blk)) + 16)
        r = L.orc_s2_encode_block(out, bytes(blk), len(blk), 0)
        return out.raw[:r]

    def dec(body, n):
        r, out = s2_decode(body, n)
        return out if r == n else None
    tw = H.golden("twain.txt")
    data = tw[:100000] + os.urandom(70000) + tw[100000:150000]
    st = R.write_stream(data, enc, extra_chunks=True)
    assert R.read_stream(st, dec) == data
    bad = bytearray(st); bad[14] ^= 1
    with pytest.raises(ValueError, match="crc"):
        R.read_stream(bytes(bad), dec)


def test_concat_blocks_matches_the_reference_rule():
    """s2.ConcatBlocks (s2/encode.go:322-361) in the host mirror: the joined block decodes (oracle s2Decode) to the joined
    content; empty input gives the one-byte block; a bad varint is ErrCorrupt."""
    import importlib.util, sys, types
    # the mirror module imports torch and the CUDA library: load only its pure helpers
    src = open(os.path.join(os.path.dirname(__file__), "..", "compress_b200", "s2.py")).read()
    ns = {}
    pre = src[src.index("class ErrCorrupt"):src.index("class Codec:")]
    exec("class B2CError(RuntimeError):\n    pass\n" + pre, ns)
    from test_oracle_s2 import s2_encode, s2_decode
    tw = H.golden("twain.txt")
    pieces = [tw[:65536], tw[65536:100000], b"", tw[100000:100007]]
    blocks = [s2_encode(p_, 0) for p_ in pieces]
    joined = ns["ConcatBlocks"](blocks)
    want = b"".join(pieces)
    r, out = s2_decode(joined, len(want))
    assert r == len(want) and out == want
    assert ns["ConcatBlocks"]([s2_encode(b"", 0)]) == b"\x00"
    with pytest.raises(ns["ErrCorrupt"]):
        ns["ConcatBlocks"]([b"\xff\xff\xff\xff\xff\xff"])


def _emu_stream(E, data, block=65536, snappy=False, better=False):
    E.emu_s2_encode_stream.restype = ctypes.c_int64
    E.emu_s2_encode_stream.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int,
                                       ctypes.c_int]
    src = np.frombuffer(bytes(data) + bytes(64), dtype=np.uint8).copy()
    cap = len(data) + len(data) // 5 + 8 * (len(data) // block + 2) + 64
    dst = np.zeros(cap, dtype=np.uint8)
    r = E.emu_s2_encode_stream(src.ctypes.data, len(data), block, dst.ctypes.data, cap, 1 if snappy else 0, 1 if better else 0)
    assert r >= 10, r
    return bytes(dst[:r])


def test_emulated_stream_writer_against_the_model(emu_lib):
    """The device's stream assembly (block encode, per-block checksum, scan, placement) under the emulator: the stream is read
    by the model of s2.Reader with the oracle's s2Decode; incompressible blocks are uncompressed chunks; the last block is ragged."""
    from test_oracle_s2 import s2_decode

    def dec(body, n):
        r, out = s2_decode(body, n)
        return out if r == n else None
    tw = H.golden("twain.txt")
    rng = np.random.default_rng(12)
    cases = [b"", b"a", tw[:100], tw[:65536], tw[:65537], tw[:200001],
             tw[:70000] + rng.integers(0, 256, 70000, dtype=np.uint8).tobytes() + tw[70000:90000]]
    for data in cases:
        for snappy, better in ((False, False), (False, True), (True, False)):
            st = _emu_stream(emu_lib, data, snappy=snappy, better=better)
            assert st[:10] == (R.MAGIC_SNAPPY if snappy else R.MAGIC_S2)
            assert R.read_stream(st, dec) == data
    st = _emu_stream(emu_lib, tw[:50000], block=4096)
    assert R.read_stream(st, dec) == tw[:50000]
    o, types = 10, []
    st = _emu_stream(emu_lib, cases[-1])
    while o < len(st):
        types.append(st[o]); o += 4 + (st[o + 1] | st[o + 2] << 8 | st[o + 3] << 16)
    assert 1 in types and 0 in types          # the random block travels uncompressed, the text blocks compressed
