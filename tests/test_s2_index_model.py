"""The S2 seek index (s2/index.go) on the host: varint known answers, a hand-assembled index, Load(appendTo) identity and
error classes, Find's rules, IndexStream over model streams, header removal, and the reference's ExampleIndex_Load scenario
(s2/index_test.go:16-104: 5 MiB of compressible bytes in 100 KiB blocks, skip forward to every 555 555th offset using only
the index and the bytes from the found chunk on) with the stream model + the oracle's block codecs standing in for the
device stream decoder.  CPU only."""
import ctypes
import io

import numpy as np
import pytest

import s2_stream_ref as R
from compress_b200 import s2_index as X


def _codecs():
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
    return enc, dec


def test_varint_known_answers():
    # encoding/binary: zig-zag then base 128
    for v, want in ((0, b"\x00"), (-1, b"\x01"), (1, b"\x02"), (-64, b"\x7f"), (64, b"\x80\x01"), (-65, b"\x81\x01"),
                    (300, b"\xd8\x04"), ((1 << 63) - 1, b"\xfe\xff\xff\xff\xff\xff\xff\xff\xff\x01"),
                    (-(1 << 63), b"\xff\xff\xff\xff\xff\xff\xff\xff\xff\x01")):
        assert X.put_varint(v) == want, v
        assert X.varint(want) == (v, len(want))
        assert X.varint(b"zz" + want + b"tail", 2) == (v, len(want))
    assert X.varint(b"") == (0, 0) and X.varint(b"\x80\x80") == (0, 0)          # short buffer
    assert X.varint(b"\xff" * 9 + b"\x02")[1] < 0 and X.varint(b"\xff" * 11)[1] < 0   # overflow


def test_hand_assembled_index():
    # a stream that is only the identifier: totals 0 / 10, no entries
    want = (b"\x99\x15\x00\x00" + b"s2idx\x00" + b"\x00" + b"\x14" + b"\x00" + b"\x00" + b"\x00" +
            (25).to_bytes(4, "little") + b"\x00xdi2s")
    assert X.IndexStream(R.MAGIC_S2) == want
    # two entries 2 MiB apart with an implied uncompressed spacing: est 2 MiB, predictions est/2 = 1 MiB
    i = X.Index(); i.reset(2 << 20)
    i.add(10, 0); i.add(10 + 700001, 2 << 20); i.add(10 + 700001 + 900000, 4 << 20)
    b = i.appendTo(b"", 5 << 20, 2000000)
    # first delta 700001 - 1048576 = -348575; Go's /2 truncates toward zero: -174287, so the next prediction is 874289
    body = (X.put_varint(5 << 20) + X.put_varint(2000000) + X.put_varint(2 << 20) + X.put_varint(3) + b"\x00" +
            X.put_varint(10) + X.put_varint(-348575) + X.put_varint(900000 - 874289))
    assert b == b"\x99" + (len(body) + 16).to_bytes(3, "little") + b"s2idx\x00" + body + (len(body) + 20).to_bytes(4, "little") + b"\x00xdi2s"
    j = X.Index()
    assert j.Load(b + b"rest") == b"rest"
    assert j.info == i.info and (j.TotalUncompressed, j.TotalCompressed, j.estBlockUncomp) == (5 << 20, 2000000, 2 << 20)


def test_load_append_identity_and_errors():
    rng = np.random.default_rng(3)
    for trial in range(30):
        i = X.Index(); i.reset(int(rng.choice([65536, 1 << 20, 4 << 20])))
        c = u = 0
        c += 10
        n = int(rng.integers(0, 400))
        for _ in range(n):
            i.add(c, u)
            u += int(rng.integers(1, 3 << 20)) if trial % 3 else i.estBlockUncomp
            c += int(rng.integers(1, 2 << 20))
        blob = i.appendTo(b"", u, c)
        j = X.Index()
        assert j.Load(blob) == b""
        assert j.info == i.info and j.TotalUncompressed == u and j.TotalCompressed == c
        assert j.appendTo(b"", u, c) == blob
        assert X.RestoreIndexHeaders(X.RemoveIndexHeaders(blob)) == blob
        assert len(X.RemoveIndexHeaders(blob)) == len(blob) - 20
        # entries are at least 1 MiB apart
        assert all(b[1] - a[1] >= X.minIndexDist for a, b in zip(j.info, j.info[1:]))
        # truncation / damage: never a wrong answer
        for cut in (0, 5, 16, len(blob) - 1):
            with pytest.raises((X.ErrUnexpectedEOF, X.ErrCorrupt)):
                X.Index().Load(blob[:cut])
        with pytest.raises(X.ErrCorrupt):
            X.Index().Load(b"\x98" + blob[1:])
        with pytest.raises(X.ErrUnsupported):
            X.Index().Load(blob[:4] + b"S2idx\x00" + blob[10:])
        with pytest.raises(X.ErrCorrupt):
            X.Index().Load(blob[:-6] + b"\x00xdi2S")
    assert X.RemoveIndexHeaders(b"short") is None and X.RestoreIndexHeaders(b"") == b""


def test_find_rules():
    i = X.Index(); i.reset(1 << 20)
    for k in range(300):
        i.add(10 + k * 400000, k << 20)
    i.TotalUncompressed, i.TotalCompressed = 300 << 20, 10 + 300 * 400000
    assert i.Find(0) == (10, 0)
    assert i.Find((5 << 20) + 17) == (10 + 5 * 400000, 5 << 20)
    assert i.Find((5 << 20) - 1) == (10 + 4 * 400000, 4 << 20)
    assert i.Find(-1) == (10 + 299 * 400000, 299 << 20)                   # from the end
    assert i.Find(300 << 20) == (10 + 299 * 400000, 299 << 20)
    with pytest.raises(X.ErrUnexpectedEOF):
        i.Find((300 << 20) + 1)
    with pytest.raises(X.ErrUnexpectedEOF):
        i.Find(-(300 << 20) - 1)
    small = X.Index(); small.TotalUncompressed = 100
    assert small.Find(50) == (0, 0)
    unknown = X.Index(); unknown.reset(65536)
    with pytest.raises(X.ErrCorrupt):
        unknown.Find(0)
    # more than 65 536 entries are thinned out on serialisation
    big = X.Index(); big.reset(1 << 20)
    big.info = [(10 + k * 1000, k << 20) for k in range(70000)]
    blob = big.appendTo(b"", 70000 << 20, 10 + 70000 * 1000)
    j = X.Index(); j.Load(blob)
    assert len(j.info) == 35000 and j.estBlockUncomp == 2 << 20 and j.info[1] == (10 + 2000, 2 << 20)


def test_index_stream_and_seek_example(oracle_lib):
    enc, dec = _codecs()
    rng = np.random.default_rng(0xBEEF)
    data = (rng.integers(0, 4, 5 << 20, dtype=np.uint8) + ord("0")).astype(np.uint8).tobytes()
    block = 100 << 10
    stream = R.write_stream(data, enc, block_size=block)
    idx_bytes = X.IndexStream(io.BytesIO(stream))
    idx = X.Index(); idx.Load(idx_bytes)
    assert idx.TotalUncompressed == len(data) and idx.TotalCompressed == len(stream) and idx.estBlockUncomp >= block
    assert idx.info[0] == (10, 0) and len(idx.info) == 5
    # every entry is the start of a chunk holding the block that begins at its uncompressed offset
    for c, u in idx.info:
        assert u % block == 0 and stream[c] in (0, 1)
    decode_stream = lambda part: R.read_stream(part, dec)
    for want in range(0, len(data), 555555):
        c, u = idx.Find(want)
        # "we cannot seek in input, but we have the index": only the bytes from the found chunk on are used
        got = R.read_stream(R.MAGIC_S2 + stream[c:], dec) if c else R.read_stream(stream, dec)
        assert got[want - u:] == data[want:]
        assert X.read_range(stream, want, 1000, decode_stream, idx_bytes) == data[want:want + 1000]
    assert X.read_range(stream, -100, None, decode_stream, idx) == data[-100:]
    assert X.read_range(stream, 3 << 20, 2 << 20, decode_stream, idx) == data[3 << 20:]
    assert X.read_range(stream, len(data), 10, decode_stream, idx) == b""
    # the range reader hands the decoder only the chunks it needs
    sizes = []
    X.read_range(stream, (2 << 20) + 5, 10, lambda part: (sizes.append(len(part)), R.read_stream(part, dec))[1], idx)
    assert sizes[0] < len(stream) // 3
    # appended to the stream it is skipped by readers and found from the end
    full = stream + idx_bytes
    assert R.read_stream(full, dec) == data
    k = X.Index(); k.LoadStream(io.BytesIO(full))
    assert k.info == idx.info
    assert X.read_range(full, 4444440, 77, decode_stream) == data[4444440:4444440 + 77]
    with pytest.raises(X.ErrUnsupported):
        X.Index().LoadStream(io.BytesIO(stream))
    # Snappy streams, uncompressed chunks, structural errors
    sn = R.write_stream(data[:300000] + bytes(rng.integers(0, 256, 200000, dtype=np.uint8)), enc, snappy=True)
    s = X.Index(); s.Load(X.IndexStream(sn))
    assert s.TotalUncompressed == 500000 and s.TotalCompressed == len(sn)
    with pytest.raises(X.ErrCorrupt):
        X.IndexStream(stream[4:])
    with pytest.raises(X.ErrUnexpectedEOF):
        X.IndexStream(stream[:-3])
    with pytest.raises(X.ErrUnsupported):
        X.IndexStream(R.MAGIC_S2 + b"\x05\x04\x00\x00abcd")
    assert b'"total_uncompressed": 5242880' in idx.JSON()
