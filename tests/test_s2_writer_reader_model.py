"""Host logic of the S2 stream Writer / Reader mirrors (s2.NewWriter / s2.NewReader) with a stand-in
codec: the stream model of tests/s2_stream_ref.py over the oracle's block codecs plays EncodeStream / DecodeStream, which the
product classes get from the GPU codec.  Batching, index and padding at Close, Skip, error order, Snappy streams,
concatenated streams.  CPU only."""
import ctypes
import io

import numpy as np
import pytest

import helpers as H
import s2_stream_ref as R
from compress_b200 import s2 as S
from compress_b200 import s2_index as X


class ModelCodec:
    def __init__(self):
        from test_oracle_s2 import s2_decode, _L
        L = _L()
        L.orc_s2_encode_block.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int64, ctypes.c_int]
        self.L, self.s2_decode = L, s2_decode
        self.enc_calls, self.dec_calls = [], []

    def _enc(self, mode):
        def enc(blk):
            out = ctypes.create_string_buffer(self.L.orc_s2_max_encoded_len(len(blk)) + 16)
            r = self.L.orc_s2_encode_block(out, bytes(blk), len(blk), mode)
            return out.raw[:r]                    # the block body (no length prefix); b"" = not compressible
        return enc

    def EncodeStream(self, src, better=False, snappy=False, block_size=65536, index=False):
        self.enc_calls.append(len(src))
        return R.write_stream(bytes(src), self._enc(2 if snappy else (1 if better else 0)), block_size=block_size, snappy=snappy)

    def DecodeStream(self, stream, max_size=None):
        self.dec_calls.append(len(stream))

        def dec(body, n):
            r, out = self.s2_decode(body, n)
            return out if r == n else None
        try:
            return R.read_stream(bytes(stream), dec)
        except ValueError as e:
            raise {"corrupt": S.ErrCorrupt, "crc": S.ErrCRC, "unsupported": S.ErrUnsupported}[str(e)](str(e))

    def close(self):
        pass


def _data(n, seed=1):
    rng = np.random.default_rng(seed)
    tw = H.golden("twain.txt")
    return (tw * (n // len(tw) + 1))[:n // 2] + (rng.integers(0, 4, n - n // 2, dtype=np.uint8) + 48).astype(np.uint8).tobytes()


def test_stand_in_codec(oracle_lib):
    c = ModelCodec()
    st = c.EncodeStream(b"hello hello hello hello hello hello hello hello hello hello hello hello hello hello hello")
    assert c.DecodeStream(st).startswith(b"hello hello")


def test_writer_batches_index_padding(oracle_lib):
    data = _data(3 << 20)
    c = ModelCodec()
    sink = io.BytesIO()
    w = S.Writer(sink, codec=c, batch_bytes=1 << 20, add_index=True, padding=4096, rand=lambda n: b"\x00" * n)
    for o in range(0, len(data), 300001):
        w.Write(data[o:o + 300001])
    w.Close()
    out = sink.getvalue()
    assert c.enc_calls[:3] == [1 << 20] * 3 and sum(c.enc_calls) == len(data)      # one device call per full batch
    assert out.count(S.MAGIC_S2) == 1 and out.startswith(S.MAGIC_S2)
    assert len(out) % 4096 == 0
    assert c.DecodeStream(out) == data
    idx = X.Index(); idx.LoadStream(io.BytesIO(out))                                # the index is the last thing in the stream
    assert idx.TotalUncompressed == len(data) and idx.TotalCompressed == -1          # unknown with padding (s2/writer.go:810-813)
    assert idx.info[0] == (10, 0) and len(idx.info) == 3
    for cmp_off, u in idx.info:
        assert out[cmp_off] in (0, 1) and u % 65536 == 0
    assert X.read_range(out, 2500000, 100, c.DecodeStream, idx) == data[2500000:2500100]
    # without padding the index equals what IndexStream makes of the finished stream; CloseIndex returns it instead of appending
    sink2 = io.BytesIO()
    w = S.Writer(sink2, codec=c, batch_bytes=1 << 20)
    w.EncodeBuffer(data)
    ib = w.CloseIndex()
    assert sink2.getvalue() + ib == sink2.getvalue() + X.IndexStream(sink2.getvalue(), est_block=65536)
    with pytest.raises(S.B2CError):
        w.Write(b"x")
    # Flush cuts a short block; ReadFrom; Reset; Snappy + better flags reach the codec
    sink3 = io.BytesIO()
    w = S.Writer(sink3, codec=c, snappy=True, block_size=32768)
    w.Write(b"abc" * 1000); w.Flush(); assert w.ReadFrom(io.BytesIO(data[:100000])) == 100000; w.Close()
    assert sink3.getvalue().startswith(S.MAGIC_SNAPPY) and c.DecodeStream(sink3.getvalue()) == b"abc" * 1000 + data[:100000]
    w.Reset(io.BytesIO()); w.Close()
    with pytest.raises(S.ErrUnsupported):
        S.Writer(io.BytesIO(), codec=c, block_size=1 << 20)


class Dribble:
    def __init__(self, data, k):
        self.d, self.k, self.p = data, k, 0

    def read(self, n=-1):
        n = self.k if n < 0 else min(n, self.k)
        out = self.d[self.p:self.p + n]
        self.p += len(out)
        return out


@pytest.mark.parametrize("piece,batch", [(13, 1 << 20), (70000, 100000), (1 << 22, 1 << 22)])
def test_reader_batches_skip_and_concat(oracle_lib, piece, batch):
    data = _data(1 << 20, seed=2)
    c = ModelCodec()
    st = c.EncodeStream(data) + X.IndexStream(c.EncodeStream(data))
    st2 = c.EncodeStream(data[:99999], snappy=True)
    c.dec_calls.clear()
    r = S.Reader(Dribble(st + st2, piece), codec=c, batch_bytes=batch, read_size=piece)
    assert r.read(1000) == data[:1000]
    r.Skip(300000)                                          # whole blocks inside are never decoded
    assert r.read(5) == data[301000:301005]
    rest = r.read()
    assert rest == data[301005:] + data[:99999]
    assert r.read() == b""
    if batch >= 1 << 22:
        assert len(c.dec_calls) == 2                        # one call per stream (the identifier ends a batch)
    out = io.BytesIO()
    r.Reset(io.BytesIO(st))
    assert r.DecodeConcurrent(out) == len(data) and out.getvalue() == data
    r.Reset(io.BytesIO(st))
    with pytest.raises(S.ErrCorrupt):
        r.Skip(len(data) + 1)


def test_reader_errors(oracle_lib):
    data = _data(400000, seed=3)
    c = ModelCodec()
    st = c.EncodeStream(data)
    chunks = list(S._walk_chunks(st, 0, len(st)))
    # damage the payload of the 4th data chunk: the three blocks before it are delivered, then the error
    typ, start, ln, d = chunks[4]
    bad = bytearray(st); bad[start + 8 + 20] ^= 0xFF
    r = S.Reader(io.BytesIO(bytes(bad)), codec=c)
    got = r.read()
    assert got == data[:3 * 65536]
    with pytest.raises((S.ErrCorrupt, S.ErrCRC)):
        r.read()
    with pytest.raises((S.ErrCorrupt, S.ErrCRC)):
        r.read()
    # truncated stream
    r = S.Reader(io.BytesIO(st[:-7]), codec=c)
    assert r.read() == data[:len(data) // 65536 * 65536]
    with pytest.raises(S.ErrCorrupt):
        r.read()
    # no identifier
    with pytest.raises(S.ErrCorrupt):
        S.Reader(io.BytesIO(st[10:]), codec=c).read()
    assert S.Reader(io.BytesIO(st[10:]), codec=c, ignore_stream_identifier=True).read() == data
    # reserved unskippable chunk
    with pytest.raises(S.ErrUnsupported):
        S.Reader(io.BytesIO(S.MAGIC_S2 + b"\x05\x04\x00\x00abcd"), codec=c).read()
    assert S.Reader(io.BytesIO(b""), codec=c).read() == b""


def test_padding_helpers():
    for written in (0, 1, 3, 4, 5, 100, 1023, 1024):
        for mult in (1, 2, 4, 8, 512, 1024):
            add = S.calcSkippableFrame(written, mult)
            assert (written + add) % mult == 0 and (add == 0 or add >= 4)
    assert S.skippableFrame(0) == b"" and S.skippableFrame(6, lambda n: b"ab") == b"\xfe\x02\x00\x00ab"
    with pytest.raises(ValueError):
        S.skippableFrame(3)


@pytest.mark.parametrize("better", [False, True])
def test_framing_format(oracle_lib, better):
    """TestFramingFormat / TestFramingFormatBetter (s2/s2_test.go:755-826): 1e6 bytes alternating 1e5 of noise and 1e5 of one
    byte value -- larger than a block -- through Writer.Write + Close and back through the Reader."""
    rng = np.random.default_rng(1)
    src = bytearray(1000000)
    for i in range(10):
        src[100000 * i:100000 * (i + 1)] = bytes(rng.integers(0, 256, 100000, dtype=np.uint8)) if i % 2 == 0 else bytes([i]) * 100000
    src = bytes(src)
    c = ModelCodec()
    buf = io.BytesIO()
    bw = S.Writer(buf, codec=c, better=better)
    assert bw.Write(src) == len(src)
    bw.Close()
    st = buf.getvalue()
    types = [t for t, _, _, _ in S._walk_chunks(st, 0, len(st))]
    assert 0x00 in types and 0x01 in types                   # both compressed and uncompressed chunks occur
    assert S.Reader(io.BytesIO(st), codec=c).read() == src


def test_snappy_package_forwarders(oracle_lib):
    """snappy.NewBufferedWriter / NewWriter / NewReader (snappy/encode.go:41-54, snappy/decode.go:53-55) forward to the S2
    classes in Snappy-compatible, better mode; the unbuffered writer leaves nothing pending after Write."""
    from compress_b200 import snappy
    c = ModelCodec()
    data = _data(300000, seed=9)
    buf = io.BytesIO()
    w = snappy.NewBufferedWriter(buf, codec=c)
    w.Write(data[:1000]); assert buf.getvalue() == b""      # buffered
    w.Write(data[1000:]); w.Close()
    assert buf.getvalue().startswith(S.MAGIC_SNAPPY)
    assert snappy.NewReader(io.BytesIO(buf.getvalue()), codec=c).read() == data
    buf2 = io.BytesIO()
    w = snappy.NewWriter(buf2, codec=c)
    w.Write(data[:1000])
    assert snappy.NewReader(io.BytesIO(buf2.getvalue()), codec=c).read() == data[:1000]   # already complete
    w.Write(data[1000:5000]); w.Close()
    assert snappy.NewReader(io.BytesIO(buf2.getvalue()), codec=c).read() == data[:5000]
    assert snappy.DecodedLen(b"\xe8\x07rest") == 1000 and snappy.MaxEncodedLen(0) >= 0


def test_reader_equals_whole_stream_decode_on_mutations(oracle_lib):
    """Cutting the input at chunk boundaries and decoding in batches must not change the verdict: for mutated streams the Reader
    (small batches, dribbled input) ends like one DecodeStream of the whole input -- same content, or an error."""
    from test_emu_decoder_mutations import _mutations
    rng = np.random.default_rng(77)
    c = ModelCodec()
    data = _data(200000, seed=5)
    streams = [c.EncodeStream(data), c.EncodeStream(data[:70000], snappy=True) + c.EncodeStream(data[:5]),
               c.EncodeStream(data[:3000], block_size=4096)]
    ok = bad = 0
    for st in streams:
        for m in [st] + _mutations(st, rng, 60):
            try:
                want = c.DecodeStream(m)
            except S.B2CError:
                want = None
            r = S.Reader(Dribble(m, 999), codec=c, batch_bytes=70000, read_size=999)
            got, failed = b"", False
            try:
                while True:
                    part = r.read(50000)
                    if not part:
                        break
                    got += part
            except S.B2CError:
                failed = True
            if want is None:
                assert failed
                bad += 1
            else:
                assert not failed and got == want
                ok += 1
    assert ok >= 3 and bad > 50                         # (the checksums catch nearly every mutation)


def test_read_seeker(oracle_lib):
    """TestSeeking (s2/index_test.go:105-300): random seeks with an index on a seekable input (from the stream's end or passed
    separately), forward-only seeking without one, ReadAt, seeks relative to the position and the end."""
    rng = np.random.default_rng(11)
    c = ModelCodec()
    data = _data(4 << 20, seed=7)
    sink = io.BytesIO()
    w = S.Writer(sink, codec=c, add_index=True)
    w.Write(data); w.Close()
    st = sink.getvalue()
    plain = c.EncodeStream(data)
    idx_bytes = X.IndexStream(plain)
    for stream, index in ((st, None), (plain, idx_bytes)):
        rs = S.Reader(io.BytesIO(stream), codec=c).ReadSeeker(random=True, index=index)
        for _ in range(25):
            off = int(rng.integers(0, len(data)))
            assert rs.Seek(off, 0) == off
            n = int(rng.integers(1, 5000))
            assert rs.read(n) == data[off:off + n]
        assert rs.Seek(-100, 2) == len(data) - 100 and rs.read() == data[-100:]
        assert rs.Seek(1000, 0) == 1000 and rs.read(10) == data[1000:1010]
        assert rs.Seek(5, 1) == 1015 and rs.read(10) == data[1015:1025]          # relative to the position
        assert rs.Seek(-1025, 1) == 0 and rs.read(3) == data[:3]                 # backwards through the index
        assert rs.ReadAt(100, 3 << 20) == data[3 << 20:(3 << 20) + 100]
        assert rs.ReadAt(100, len(data) - 10) == data[-10:]
        with pytest.raises(ValueError):
            rs.Seek(-1, 0)
        with pytest.raises(S.ErrCorrupt):
            rs.Seek(len(data) + 1, 0)
    # seeks decode only what they need: far less input than the stream is handed to the codec
    c.dec_calls.clear()
    rs = S.Reader(io.BytesIO(st), codec=c, batch_bytes=1 << 20).ReadSeeker(random=True)
    rs.Seek(3500000, 0); rs.read(10)
    assert sum(c.dec_calls) < len(st) // 2
    # no index: forward only
    with pytest.raises(S.ErrCantSeek):
        S.Reader(io.BytesIO(plain), codec=c).ReadSeeker(random=True)
    rs = S.Reader(io.BytesIO(plain), codec=c).ReadSeeker(random=False)
    assert rs.Seek(100000, 0) == 100000 and rs.read(5) == data[100000:100005]
    with pytest.raises(S.ErrUnsupported):
        rs.Seek(10, 0)
    with pytest.raises(S.ErrUnsupported):
        rs.Seek(-1, 2)
    # an input that cannot seek
    with pytest.raises(S.ErrCantSeek):
        S.Reader(Dribble(st, 1000), codec=c).ReadSeeker(random=True)
    rs = S.Reader(Dribble(st, 100000), codec=c).ReadSeeker(random=False, index=idx_bytes)
    assert rs.Seek(200000, 0) == 200000 and rs.read(4) == data[200000:200004]
    with pytest.raises(S.ErrCantSeek):
        S.Reader(io.BytesIO(plain), codec=c).ReadSeeker(index=b"\x99garbage-that-is-long-enough-to-parse")
