"""Host logic of the stream reader and the zip adapters with stand-in codecs: the frame walk
(frame_span) against every frame the golden vectors hold, the Reader's batching / error order / truncation behaviour, the
Writer -> Reader pipe, and zip archives written and read through Python's zipfile with method 93.  The stand-ins (libzstd as
the encoder, the decoder oracle as the decoder) exist only here: the product classes construct the GPU codecs.  CPU only."""
import io
import os
import zipfile

import numpy as np
import pytest

import helpers as H
from compress_b200 import zstd as Z
from compress_b200 import zstd_zip as ZZ


class OracleDecoder:
    """decode_chunks of compress_b200.zstd.Decoder, served by the decoder oracle."""

    def __init__(self):
        self.calls = []

    def decode_chunks(self, streams, caps=None):
        self.calls.append(len(streams))
        outs, codes = [], []
        for s, cap in zip(streams, caps):
            r, got = H.oracle_decode(bytes(s), cap)
            codes.append(int(r)); outs.append(got if r >= 0 else None)
        return outs, codes

    def close(self):
        pass


class LibzstdEncoder:
    """encode_frames of compress_b200.zstd.Encoder, served by libzstd (one frame per input)."""

    def encode_frames(self, inputs):
        import ctypes
        L = H.libzstd()
        out = []
        for x in inputs:
            cap = L.ZSTD_compressBound(len(x))
            buf = ctypes.create_string_buffer(cap)
            r = L.ZSTD_compress(buf, cap, bytes(x), len(x), 1)
            out.append(buf.raw[:r])
        return out

    def close(self):
        pass


def _frames(oracle_lib):
    tw = H.golden("twain.txt")
    parts = [tw[:70000], b"", tw[70000:70010], bytes(5000), tw[100000:300000]]
    return parts, [H.oracle_encode(p, 1 + (i % 3))[1] for i, p in enumerate(parts)]


def test_frame_span_on_golden_frames(oracle_lib):
    zf = zipfile.ZipFile(os.path.join(H.GOLDEN, "zstd_decoder.zip"))
    seen = 0
    for nm in zf.namelist():
        if not nm.endswith(".zst"):
            continue
        comp = zf.read(nm)
        off = total = 0
        while off < len(comp):
            sp = Z.frame_span(comp, off)
            assert sp is not None, nm
            r, got = H.oracle_decode(comp[off:off + sp.length], 64 << 20)
            assert r >= 0, nm                               # exactly one frame was cut out
            if not sp.skippable:
                if sp.content_size is not None:
                    assert sp.content_size == r, nm
                assert sp.bound >= r, nm
            off += sp.length; total += r
            seen += 1
        assert off == len(comp)
        # any proper prefix is "need more input", never a wrong length
        for cut in (1, 3, 4, 5, 6, 11, len(comp) // 2, len(comp) - 1):
            if cut < Z.frame_span(comp, 0).length:
                assert Z.frame_span(comp[:cut], 0) is None, (nm, cut)
    assert seen >= 94


def test_frame_span_errors():
    with pytest.raises(Z.ZstdError) as e:
        Z.frame_span(b"\x00\x01\x02\x03\x04")
    assert e.value.code == Z.ErrMagicMismatch
    hdr = b"\x28\xb5\x2f\xfd"
    with pytest.raises(Z.ZstdError):
        Z.frame_span(hdr + b"\x08\x00" + bytes(8))          # reserved bit
    with pytest.raises(Z.ZstdError):
        Z.frame_span(hdr + b"\x20\x01" + b"\x07\x00\x00")    # reserved block type 3
    with pytest.raises(Z.ZstdError) as e:
        Z.frame_span(hdr + b"\x00\xf8" + bytes(8), max_window=1 << 20)
    assert e.value.code == Z.ErrWindowSizeExceeded
    sk = b"\x53\x2a\x4d\x18" + (5).to_bytes(4, "little") + b"abcde"
    sp = Z.frame_span(sk + hdr)
    assert sp.skippable and sp.length == 13
    assert Z.frame_span(sk[:10]) is None


class Dribble(io.RawIOBase):
    """An io.Reader that returns at most k bytes per read."""

    def __init__(self, data, k):
        self.d, self.k, self.p = data, k, 0

    def read(self, n=-1):
        n = self.k if n < 0 else min(n, self.k)
        out = self.d[self.p:self.p + n]
        self.p += len(out)
        return out


@pytest.mark.parametrize("piece,batch", [(7, 1 << 20), (4096, 3000), (1 << 20, 1 << 20), (1000, 1)])
def test_reader_batches_and_order(oracle_lib, piece, batch):
    parts, frames = _frames(oracle_lib)
    skip = b"\x50\x2a\x4d\x18" + (3).to_bytes(4, "little") + b"xyz"
    stream = frames[0] + skip + b"".join(frames[1:])
    dec = OracleDecoder()
    r = Z.Reader(Dribble(stream, piece), decoder=dec, batch_bytes=batch, read_size=piece)
    got = bytearray()
    while True:
        b = r.read(10000)
        if not b:
            break
        assert len(b) <= 10000
        got += b
    assert bytes(got) == b"".join(parts)
    assert r.frames == len(frames)
    if batch >= 1 << 20 and piece >= 1 << 20:
        assert dec.calls == [len(frames)]                    # the whole stream was one batch: one device call
    assert r.read() == b""
    r.Reset(io.BytesIO(frames[3]))
    assert r.read() == parts[3]
    assert Z.Reader(io.BytesIO(b""), decoder=dec).read() == b""


def test_reader_write_to_and_errors(oracle_lib):
    parts, frames = _frames(oracle_lib)
    dec = OracleDecoder()
    out = io.BytesIO()
    n = Z.Reader(io.BytesIO(b"".join(frames)), decoder=dec).WriteTo(out)
    assert n == sum(map(len, parts)) and out.getvalue() == b"".join(parts)
    # truncated inside the last frame: everything before it is delivered, then the error
    r = Z.Reader(io.BytesIO(b"".join(frames)[:-9]), decoder=dec)
    assert r.read() == b"".join(parts[:-1])
    with pytest.raises(Z.ZstdError):
        r.read()
    with pytest.raises(Z.ZstdError):
        r.read()                                             # errors stick (the reference keeps d.current.err)
    # garbage after a good frame
    r = Z.Reader(io.BytesIO(frames[0] + b"garbage!"), decoder=dec)
    assert r.read(100) == parts[0][:100]
    assert r.read() == parts[0][100:]
    with pytest.raises(Z.ZstdError) as e:
        r.read()
    assert e.value.code == Z.ErrMagicMismatch
    # a corrupted block inside the second frame: the decoder's code comes through, the first frame is delivered
    bad = bytearray(frames[0] + frames[4]); bad[len(frames[0]) + 40] ^= 0x55
    r = Z.Reader(io.BytesIO(bytes(bad)), decoder=dec)
    first = r.read(len(parts[0]))
    assert first == parts[0]
    with pytest.raises(Z.ZstdError):
        r.read()
    # frame larger than the per-frame limit
    with pytest.raises(Z.ZstdError) as e:
        Z.Reader(io.BytesIO(frames[4]), decoder=dec, max_frame=1000).read()
    assert e.value.code == Z.ErrDecoderSizeExceeded


def test_writer_reader_pipe(oracle_lib):
    rng = np.random.default_rng(5)
    data = H.golden("twain.txt") + bytes(rng.integers(0, 256, 100000, dtype=np.uint8))
    sink = io.BytesIO()
    w = Z.Writer(sink, encoder=LibzstdEncoder(), frame_bytes=50000)
    for o in range(0, len(data), 33333):
        w.Write(data[o:o + 33333])
    w.Flush()
    assert w.ReadFrom(io.BytesIO(b"tail")) == 4
    w.Close()
    stream = sink.getvalue()
    r = Z.Reader(io.BytesIO(stream), decoder=OracleDecoder())
    assert r.read() == data + b"tail"
    assert r.frames == len(data) // 50000 + 2
    assert H.libzstd_decode(stream, len(data) + 4) == data + b"tail"
    # push form
    r = Z.Reader(None, decoder=OracleDecoder())
    got = b"".join(r.Feed(stream[o:o + 7777]) for o in range(0, len(stream), 7777))
    assert got == data + b"tail" and r.Pending() == 0
    # Reset on the writer: a fresh stream, an empty one is still a frame
    sink2 = io.BytesIO()
    w = Z.Writer(sink, encoder=LibzstdEncoder()); w.Write(b"dropped"); w.Reset(sink2); w.Close()
    assert Z.Reader(io.BytesIO(sink2.getvalue()), decoder=OracleDecoder()).read() == b"" and len(sink2.getvalue()) > 4


def test_zip_adapters(oracle_lib, tmp_path):
    tw = H.golden("twain.txt")
    files = {"twain.txt": tw, "empty": b"", "small.bin": b"abc" * 10, "dir/html.txt": H.golden("html.txt")}
    enc, dec = LibzstdEncoder(), OracleDecoder()
    ZZ.Register(ZZ.ZipCompressor(encoder=enc), ZZ.ZipDecompressor(decoder=dec))
    try:
        ZZ.Register()                                        # idempotent
        path = str(tmp_path / "a.zip")
        with zipfile.ZipFile(path, "w", compression=ZZ.ZipMethodWinZip) as z:
            for k, v in files.items():
                z.writestr(k, v)
            z.writestr("stored", b"plain", compress_type=zipfile.ZIP_STORED)
            with z.open("streamed", "w") as f:               # entry written piecewise
                for o in range(0, len(tw), 50001):
                    f.write(tw[o:o + 50001])
        with zipfile.ZipFile(path) as z:
            assert z.testzip() is None                       # every CRC-32 holds
            for k, v in files.items():
                assert z.getinfo(k).compress_type == 93
                assert z.read(k) == v
            assert z.read("stored") == b"plain" and z.read("streamed") == tw
            with z.open("twain.txt") as f:                   # piecewise read
                assert f.read(1000) == tw[:1000] and f.read() == tw[1000:]
            info = z.getinfo("twain.txt")
            assert info.compress_size < len(tw) // 2
        # the entry's bytes are a plain zstd stream: libzstd reads them straight from the archive
        raw = open(path, "rb").read()
        with zipfile.ZipFile(path) as z:
            info = z.getinfo("twain.txt")
        nl, xl = int.from_bytes(raw[info.header_offset + 26:info.header_offset + 28], "little"), int.from_bytes(
            raw[info.header_offset + 28:info.header_offset + 30], "little")
        body = raw[info.header_offset + 30 + nl + xl:][:info.compress_size]
        assert H.libzstd_decode(body, len(tw)) == tw
        # an archive made by someone else: one libzstd level-19 frame as method 93
        other = str(tmp_path / "b.zip")
        with zipfile.ZipFile(other, "w", compression=zipfile.ZIP_DEFLATED) as z:
            z.writestr("d.txt", tw[:5000])
        with zipfile.ZipFile(other) as z:
            assert z.read("d.txt") == tw[:5000]              # other methods still go to zipfile's own codecs
    finally:
        ZZ.Unregister()
    with pytest.raises(NotImplementedError):
        zipfile.ZipFile(str(tmp_path / "c.zip"), "w", compression=93)


def test_header_decode_golden_vectors():
    """TestHeader_Decode (zstd/decodeheader_test.go:12): every entry of headers.zip gives exactly the Header of
    headers-want.json.zst through the host mirror's Header.Decode, or fails where the reference has no entry; AppendTo of a
    decoded header decodes to the same fields again."""
    import json
    golden = json.loads(H.libzstd_decode(open(os.path.join(H.GOLDEN, "zstd_headers-want.json.zst"), "rb").read(), 32 << 20))
    zh = zipfile.ZipFile(os.path.join(H.GOLDEN, "zstd_headers.zip"))
    ok = bad = 0
    for nm in zh.namelist():
        b = zh.read(nm)
        h = Z.Header()
        try:
            h.Decode(b)
        except (Z.ZstdError, Z.ErrUnexpectedEOFHeader):
            assert nm not in golden, nm
            bad += 1
            continue
        assert h.as_dict() == golden[nm], nm
        ok += 1
        if not h.Skippable and h.WindowSize <= 0xFFFFFFFF and not (h.SingleSegment and not h.HasFCS):
            again = Z.Header()
            again.Decode(h.AppendTo())
            for f in ("SingleSegment", "DictionaryID", "HasCheckSum"):
                assert getattr(again, f) == getattr(h, f), (nm, f)
            if h.SingleSegment or h.FrameContentSize >= 256:     # smaller sizes are not stored (zstd/frameenc.go:75-79)
                assert again.FrameContentSize == h.FrameContentSize, nm
            if not h.SingleSegment and h.WindowSize >= 1 << 10:
                assert again.WindowSize >= h.WindowSize // 2, nm   # the writer stores a power of two that covers it
    assert ok == len(golden) and ok + bad == len(zh.namelist()) and ok > 1000 and bad > 1000


def test_frame_header_and_padding(oracle_lib):
    # frame_header_bytes == the header of the oracle's frames (frameHeader.appendTo) for sizes around every form change
    for n in (0, 1, 255, 256, 257, 1024, 1025, 65535, 65536, 65791, 65792, 70000):
        data = bytes(n)
        frame = H.oracle_encode(data, 1)[1]
        h = Z.Header(); h.Decode(frame)
        assert frame.startswith(h.AppendTo()), n
        assert Z.frame_span(frame).content_size in (n, None)
    # calcSkippableFrame (zstd/frameenc.go:96-116): result is 0 or >= 8, and lands on the multiple
    for written in (0, 1, 7, 8, 100, 1023, 1024, 1025, 4095):
        for mult in (1, 2, 8, 16, 1000, 1024, 4096):
            add = Z.calcSkippableFrame(written, mult)
            assert (written + add) % mult == 0 and (add == 0 or add >= 8)
            if written % mult == 0:
                assert add == 0
            elif mult - written % mult >= 8:
                assert add == mult - written % mult
    with pytest.raises(ValueError):
        Z.calcSkippableFrame(10, 0)
    with pytest.raises(ValueError):
        Z.calcSkippableFrame(-1, 8)
    # skippableFrame: invisible to decoders
    frame = H.oracle_encode(b"hello world" * 50, 1)[1]
    add = Z.calcSkippableFrame(len(frame), 512)
    padded = Z.skippableFrame(frame, add)
    assert len(padded) == 512 and padded[len(frame):len(frame) + 4] == b"\x50\x2a\x4d\x18"
    assert H.libzstd_decode(padded, 1000) == b"hello world" * 50
    r, got = H.oracle_decode(padded, 1000)
    assert got == b"hello world" * 50
    sp = Z.frame_span(padded, len(frame))
    assert sp.skippable and sp.length == add
    assert Z.skippableFrame(b"x", 0) == b"x"
    with pytest.raises(ValueError):
        Z.skippableFrame(b"", 7)
    assert Z.skippableFrame(b"", 12, fill=lambda n: b"\x01" * n) == b"\x50\x2a\x4d\x18\x04\x00\x00\x00\x01\x01\x01\x01"


def test_writer_padding(oracle_lib):
    """WithEncoderPadding on the stream writer: the total written is a multiple, the padding a skippable frame."""
    enc = LibzstdEncoder(); enc.padding = 4096
    sink = io.BytesIO()
    w = Z.Writer(sink, encoder=enc, frame_bytes=30000)
    data = H.golden("twain.txt")[:100000]
    w.Write(data); w.Close()
    out = sink.getvalue()
    assert len(out) % 4096 == 0
    assert Z.Reader(io.BytesIO(out), decoder=OracleDecoder()).read() == data
    assert H.libzstd_decode(out, len(data)) == data
    h = Z.Header(); h.Decode(out[out.rindex(b"\x50\x2a\x4d\x18"):])
    assert h.Skippable and h.HeaderSize == 8


def test_reader_on_reference_fuzz_seeds(oracle_lib):
    """The stream reader's host walk on the reference's FuzzDecodeAll seeds (every 16th, tests/golden): with the decoder oracle
    behind it, Reader.read of a whole input ends the way DecodeAll of the same bytes does -- the same content, or an error
    where DecodeAll fails (frames whose declared or bounded size exceeds the limit are errors on both sides)."""
    zf = zipfile.ZipFile(os.path.join(H.GOLDEN, "zstd_fuzz_decode_subset.zip"))
    cap = 1 << 20
    ok = err = 0
    for nm in zf.namelist():
        b = zf.read(nm)
        ro, want = H.oracle_decode(b, cap)
        r = Z.Reader(io.BytesIO(b), decoder=OracleDecoder(), max_frame=cap, max_window=1 << 30)
        got, failed = b"", False
        try:
            while True:                      # (what precedes a bad frame is delivered first; the error comes with the next read)
                part = r.read()
                if not part:
                    break
                got += part
        except Z.B2CError:
            failed = True
        if ro >= 0:
            assert not failed and got == want, nm
            ok += 1
        else:
            assert failed, (nm, ro)
            err += 1
    assert ok > 3 and err > 300
