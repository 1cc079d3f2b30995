"""Stream reader and zip adapters on the device: the Writer's frames through Reader (pull and push
forms) and a zip archive with method 93 written and read back by Python's zipfile over the GPU codecs; libzstd reads the same
bytes.  (Named to run after the codec tests it builds on.)"""
import io
import zipfile

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def test_writer_reader_zip_on_device(tmp_path):
    from compress_b200 import zstd as Z
    from compress_b200 import zstd_zip as ZZ
    rng = np.random.default_rng(9)
    tw = H.golden("twain.txt")
    data = tw + bytes(rng.integers(0, 256, 70000, dtype=np.uint8)) + tw[:100000]
    for level in (Z.SpeedFastest, Z.SpeedDefault):
        sink = io.BytesIO()
        w = Z.Writer(sink, level=level)
        for o in range(0, len(data), 100003):
            w.Write(data[o:o + 100003])
        w.Close()
        stream = sink.getvalue()
        assert H.libzstd_decode(stream, len(data)) == data
        r = Z.Reader(io.BytesIO(stream), batch_bytes=1 << 30)
        assert r.read(12345) == data[:12345] and r.read() == data[12345:]
        assert r.frames >= 2
        r.Reset(io.BytesIO(stream[:-5]))
        got = r.read()
        assert data.startswith(got) and len(got) < len(data)
        with pytest.raises(Z.ZstdError):
            r.read()
        r.Reset(None)
        assert b"".join(r.Feed(stream[o:o + 50000]) for o in range(0, len(stream), 50000)) == data
        r.Close()
    ZZ.Register()
    try:
        path = str(tmp_path / "a.zip")
        with zipfile.ZipFile(path, "w", compression=ZZ.ZipMethodWinZip) as z:
            z.writestr("twain.txt", tw)
            z.writestr("empty", b"")
            z.writestr("mixed.bin", data)
        with zipfile.ZipFile(path) as z:
            assert z.testzip() is None
            assert z.read("twain.txt") == tw and z.read("empty") == b"" and z.read("mixed.bin") == data
            assert z.getinfo("twain.txt").compress_type == 93 and z.getinfo("twain.txt").compress_size < len(tw) // 2
    finally:
        ZZ.Unregister()
