"""zstd in zip archives: the mirror of zstd.ZipCompressor / zstd.ZipDecompressor (zstd/zip.go:117-141), compression method
93 (ZipMethodWinZip, zstd/zip.go:17; 20 = ZipMethodPKWare is accepted on reading).

The reference hands `archive/zip` a constructor pair through zip.RegisterCompressor / RegisterDecompressor.  Python's `zipfile`
has the same seam but no public registration call (it gains native method 93 in 3.14), so `Register()` installs the pair on
the three module-level hooks zipfile consults -- `_check_compression`, `_get_compressor`, `_get_decompressor` -- for methods
93 and 20 only; every other method goes to the original functions.  The objects handed out follow zipfile's own codec
protocol (`compress(data)` / `flush()`; `decompress(data)` / `eof`), and sit on the package's stream Writer / Reader, so the
bytes are produced and consumed by the GPU codecs: there is no CPU path here.

    from compress_b200 import zstd_zip
    zstd_zip.Register()
    with zipfile.ZipFile(path, "w", compression=zstd_zip.ZipMethodWinZip) as z: z.writestr("a.txt", data)
"""
from . import zstd as _zstd

ZipMethodWinZip = 93          # zstd/zip.go:17
ZipMethodPKWare = 20          # zstd/zip.go:21 (deprecated id; read only)


class _Sink:
    """The io.Writer the stream Writer emits frames into; drained by the zip codec object."""

    def __init__(self):
        self.buf = bytearray()

    def write(self, b):
        self.buf += b
        return len(b)

    def take(self):
        out = bytes(self.buf)
        self.buf.clear()
        return out


class _ZipWriter:
    """pooledZipWriter (zstd/zip.go:87-113) in zipfile's compressor protocol."""

    def __init__(self, make_writer):
        self._sink = _Sink()
        self._w = make_writer(self._sink)

    def compress(self, data):
        self._w.Write(bytes(data))
        return self._sink.take()

    def flush(self):
        self._w.Close()                   # the last frame (an empty entry is still one frame)
        return self._sink.take()


class _ZipReader:
    """pooledZipReader (zstd/zip.go:40-77) in zipfile's decompressor protocol.  zipfile feeds the entry's compressed bytes in
    pieces and stops once they are exhausted; complete frames are decoded as they arrive (Reader.Feed), so everything has
    been returned by the time the last piece is in."""

    def __init__(self, make_reader):
        self._r = make_reader(None)
        self.eof = False

    def decompress(self, data):
        return self._r.Feed(data)

    def flush(self):
        return b""


def ZipCompressor(level=_zstd.SpeedFastest, device=0, encoder=None):
    """-> constructor of zip entry compressors, all sharing one Encoder (the reference pools them, zstd/zip.go:117-133).
    Entries carry no frame checksum, as zip has its own CRC-32 (zstd/zip.go:119)."""
    shared = []

    def make():
        if not shared:
            shared.append(encoder if encoder is not None else _zstd.Encoder(level=level, crc=False, device=device, max_chunks=64))
        return _ZipWriter(lambda sink: _zstd.Writer(sink, level=level, crc=False, encoder=shared[0]))
    return make


def ZipDecompressor(device=0, max_window=128 << 20, decoder=None):
    """-> constructor of zip entry decompressors sharing one Decoder; 128 MiB window limit by default (zstd/zip.go:135-141)."""
    shared = []

    def make():
        if not shared:
            shared.append(decoder if decoder is not None else _zstd.Decoder(device=device))
        return _ZipReader(lambda src: _zstd.Reader(src, max_window=max_window, decoder=shared[0]))
    return make


_installed = {}


def Register(compressor=None, decompressor=None, module=None):
    """zip.RegisterCompressor(zstd.ZipMethodWinZip, zstd.ZipCompressor()) + zip.RegisterDecompressor(...) for Python's
    zipfile (or a module with the same three hooks).  Idempotent; Unregister() restores the original hooks."""
    import zipfile as _zipfile
    zf = module or _zipfile
    if zf in _installed:
        return
    for hook in ("_check_compression", "_get_compressor", "_get_decompressor"):
        if not hasattr(zf, hook):
            raise RuntimeError("zipfile has no %s hook in this Python" % hook)
    comp = compressor or ZipCompressor()
    decomp = decompressor or ZipDecompressor()
    orig = (zf._check_compression, zf._get_compressor, zf._get_decompressor)
    ours = (ZipMethodWinZip, ZipMethodPKWare)

    def check(compression):
        if compression in ours:
            return
        return orig[0](compression)

    def get_compressor(compress_type, compresslevel=None):
        if compress_type in ours:
            return comp()
        return orig[1](compress_type, compresslevel)

    def get_decompressor(compress_type):
        if compress_type in ours:
            return decomp()
        return orig[2](compress_type)

    zf._check_compression, zf._get_compressor, zf._get_decompressor = check, get_compressor, get_decompressor
    if hasattr(zf, "compressor_names"):
        zf.compressor_names.setdefault(ZipMethodWinZip, "zstd")
    _installed[zf] = orig


def Unregister(module=None):
    import zipfile as _zipfile
    zf = module or _zipfile
    orig = _installed.pop(zf, None)
    if orig:
        zf._check_compression, zf._get_compressor, zf._get_decompressor = orig
