"""The reference's `snappy` package (snappy/encode.go:20-56, snappy/decode.go:35-56): a drop-in for golang/snappy whose calls
forward to s2 in its Snappy-compatible modes -- here to the GPU codec's.  Block calls use one lazily created Codec per
process (the Go functions are stateless); pass `codec=` to share one."""
from . import s2 as _s2

_codec = None


def _default(codec):
    global _codec
    if codec is not None:
        return codec
    if _codec is None:
        _codec = _s2.Codec()
    return _codec


def Encode(src, codec=None):
    """snappy.Encode = s2.EncodeSnappyBetter (snappy/encode.go:20-22): a Snappy block any Snappy decoder reads."""
    return _default(codec).EncodeSnappyBetter(src)


def Decode(src, codec=None):
    """snappy.Decode = s2.Decode (snappy/decode.go:46-48)."""
    return _default(codec).Decode(src)


def DecodedLen(src):
    """snappy.DecodedLen (snappy/decode.go:35-37)."""
    return _s2._decoded_len(src)[0]


def MaxEncodedLen(src_len):
    """snappy.MaxEncodedLen (snappy/encode.go:28-30)."""
    return _s2.MaxEncodedLen(src_len)


def NewBufferedWriter(w, codec=None, **kw):
    """Stream writer in the Snappy framing format: s2.NewWriter(w, WriterSnappyCompat(), WriterBetterCompression())
    (snappy/encode.go:52-54)."""
    return _s2.Writer(w, codec=_default(codec), snappy=True, better=True, **kw)


def NewWriter(w, codec=None, **kw):
    """The deprecated unbuffered writer: as NewBufferedWriter with WriterFlushOnWrite (snappy/encode.go:41-43)."""
    return _s2.Writer(w, codec=_default(codec), snappy=True, better=True, flush_on_write=True, **kw)


def NewReader(r, codec=None, **kw):
    """Stream reader (snappy/decode.go:53-55; the 64 KiB block limit of Snappy streams is enforced by the stream decoder
    once it has seen the Snappy identifier)."""
    return _s2.Reader(r, codec=_default(codec), **kw)
