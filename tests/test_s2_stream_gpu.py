"""GPU tests of the S2 / Snappy framing format: streams written by the device are read by the
reference-format reader model (tests/s2_stream_ref.py: chunk walk of s2/reader.go, masked CRC32-C of s2/s2.go:118-126,
blocks decoded by the oracle's s2Decode) and by the device's own reader; streams written by the model with the oracle's
block encoders -- including 1 MiB blocks, padding and skippable chunks, uncompressed chunks and Snappy streams -- are read by
the device; damaged streams give the reader's error classes."""
import numpy as np
import pytest
import torch

import helpers as H
import s2_stream_ref as R
from test_oracle_s2 import s2_decode as orc_s2_decode

pytestmark = pytest.mark.gpu


def _orc_decode_block(body, n):
    r, out = orc_s2_decode(body, n)
    return out if r == n else None


def _orc_encoder(better=False, snappy=False):
    """encodeBlock / encodeBlockBetter / encodeBlockSnappy of the oracle: the block body, b'' = not compressible."""
    import ctypes
    import test_oracle_s2 as T
    L = T._L()
    L.orc_s2_encode_block.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int64, ctypes.c_int]
    mode = 2 if snappy else (1 if better else 0)

    def enc(blk):
        out = ctypes.create_string_buffer(L.orc_s2_max_encoded_len(len(blk)) + 16)
        r = L.orc_s2_encode_block(out, bytes(blk), len(blk), mode)
        assert r >= 0
        return out.raw[:r]
    return enc


@pytest.fixture(scope="module")
def codec():
    from compress_b200 import s2
    c = s2.Codec()
    yield c
    c.close()


@pytest.mark.parametrize("better,snappy", [(False, False), (True, False), (False, True), (True, True)])
def test_device_streams_read_by_the_reference_format_reader(codec, better, snappy):
    tw = H.golden("twain.txt")
    rng = np.random.default_rng(3)
    for data in (b"", b"a", tw[:100], tw[:65536], tw[:65537], tw, rng.integers(0, 256, 200000, dtype=np.uint8).tobytes() + tw[:70000],
                 bytes(300000)):
        st = codec.EncodeStream(data, better=better, snappy=snappy)
        assert st[:10] == (R.MAGIC_SNAPPY if snappy else R.MAGIC_S2)
        assert R.read_stream(st, _orc_decode_block) == data
        assert len(st) <= codec_bound(len(data))
        assert codec.DecodeStream(st, max_size=len(data) + 64) == data
    # smaller blocks (WriterBlockSize)
    st = codec.EncodeStream(tw, better=better, snappy=snappy, block_size=4096)
    assert R.read_stream(st, _orc_decode_block) == tw and codec.DecodeStream(st, max_size=len(tw)) == tw


def codec_bound(n, block=65536):
    from compress_b200._lib import lib
    return int(lib.b2c_s2_stream_bound(n, block))


def test_incompressible_blocks_are_uncompressed_chunks(codec):
    rng = np.random.default_rng(9)
    data = rng.integers(0, 256, 3 * 65536 + 100, dtype=np.uint8).tobytes()
    st = codec.EncodeStream(data)
    # every chunk after the identifier is type 0x01 with the raw bytes
    o, types = 10, []
    while o < len(st):
        types.append(st[o]); ln = st[o + 1] | st[o + 2] << 8 | st[o + 3] << 16; o += 4 + ln
    assert types == [1, 1, 1, 1] and len(st) == 10 + 4 * 8 + len(data)
    assert R.read_stream(st, _orc_decode_block) == data


def test_model_streams_read_by_the_device(codec):
    tw = H.golden("twain.txt")
    rng = np.random.default_rng(4)
    mixed = tw[:150000] + rng.integers(0, 256, 70000, dtype=np.uint8).tobytes() + tw[150000:300000]
    for snappy in (False, True):
        enc = _orc_encoder(snappy=snappy)
        st = R.write_stream(mixed, enc, block_size=65536, snappy=snappy, extra_chunks=True)
        assert codec.DecodeStream(st, max_size=len(mixed)) == mixed
    # the reference writer's default 1 MiB blocks (copies with 4-byte offsets can appear)
    big = (tw * 3)[:1 << 20] + tw[:12345]
    st = R.write_stream(big, _orc_encoder(), block_size=1 << 20)
    assert codec.DecodeStream(st, max_size=len(big)) == big


def test_reader_errors(codec):
    from compress_b200 import s2
    tw = H.golden("twain.txt")
    st = bytearray(codec.EncodeStream(tw[:100000]))
    bad = bytearray(st); bad[10 + 4] ^= 1                      # checksum of the first chunk
    with pytest.raises(s2.ErrCRC):
        codec.DecodeStream(bytes(bad), max_size=200000)
    bad = bytearray(st); bad[10 + 8 + 20] ^= 0x55              # payload
    with pytest.raises((s2.ErrCRC, s2.ErrCorrupt)):
        codec.DecodeStream(bytes(bad), max_size=200000)
    with pytest.raises(s2.ErrCorrupt):
        codec.DecodeStream(bytes(st[10:]), max_size=200000)    # no stream identifier
    with pytest.raises(s2.ErrCorrupt):
        codec.DecodeStream(bytes(st[:-3]), max_size=200000)    # cut short
    with pytest.raises(s2.ErrUnsupported):
        codec.DecodeStream(bytes(st[:10]) + bytes([0x02, 1, 0, 0, 0]) + bytes(st[10:]), max_size=200000)
    for b_ in (bad,):
        for fn in (lambda: R.read_stream(bytes(b_), _orc_decode_block),):
            with pytest.raises(ValueError):
                fn()


def test_stream_device_one_gib_shape():
    """Device-resident stream of 64 MiB of synthetic text: total, error flag, decode through the model of a prefix."""
    from compress_b200 import s2
    c = s2.Codec()
    src = H.synth_text_torch(64 << 20, "cuda", seed=8)
    dst, total, err = c.encode_stream_device(src)
    torch.cuda.synchronize()
    assert int(err.item()) == 0
    n = int(total.cpu().numpy()[0])
    st = bytes(dst[:n].cpu().numpy())
    assert c.DecodeStream(st, max_size=64 << 20) == bytes(src.cpu().numpy())
    c.close()


def test_encode_any_size_is_one_block(codec):
    """s2.Encode of an input larger than the device's 64 KiB block: pieces encoded as one batch, joined with ConcatBlocks
    (s2/encode.go:322-361) -- one block the oracle's s2Decode and the device decode."""
    tw = H.golden("twain.txt")
    for fn in (codec.Encode, codec.EncodeBetter, codec.EncodeSnappy):
        blk = fn(tw)
        r, out = orc_s2_decode(blk, len(tw))
        assert r == len(tw) and out == tw
        assert codec.Decode(blk) == tw


def test_device_stream_equals_emulated_kernels(codec, emu_lib):
    from test_s2_stream_model import _emu_stream
    tw = H.golden("twain.txt")
    for data in (tw[:200001], tw[:65536], b"xy" * 40000):
        for snappy, better in ((False, False), (False, True), (True, False)):
            assert codec.EncodeStream(data, better=better, snappy=snappy) == _emu_stream(emu_lib, data, snappy=snappy, better=better)
