"""GPU parity tests for the zstd decoder (staged kernels + the one-warp decoder for what they mark), through the C ABI
(libb200comp.so).
Golden vectors are the reference's own (zstd/testdata/*.zip, committed under tests/golden/)."""
import os
import zipfile

import numpy as np
import pytest
import torch

import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dec():
    from compress_b200 import zstd
    d = zstd.Decoder()
    yield d
    d.close()


def _pairs(zf):
    names = set(zf.namelist())
    for nm in sorted(names):
        if nm.endswith(".zst") and nm[:-4] in names:
            yield nm, zf.read(nm), zf.read(nm[:-4])


def test_decoder_zip_full(dec):
    # zstd/decoder_test.go:201-216 TestNewDecoder: all 94 pairs of the reference's decoder.zip
    items = list(_pairs(zipfile.ZipFile(os.path.join(H.GOLDEN, "zstd_decoder.zip"))))
    assert len(items) == 94
    outs, codes = dec.decode_chunks([c for _, c, _ in items], [len(w) + 64 for _, _, w in items])
    for (nm, _, want), got, code in zip(items, outs, codes):
        assert code == len(want) and got == want, (nm, code)


def test_good_and_bad_zip(dec, oracle_lib):
    # zstd/decoder_test.go:393 TestNewDecoderGood, :409 TestNewDecoderBad; the oracle gives the expected bytes
    zf = zipfile.ZipFile(os.path.join(H.GOLDEN, "zstd_good.zip"))
    comps = [zf.read(nm) for nm in zf.namelist() if nm.endswith(".zst")]
    wants = []
    for c in comps:
        r, got = H.oracle_decode(c, 64 << 20)
        assert r >= 0
        wants.append(got)
    outs, codes = dec.decode_chunks(comps, [len(w) + 64 for w in wants])
    for want, got, code in zip(wants, outs, codes):
        assert code == len(want) and got == want
    zf = zipfile.ZipFile(os.path.join(H.GOLDEN, "zstd_bad.zip"))
    bad = [zf.read(nm) for nm in zf.namelist() if nm.endswith(".zst")]
    assert len(bad) == 32
    outs, codes = dec.decode_chunks(bad, [4 << 20] * len(bad))
    assert all(c < 0 for c in codes), codes
    # same error class as the oracle (which restates the reference's checks)
    for c, code in zip(bad, codes):
        ro, _ = H.oracle_decode(c, 4 << 20)
        assert ro == code, (ro, code)


def test_decode_oracle_frames(dec, oracle_lib):
    rng = np.random.default_rng(7)
    srcs = [b"", b"a", b"abc" * 7, bytes(1000), bytes(rng.integers(0, 256, 3000, dtype=np.uint8)),
            H.golden("twain.txt"), H.golden("html.txt"), H.golden("e.txt"), H.synth_text(1 << 20),
            bytes(rng.integers(0, 4, 300000, dtype=np.uint8))]
    enc = lambda b: H.oracle_encode(b)[1]
    comps = [enc(s) for s in srcs]
    srcs.append(srcs[5][:5000] + srcs[6][:5000])
    comps.append(enc(srcs[5][:5000]) + b"\x50\x2a\x4d\x18\x03\x00\x00\x00xyz" + enc(srcs[6][:5000]))
    outs, codes = dec.decode_chunks(comps, [len(s) + 16 for s in srcs])
    for i, (s, got, code) in enumerate(zip(srcs, outs, codes)):
        assert code == len(s) and got == s, (i, code)
    # DecodeAll mirror + error values
    from compress_b200 import zstd
    assert dec.DecodeAll(comps[5]) == srcs[5]
    with pytest.raises(zstd.ZstdError) as ei:
        dec.DecodeAll(comps[5][:-1] + bytes([comps[5][-1] ^ 1]))
    assert ei.value.code == zstd.ErrCRCMismatch
    with pytest.raises(zstd.ZstdError) as ei:
        dec.DecodeAll(b"\x00\x01\x02\x03\x04\x05")
    assert ei.value.code == zstd.ErrMagicMismatch
    with pytest.raises(zstd.ZstdError):
        dec.DecodeAll(comps[5][:1000])
    with pytest.raises(zstd.ZstdError) as ei:
        dec.DecodeAll(comps[5], size_hint=1000)
    assert ei.value.code == zstd.ErrDecoderSizeExceeded


def test_roundtrip_device_256mib(dec):
    # encode -> decode on the device, 4096 chunks of 64 KiB of the bench workload + mixed corpora
    from compress_b200 import zstd
    e = zstd.Encoder(max_chunks=64)
    n = 4096
    src = H.synth_text_torch(n * 65536, "cuda")
    tw = np.frombuffer(H.golden("twain.txt")[:5 * 65536], dtype=np.uint8)
    src[:tw.size] = torch.from_numpy(tw.copy()).cuda()
    src[10 * 65536:11 * 65536] = 0
    src[11 * 65536:12 * 65536] = torch.randint(0, 256, (65536,), dtype=torch.uint8, device="cuda")
    frames, sizes = e.encode_device(src)
    torch.cuda.synchronize()
    assert int(sizes.min()) > 0
    out, osz = dec.decode_device(frames, sizes.to(torch.int32), src_stride=frames.stride(0), dst_cap=65536)
    torch.cuda.synchronize()
    assert bool((osz == 65536).all()), osz[osz != 65536][:10]
    assert torch.equal(out.view(-1), src)
    assert dec.staged_count(n) == n                      # every one of these frames is the staged kernels' work
    # too small destination -> every stream reports an error, nothing is written past the capacity
    out2 = torch.full((n, 1024 + 64), 0x5A, dtype=torch.uint8, device="cuda")
    _, osz2 = dec.decode_device(frames, sizes.to(torch.int32), src_stride=frames.stride(0), dst=out2, dst_cap=1024,
                                dst_stride=1024 + 64)
    torch.cuda.synchronize()
    assert bool((osz2 < 0).all())
    assert bool((out2[:, 1024:] == 0x5A).all())
    assert dec.staged_count(n) == 0                      # ... and none of these: the one-warp decoder names the error
    e.close()


def test_staged_and_onewarp_agree(dec, oracle_lib):
    """Levels 2 and 3 (128 KiB blocks), damaged frames and frames the staged kernels do not take (two frames in one input,
    more than four blocks): staged + one-warp results equal the one-warp decoder's alone (B2C_DEC=onewarp context)."""
    import os
    from compress_b200 import zstd
    rng = np.random.default_rng(5)
    base = H.synth_text(6 * 131072, seed=9)
    comps, wants = [], []
    for level in (zstd.SpeedDefault, zstd.SpeedBetterCompression):
        e = zstd.Encoder(level=level, max_chunks=16)
        for k in range(6):
            c = base[k * 131072:(k + 1) * 131072 - 17 * k]
            comps.append(e.EncodeAll(c)); wants.append(c)
        e.close()
    long_src = H.synth_text(700000, seed=3)
    r, long_frame = H.oracle_encode(long_src, level=2)          # one frame, six blocks
    comps.append(bytes(long_frame)); wants.append(long_src)
    comps.append(comps[0] + comps[1]); wants.append(wants[0] + wants[1])
    for k in range(12):                                          # damage
        b = bytearray(comps[k])
        pos = int(rng.integers(8, len(b)))
        b[pos] ^= 1 << int(rng.integers(0, 8))
        comps.append(bytes(b)); wants.append(None)
    caps = [len(w) + 32 if w is not None else 131072 + 32 for w in wants]
    outs, codes = dec.decode_chunks(comps, caps)
    os.environ["B2C_DEC"] = "onewarp"
    try:
        d1 = zstd.Decoder()
    finally:
        del os.environ["B2C_DEC"]
    outs1, codes1 = d1.decode_chunks(comps, caps)
    d1.close()
    assert list(codes) == list(codes1)
    for w, o, o1, c in zip(wants, outs, outs1, codes):
        if c >= 0:
            assert o == o1
        if w is not None:
            assert c == len(w) and o == w


def test_decode_all_splits_frames(dec, oracle_lib):
    """DecodeAll of a stream of many frames (what EncodeAll of a large input is): the host pre-scan hands every frame to
    its own warp and sizes the device buffer from the declared content sizes; results equal the oracle's, including
    skippable frames in between, an empty frame, a frame without a content size (serial fallback) and a corrupt frame."""
    from compress_b200 import zstd
    enc = zstd.Encoder(max_chunks=256)
    data = H.synth_text(40 * 65536 + 1234, seed=21)
    stream = enc.EncodeAll(data)                       # 41 frames
    assert dec.DecodeAll(stream, size_hint=len(data) + 64) == data
    skip = bytes([0x5A, 0x2A, 0x4D, 0x18, 5, 0, 0, 0]) + b"hello"
    empty = enc.EncodeAll(b"")
    mixed = skip + stream[:0] + empty + stream + skip
    assert dec.DecodeAll(mixed, size_hint=len(data) + 64) == data
    # too small a destination: the reference's "decoder size exceeded" class, not a partial result
    with pytest.raises(zstd.ZstdError):
        dec.DecodeAll(stream, size_hint=len(data) - 1)
    # a libzstd streaming-style frame without content size in the middle: whole input goes to one warp, same bytes
    import ctypes
    Z = H.libzstd()
    raw = data[:200000]
    r, of = H.oracle_encode(raw, level=1)
    nofcs = bytearray(of)
    outs, codes = dec.decode_chunks([stream + bytes(of)], [len(data) + len(raw) + 64])
    assert codes[0] == len(data) + len(raw) and outs[0] == data + raw
    # corrupt one frame in the middle: the first failing frame decides, as in a serial decode
    bad = bytearray(stream)
    bad[len(stream) // 2] ^= 0x55
    outs, codes = dec.decode_chunks([bytes(bad)], [len(data) + 64])
    ro, _ = H.oracle_decode(bytes(bad), len(data) + 64)
    assert codes[0] < 0 and ro < 0
    enc.close()
