"""GPU parity tests for the standalone huff0 block coder through the C ABI: bytes equal to the oracle's
huff0.Compress4X / Compress1X, error values, decompress inverse (BASELINE config 4)."""
import numpy as np
import pytest
import torch

import helpers as H
from test_emu_huf0 import _blocks, orc_compress, orc_decompress

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def codec():
    from compress_b200 import huff0
    c = huff0.Codec()
    yield c
    c.close()


def test_huf_compress_parity(codec, oracle_lib):
    from compress_b200 import huff0
    blocks = _blocks()
    for four in (True, False):
        got = codec.compress_blocks(blocks, four)
        for i, (g, b) in enumerate(zip(got, blocks)):
            w = orc_compress(b, four)
            assert g[1] == w[1] and g[0] == w[0], (four, i, g[1], w[1])
    with pytest.raises(huff0.ErrUseRLE):
        codec.Compress4X(bytes(1000))
    with pytest.raises(huff0.ErrIncompressible):
        codec.Compress1X(bytes(np.random.default_rng(0).integers(0, 256, 5000, dtype=np.uint8)))
    with pytest.raises(huff0.ErrTooBig):
        codec.Compress4X(bytes(np.random.default_rng(0).integers(0, 4, 262144, dtype=np.uint8)))


def test_huf_decompress(codec, oracle_lib):
    from compress_b200 import huff0
    blocks = _blocks()
    for four in (True, False):
        ok = [(orc_compress(b, four)[0], b) for b in blocks if orc_compress(b, four)[1] > 0]
        got = codec.decompress_blocks([c for c, _ in ok], [len(b) for _, b in ok], four)
        for (g, code), (_, b) in zip(got, ok):
            assert code == len(b) and g == b
        c0, b0 = ok[-1]
        cases = [(c0, len(b0) - 1), (c0[:-1], len(b0)), (c0[:40], len(b0)), (bytes([200]) * 50, 100)]
        res = codec.decompress_blocks([c for c, _ in cases], [d for _, d in cases], four)
        for (g, code), (c, d) in zip(res, cases):
            wcode, wout = orc_decompress(c, d, four)
            assert (code < 0) == (wcode < 0)
            if code >= 0:
                assert g == wout
    tw = H.golden("twain.txt")[:100000]
    assert codec.Decompress4X(codec.Compress4X(tw), len(tw)) == tw
    assert codec.Decompress1X(codec.Compress1X(tw[:5000]), 5000) == tw[:5000]
    with pytest.raises(huff0.ErrCorrupt):
        codec.Decompress4X(codec.Compress4X(tw)[:-7], len(tw))


def test_huf_roundtrip_config4(codec):
    # 512 blocks of 262143 bytes of the bench text (BASELINE config 4 shape, 128 MiB)
    n, bs = 512, 262143
    stride = 262144
    src = H.synth_text_torch(n * stride, "cuda").view(n, stride)
    sizes = torch.full((n,), bs, dtype=torch.int32, device="cuda")
    comp, csz = codec.compress_device(src.view(-1), stride, sizes, four=True)
    torch.cuda.synchronize()
    assert int(csz.min()) > 0 and float(csz.sum()) / (n * bs) < 0.75
    out, osz = codec.decompress_device(comp.view(-1), comp.stride(0), csz.to(torch.int32), sizes, stride, four=True)
    torch.cuda.synchronize()
    assert bool((osz == bs).all())
    assert torch.equal(out[:, :bs], src[:, :bs])


def test_huf_reference_error_table(codec, oracle_lib):
    # huff0/compress_test.go:20-52: expected error class of Compress1X / Compress4X per input
    from test_reference_tables import HUF_TABLE, huf_inputs, _cls
    inputs = huf_inputs()
    names = list(inputs)
    for four in (False, True):
        got = codec.compress_blocks([inputs[n] for n in names], four)
        for n, (out, code) in zip(names, got):
            assert _cls(code) == HUF_TABLE[n][1 if four else 0], (n, four, code)
            assert (out, code) == orc_compress(inputs[n], four), n


def test_zstd_large_zeros_gpu():
    # zstd/testdata/large.zip (TestNewDecoderLarge): 100 KiB and 10 MiB of zeros
    import os
    import zipfile
    from compress_b200 import zstd
    dec = zstd.Decoder()
    zf = zipfile.ZipFile(os.path.join(H.GOLDEN, "zstd_large.zip"))
    for nm in zf.namelist():
        if nm.endswith(".zst"):
            size = int(zf.read(nm + ".size"))
            assert dec.DecodeAll(zf.read(nm), size_hint=size + 8) == bytes(size), nm
    dec.close()


def test_huf_host_buffer_calls_and_read_table(oracle_lib):
    """b2c_huf_compress_chunks / b2c_huf_decompress_chunks / b2c_huf_read_table (the calls a cgo shim binds for
    huff0.Compress4X, Decoder.Decompress4X and huff0.ReadTable) against the device-resident path and the oracle."""
    import ctypes
    from compress_b200 import huff0
    hc = huff0.Codec()
    rng = np.random.default_rng(9)
    blocks = [H.golden("twain.txt")[:100000], H.golden("e.txt")[:262143], bytes(rng.integers(0, 5, 50000, dtype=np.uint8)),
              H.golden("html.txt"), bytes(rng.integers(0, 256, 4000, dtype=np.uint8)), bytes(1000)]
    for four in (True, False):
        a = hc.compress_chunks(blocks, four)
        b = hc.compress_blocks(blocks, four)
        assert a == b
        good = [(blk, c) for blk, (c, code) in zip(blocks, a) if code > 0]
        back = hc.decompress_chunks([c for _, c in good], [len(blk) for blk, _ in good], four)
        assert [x[0] for x in back] == [blk for blk, _ in good]
    comp = hc.Compress4X(blocks[0])

    class DT(ctypes.Structure):
        _fields_ = [("dt", ctypes.c_uint16 * 2048), ("actualTableLog", ctypes.c_uint), ("loaded", ctypes.c_int)]
    L = oracle_lib
    L.orc_huf_read_table.restype = ctypes.c_int64
    L.orc_huf_read_table.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t]
    d = DT()
    used = L.orc_huf_read_table(ctypes.byref(d), comp, len(comp))
    nbits, tlog, rest = hc.ReadTable(comp)
    assert tlog == d.actualTableLog and rest == comp[used:]
    want = [0] * 256
    for e in d.dt[: 1 << d.actualTableLog]:
        want[e >> 8] = e & 0xff
    assert nbits == want
    with pytest.raises(huff0.ErrCorrupt):
        hc.ReadTable(comp[:2])
    hc.close()
