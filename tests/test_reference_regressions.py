"""The reference's own ENCODER regression inputs pushed through every encoder of this repo.

  zstd/testdata/comp-crashers.zip  1657 inputs that once broke a zstd encoder (TestEncoderRegression,
                                   zstd/encoder_test.go:218-275: encode at every level, decode, compare)
  s2/testdata/enc_regressions.zip  51 inputs for the S2 encoders (TestEncoderRegression, s2/s2_test.go:2133-2200:
                                   Encode / EncodeBetter / EncodeSnappy, decode, compare, MaxEncodedLen respected)

CPU part: the oracle restatements at levels 1-3 / S2 modes, and the kernels under the SIMT emulator on the small
entries.  GPU part (-m gpu): every entry, cut into blocks of the level's size, through the C ABI; decoded by the oracle
decoder and libzstd (zstd) or the oracle's s2Decode and pyarrow (Snappy)."""
import os
import zipfile

import numpy as np
import pytest

import helpers as H
from test_oracle_s2 import s2_decode as orc_s2_decode, s2_encode as orc_s2_encode, _L as _s2L


def _entries(name):
    zf = zipfile.ZipFile(os.path.join(H.GOLDEN, name))
    return [(nm, zf.read(nm)) for nm in zf.namelist() if not nm.endswith("/")]


def test_fixture_counts():
    assert len(_entries("zstd_comp_crashers.zip")) == 1657
    assert len(_entries("s2_enc_regressions.zip")) == 51


@pytest.mark.parametrize("level", [1, 2, 3])
def test_oracle_zstd_comp_crashers(oracle_lib, level):
    for nm, data in _entries("zstd_comp_crashers.zip"):
        r, enc = H.oracle_encode(data, level=level)
        assert 0 < r <= oracle_lib.orc_zstd_max_encoded_size(len(data), level, 1), nm
        n, dec = H.oracle_decode(enc, len(data) + 16)
        assert n == len(data) and dec == data, nm
        assert H.libzstd_decode(enc, max(len(data), 1)) == data, nm


def test_oracle_s2_enc_regressions(oracle_lib):
    L = _s2L()
    for nm, data in _entries("s2_enc_regressions.zip"):
        for mode in (0, 1, 2):          # s2.Encode, EncodeBetter, EncodeSnappy
            enc = orc_s2_encode(data, mode)
            assert len(enc) <= L.orc_s2_max_encoded_len(len(data)), (nm, mode)
            n, dec = orc_s2_decode(enc, len(data))
            assert n == len(data) and dec == data, (nm, mode)


def _blocks_of(entries, block):
    out = []
    for _, data in entries:
        out += [data[i:i + block] for i in range(0, max(len(data), 1), block)]
    return out


@pytest.mark.parametrize("level", [1, 2])
def test_emu_zstd_comp_crashers_small(emu_lib, level):
    """The kernels' sources under the SIMT emulator on the small entries (< 512 bytes: all but three); the three large ones
    run on the device."""
    from emu_util import emu_encode
    chunks = [d for _, d in _entries("zstd_comp_crashers.zip") if len(d) < 512]
    assert len(chunks) >= 1650
    frames, outs, _, _, _ = emu_encode(emu_lib, chunks, level=level, seq_cap=256)
    for c, f, r in zip(chunks, frames, outs):
        assert r == len(f) > 0
        n, dec = H.oracle_decode(f, len(c) + 16)
        assert n == len(c) and dec == c
        assert H.libzstd_decode(f, max(len(c), 1)) == c


def test_emu_s2_enc_regressions_small(emu_lib):
    from emu_util import emu_s2_encode
    blocks = _blocks_of([e for e in _entries("s2_enc_regressions.zip") if len(e[1]) <= 65536 + 13], 65536)
    for better in (False, True):
        for snappy in (False, True):
            enc, _ = emu_s2_encode(emu_lib, blocks, snappy=snappy, better=better)
            for b, e in zip(blocks, enc):
                n, dec = orc_s2_decode(e, len(b))
                assert n == len(b) and dec == b


@pytest.mark.gpu
@pytest.mark.parametrize("level", [1, 2])
def test_gpu_zstd_comp_crashers(level):
    from compress_b200 import zstd
    enc = zstd.Encoder(level=level, max_chunks=512)
    entries = _entries("zstd_comp_crashers.zip")
    chunks = _blocks_of(entries, enc.block)
    frames = enc.encode_chunks(chunks)
    for c, f in zip(chunks, frames):
        assert 0 < len(f) <= enc.MaxEncodedSize(len(c))
        n, dec = H.oracle_decode(f, len(c) + 16)
        assert n == len(c) and dec == c
        assert H.libzstd_decode(f, max(len(c), 1)) == c
    # whole entries through EncodeAll: concatenated frames decode to the entry
    for nm, data in entries[-40:] + sorted(entries, key=lambda e: -len(e[1]))[:3]:
        out = enc.EncodeAll(data)
        n, dec = H.oracle_decode(out, len(data) + 16)
        assert n == len(data) and dec == data, nm
    enc.close()


@pytest.mark.gpu
def test_gpu_s2_enc_regressions():
    from compress_b200 import s2
    pa = pytest.importorskip("pyarrow")
    pc = pa.Codec("snappy")
    codec = s2.Codec()
    blocks = _blocks_of(_entries("s2_enc_regressions.zip"), 65536)
    for better in (False, True):
        for snappy in (False, True):
            enc = codec.encode_blocks(blocks, snappy=snappy, better=better)
            for b, e in zip(blocks, enc):
                assert 0 < len(e) <= s2.MaxEncodedLen(len(b))
                n, dec = orc_s2_decode(e, len(b))
                assert n == len(b) and dec == b
                if snappy and len(b):
                    assert pc.decompress(e, decompressed_size=len(b)).to_pybytes() == b
    codec.close()
