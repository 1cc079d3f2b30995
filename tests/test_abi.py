"""The C-ABI library loads and exports every symbol include/b2c.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "b2c.h")).read()
    return sorted(set(re.findall(r"B2C_API[^;(]*?\b(b2c_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    so = os.path.join(ROOT, "compress_b200", "_lib", "libb200comp.so")
    if not os.path.exists(so):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(so)
    names = _declared()
    assert len(names) >= 10
    for nm in names:
        assert hasattr(lib, nm), nm


def test_binding_matches_header():
    from compress_b200 import _lib
    assert sorted(_lib.EXPORTED_SYMBOLS) == _declared()


def test_no_device_fails_loudly():
    import torch
    from compress_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert _lib.lib.b2c_device_count() == 0
    assert not _lib.lib.b2c_ctx_create(0, 16)
    from compress_b200 import zstd, s2, huff0
    for ctor in (zstd.Encoder, zstd.Decoder, s2.Codec, huff0.Codec):
        with pytest.raises(_lib.B2CError):
            ctor()
    # the entry points themselves refuse to run without a context (no silent CPU path)
    assert _lib.lib.b2c_zstd_encode_device(None, 1, 3, None, 0, None, 0, None, 0, None, 1, None) == -100
    assert _lib.lib.b2c_s2_decode_device(None, None, 0, None, None, None, 0, None, 0, None, 1, None) == -100
    assert _lib.lib.b2c_huf_compress_device(None, 1, None, 0, None, 0, None, 0, None, 1, None) == -100


def test_bound_matches_reference_formula():
    # Encoder.MaxEncodedSize (zstd/encoder.go:843-873) vs the oracle's restatement
    import helpers as H
    from compress_b200 import _lib
    L = H.oracle()
    for n in [0, 1, 255, 256, 1000, 65535, 65536]:
        assert _lib.lib.b2c_zstd_bound(n, 1) == L.orc_zstd_max_encoded_size(n, 1, 1)


def test_s2_bound_matches_reference_formula():
    # s2.MaxEncodedLen (s2/encode.go:389-418) vs the oracle's restatement
    import ctypes as c
    import helpers as H
    from compress_b200 import _lib
    L = H.oracle()
    L.orc_s2_max_encoded_len.restype = c.c_int64
    L.orc_s2_max_encoded_len.argtypes = [c.c_int64]
    for n in [0, 1, 59, 60, 61, 255, 256, 257, 65535, 65536, 1 << 24, (1 << 32) - 11]:
        assert _lib.lib.b2c_s2_bound(n) == L.orc_s2_max_encoded_len(n), n
    assert _lib.lib.b2c_s2_bound((1 << 32) - 5) == 0 and L.orc_s2_max_encoded_len((1 << 32) - 5) == -1
