import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_lib():
    import helpers
    helpers.build_oracle()
    return helpers.oracle()


@pytest.fixture(scope="session")
def emu_lib():
    import ctypes
    import helpers
    helpers.build_emu()
    E = ctypes.CDLL(helpers.EMU_SO)
    c = ctypes
    E.emu_zstd_encode.argtypes = [c.c_void_p, c.c_uint64, c.c_void_p, c.c_uint32, c.c_void_p, c.c_uint64, c.c_void_p,
                                  c.c_uint32, c.c_void_p, c.c_void_p, c.c_void_p, c.c_uint32]
    E.emu_zstd_decode.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p, c.c_uint32, c.c_void_p, c.c_void_p, c.c_void_p,
                                  c.c_void_p]
    E.emu_s2_encode.argtypes = [c.c_void_p, c.c_uint64, c.c_void_p, c.c_uint32, c.c_void_p, c.c_uint64, c.c_void_p, c.c_int]
    E.emu_s2_decode.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p, c.c_uint32, c.c_void_p, c.c_void_p, c.c_void_p,
                                c.c_void_p]
    E.emu_huf_compress.argtypes = [c.c_void_p, c.c_uint64, c.c_void_p, c.c_uint32, c.c_void_p, c.c_uint64, c.c_void_p, c.c_int]
    E.emu_huf_decompress.argtypes = [c.c_void_p, c.c_uint64, c.c_void_p, c.c_uint32, c.c_void_p, c.c_uint64, c.c_void_p,
                                     c.c_void_p, c.c_int]
    return E
