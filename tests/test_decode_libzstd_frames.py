"""Frames written by an independent encoder (the system libzstd, all its match finders and block kinds) must decode
to the source through the decoder oracle, through the kernels under the CPU SIMT emulator, and -- `-m gpu` --
through the C ABI on the device.

The reference's own decoder tests feed it third-party frames too (zstd/testdata/good.zip, decoder.zip); this widens
that to frames with multi-block history, long windows, repeat offsets, treeless / RLE / raw literal sections and
RLE sequence tables that the repo's own level-1 encoder never emits.
"""
import ctypes

import numpy as np
import pytest

import helpers as H
from emu_util import emu_decode

LEVELS = [-5, 1, 3, 6, 12, 19]


def _compress(data, level):
    Z = H.libzstd()
    cap = Z.ZSTD_compressBound(len(data))
    out = ctypes.create_string_buffer(cap)
    r = Z.ZSTD_compress(out, cap, bytes(data), len(data), level)
    assert not Z.ZSTD_isError(r)
    return out.raw[:r]


def _inputs():
    rng = np.random.Generator(np.random.PCG64(77))
    text = H.synth_text(300 << 10, seed=21)
    tw = H.golden("twain.txt")
    unit = rng.integers(0, 256, size=29, dtype=np.uint8).tobytes()
    few = rng.integers(0, 3, size=200_000, dtype=np.uint8).tobytes()          # tiny alphabet: 1X/4X + treeless
    return {
        "text300k": text,                                   # three blocks, offsets reaching back over block borders
        "twain": tw,
        "periodic": unit * 9000 + text[:5000] + unit * 50,  # very long matches, repeat offsets
        "zeros": bytes(150_000),                            # RLE blocks
        "random": rng.integers(0, 256, size=140_000, dtype=np.uint8).tobytes(),   # raw blocks
        "few_symbols": few,
        "tiny": b"abc",
        "empty": b"",
        "mixed": text[:70_000] + bytes(3000) + tw[:90_000] + few[:40_000] + text[10_000:60_000],
    }


@pytest.fixture(scope="module")
def frames():
    try:
        H.libzstd()
    except OSError:
        pytest.skip("no system libzstd")
    out = []
    for name, data in _inputs().items():
        for lv in LEVELS:
            out.append((f"{name}@{lv}", _compress(data, lv), data))
    return out


def test_oracle_decodes_libzstd_frames(frames, oracle_lib):
    for name, comp, want in frames:
        n, got = H.oracle_decode(comp, len(want))
        assert n == len(want) and got == want, name


def test_emulated_kernel_decodes_libzstd_frames(frames, emu_lib):
    # every stream is one warp of the decode kernel; both lane orders of the emulator
    small = [(n, c, w) for n, c, w in frames if len(w) <= 160_000]
    for desc in (0, 1):
        outs, res = emu_decode(emu_lib, [c for _, c, _ in small], [len(w) + 32 for _, _, w in small], desc=desc)
        for (name, _, want), code, got in zip(small, outs, res):
            assert code == len(want) and got == want, (name, int(code), desc)


@pytest.mark.gpu
def test_gpu_decodes_libzstd_frames(frames):
    from compress_b200 import zstd
    dec = zstd.Decoder()
    outs, codes = dec.decode_chunks([c for _, c, _ in frames], [len(w) + 32 for _, _, w in frames])
    for (name, _, want), got, code in zip(frames, outs, codes):
        assert code == len(want) and got == want, (name, code)
    # concatenated frames in one stream (Decoder.DecodeAll accepts several frames back to back, zstd/decoder.go:319)
    a, b = frames[0], frames[7]
    assert dec.DecodeAll(a[1] + b[1]) == a[2] + b[2]
    dec.close()
