"""Level 2 (doubleFastEncoder, zstd/enc_dfast.go) and level 3 (betterFastEncoder, zstd/enc_better.go) oracle
restatements: SURVEY §8 rows a-3 and a-4.

No encoder golden vectors exist in the reference, so the restatements are pinned functionally: every frame
must decode to the input with the decoder oracle (itself pinned against the reference's decoder fixtures)
and with the system libzstd, and the levels must order like the reference's (SpeedFastest > SpeedDefault >
SpeedBetterCompression in output size on compressible input, zstd/encoder_options.go:163-190).
"""
import numpy as np
import pytest

from helpers import golden, libzstd_decode, oracle_decode, oracle_encode, synth_chunks, synth_text


def _roundtrip(data, crc=True, level=2):
    r, enc = oracle_encode(data, level=level, crc=crc)
    assert r >= 0, f"encode failed {r}"
    n, dec = oracle_decode(enc, len(data))
    assert n == len(data) and dec == bytes(data)
    try:
        z = libzstd_decode(enc, len(data))
    except OSError:
        z = None
    else:
        assert z == bytes(data), "libzstd disagrees"
    return enc


@pytest.mark.parametrize("n", [0, 1, 7, 15, 16, 17, 25, 100, 1000, 1023, 1024, 1025, 4096, 65535, 65536, 65537])
@pytest.mark.parametrize("level", [2, 3])
def test_small_and_edge_sizes(n, level):
    data = synth_text(max(n, 1), seed=n + 1)[:n]
    _roundtrip(data, level=level)
    _roundtrip(bytes(n), crc=False, level=level)


def test_one_block_128k_text_level_order():
    data = synth_text(128 << 10, seed=3)
    e2 = _roundtrip(data)
    e3 = _roundtrip(data, level=3)
    r1, e1 = oracle_encode(data, level=1)
    # level 1 cuts the input into 64 KiB blocks inside one frame; levels 2/3 search harder over 128 KiB blocks
    assert len(e3) < len(e2) < len(e1)
    assert len(e2) < 0.5 * len(data)


@pytest.mark.parametrize("level", [2, 3])
def test_multi_block_history(level):
    # > 128 KiB: the history variant (Encode, enc_dfast.go:38 / enc_better.go:56) with offsets carried across blocks
    base = synth_text(300 << 10, seed=9)
    data = base + base[1000:200000] + bytes(5000) + base[:70000]
    enc = _roundtrip(data, level=level)
    assert len(enc) < 0.45 * len(data)
    if level == 3:
        # the long table of the level-3 finder reaches back over the whole 8 MiB window
        assert len(enc) < len(_roundtrip(data, level=2))


@pytest.mark.parametrize("level", [2, 3])
@pytest.mark.parametrize("kind", ["random", "zeros"])
def test_incompressible_and_rle(kind, level):
    for size in (100, 65536, 131072, 200000):
        data = synth_chunks(kind, 1, size=size, seed=5)[0]
        enc = _roundtrip(data, level=level)
        if kind == "random":
            assert len(enc) <= len(data) + 3 * (size // (128 << 10) + 1) + 18
        else:
            assert len(enc) < 64


@pytest.mark.parametrize("level", [2, 3])
def test_repeats_and_long_matches(level):
    rng = np.random.Generator(np.random.PCG64(11))
    unit = rng.integers(0, 256, size=37, dtype=np.uint8).tobytes()
    data = unit * 4000 + synth_text(5000, seed=2) + unit * 100
    _roundtrip(data, level=level)
    # short period: offset-2 loop and repeat code paths
    data = (b"abcdefgh" * 50 + b"XY") * 300
    _roundtrip(data, level=level)


@pytest.mark.parametrize("level", [2, 3])
def test_reference_corpus_fixture(level):
    # a real text sample from the reference's test data (committed fixture)
    import io
    import zipfile
    zf = zipfile.ZipFile(io.BytesIO(golden("huff0_inputs.zip")))
    for name in zf.namelist():
        data = zf.read(name)
        if len(data) == 0:
            continue
        _roundtrip(data, level=level)


def _lz_like(rng, n):
    """Bytes built from random literals and copies of earlier data at random distances (short and far)."""
    out = bytearray(rng.integers(0, 64, size=64, dtype=np.uint8).tobytes())
    while len(out) < n:
        k = int(rng.integers(0, 4))
        if k == 0:
            out += rng.integers(0, 256, size=int(rng.integers(1, 40)), dtype=np.uint8).tobytes()
        else:
            dist = int(rng.integers(1, min(len(out), 1 << int(rng.integers(1, 18))) + 1))
            ln = int(rng.integers(3, 300 if k == 3 else 24))
            start = len(out) - dist
            for i in range(ln):                    # overlapping copies allowed, as in the format
                out.append(out[start + i])
    return bytes(out[:n])


@pytest.mark.parametrize("level", [1, 2, 3])
@pytest.mark.parametrize("seed", range(6))
def test_random_lz_structured_inputs(level, seed):
    rng = np.random.Generator(np.random.PCG64(1000 + seed))
    n = int(rng.integers(1, 400_000))
    data = _lz_like(rng, n)
    enc = _roundtrip(data, level=level, crc=bool(seed & 1))
    assert len(enc) <= len(data) + 3 * (n // (64 << 10) + 1) + 18
