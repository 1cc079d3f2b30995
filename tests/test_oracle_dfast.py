"""Level 2 (doubleFastEncoder, zstd/enc_dfast.go) oracle restatement: SURVEY §8 row a-3.

No encoder golden vectors exist in the reference, so the restatement is pinned functionally: every frame
must decode to the input with the decoder oracle (itself pinned against the reference's decoder fixtures)
and with the system libzstd, and it must not be worse than the level-1 restatement on compressible input
(the reference's own ordering of SpeedFastest / SpeedDefault, zstd/encoder_options.go:163-190).
"""
import numpy as np
import pytest

from helpers import golden, libzstd_decode, oracle_decode, oracle_encode, synth_chunks, synth_text


def _roundtrip(data, crc=True):
    r, enc = oracle_encode(data, level=2, crc=crc)
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
def test_small_and_edge_sizes(n):
    data = synth_text(max(n, 1), seed=n + 1)[:n]
    _roundtrip(data)
    _roundtrip(bytes(n), crc=False)


def test_one_block_128k_text_beats_level1():
    data = synth_text(128 << 10, seed=3)
    e2 = _roundtrip(data)
    r1, e1 = oracle_encode(data, level=1)
    # level 1 cuts the input into 64 KiB blocks inside one frame; level 2 has longer search and 128 KiB blocks
    assert len(e2) < len(e1)
    assert len(e2) < 0.5 * len(data)


def test_multi_block_history():
    # > 128 KiB: the history variant (Encode, enc_dfast.go:38) with offsets carried across blocks
    base = synth_text(300 << 10, seed=9)
    data = base + base[1000:200000] + bytes(5000) + base[:70000]
    enc = _roundtrip(data)
    assert len(enc) < 0.45 * len(data)


@pytest.mark.parametrize("kind", ["random", "zeros"])
def test_incompressible_and_rle(kind):
    for size in (100, 65536, 131072, 200000):
        data = synth_chunks(kind, 1, size=size, seed=5)[0]
        enc = _roundtrip(data)
        if kind == "random":
            assert len(enc) <= len(data) + 3 * (size // (128 << 10) + 1) + 18
        else:
            assert len(enc) < 64


def test_repeats_and_long_matches():
    rng = np.random.Generator(np.random.PCG64(11))
    unit = rng.integers(0, 256, size=37, dtype=np.uint8).tobytes()
    data = unit * 4000 + synth_text(5000, seed=2) + unit * 100
    _roundtrip(data)
    # short period: offset-2 loop and repeat code paths
    data = (b"abcdefgh" * 50 + b"XY") * 300
    _roundtrip(data)


def test_reference_corpus_fixture():
    # a real text sample from the reference's test data (committed fixture)
    import io
    import zipfile
    zf = zipfile.ZipFile(io.BytesIO(golden("huff0_inputs.zip")))
    for name in zf.namelist():
        data = zf.read(name)
        if len(data) == 0:
            continue
        _roundtrip(data)
