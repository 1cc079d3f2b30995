"""The tile-ordered parse (compress_b200/csrc/b2c_lz.cuh: zstd levels 1 and 2) under the SIMT emulator.

SURVEY section 8 rows a-2 (fastEncoder), a-3 (doubleFastEncoder), a-4 (betterFastEncoder), a-5/a-6: the GPU match finders are a different,
deterministic, parallel parse, so parity is layered: the entropy stage is byte-identical to the
oracle's blockEnc.encode for the same parse (raw / RLE decisions included), every frame decodes with the pinned
decoder oracle and libzstd, sizes respect MaxEncodedSize, output size is within +3 % of the reference algorithm at the
same level and block size, and the bytes do not depend on lane scheduling.  CPU only."""
import numpy as np
import pytest

import helpers as H
from check_util import check_frames
from emu_util import emu_encode
from test_emu_encoder_random import _structured


def _edge(block):
    rng = np.random.Generator(np.random.PCG64(7))
    tw = H.golden("twain.txt")
    return [b"", b"a", b"abcdefgh", b"a" * 9, b"a" * 12, b"a" * 13, b"a" * 16, b"ab" * 9, b"a" * 100, b"a" * block,
            bytes(range(256)) * 16, tw[:300], tw[:1023], tw[:1024], tw[:1025], tw[1000:1000 + 4097],
            rng.integers(0, 256, 3000, dtype=np.uint8).tobytes(), rng.integers(0, 3, 5000, dtype=np.uint8).tobytes(),
            b"abcd" * 5000, b"0123456789" * 300, tw[:block - 1], tw[:block - 15], bytes(block)]


@pytest.mark.parametrize("level", [1, 2, 3])
def test_lz_edge_cases(emu_lib, level):
    chunks = _edge(65536 if level == 1 else 131072)
    frames, outs, hdr, seqs, lits = emu_encode(emu_lib, chunks, level=level)
    assert (outs > 0).all()
    check_frames(chunks, frames, hdr, seqs, lits, label="lz-edge-L%d" % level, level=level)


@pytest.mark.parametrize("level", [1, 2, 3])
def test_lz_corpora_ratio(emu_lib, level):
    """Ratio tolerance against the reference algorithm (oracle restatement) at the same level and chunking: <= +3 %
    per corpus -- Twain, HTML, e.txt, synthetic text (no corpus is excluded)."""
    block = 65536 if level == 1 else 131072
    tw = H.golden("twain.txt")
    corp = {"twain": [tw[i:i + block] for i in range(0, 2 * block, block)], "html": [H.golden("html.txt")],
            "e": [H.golden("e.txt")[:block]], "synth": [H.synth_text(block, 3)]}
    for name, chunks in corp.items():
        frames, outs, hdr, seqs, lits = emu_encode(emu_lib, chunks, level=level)
        check_frames(chunks, frames, hdr, seqs, lits, label="lz-%s-L%d" % (name, level), level=level)
        ref = sum(H.oracle_encode(c, level=level)[0] for c in chunks)
        got = sum(len(f) for f in frames)
        assert got <= ref * 1.03, (name, level, got, ref)


@pytest.mark.parametrize("level", [1, 2, 3])
def test_lz_deterministic_lane_order(emu_lib, level):
    tw = H.golden("twain.txt")
    chunks = [tw[200000:200000 + 20000], b"xyz" * 3000, H.synth_text(30000, 11), bytes(5000) + tw[:3000] + bytes(70)]
    a = emu_encode(emu_lib, chunks, desc=0, level=level)[0]
    b = emu_encode(emu_lib, chunks, desc=1, level=level)[0]
    assert a == b


@pytest.mark.parametrize("level", [1, 2])
def test_lz_random_structures(emu_lib, level):
    rng = np.random.default_rng(4242 + level)
    block = 65536 if level == 1 else 131072
    sizes = [block, block - 1, block - 15, 40000, 8191, 4097, 1000, 333, 129, 128, 127, 33, 9, 8, 7]
    chunks = [_structured(rng, n) for n in sizes] + [_structured(rng, block) for _ in range(2)]
    frames, outs, hdr, seqs, lits = emu_encode(emu_lib, chunks, level=level)
    check_frames(chunks, frames, hdr, seqs, lits, label="lz-rand-L%d" % level, level=level)


def test_lz_too_big_chunk_is_reported(emu_lib):
    frames, outs, _, _, _ = emu_encode(emu_lib, [b"x" * 65537], level=1)
    assert outs[0] == -3
