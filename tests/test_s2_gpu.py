"""GPU parity tests for the S2 / Snappy block codec through the C ABI (libb200comp.so)."""
import numpy as np
import pytest
import torch

import helpers as H
from s2_vectors import DECODE_TABLE, INVALID_VARINT
from test_oracle_s2 import s2_decode as orc_decode, s2_encode as orc_encode, _L
from test_emu_s2 import _blocks

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def codec():
    from compress_b200 import s2
    c = s2.Codec()
    yield c
    c.close()


def test_s2_encode_blocks(codec, oracle_lib):
    from compress_b200 import s2
    blocks = _blocks()
    L = _L()
    for better in (False, True):
        for snappy in (False, True):
            enc = codec.encode_blocks(blocks, snappy=snappy, better=better)
            for i, (b, c) in enumerate(zip(blocks, enc)):
                assert 0 < len(c) <= L.orc_s2_max_encoded_len(len(b)) == s2.MaxEncodedLen(len(b)), i
                n, got = orc_decode(c, len(b))
                assert n == len(b) and got == b, (better, snappy, i)
    with pytest.raises(s2.ErrTooLarge):
        codec.encode_blocks([bytes(65537)])           # one device block is at most 64 KiB ...
    big = codec.Encode(bytes(65537))                  # ... s2.Encode joins pieces into one block (ConcatBlocks)
    n, got = orc_decode(big, 65537)
    assert n == 65537 and got == bytes(65537)
    assert s2.MaxEncodedLen(0) == 1 and s2.MaxEncodedLen(0xffffffff) == -1


def test_s2_gpu_matches_emulator(codec, emu_lib):
    from emu_util import emu_s2_encode
    blocks = _blocks()
    for better in (False, True):
        for snappy in (False, True):
            emu, _ = emu_s2_encode(emu_lib, blocks, snappy=snappy, better=better)
            assert codec.encode_blocks(blocks, snappy=snappy, better=better) == emu


def test_s2_ratio_per_corpus(codec, oracle_lib):
    """Per corpus, each class is within +5 % of the reference algorithm of the same class (s2.Encode / EncodeBetter /
    EncodeSnappy restated in the oracle) -- BASELINE config 3 "+Better" included."""
    tw = H.golden("twain.txt")
    corp = {"twain": [tw[i:i + 65536] for i in range(0, len(tw), 65536)], "html": [H.golden("html.txt")],
            "e": [H.golden("e.txt")[:65536], H.golden("e.txt")[65536:]], "synth": H.synth_chunks("text", 8, seed=5)}
    for name, chunks in corp.items():
        for better, snappy, mode in ((False, False, 0), (True, False, 1), (False, True, 2), (True, True, 2)):
            ours = sum(len(x) for x in codec.encode_blocks(chunks, snappy=snappy, better=better))
            ref = sum(len(orc_encode(c, mode)) for c in chunks)
            assert ours <= 1.05 * ref, (name, better, snappy, ours, ref)


def test_snappy_interop(codec):
    pa = pytest.importorskip("pyarrow")
    pc = pa.Codec("snappy")
    blocks = [b for b in _blocks() if len(b)]
    for better in (False, True):
        enc = codec.encode_blocks(blocks, snappy=True, better=better)
        for b, c in zip(blocks, enc):
            assert pc.decompress(c, decompressed_size=len(b)).to_pybytes() == b
    # blocks written by an independent Snappy encoder decode on the GPU
    comp = [pc.compress(b).to_pybytes() for b in blocks]
    outs, codes = codec.decode_blocks(comp, [len(b) for b in blocks])
    assert outs == blocks and codes == [len(b) for b in blocks]


def test_s2_decode_vectors(codec, oracle_lib):
    from compress_b200 import s2
    ins = [v[0] for v in DECODE_TABLE] + INVALID_VARINT
    outs, codes = codec.decode_blocks(ins, [100] * len(ins))
    for i, (inp, want, ok) in enumerate(DECODE_TABLE):
        if ok:
            assert codes[i] == len(want) and outs[i] == want, i
        else:
            assert codes[i] == -5, i
    assert all(c == -5 for c in codes[len(DECODE_TABLE):])
    blocks = _blocks()
    comp = [orc_encode(b, m) for b in blocks for m in (0, 1, 2)] + [H.golden("s2_twain.txt.rawsnappy")]
    want = [b for b in blocks for m in (0, 1, 2)] + [H.golden("s2_twain.txt")]
    outs, codes = codec.decode_blocks(comp, [len(w) for w in want])
    assert outs == want
    assert codec.Decode(comp[24]) == want[24]
    with pytest.raises(s2.ErrCorrupt):
        codec.Decode(comp[24][:-3])
    with pytest.raises(s2.ErrTooLarge):
        codec.Decode(comp[24], max_len=100)


def test_s2_roundtrip_device_256mib(codec):
    n = 4096
    src = H.synth_text_torch(n * 65536, "cuda")
    src[10 * 65536:11 * 65536] = 0
    src[11 * 65536:12 * 65536] = torch.randint(0, 256, (65536,), dtype=torch.uint8, device="cuda")
    for snappy, better in ((False, False), (True, False), (False, True), (True, True)):
        enc, sizes = codec.encode_device(src, snappy=snappy, better=better)
        torch.cuda.synchronize()
        assert int(sizes.min()) > 0
        out, osz = codec.decode_device(enc, sizes.to(torch.int32), src_stride=enc.stride(0))
        torch.cuda.synchronize()
        assert bool((osz == 65536).all())
        assert torch.equal(out.view(-1), src)
        ratio = float(sizes.sum()) / src.numel()
        assert 0.3 < ratio < 0.8, ratio


def test_s2_staged_decode_is_used_and_equals_one_warp(codec, emu_lib):
    """The staged S2 decoder (tag walk per lane, execution per warp) takes the encoders' blocks on the device, and its bytes
    equal the emulated kernels' (which tests/test_emu_s2.py compares with the one-warp decoder and the oracle)."""
    from emu_util import emu_s2_decode
    tw = H.golden("twain.txt")
    blocks = [tw[i:i + 65536] for i in range(0, 5 * 65536, 65536)] + [b"ab" * 20000, bytes(50000), tw[:1000]]
    for better in (False, True):
        comps = codec.encode_blocks(blocks, better=better)
        outs, codes = codec.decode_blocks(comps, [len(b) for b in blocks])
        assert outs == blocks, codes
        assert codec.staged_count(len(blocks)) == len(blocks)
        _, res, _, _ = emu_s2_decode(emu_lib, comps, [len(b) for b in blocks])
        assert res == blocks
