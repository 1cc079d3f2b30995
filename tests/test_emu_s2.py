"""S2 / Snappy block kernels (encode tail of b2c_zstd_enc.cuh, b2c_s2_dec.cuh) under the SIMT emulator, checked
against the S2 oracle, the reference's decode vectors and pyarrow's Snappy codec.  CPU only."""
import numpy as np
import pytest

import helpers as H
from emu_util import emu_s2_decode, emu_s2_encode
from s2_vectors import DECODE_TABLE, INVALID_VARINT
from test_oracle_s2 import s2_decode as orc_decode, s2_encode as orc_encode, _L


def _blocks():
    rng = np.random.default_rng(11)
    tw = H.golden("twain.txt")
    return [b"", b"a", b"abc" * 5, bytes(31), bytes(32), bytes(100), tw[:33], tw[:1000], tw[:65536], tw[70000:70000 + 65521],
            H.golden("html.txt")[:65536], H.golden("e.txt")[:65536], H.synth_text(65536), bytes(65536), b"ab" * 32768,
            b"abc" * 21845, bytes(rng.integers(0, 256, 5000, dtype=np.uint8)), bytes(rng.integers(0, 256, 65536, dtype=np.uint8)),
            bytes(rng.integers(0, 3, 65536, dtype=np.uint8)), tw[:3000] + bytes(rng.integers(0, 256, 60000, dtype=np.uint8)),
            (tw[:700] + bytes(rng.integers(0, 256, 300, dtype=np.uint8))) * 60]


def test_emu_s2_encode_roundtrip(emu_lib, oracle_lib):
    blocks = _blocks()
    L = _L()
    for snappy in (False, True):
        for desc in (0, 1):
            enc, outs = emu_s2_encode(emu_lib, blocks, snappy=snappy, desc=desc)
            for i, (b, c, r) in enumerate(zip(blocks, enc, outs)):
                assert r == len(c) > 0 and r <= L.orc_s2_max_encoded_len(len(b)), (i, r)
                n, got = orc_decode(c, len(b))
                assert n == len(b) and got == b, (snappy, desc, i)
    # lane order must not change the bytes
    a, _ = emu_s2_encode(emu_lib, blocks, desc=0)
    b, _ = emu_s2_encode(emu_lib, blocks, desc=1)
    assert a == b
    # ratio next to the reference algorithm (oracle) on text: within 5 % of s2.Encode
    tw = H.golden("twain.txt")[:65536]
    ours = len(emu_s2_encode(emu_lib, [tw])[0][0])
    ref = len(orc_encode(tw, 0))
    assert ours < 1.05 * ref, (ours, ref)
    # incompressible and tiny blocks are stored as one literal
    enc, _ = emu_s2_encode(emu_lib, [blocks[17], blocks[2]])
    assert len(enc[0]) == 65536 + 3 + 3 and enc[1] == bytes([15, 14 << 2]) + blocks[2]


def test_emu_s2_better_and_ratios(emu_lib, oracle_lib):
    """s2.EncodeBetter / EncodeSnappyBetter class: blocks decode with the oracle's s2Decode in
    both lane orders, and per corpus the size is within +5 % of the reference algorithm of the same class
    (all corpora, not one chunk); the fast class is checked the same way against s2.Encode."""
    blocks = _blocks()
    L = _L()
    for snappy in (False, True):
        a, outs = emu_s2_encode(emu_lib, blocks, snappy=snappy, better=True, desc=0)
        b2, _ = emu_s2_encode(emu_lib, blocks, snappy=snappy, better=True, desc=1)
        assert a == b2
        for i, (blk, c, r) in enumerate(zip(blocks, a, outs)):
            assert r == len(c) > 0 and r <= L.orc_s2_max_encoded_len(len(blk)), (i, r)
            n, got = orc_decode(c, len(blk))
            assert n == len(blk) and got == blk, (snappy, i)
    tw = H.golden("twain.txt")
    corp = {"twain": [tw[i:i + 65536] for i in range(0, 3 * 65536, 65536)], "html": [H.golden("html.txt")],
            "e": [H.golden("e.txt")[:65536]], "synth": [H.synth_text(65536, 3)]}
    for name, chunks in corp.items():
        for better, mode in ((False, 0), (True, 1)):
            ours = sum(len(x) for x in emu_s2_encode(emu_lib, chunks, better=better)[0])
            ref = sum(len(orc_encode(c, mode)) for c in chunks)
            assert ours <= 1.05 * ref, (name, better, ours, ref)
        # Snappy-compatible output of both classes against the oracle's EncodeSnappy (it has one Snappy match finder class)
        ours = sum(len(x) for x in emu_s2_encode(emu_lib, chunks, snappy=True)[0])
        ref = sum(len(orc_encode(c, 2)) for c in chunks)
        assert ours <= 1.05 * ref, (name, "snappy", ours, ref)


def test_emu_snappy_output_is_snappy(emu_lib):
    pa = pytest.importorskip("pyarrow")
    codec = pa.Codec("snappy")
    blocks = [b for b in _blocks() if len(b)]
    for better in (False, True):
        enc, _ = emu_s2_encode(emu_lib, blocks, snappy=True, better=better)
        for i, (b, c) in enumerate(zip(blocks, enc)):
            assert codec.decompress(c, decompressed_size=len(b)).to_pybytes() == b, (better, i)


def test_emu_s2_decode(emu_lib, oracle_lib):
    # reference vectors (s2/s2_test.go TestDecode / TestInvalidVarint): output, error and no write past dLen
    ins = [v[0] for v in DECODE_TABLE] + INVALID_VARINT
    for desc in (0, 1):
        outs, res, dst, dst_off = emu_s2_decode(emu_lib, ins, [100] * len(ins), desc)
        for i, (inp, want, ok) in enumerate(DECODE_TABLE):
            if ok:
                assert outs[i] == len(want) and res[i] == want, i
            else:
                assert outs[i] == -5, i
            dlen = inp[0]
            tail = dst[int(dst_off[i]) + dlen:int(dst_off[i]) + 112]
            assert (tail == 0x5A).all(), i
        assert (outs[len(DECODE_TABLE):] == -5).all()
    # blocks from the oracle's encoders (repeat tags, long copies, Snappy form) and the golden Snappy block
    blocks = _blocks()
    comp = [orc_encode(b, m) for b in blocks for m in (0, 1, 2)]
    want = [b for b in blocks for m in (0, 1, 2)]
    comp.append(H.golden("s2_twain.txt.rawsnappy"))
    want.append(H.golden("s2_twain.txt"))
    outs, res, _, _ = emu_s2_decode(emu_lib, comp, [len(w) for w in want])
    for i, (w, r, g) in enumerate(zip(want, outs, res)):
        assert r == len(w) and g == w, i
    # too small destination / truncated input
    outs, _, _, _ = emu_s2_decode(emu_lib, [comp[24], comp[24][:-3]], [100, 65536])
    assert outs[0] == -4 and outs[1] == -5


def test_s2_decode_copy4_and_length_offset(emu_lib, oracle_lib):
    # s2/s2_test.go:472-597 TestDecodeCopy4, TestDecodeLengthOffset (output and no write past the decoded length)
    from s2_vectors import decode_copy4_case, decode_length_offset_cases
    cases = [decode_copy4_case()] + decode_length_offset_cases()
    for inp, want in cases[:200] + cases[::37]:
        n, got = orc_decode(inp, len(want))
        assert n == len(want) and got == want
    outs, res, dst, dst_off = emu_s2_decode(emu_lib, [c[0] for c in cases], [len(c[1]) + 40 for c in cases])
    for i, ((inp, want), r, got) in enumerate(zip(cases, outs, res)):
        assert r == len(want) and got == want, i
        tail = dst[int(dst_off[i]) + len(want):int(dst_off[i]) + len(want) + 40]
        assert (tail == 0x5A).all(), i


def test_emu_s2_staged_decode_forms(emu_lib, oracle_lib):
    """The staged S2 block decoder (tag walk one lane per block, execution one warp per block) and the one-warp decoder give
    the same bytes and the same results; the staged kernels really take the encoders' blocks (all tag kinds: literals of every
    header size, copy1 / copy2, repeats with extended lengths, long runs), in both lane orders; what they cannot take (a block
    above 64 KiB, damaged blocks) is answered by the one-warp kernel with the reference's error."""
    from emu_util import emu_s2_decode, emu_s2_encode
    from test_oracle_s2 import s2_encode
    import numpy as np
    tw = H.golden("twain.txt")
    rng = np.random.default_rng(21)
    blocks = [tw[:65536], tw[100000:100000 + 30011], b"", b"a", b"ab" * 30000, bytes(65536), H.golden("html.txt"),
              rng.integers(0, 256, 5000, dtype=np.uint8).tobytes() + tw[:20000] + rng.integers(0, 256, 300, dtype=np.uint8).tobytes(),
              tw[:100] + bytes(40000) + tw[:100]]
    comps = []
    for better in (False, True):
        for snappy in (False, True):
            comps += emu_s2_encode(emu_lib, blocks, snappy=snappy, better=better)[0]
    srcs = blocks * 4
    comps += [s2_encode(b, m) for b in blocks[:2] for m in (0, 1, 2)]          # the reference encoders' blocks too
    srcs += [b for b in blocks[:2] for _ in range(3)]
    big = s2_encode(tw[:70000], 0)                                                # > 64 KiB decoded: one-warp kernel
    damaged = bytearray(comps[0]); damaged[len(damaged) // 2] ^= 0x41
    comps_all = comps + [big, bytes(damaged), comps[0][:-7]]
    caps = [len(x) for x in srcs] + [70000, 65536, 65536]
    results = {}
    for staged in (1, 0):
        for desc in (0, 1):
            emu_lib.emu_set_s2_staged(staged)
            outs, res, _, _ = emu_s2_decode(emu_lib, comps_all, caps, desc=desc)
            results[(staged, desc)] = (list(outs), res)
            if staged:
                k = emu_lib.emu_get_s2_staged_count()
                assert k >= len(comps) - 8, k          # (empty / tiny blocks may go either way)
    emu_lib.emu_set_s2_staged(1)
    base = results[(0, 0)]
    for key, val in results.items():
        assert val == base, key
    outs, res = base
    for i, sblk in enumerate(srcs):
        assert outs[i] == len(sblk) and res[i] == sblk, i
    assert outs[len(srcs)] == 70000 and res[len(srcs)] == tw[:70000]
    assert outs[-1] < 0


def test_emu_s2_staged_decode_random_structures(emu_lib, oracle_lib):
    """Random LZ-structured blocks (runs, periodic patterns, far copies, noise) through every encoder mode, decoded by the staged
    kernels and by the one-warp kernel: same bytes, equal to the source; the oracle's s2Decode agrees."""
    from emu_util import emu_s2_decode, emu_s2_encode
    from test_emu_encoder_random import _structured
    from test_oracle_s2 import s2_decode
    import numpy as np
    rng = np.random.default_rng(77)
    blocks = [_structured(rng, int(rng.integers(1, 65537))) for _ in range(10)] + [_structured(rng, 65536) for _ in range(3)]
    for better in (False, True):
        for snappy in (False, True):
            comps = emu_s2_encode(emu_lib, blocks, snappy=snappy, better=better)[0]
            caps = [len(b) for b in blocks]
            emu_lib.emu_set_s2_staged(1)
            o1, r1, _, _ = emu_s2_decode(emu_lib, comps, caps)
            k = emu_lib.emu_get_s2_staged_count()
            emu_lib.emu_set_s2_staged(0)
            o0, r0, _, _ = emu_s2_decode(emu_lib, comps, caps)
            emu_lib.emu_set_s2_staged(1)
            assert r1 == blocks and r0 == blocks and list(o1) == list(o0), (better, snappy)
            assert k >= len(blocks) - 2, k
            for b, c in zip(blocks[:3], comps[:3]):
                n, got = s2_decode(c, len(b))
                assert n == len(b) and got == b
