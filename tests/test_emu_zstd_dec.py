"""The zstd decode kernel's source (compress_b200/csrc/b2c_zstd_dec.cuh) run under the SIMT emulator against the
reference's golden vectors and the oracle decoder.  CPU only; the same cases run on the device in test_zstd_gpu.py."""
import os
import zipfile

import numpy as np
import pytest

import helpers as H
from emu_util import emu_decode, emu_encode


def _pairs(zf):
    names = set(zf.namelist())
    for nm in sorted(names):
        if nm.endswith(".zst") and nm[:-4] in names:
            yield nm, zf.read(nm), zf.read(nm[:-4])


def test_emu_decoder_zip_subset(emu_lib, staged_form):
    # zstd/decoder_test.go:201-216 TestNewDecoder
    zf = zipfile.ZipFile(os.path.join(H.GOLDEN, "zstd_decoder.zip"))
    items = sorted(_pairs(zf), key=lambda it: len(it[2]))[:24]      # the emulator is slow: the 24 smallest pairs
    for desc in (0, 1):
        outs, res = emu_decode(emu_lib, [c for _, c, _ in items], [len(w) + 64 for _, _, w in items], desc)
        for (nm, _, want), r, got in zip(items, outs, res):
            assert r == len(want) and got == want, nm


def test_emu_good_zip(emu_lib, oracle_lib):
    # zstd/decoder_test.go:393 TestNewDecoderGood
    zf = zipfile.ZipFile(os.path.join(H.GOLDEN, "zstd_good.zip"))
    comps, wants, names = [], [], []
    for nm in zf.namelist():
        if not nm.endswith(".zst"):
            continue
        comp = zf.read(nm)
        r, got = H.oracle_decode(comp, 64 << 20)
        assert r >= 0
        comps.append(comp); wants.append(got); names.append(nm)
    outs, res = emu_decode(emu_lib, comps, [len(w) + 64 for w in wants])
    for nm, want, r, got in zip(names, wants, outs, res):
        assert r == len(want) and got == want, nm


def test_emu_bad_zip(emu_lib, oracle_lib, staged_form):
    # zstd/decoder_test.go:409-455 TestNewDecoderBad: every input must be rejected, as the oracle does
    zf = zipfile.ZipFile(os.path.join(H.GOLDEN, "zstd_bad.zip"))
    comps, names = [], []
    for nm in zf.namelist():
        if nm.endswith(".zst"):
            comps.append(zf.read(nm)); names.append(nm)
    outs, _ = emu_decode(emu_lib, comps, [4 << 20] * len(comps))
    for nm, comp, r in zip(names, comps, outs):
        ro, _ = H.oracle_decode(comp, 4 << 20)
        assert r < 0 and ro == r, (nm, ro, r)


def test_emu_decode_encoder_output(emu_lib, oracle_lib):
    # frames produced by the oracle encoder (Huffman 1X/4X, RLE/raw/treeless literals, FSE/RLE/predefined/repeat
    # tables, multi-block frames with repeat offsets, concatenated + skippable frames) decode to the source
    rng = np.random.default_rng(7)
    srcs = [b"", b"a", b"abc" * 7, bytes(1000), bytes(rng.integers(0, 256, 3000, dtype=np.uint8)),
            H.golden("twain.txt")[:200000], H.golden("html.txt")[:70000], H.golden("e.txt")[:66000],
            H.synth_text(150000), bytes(rng.integers(0, 4, 90000, dtype=np.uint8))]
    enc = lambda b: H.oracle_encode(b)[1]
    comps = [enc(s) for s in srcs]
    # concatenated frames with a skippable frame in between
    srcs.append(srcs[5][:5000] + srcs[6][:5000])
    comps.append(enc(srcs[5][:5000]) + b"\x50\x2a\x4d\x18\x03\x00\x00\x00xyz" + enc(srcs[6][:5000]))
    staged = []
    outs, res = emu_decode(emu_lib, comps, [len(s) + 16 for s in srcs], staged=staged)
    for i, (s, r, got) in enumerate(zip(srcs, outs, res)):
        assert r == len(s) and got == s, i
    # the staged kernels (b2c_zstd_dec_staged.cuh) take single frames of at most four blocks; the concatenated input and
    # the longer frames are left to the one-warp decoder
    assert staged[:5] == [1] * 5 and staged[6] == 1 and staged[7] == 1 and staged[10] == 0, staged
    outs1, res1 = emu_decode(emu_lib, comps, [len(s) + 16 for s in srcs], mode=1)
    assert list(outs1) == list(outs) and res1 == res
    # too-small destination and truncation are reported, not written past
    outs, _ = emu_decode(emu_lib, [comps[5], comps[5][:-5], comps[8][:1000]], [1000, 300000, 300000])
    assert all(o < 0 for o in outs)


def test_emu_decode_own_frames(emu_lib):
    chunks = H.synth_chunks("text", 3) + [H.golden("twain.txt")[:65536], bytes(65536), b"xy" * 100]
    frames, outs, *_ = emu_encode(emu_lib, chunks)
    staged = []
    souts, res = emu_decode(emu_lib, frames, [65536] * len(frames), staged=staged)
    for c, r, got in zip(chunks, souts, res):
        assert r == len(c) and got == c
    assert staged == [1] * len(frames)


def test_emu_decode_staged_levels_and_marks(emu_lib, staged_form):
    # frames of every encoder level through the staged kernels (reversed lane order too); a flipped content checksum, a
    # flipped payload bit and a short destination are marked by a staged stage and answered by the one-warp decoder
    # exactly as it answers them alone
    chunks = [H.golden("twain.txt")[:131072], H.synth_chunks("text", 1, size=131072)[0][:100000], bytes(131072),
              H.golden("html.txt")[:70000]]
    for level in (2, 3):
        frames, outs, *_ = emu_encode(emu_lib, chunks, level=level)
        for desc in (0, 1):
            staged = []
            souts, res = emu_decode(emu_lib, frames, [len(c) for c in chunks], desc=desc, staged=staged)
            for c, r, got in zip(chunks, souts, res):
                assert r == len(c) and got == c
            assert staged == [1] * len(frames), (level, desc, staged)
    f = frames[0]
    bad_crc = f[:-1] + bytes([f[-1] ^ 1])
    bad_bit = f[:len(f) // 2] + bytes([f[len(f) // 2] ^ 0x10]) + f[len(f) // 2 + 1:]
    cases = [bad_crc, bad_bit, f, f[:-3]]
    caps = [131072, 131072, 1000, 131072]
    staged = []
    outs0, _ = emu_decode(emu_lib, cases, caps, staged=staged)
    outs1, _ = emu_decode(emu_lib, cases, caps, mode=1)
    assert list(outs0) == list(outs1) and all(o < 0 for o in outs0), (outs0, outs1)
    assert staged == [0, 0, 0, 0]
