"""Frame mode of the zstd encoder (b2c_zstd_encode_frames_*) under the SIMT emulator: SURVEY section 8 rows a-7 (history:
fastBase.addBlock / ensureHist, zstd/enc_base.go:57-199), a-15 (frame header for any content size) and f-1 (EncodeAll of
an input larger than one block is ONE frame, zstd/encoder.go:796-830).

Checked: every frame decodes with the pinned decoder oracle and with libzstd; the frame header bytes equal the
reference's for the same input length and level (oracle restatement of frameHeader.appendTo via its EncodeAll); every
block is byte-identical to the oracle's blockEnc.encode for the kernel's parse (fresh block encoder, last-block flag);
the sequences replayed over the history reproduce the input (so offsets that reach into earlier blocks are real); frame
mode is not larger than independent chunks and stays within a stated distance of the reference's multi-block output.
CPU only."""
import numpy as np
import pytest

import helpers as H
from check_util import _lits_from_seqs
from emu_util import emu_encode_frames, emu_encode


def split_blocks(frame, hdr_len, crc):
    """[(last, type, size, payload bytes)] of a frame's blocks."""
    out = []
    o = hdr_len
    end = len(frame) - (4 if crc else 0)
    while o < end:
        bh = frame[o] | (frame[o + 1] << 8) | (frame[o + 2] << 16)
        last, typ, size = bh & 1, (bh >> 1) & 3, bh >> 3
        body = 1 if typ == 1 else size
        out.append((last, typ, size, frame[o:o + 3 + body]))
        o += 3 + body
        if last:
            break
    assert o == end, (o, end)
    return out


def replay_frame(data, blocks):
    """Execute every block's sequences over the frame so far (repeat offsets carried across blocks, seqdec.go:463-500)."""
    out = bytearray()
    rep = [1, 4, 8]
    for b in blocks:
        org = data[len(out):len(out) + b["len"]]
        if b["nseq"] == 0:
            out += org
            continue
        tri = b["tri"]
        lits = b["lits"] if b["lits"] is not None else _lits_from_seqs(tri, org)
        lp = 0
        for ll, ml3, ofv in tri:
            ll = int(ll); ml = int(ml3) + 3; ofv = int(ofv)
            out += lits[lp:lp + ll]
            lp += ll
            if ofv > 3:
                off = ofv - 3
                rep = [off, rep[0], rep[1]]
            else:
                idx = ofv - 1 + (1 if ll == 0 else 0)
                assert idx == 0, "the kernels only emit repeat code 1 with literals"
                off = rep[0]
            start = len(out) - off
            assert start >= 0
            for k in range(ml):
                out.append(out[start + k])
        out += lits[lp:]
    return bytes(out)


def check_frame_mode(inputs, frames, blocks, level, label):
    L = H.oracle()
    bi = 0
    fblock = 49152 if level == 1 else 98304
    for i, (data, fr) in enumerate(zip(inputs, frames)):
        r, dec = H.oracle_decode(fr, len(data) + 64)
        assert r == len(data) and dec == data, f"{label} frame {i}: oracle decode mismatch (r={r})"
        assert H.libzstd_decode(fr, len(data)) == data, f"{label} frame {i}: libzstd decode mismatch"
        # header = the reference's for this length and level (the oracle's EncodeAll writes frameHeader.appendTo)
        _, ref = H.oracle_encode(data, level=level)
        hl = 6 if len(data) == 0 else (5 + (0 if 1024 < len(data) <= ((4 << 20) if level == 1 else (8 << 20)) else 1) +
                                       (0 if len(data) < 256 and not len(data) > 1024 else (2 if len(data) < 65792 else (4 if len(data) < 0xffffffff else 8))))
        if 0 < len(data) < 256:
            hl = 6                      # window byte, no content size
        assert fr[:hl] == ref[:hl], f"{label} frame {i}: header differs from the reference's ({fr[:hl].hex()} vs {ref[:hl].hex()})"
        nb = max(1, (len(data) + fblock - 1) // fblock)
        mine = blocks[bi:bi + nb]
        bi += nb
        parts = split_blocks(fr, hl, crc=len(data) > 0)
        assert len(parts) == nb, f"{label} frame {i}: {len(parts)} blocks, planned {nb}"
        pos = 0
        for j, (b, (last, typ, size, raw)) in enumerate(zip(mine, parts)):
            assert last == (1 if j == nb - 1 else 0) == b["last"]
            org = data[pos:pos + b["len"]]
            assert b["hist"] == min(pos, (65536 if level == 1 else 131072) - fblock)
            if b["nseq"] > 0:
                tri = np.asarray(b["tri"]).astype(np.uint32)
                lb = b["lits"] if b["kind"] == 0 else _lits_from_seqs(tri, org)
                rb, ob = H.oracle_encode_block(org, lb, tri, last)
                assert rb == len(raw) and ob == raw, f"{label} frame {i} block {j}: entropy stage differs from oracle blockEnc.encode (kind {b['kind']})"
            else:
                assert typ == 0 and raw[3:] == org
            pos += b["len"]
        assert replay_frame(data, mine) == data, f"{label} frame {i}: sequences do not reproduce the input"
        assert len(fr) <= L.orc_zstd_max_encoded_size(len(data), level, 1) + 3 * nb


@pytest.mark.parametrize("level", [1, 2, 3])
def test_frames_sizes_and_history(emu_lib, level):
    tw = H.golden("twain.txt")
    fblock = 49152 if level == 1 else 98304
    inputs = [b"", b"a", tw[:200], tw[:1024], tw[:1025], tw[:fblock], tw[:fblock + 1], tw[:2 * fblock - 1], tw[:3 * fblock + 777],
              bytes(2 * fblock + 5), b"abcd" * (fblock // 2), H.golden("html.txt"), H.golden("e.txt")[:70000]]
    frames, blocks, _ = emu_encode_frames(emu_lib, inputs, level=level)
    check_frame_mode(inputs, frames, blocks, level, "frames-L%d" % level)
    # offsets reach into earlier blocks: some block beyond the first has a match starting before the block
    multi = [b for b in blocks if b["hist"] > 0 and b["nseq"] > 0]
    assert multi
    deep = 0
    for b in multi:
        pos = 0
        for ll, ml3, ofv in b["tri"]:
            pos += int(ll)
            if int(ofv) > 3 and int(ofv) - 3 > pos:
                deep += 1
            pos += int(ml3) + 3
    assert deep > 0, "no match refers to the history"


@pytest.mark.parametrize("level", [1, 2])
def test_frames_ratio(emu_lib, level):
    """One frame with history vs (a) independent one-block frames (round-1 EncodeAll) and (b) the reference's multi-block
    EncodeAll (oracle: full window, 64 / 128 KiB blocks, repeat-mode tables).  Stated tolerance: frame mode <= independent
    chunks, and <= +4 % of the reference's single frame on text."""
    tw = H.golden("twain.txt")
    data = tw[:6 * 65536]
    frames, _, _ = emu_encode_frames(emu_lib, [data], level=level, dump=False)
    got = len(frames[0])
    block = 65536 if level == 1 else 131072
    indep = sum(len(f) for f in emu_encode(emu_lib, [data[i:i + block] for i in range(0, len(data), block)], level=level)[0])
    ref = H.oracle_encode(data, level=level)[0]
    print("level %d: frame mode %d, independent chunks %d, reference single frame %d" % (level, got, indep, ref))
    assert got <= indep * 1.005
    assert got <= ref * 1.04


def test_frames_deterministic_lane_order(emu_lib):
    tw = H.golden("twain.txt")
    inputs = [tw[1000:1000 + 90000], b"xyz" * 30000]
    a = emu_encode_frames(emu_lib, inputs, level=1, desc=0, dump=False)[0]
    b = emu_encode_frames(emu_lib, inputs, level=1, desc=1, dump=False)[0]
    assert a == b


def test_frames_unaligned_inputs(emu_lib):
    """Frames that do not start on a 16-byte boundary: the staged copy and the checksum take their unaligned forms; the
    bytes must not change."""
    tw = H.golden("twain.txt")
    inputs = [tw[:150001], tw[5:70000]]
    a = emu_encode_frames(emu_lib, inputs, level=1, dump=False)[0]
    for mis in (1, 8, 13):
        b = emu_encode_frames(emu_lib, inputs, level=1, dump=False, misalign=mis)[0]
        assert a == b
    for data, fr in zip(inputs, a):
        assert H.libzstd_decode(fr, len(data)) == data


def test_long_frames_decode_block_parallel(emu_lib):
    """Frames with more than four blocks on the staged decoder: in its per-block form (one lane / quad per (input, block),
    the form the device picks for few inputs) frame mode's own output stays on the staged path -- its blocks never use a
    repeat offset of an earlier block -- while the per-input form (at most four blocks) hands it to the one-warp decoder.
    Both give the input back; single-block frames are staged either way.  libzstd frames, which do use repeat offsets across
    blocks, are decoded correctly whichever path takes them."""
    from emu_util import emu_decode
    tw = H.golden("twain.txt")
    inputs = [tw, tw[:49152 * 6 + 5], tw[:1000], bytes(200000) + tw[:100000]]
    frames = emu_encode_frames(emu_lib, inputs, level=1, dump=False)[0]
    z = H.libzstd()
    import ctypes
    big = tw[:300000]
    cap = z.ZSTD_compressBound(len(big))
    buf = ctypes.create_string_buffer(cap)
    zn = z.ZSTD_compress(buf, cap, big, len(big), 3)
    frames_all = frames + [buf.raw[:zn]]
    inputs_all = inputs + [big]
    caps = [len(x) + 16 for x in inputs_all]
    try:
        emu_lib.emu_set_dec_maxb(32)
        staged = []
        sizes, outs = emu_decode(emu_lib, frames_all, caps, staged=staged)
        assert outs == inputs_all, sizes
        assert staged[:4] == [1, 1, 1, 1], staged          # 8, 7, 1 and 7 blocks: all staged, block-parallel
        emu_lib.emu_set_dec_maxb(4)
        staged4 = []
        sizes, outs = emu_decode(emu_lib, frames_all, caps, staged=staged4)
        assert outs == inputs_all, sizes
        assert staged4[:4] == [0, 0, 1, 0], staged4
        for desc in (1,):
            emu_lib.emu_set_dec_maxb(32)
            sizes, outs = emu_decode(emu_lib, frames_all, caps, desc=desc)
            assert outs == inputs_all
    finally:
        emu_lib.emu_set_dec_maxb(0)


@pytest.mark.parametrize("level", [1, 2])
def test_frames_random_structures(emu_lib, level):
    """Frame mode on random LZ-structured inputs of random sizes (block boundaries land anywhere, long runs and far copies cross
    them): every block equals the oracle's blockEnc.encode for the kernel's parse, frames decode with the oracle and libzstd,
    and both forms of the staged decoder give the input back."""
    from emu_util import emu_decode
    from test_emu_encoder_random import _structured
    rng = np.random.Generator(np.random.PCG64(100 + level))
    fblock = 49152 if level == 1 else 98304
    inputs = []
    for k in range(6):
        n = int(rng.integers(1, 4 * fblock))
        inputs.append(_structured(rng, n))
    inputs += [_structured(rng, fblock), _structured(rng, 2 * fblock + 1)]
    frames, blocks, _ = emu_encode_frames(emu_lib, inputs, level=level)
    check_frame_mode(inputs, frames, blocks, level, "frames-rnd-L%d" % level)
    caps = [len(x) + 8 for x in inputs]
    try:
        for maxb in (0, 4):
            emu_lib.emu_set_dec_maxb(maxb)
            sizes, outs = emu_decode(emu_lib, frames, caps)
            assert outs == inputs, (maxb, list(sizes))
    finally:
        emu_lib.emu_set_dec_maxb(0)
