"""Parity checks shared by the emulated-kernel tests and the GPU tests."""
import numpy as np

import helpers as H
from emu_util import frame_header_len


def check_frames(chunks, frames, hdr=None, seqs=None, lits=None, crc=True, label="", level=1, check_raw=True):
    """Every frame must (1) decode to its chunk with the oracle's restatement of the reference decoder
    and with libzstd, (2) respect MaxEncodedSize, and -- when the parse dump is given -- (3) carry a block
    that is byte-identical to the oracle's blockEnc.encode for the same literals + sequences."""
    L = H.oracle()
    tot = 0
    for i, c in enumerate(chunks):
        enc = frames[i]
        assert len(enc) > 0, f"{label} chunk {i}: empty output"
        assert len(enc) <= L.orc_zstd_max_encoded_size(len(c), level, 1), f"{label} chunk {i}: exceeds MaxEncodedSize"
        r, dec = H.oracle_decode(enc, len(c) + 64)
        assert r == len(c) and dec == c, f"{label} chunk {i}: oracle decode mismatch (r={r})"
        z = H.libzstd_decode(enc, len(c))
        assert z == c, f"{label} chunk {i}: libzstd decode mismatch"
        tot += len(enc)
        if hdr is not None:
            nseq, nlit, kind, _ = [int(x) for x in hdr[i]]
            if nseq > 0 and (kind == 0 or check_raw):
                # kind 0 = compressed candidate, 1 = raw fallback, 2 = RLE: in every case the block must be what the
                # oracle's blockEnc.encode makes of the same literals + sequences (raw / RLE decisions included)
                tri = np.asarray(seqs[i][:nseq]).astype(np.uint32)
                if kind == 0:
                    lb = bytes(np.asarray(lits[i][:nlit]).astype(np.uint8))
                else:
                    lb = _lits_from_seqs(tri, c)       # raw / RLE blocks: the kernel does not produce the literals
                    assert len(lb) == nlit, f"{label} chunk {i}: literal count {nlit} vs sequences {len(lb)}"
                # the sequences must reproduce the chunk (independent of the entropy stage)
                assert _replay(tri, lb, len(c)) == c, f"{label} chunk {i}: sequences do not reproduce the input"
                rb, ob = H.oracle_encode_block(c, lb, tri, 1)
                fh = frame_header_len(len(c))
                blk = enc[fh:len(enc) - (4 if crc else 0)]
                assert rb == len(blk) and ob == blk, (
                    f"{label} chunk {i}: entropy stage differs from oracle blockEnc.encode "
                    f"(kind {kind}, oracle {rb} B, gpu {len(blk)} B, first diff {_first_diff(ob, blk)})")
    return tot


def _lits_from_seqs(tri, c):
    """Literal bytes implied by (litLen, matchLen-3, offset) triples over the chunk c."""
    out = bytearray()
    pos = 0
    for ll, ml3, _ in tri:
        out += c[pos:pos + int(ll)]
        pos += int(ll) + int(ml3) + 3
    out += c[pos:]
    return bytes(out)


def _first_diff(a, b):
    for j in range(min(len(a), len(b))):
        if a[j] != b[j]:
            return j
    return min(len(a), len(b))


def _replay(tri, lits, n):
    """Execute (litLen, matchLen-3, offsetValue) with zstd repeat-offset rules (seqdec.go:463-500)."""
    out = bytearray()
    rep = [1, 4, 8]
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
            if idx == 0:
                off = rep[0]
            else:
                off = rep[0] - 1 if idx == 3 else rep[idx]
                if idx == 1:
                    rep = [off, rep[0], rep[2]]
                else:
                    rep = [off, rep[0], rep[1]]
        start = len(out) - off
        if start < 0:
            return None
        for k in range(ml):
            out.append(out[start + k])
    out += lits[lp:]
    return bytes(out)
