#!/usr/bin/env python
"""The reference's own fuzz seed corpora through the oracle and the emulated kernels (CPU only; run in the build container,
where /root/reference is mounted -- the corpora are too large to commit; tests/golden holds a size-capped subset, see
tests/golden/make_fixtures.py).

  decode:  zstd/testdata/fuzz/decode-corpus-{raw,encoded}.zip  (FuzzDecodeAll / FuzzDecoder seeds, zstd/fuzz_test.go:17-19)
           every input: oracle verdict == emulated kernels' verdict (both staged forms), bytes equal when accepted;
           libzstd's bytes equal the oracle's wherever both accept.
  encode:  zstd/testdata/fuzz/encode-corpus-{raw,encoded}.zip  (FuzzEncoding seeds, zstd/fuzz_test.go:154-170)
           every input through the emulated encoders (one-block frames when it fits a block, frame mode otherwise; levels 1-3)
           and back through the oracle decoder, libzstd and the emulated decoder.
  s2:      s2/testdata/fuzz/block-corpus-{raw,enc}.zip         (FuzzEncodingBlocks seeds, s2/fuzz_test.go:16-18)
           every input (cut to 64 KiB blocks) through the four emulated S2 / Snappy encoders and back through the oracle's
           and the emulated decoder.

  huff0:   huff0/testdata/{fse_compress,regression}.zip (FuzzCompress seeds, TestCompressRegression) -- emulated Compress1X / 4X
           output bytes equal the oracle's; huff0/testdata/{huff0_decompress1x,decompress1x_regression}.zip -- ReadTable parity.

  python tools/fuzz_ref_corpora.py [decode] [encode] [s2] [huff0] [--ref /root/reference] [--limit N]"""
import argparse
import os
import sys
import time
import zipfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

import helpers as H                                                                  # noqa: E402
from emu_util import emu_encode, emu_decode, emu_s2_encode, emu_s2_decode, emu_encode_frames   # noqa: E402
from test_oracle_s2 import s2_decode as orc_s2_decode                               # noqa: E402


def go_unquote(s):
    """The body of a Go interpreted string literal (strconv.Quote output) -> bytes."""
    out = bytearray()
    i, n = 0, len(s)
    simple = {"a": 7, "b": 8, "f": 12, "n": 10, "r": 13, "t": 9, "v": 11, "\\": 92, "'": 39, '"': 34}
    while i < n:
        c = s[i]
        if c != "\\":
            out += c.encode("utf-8")
            i += 1
            continue
        e = s[i + 1]
        if e in simple:
            out.append(simple[e]); i += 2
        elif e == "x":
            out.append(int(s[i + 2:i + 4], 16)); i += 4
        elif e == "u":
            out += chr(int(s[i + 2:i + 6], 16)).encode("utf-8"); i += 6
        elif e == "U":
            out += chr(int(s[i + 2:i + 10], 16)).encode("utf-8"); i += 10
        elif e in "01234567":
            out.append(int(s[i + 1:i + 4], 8)); i += 4
        else:
            raise ValueError("bad escape \\%s" % e)
    return bytes(out)


def read_corpus(path, limit=0):
    """Entries of a corpus zip as bytes: raw files as they are, `go test fuzz v1` files -> their first []byte argument."""
    zf = zipfile.ZipFile(path)
    out = []
    for info in zf.infolist():
        if info.is_dir():
            continue
        b = zf.read(info)
        if b.startswith(b"go test fuzz v1"):
            val = None
            for line in b.decode("utf-8", "surrogateescape").split("\n")[1:]:
                line = line.strip()
                if line.startswith('[]byte("') and line.endswith('")'):
                    val = go_unquote(line[8:-2])
                    break
            if val is None:
                continue
            b = val
        out.append((info.filename, b))
        if limit and len(out) >= limit:
            break
    return out


def run_decode(E, ref, limit, log):
    cap = 1 << 20                                   # every input is <= 31 KB; the same capacity for every decoder
    tot = acc = libz_both = 0
    unsup, unsup_valid = {}, {}
    t0 = time.time()
    for name in ("decode-corpus-raw.zip", "decode-corpus-encoded.zip"):
        items = read_corpus(os.path.join(ref, "zstd/testdata/fuzz", name), limit)
        for base in range(0, len(items), 200):
            grp = items[base:base + 200]
            want = [H.oracle_decode(b, cap) for _, b in grp]
            for form, maxb in (("per-block", 0), ("per-input", 4)):
                E.emu_set_dec_maxb(maxb)
                sizes, outs = emu_decode(E, [b for _, b in grp], [cap] * len(grp))
                for (nm, b), (ro, wb), r, got in zip(grp, want, sizes, outs):
                    if int(r) == -11 and ro != -11:
                        # documented deviation: a Huffman-weight FSE table with tableLog > 9 is answered
                        # "unsupported" at once; the reference reads on (and here always ends in an error of its own)
                        unsup[form] = unsup.get(form, 0) + 1
                        if ro >= 0:
                            unsup_valid[form] = unsup_valid.get(form, 0) + 1
                        continue
                    if int(r) != ro or (ro >= 0 and got != wb):
                        path = "/tmp/refcorpus_fail_%s.zst" % nm[:40]
                        open(path, "wb").write(b)
                        raise SystemExit("MISMATCH decode %s %s: oracle %d emu %d -> %s" % (name, form, ro, int(r), path))
            E.emu_set_dec_maxb(0)
            for (nm, b), (ro, wb) in zip(grp, want):
                tot += 1
                if ro >= 0:
                    acc += 1
                    z = H.libzstd_decode(b, cap)
                    if z is not None:
                        libz_both += 1
                        if z != wb:
                            raise SystemExit("MISMATCH oracle vs libzstd on %s/%s" % (name, nm))
        log("decode %-28s %5d inputs   (running: %d checked, %d accepted, %d also by libzstd with equal bytes; answered "
            "'unsupported' where the oracle says otherwise: %s, of those valid for the oracle: %s)  %.0f s"
            % (name, len(items), tot, acc, libz_both, unsup, unsup_valid, time.time() - t0))


def check_frames(E, what, inputs, frames):
    for i, (c, f) in enumerate(zip(inputs, frames)):
        f = bytes(f)
        n, got = H.oracle_decode(f, len(c) + 16)
        if n != len(c) or got != c or H.libzstd_decode(f, max(len(c), 1)) != c:
            open("/tmp/refcorpus_fail_enc.bin", "wb").write(c)
            raise SystemExit("MISMATCH %s input %d (%d bytes) -> /tmp/refcorpus_fail_enc.bin" % (what, i, len(c)))
    sizes, outs = emu_decode(E, [bytes(f) for f in frames], [len(c) + 16 for c in inputs])
    for i, c in enumerate(inputs):
        if int(sizes[i]) != len(c) or outs[i] != c:
            open("/tmp/refcorpus_fail_enc.bin", "wb").write(c)
            raise SystemExit("MISMATCH %s (emulated decoder) input %d -> /tmp/refcorpus_fail_enc.bin" % (what, i))


def run_encode(E, ref, limit, log):
    t0 = time.time()
    for name in ("encode-corpus-raw.zip", "encode-corpus-encoded.zip"):
        items = [b for _, b in read_corpus(os.path.join(ref, "zstd/testdata/fuzz", name), limit)]
        nchunk = nframe = 0
        for level, blk in ((1, 65536), (2, 131072), (3, 131072)):
            small = [b for b in items if len(b) <= blk]
            big = [b for b in items if len(b) > blk]
            for base in range(0, len(small), 64):
                grp = small[base:base + 64]
                check_frames(E, "%s chunks L%d" % (name, level), grp, emu_encode(E, grp, level=level)[0])
                nchunk += len(grp)
            for base in range(0, len(big), 8):
                grp = big[base:base + 8]
                check_frames(E, "%s frames L%d" % (name, level), grp, emu_encode_frames(E, grp, level=level, dump=False)[0])
                nframe += len(grp)
            log("encode %-28s level %d: %d one-block inputs, %d frame-mode inputs so far  %.0f s"
                % (name, level, nchunk, nframe, time.time() - t0))


def run_s2(E, ref, limit, log):
    t0 = time.time()
    for name in ("block-corpus-raw.zip", "block-corpus-enc.zip"):
        items = [b for _, b in read_corpus(os.path.join(ref, "s2/testdata/fuzz", name), limit)]
        blocks = []
        for b in items:
            blocks += [b[o:o + 65536] for o in range(0, max(len(b), 1), 65536)][:4]      # at most 256 KiB of each input
        done = 0
        for base in range(0, len(blocks), 64):
            grp = blocks[base:base + 64]
            for snappy in (False, True):
                for better in (False, True):
                    enc = emu_s2_encode(E, grp, snappy=snappy, better=better)[0]
                    sizes, outs, _, _ = emu_s2_decode(E, [bytes(e) for e in enc], [len(b) for b in grp])
                    for i, (b, e) in enumerate(zip(grp, enc)):
                        n, got = orc_s2_decode(bytes(e), len(b))
                        if n != len(b) or got != b or int(sizes[i]) != len(b) or outs[i] != b:
                            open("/tmp/refcorpus_fail_s2.bin", "wb").write(b)
                            raise SystemExit("MISMATCH s2 %s snappy=%d better=%d -> /tmp/refcorpus_fail_s2.bin" % (name, snappy, better))
            done += len(grp)
        log("s2     %-28s %d inputs, %d blocks x 4 modes  %.0f s" % (name, len(items), done, time.time() - t0))


def run_huff0(E, ref, limit, log):
    import ctypes
    import numpy as np
    from emu_util import emu_huf_compress, emu_huf_decompress
    from test_emu_huf0 import orc_compress, orc_decompress
    t0 = time.time()
    # FuzzCompress seeds + TestCompressRegression inputs: output bytes and result class equal the oracle's, 1X and 4X
    items = []
    for name in ("fse_compress.zip", "regression.zip"):
        items += [b[:262143] for _, b in read_corpus(os.path.join(ref, "huff0/testdata", name), limit) if len(b)]
    ok = rej = 0
    for four in (False, True):
        for base in range(0, len(items), 32):
            grp = items[base:base + 32]
            got = emu_huf_compress(E, grp, four)
            back = []
            for b, (comp, code) in zip(grp, got):
                wcomp, wcode = orc_compress(b, four)
                if comp != wcomp or (code if code < 0 else 0) != (wcode if wcode < 0 else 0):
                    open("/tmp/refcorpus_fail_huf.bin", "wb").write(b)
                    raise SystemExit("MISMATCH huff0 compress four=%d: emu %d oracle %d -> /tmp/refcorpus_fail_huf.bin" % (four, code, wcode))
                if code > 0:
                    back.append((comp, b)); ok += 1
                else:
                    rej += 1
            if back:
                dec = emu_huf_decompress(E, [c for c, _ in back], [len(b) for _, b in back], four)
                for (c, b), (out, code) in zip(back, dec):
                    if out != b or orc_decompress(c, len(b), four)[1] != b:
                        raise SystemExit("MISMATCH huff0 round trip four=%d" % four)
    log("huff0  compress: %d inputs x {1X, 4X}: %d compressed (bytes equal to the oracle's, decompressed back), %d rejected "
        "with the oracle's error class  %.0f s" % (len(items), ok, rej, time.time() - t0))
    # FuzzDecompress1x seeds + the 1X regression: ReadTable verdict, bytes consumed and code lengths equal the oracle's
    seeds = []
    for name in ("huff0_decompress1x.zip", "decompress1x_regression.zip"):
        seeds += [b for _, b in read_corpus(os.path.join(ref, "huff0/testdata", name), limit) if len(b)]

    class DT(ctypes.Structure):
        _fields_ = [("dt", ctypes.c_uint16 * 2048), ("actualTableLog", ctypes.c_uint), ("loaded", ctypes.c_int)]
    L = H.oracle()
    L.orc_huf_read_table.restype = ctypes.c_int64
    L.orc_huf_read_table.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t]
    n = len(seeds)
    stride = max(len(c) for c in seeds) + 16
    src = np.zeros((n, stride), dtype=np.uint8)
    for i, c in enumerate(seeds):
        src[i, :len(c)] = np.frombuffer(c, dtype=np.uint8)
    sizes = np.array([len(c) for c in seeds], dtype=np.uint32)
    rows = np.zeros((n, 260), dtype=np.uint8)
    outs = np.zeros(n, dtype=np.int64)
    E.emu_huf_read_table(src.ctypes.data, stride, sizes.ctypes.data, n, rows.ctypes.data, outs.ctypes.data)
    good = bad = unsup = 0
    for i, c in enumerate(seeds):
        d = DT()
        used = L.orc_huf_read_table(ctypes.byref(d), c, len(c))
        if outs[i] == -11 and used != -11:
            unsup += 1                       # weight table with tableLog > 9: the documented deviation
            continue
        if used < 0:
            if outs[i] >= 0:
                raise SystemExit("MISMATCH huff0 ReadTable seed %d: oracle rejects, emu %d" % (i, outs[i]))
            bad += 1
            continue
        want = [0] * 256
        for e in d.dt[: 1 << d.actualTableLog]:
            want[e >> 8] = e & 0xff
        if outs[i] != used or rows[i, 0] != d.actualTableLog or list(rows[i, 4:260]) != want:
            raise SystemExit("MISMATCH huff0 ReadTable seed %d: used %d vs %d" % (i, outs[i], used))
        good += 1
    log("huff0  ReadTable: %d seeds: %d tables equal to the oracle's, %d rejected by both, %d 'unsupported' (tableLog > 9)  %.0f s"
        % (n, good, bad, unsup, time.time() - t0))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", nargs="*", default=["decode", "encode", "s2", "huff0"])
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--limit", type=int, default=0)
    a = ap.parse_args()
    H.build_oracle()
    E = H.emu()

    def log(msg):
        print(msg, flush=True)
    if "decode" in a.what:
        run_decode(E, a.ref, a.limit, log)
    if "encode" in a.what:
        run_encode(E, a.ref, a.limit, log)
    if "s2" in a.what:
        run_s2(E, a.ref, a.limit, log)
    if "huff0" in a.what:
        run_huff0(E, a.ref, a.limit, log)
    print("clean")


if __name__ == "__main__":
    main()
