"""More of the reference's own expectations, applied to the oracle and to the kernel sources under the emulator:
huff0's per-input error classes (huff0/compress_test.go:20-52), the zeros frames of zstd/testdata/large.zip
(TestNewDecoderLarge), the regression inputs of zstd and s2 (verdict of the oracle == verdict of the kernels).
CPU only; tests/test_*_gpu.py repeat the table checks on the device."""
import os
import zipfile

import numpy as np

import helpers as H
from emu_util import emu_decode, emu_huf_compress, emu_huf_decompress, emu_s2_decode
from test_emu_huf0 import orc_compress, orc_decompress
from test_oracle_s2 import s2_decode as orc_s2_decode

OK, INC, RLE = 0, -1, -2
# name -> (err1X, err4X)   [huff0/compress_test.go:20-52; inputs longer than BlockSizeMax are cut to it, :232]
HUF_TABLE = {
    "digits": (OK, OK), "gettysburg": (OK, OK), "twain": (OK, OK), "random": (INC, INC), "low-ent.10k": (OK, OK),
    "superlow-ent-10k": (OK, OK), "zeroes": (RLE, RLE), "crash1": (INC, INC), "crash2": (OK, INC), "crash3": (INC, INC),
    "endzerobits": (OK, INC), "endnonzero": (OK, INC), "case1": (OK, OK), "case2": (OK, OK), "case3": (OK, OK),
    "pngdata.001": (OK, OK), "normcount2": (OK, OK),
}


def huf_inputs():
    zf = zipfile.ZipFile(os.path.join(H.GOLDEN, "huff0_inputs.zip"))
    rd = lambda nm: zf.read(nm)
    d = {
        "digits": H.golden("e.txt"), "gettysburg": rd("gettysburg.txt"), "twain": H.golden("twain.txt")[:(1 << 18) - 1],
        "random": rd("sharnd.out"), "low-ent.10k": b"1221" * 10000, "superlow-ent-10k": b"1" * 10000 + b"2" * 500,
        "zeroes": bytes(10000), "crash1": rd("crash1.bin"), "crash2": rd("crash2.bin"), "crash3": rd("crash3.bin"),
        "endzerobits": rd("endzerobits.bin"), "endnonzero": rd("endnonzero.bin"), "case1": rd("case1.bin"),
        "case2": rd("case2.bin"), "case3": rd("case3.bin"), "pngdata.001": rd("pngdata.bin"), "normcount2": rd("normcount2.bin"),
    }
    assert set(d) == set(HUF_TABLE)
    return d


def _cls(code):
    return OK if code >= 0 else code


def test_huff0_error_table_oracle(oracle_lib):
    for name, data in huf_inputs().items():
        for four, want in ((False, HUF_TABLE[name][0]), (True, HUF_TABLE[name][1])):
            out, code = orc_compress(data, four)
            assert _cls(code) == want, (name, four, code)
            if code >= 0:   # and it must come back (TestCompress*/TestDecompress* round trips)
                wcode, back = orc_decompress(out, len(data), four)
                assert wcode == len(data) and back == data, name


def test_huff0_error_table_kernels(emu_lib, oracle_lib):
    inputs = huf_inputs()
    names = list(inputs)
    for four in (False, True):
        got = emu_huf_compress(emu_lib, [inputs[n] for n in names], four)
        for n, (out, code) in zip(names, got):
            assert _cls(code) == HUF_TABLE[n][1 if four else 0], (n, four, code)
            assert (out, code) == orc_compress(inputs[n], four), n
        okn = [n for n, (o, c) in zip(names, got) if c >= 0]
        back = emu_huf_decompress(emu_lib, [o for o, c in got if c >= 0], [len(inputs[n]) for n in okn], four)
        for n, (b, c) in zip(okn, back):
            assert c == len(inputs[n]) and b == inputs[n], n


def test_zstd_large_zeros(emu_lib, oracle_lib):
    zf = zipfile.ZipFile(os.path.join(H.GOLDEN, "zstd_large.zip"))
    for nm in zf.namelist():
        if not nm.endswith(".zst"):
            continue
        comp, size = zf.read(nm), int(zf.read(nm + ".size"))
        r, got = H.oracle_decode(comp, size + 8)
        assert r == size and got == bytes(size), nm
        if size <= (1 << 20):    # the 10 MiB frame is decoded on the device test only (slow under emulation)
            outs, res = emu_decode(emu_lib, [comp], [size + 8])
            assert outs[0] == size and res[0] == bytes(size), nm


def test_regression_inputs_same_verdict(emu_lib, oracle_lib):
    zf = zipfile.ZipFile(os.path.join(H.GOLDEN, "zstd_decode_regression.zip"))
    blobs = [zf.read(n) for n in zf.namelist()]
    outs, res = emu_decode(emu_lib, blobs, [4 << 20] * len(blobs))
    for b, r, got in zip(blobs, outs, res):
        ro, want = H.oracle_decode(b, 4 << 20)
        assert ro == r and (ro < 0 or got == want)
    zf = zipfile.ZipFile(os.path.join(H.GOLDEN, "s2_dec_block_regressions.zip"))
    blobs = [zf.read(n) for n in zf.namelist()]
    outs, res, _, _ = emu_s2_decode(emu_lib, blobs, [1 << 20] * len(blobs))
    for b, r, got in zip(blobs, outs, res):
        ro, want = orc_s2_decode(b, 1 << 20)
        assert ro == r and (ro < 0 or got == want)
