#!/usr/bin/env python3
"""Regenerates tests/golden/ from the reference's own test data (run in the dev container where
/root/reference is mounted).  Only DATA fixtures are copied -- golden decoder vectors and plain
text corpora the reference's tests use -- never source code.  The GPU box has no /root/reference, so
everything the -m gpu tests, smoke() and bench.py need lives here.

  zstd_good.zip / zstd_bad.zip   zstd/testdata/good.zip, bad.zip (TestNewDecoderGood/Bad, decoder_test.go:393-455)
  zstd_decoder_subset.zip        the smaller pairs of zstd/testdata/decoder.zip (TestNewDecoder, :201-216)
  twain.txt, html.txt, e.txt     testdata/* corpora used by the encoder round-trip tests
  s2_twain.txt(.rawsnappy)       s2/testdata golden Snappy block (TestDecodeGoldenInput, s2_test.go:599)
  huff0_inputs.zip               inputs of huff0's compress tests whose error class is tabulated (huff0/compress_test.go:20-52)
  zstd_large.zip                 the .zst members of zstd/testdata/large.zip (TestNewDecoderLarge; contents are zeros)
  zstd_decode_regression.zip     zstd/testdata/decode-regression.zip
  s2_dec_block_regressions.zip   s2/testdata/dec-block-regressions.zip (TestDecodeRegression, s2/decode_test.go:19)
"""
import io, os, shutil, zipfile

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    shutil.copy(f"{REF}/zstd/testdata/good.zip", f"{HERE}/zstd_good.zip")
    shutil.copy(f"{REF}/zstd/testdata/bad.zip", f"{HERE}/zstd_bad.zip")
    shutil.copy(f"{REF}/testdata/Mark.Twain-Tom.Sawyer.txt", f"{HERE}/twain.txt")
    shutil.copy(f"{REF}/testdata/html.txt", f"{HERE}/html.txt") if os.path.exists(f"{REF}/testdata/html.txt") else None
    shutil.copy(f"{REF}/testdata/e.txt", f"{HERE}/e.txt")
    shutil.copy(f"{REF}/s2/testdata/Mark.Twain-Tom.Sawyer.txt", f"{HERE}/s2_twain.txt")
    shutil.copy(f"{REF}/s2/testdata/Mark.Twain-Tom.Sawyer.txt.rawsnappy", f"{HERE}/s2_twain.txt.rawsnappy")
    zf = zipfile.ZipFile(f"{REF}/zstd/testdata/decoder.zip")
    pairs = []
    for nm in zf.namelist():
        if nm.endswith(".zst") and nm[:-4] in zf.namelist():
            pairs.append((zf.getinfo(nm).file_size + zf.getinfo(nm[:-4]).compress_size, nm))
    pairs.sort()
    out = zipfile.ZipFile(f"{HERE}/zstd_decoder_subset.zip", "w", zipfile.ZIP_DEFLATED, compresslevel=9)
    total = 0
    for sz, nm in pairs:
        if total + sz > 700_000:
            break
        out.writestr(nm, zf.read(nm))
        out.writestr(nm[:-4], zf.read(nm[:-4]))
        total += sz
    out.close()
    hz = zipfile.ZipFile(f"{HERE}/huff0_inputs.zip", "w", zipfile.ZIP_DEFLATED, compresslevel=9)
    for nm in ("gettysburg.txt", "sharnd.out", "crash1.bin", "crash2.bin", "crash3.bin", "endzerobits.bin", "endnonzero.bin",
               "case1.bin", "case2.bin", "case3.bin", "pngdata.bin", "normcount2.bin"):
        hz.writestr(nm, open(f"{REF}/testdata/{nm}", "rb").read())
    hz.close()
    lz = zipfile.ZipFile(f"{REF}/zstd/testdata/large.zip")
    oz = zipfile.ZipFile(f"{HERE}/zstd_large.zip", "w", zipfile.ZIP_DEFLATED, compresslevel=9)
    for nm in lz.namelist():
        if nm.endswith(".zst"):
            oz.writestr(nm, lz.read(nm))
            oz.writestr(nm + ".size", str(lz.getinfo(nm[:-4]).file_size))   # the expected output is that many zero bytes
    oz.close()
    shutil.copy(f"{REF}/zstd/testdata/decode-regression.zip", f"{HERE}/zstd_decode_regression.zip")
    shutil.copy(f"{REF}/s2/testdata/dec-block-regressions.zip", f"{HERE}/s2_dec_block_regressions.zip")
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
