#!/usr/bin/env python3
"""Regenerates tests/golden/ from the reference's own test data (run in the dev container where
/root/reference is mounted).  Only DATA fixtures are copied -- golden decoder vectors and plain
text corpora the reference's tests use -- never source code.  The GPU box has no /root/reference, so
everything the -m gpu tests, smoke() and bench.py need lives here.

  zstd_good.zip / zstd_bad.zip   zstd/testdata/good.zip, bad.zip (TestNewDecoderGood/Bad, decoder_test.go:393-455)
  zstd_decoder.zip               zstd/testdata/decoder.zip, all 94 pairs (TestNewDecoder, :201-216)
  zstd_comp_crashers.zip         zstd/testdata/comp-crashers.zip: 1657 inputs that once broke an encoder (TestEncoderRegression, encoder_test.go:218)
  zstd_seqs.zip / _want.zip      sequence-decoder golden vectors (Test_seqdec_decoder, seqdec_test.go:199-302)
  zstd_headers.zip / -want.json.zst  frame-header golden vectors (TestHeader_Decode, decodeheader_test.go)
  s2_enc_regressions.zip         s2/testdata/enc_regressions.zip (TestEncoderRegression, s2/s2_test.go:2133)
  twain.txt, html.txt, e.txt     testdata/* corpora used by the encoder round-trip tests
  s2_twain.txt(.rawsnappy)       s2/testdata golden Snappy block (TestDecodeGoldenInput, s2_test.go:599)
  huff0_inputs.zip               inputs of huff0's compress tests whose error class is tabulated (huff0/compress_test.go:20-52)
  zstd_large.zip                 the .zst members of zstd/testdata/large.zip (TestNewDecoderLarge; contents are zeros)
  zstd_decode_regression.zip     zstd/testdata/decode-regression.zip
  s2_dec_block_regressions.zip   s2/testdata/dec-block-regressions.zip (TestDecodeRegression, s2/decode_test.go:19)
  zstd_regression.zip            zstd/testdata/regression.zip: 36 decoder regression inputs (TestDecoderRegression, decoder_test.go:682)
  zstd_benchdecoder.zip          zstd/testdata/benchdecoder.zip: 12 .zst files (TestDecoderMultiFrame :911, TestDecoder_Reset :964, benchmarks)
  zstd_z000028, zstd_z000028.zst the decoded / encoded pair of TestPredefTables / TestDecoderDrain (decoder_test.go:539-624, :845)
  zstd_xml.zst                   zstd/testdata/xml.zst (5 345 280 bytes decoded; benchmarks)
  zstd_fuzz_decode_subset.zip    every 16th entry of zstd/testdata/fuzz/decode-corpus-raw.zip (the reference's FuzzDecodeAll seeds,
                                 zstd/fuzz_test.go:17-19; the whole corpora run through tools/fuzz_ref_corpora.py here)
"""
import io, os, shutil, zipfile

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def fuzz_subset():
    src = zipfile.ZipFile(f"{REF}/zstd/testdata/fuzz/decode-corpus-raw.zip")
    out = zipfile.ZipFile(f"{HERE}/zstd_fuzz_decode_subset.zip", "w", zipfile.ZIP_DEFLATED, compresslevel=9)
    for k, info in enumerate(sorted(src.infolist(), key=lambda i: i.filename)):
        if k % 16 == 0 and not info.is_dir():
            zi = zipfile.ZipInfo(info.filename, date_time=(2020, 1, 1, 0, 0, 0))     # fixed stamp: the file is reproducible
            zi.compress_type = zipfile.ZIP_DEFLATED
            out.writestr(zi, src.read(info))
    out.close()


def main():
    fuzz_subset()
    shutil.copy(f"{REF}/zstd/testdata/good.zip", f"{HERE}/zstd_good.zip")
    shutil.copy(f"{REF}/zstd/testdata/bad.zip", f"{HERE}/zstd_bad.zip")
    shutil.copy(f"{REF}/testdata/Mark.Twain-Tom.Sawyer.txt", f"{HERE}/twain.txt")
    shutil.copy(f"{REF}/testdata/html.txt", f"{HERE}/html.txt") if os.path.exists(f"{REF}/testdata/html.txt") else None
    shutil.copy(f"{REF}/testdata/e.txt", f"{HERE}/e.txt")
    shutil.copy(f"{REF}/s2/testdata/Mark.Twain-Tom.Sawyer.txt", f"{HERE}/s2_twain.txt")
    shutil.copy(f"{REF}/s2/testdata/Mark.Twain-Tom.Sawyer.txt.rawsnappy", f"{HERE}/s2_twain.txt.rawsnappy")
    shutil.copy(f"{REF}/zstd/testdata/decoder.zip", f"{HERE}/zstd_decoder.zip")
    shutil.copy(f"{REF}/zstd/testdata/comp-crashers.zip", f"{HERE}/zstd_comp_crashers.zip")
    shutil.copy(f"{REF}/zstd/testdata/seqs.zip", f"{HERE}/zstd_seqs.zip")
    shutil.copy(f"{REF}/zstd/testdata/seqs-want.zip", f"{HERE}/zstd_seqs_want.zip")
    shutil.copy(f"{REF}/zstd/testdata/headers.zip", f"{HERE}/zstd_headers.zip")
    shutil.copy(f"{REF}/zstd/testdata/headers-want.json.zst", f"{HERE}/zstd_headers-want.json.zst")
    shutil.copy(f"{REF}/s2/testdata/enc_regressions.zip", f"{HERE}/s2_enc_regressions.zip")
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
    shutil.copy(f"{REF}/zstd/testdata/regression.zip", f"{HERE}/zstd_regression.zip")
    shutil.copy(f"{REF}/zstd/testdata/benchdecoder.zip", f"{HERE}/zstd_benchdecoder.zip")
    shutil.copy(f"{REF}/zstd/testdata/z000028", f"{HERE}/zstd_z000028")
    shutil.copy(f"{REF}/zstd/testdata/z000028.zst", f"{HERE}/zstd_z000028.zst")
    shutil.copy(f"{REF}/zstd/testdata/xml.zst", f"{HERE}/zstd_xml.zst")
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
