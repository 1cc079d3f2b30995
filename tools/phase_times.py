#!/usr/bin/env python3
"""Per-phase cycle breakdown of the zstd encode kernel (clock64 stamps of thread 0 of each CTA)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import helpers as H
from compress_b200 import zstd
from compress_b200._lib import lib, check

NAMES = ["load", "Ebuild", "parse", "layout+gather", "hist", "huf stats/sort", "tree+fse tables", "bits",
         "vals+write+chains", "lit sizes", "seq sizes", "zero+lit pack", "seq pack", "headers", "writeback"]

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 8
    enc = zstd.Encoder()
    src = H.synth_text_torch(n * 65536, "cuda", seed=42)
    dst = torch.empty((n, zstd.SLOT), dtype=torch.uint8, device="cuda")
    outs = torch.empty(n, dtype=torch.int64, device="cuda")
    cyc = torch.zeros((n, 16, 32), dtype=torch.int64, device="cuda")
    for _ in range(2):
        rc = lib.b2c_zstd_encode_device_timed(enc._ctx, 3, src.data_ptr(), 65536, 65536, dst.data_ptr(), zstd.SLOT,
                                              outs.data_ptr(), n, cyc.data_ptr(), None)
        check(rc, enc._ctx)
        torch.cuda.synchronize()
    c = cyc.cpu().numpy().astype(np.int64)       # [n, 16, 32] arrival of warp w at the barrier before stamp k
    rel = c[:, :6, :].max(axis=2)                  # release time of each barrier (K1 has stamps 0..5)
    tot = rel[:, 5] - c[:, 0, :].min(axis=1)
    print("K1 parse kernel: chunks", n, "mean cycles/chunk %.0f" % tot.mean(), "min", tot.min(), "max", tot.max())
    names = ["load", "Ebuild", "parse", "layout+gather+rle", "hist+writeout"]
    for k in range(5):
        dur = rel[:, k + 1] - rel[:, k]
        work = c[:, k + 1, :] - rel[:, k][:, None]          # per-warp busy time in this phase
        wmean = work.mean(axis=0)
        top = np.argsort(-wmean)[:3]
        print("%-26s %9.0f %5.1f%%   busy: mean %7.0f  slowest warps %s" % (
            names[k], dur.mean(), 100 * dur.mean() / tot.mean(), work.mean(),
            ", ".join("w%d=%.0f" % (w, wmean[w]) for w in top)))


    # finer stamps (per-warp arrival, no barrier implied): mean over chunks and warps of the time between stamps
    chain = [(2, 6, "per-thread scan (until the warp leaves the loop)"), (6, 3, "wait + long-match resolution"),
             (3, 8, "prefix-max of match ends"), (8, 9, "trim loop"), (9, 10, "two block scans"),
             (10, 11, "emit loop (literal gather, codes, stores)"), (11, 4, "tail literals + RLE test barriers"),
             (4, 12, "zero counters + barrier"), (12, 13, "histograms + literal copy-out"), (13, 14, "barrier wait"),
             (14, 5, "reduce + maxSym + end")]
    print("-- detail (mean cycles per warp; stamps are arrival times) --")
    for a, b, nm in chain:
        d = (c[:, b, :] - c[:, a, :]).astype(np.float64)
        print("%-52s mean %8.0f   max-warp mean %8.0f" % (nm, d.mean(), d.mean(axis=0).max()))


if __name__ == "__main__":
    main()
