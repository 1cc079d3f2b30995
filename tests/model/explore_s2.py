"""Scores parse variants of tests/model/parse_model.c with the S2 block format's byte costs (design tool, CPU only)."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import helpers as H
from test_oracle_s2 import s2_encode as orc_s2_encode
import explore as X


def lit_hdr(ll):
    return 0 if ll == 0 else (1 if ll <= 60 else (2 if ll <= 256 else 3))


def rep_size(off, ln):
    ln -= 4
    if ln <= 4: return 2
    if ln < 8 and off < 2048: return 2
    if ln < (1 << 8) + 4: return 3
    if ln < (1 << 16) + (1 << 8): return 4
    return 5


def copy_size(off, ln):
    if ln > 64:
        return (2 + rep_size(off, ln - 8)) if off < 2048 else (3 + rep_size(off, ln - 60))
    return 3 if (ln >= 12 or off >= 2048) else 2


def s2_size(chunk, cfg):
    n = len(chunk)
    buf = chunk + bytes(32)
    tri = np.zeros((n // 4 + 16, 3), dtype=np.uint32)
    lits = ctypes.create_string_buffer(n + 64)
    nl = ctypes.c_uint32(0)
    ns = X.M.pm_parse(buf, n, ctypes.byref(cfg), tri.ctypes.data, lits, ctypes.byref(nl))
    hdr = 1 if n < 128 else (2 if n < 16384 else 3)
    body, pos, prev = 0, 0, 0
    for i in range(ns):
        ll, ml, ofv = int(tri[i, 0]), int(tri[i, 1]) + 3, int(tri[i, 2])
        # repCodes=1 model output: ofv == 1 means "same offset as the previous sequence"
        off = prev if ofv == 1 else ofv - 3
        body += lit_hdr(ll) + ll
        body += rep_size(off, ml) if (i > 0 and off == prev) else copy_size(off, ml)
        prev = off
        pos += ll + ml
    tl = n - pos
    body += lit_hdr(tl) + tl
    if ns == 0 or body > n - (n >> 5) - 5:
        return hdr + lit_hdr(n) + n
    return hdr + body


def main():
    block = 65536
    C = X.corpora(block)
    variants = eval(open(sys.argv[1]).read(), {"mk": X.mk})
    base = {}
    for name, chunks in C.items():
        o = [sum(len(orc_s2_encode(c, m)) for c in chunks) for m in (0, 1)]
        base[name] = o
        print(f"{name:10s} oracle s2 {o[0]:8d} better {o[1]:8d} ({o[1]/o[0]-1:+.2%})")
    for vn, cfg in variants.items():
        row = []
        for name, chunks in C.items():
            s = sum(s2_size(c, cfg) for c in chunks)
            row.append(f"{name[:6]} {s:8d} vsS2 {s/base[name][0]-1:+.2%} vsBetter {s/base[name][1]-1:+.2%}")
        print(f"{vn:30s} " + " | ".join(row))


if __name__ == "__main__":
    main()
