"""Scores parse variants of tests/model/parse_model.c with the oracle's entropy stage (design tool, CPU only)."""
import ctypes, os, subprocess, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import helpers as H
from emu_util import frame_header_len

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libparse_model.so")
subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-o", SO, os.path.join(HERE, "parse_model.c")], check=True)
M = ctypes.CDLL(SO)

class Cfg(ctypes.Structure):
    _fields_ = [(k, ctypes.c_uint32) for k in
                "rangeBytes shortBits shortMls longBits longMls insStride lazy repProbe repCodes joinCont minLong minShort maxDist hist latest tile local lag probeStride both dyn".split()]

def mk(**kw):
    d = dict(rangeBytes=68, shortBits=15, shortMls=6, longBits=0, longMls=8, insStride=1, lazy=0, repProbe=0,
             repCodes=1, joinCont=0, minLong=4, minShort=4, maxDist=0, hist=0, latest=0, tile=0, local=0, lag=0, probeStride=0, both=0, dyn=0)
    d.update(kw)
    return Cfg(**d)

def model_size(chunk, cfg):
    n = len(chunk)
    buf = chunk + bytes(32)
    tri = np.zeros((n // 4 + 16, 3), dtype=np.uint32)
    lits = ctypes.create_string_buffer(n + 64)
    nl = ctypes.c_uint32(0)
    ns = M.pm_parse(buf, n, ctypes.byref(cfg), tri.ctypes.data, lits, ctypes.byref(nl))
    if ns == 0:
        return n + 3 + frame_header_len(n) + 4
    r, ob = H.oracle_encode_block(chunk, lits.raw[:nl.value], tri[:ns], 1)
    assert r > 0
    return r + frame_header_len(n) + 4

def corpora(block):
    out = {}
    for name in ["twain.txt", "html.txt", "e.txt"]:
        d = H.golden(name)
        out[name] = [d[i:i + block] for i in range(0, len(d), block)]
    t = H.synth_text(16 * block, seed=1234)
    out["synth"] = [t[i:i + block] for i in range(0, len(t), block)]
    return out

def main():
    block = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    C = corpora(block)
    variants = eval(open(sys.argv[2]).read()) if len(sys.argv) > 2 else {"r1": mk()}
    print("block", block)
    base = {}
    for name, chunks in C.items():
        tot = sum(len(c) for c in chunks)
        o = [sum(len(H.oracle_encode(c, level=l)[1]) for c in chunks) for l in (1, 2, 3)]
        base[name] = o
        print(f"{name:10s} in={tot:8d} oracle L1={o[0]:8d} L2={o[1]:8d} ({o[1]/o[0]-1:+.3%}) L3={o[2]:8d} ({o[2]/o[0]-1:+.3%})")
    for vn, cfg in variants.items():
        row = []
        for name, chunks in C.items():
            s = sum(model_size(c, cfg) for c in chunks)
            row.append(f"{name[:6]} {s:8d} vsL1 {s/base[name][0]-1:+.2%} vsL2 {s/base[name][1]-1:+.2%}")
        print(f"{vn:28s} " + " | ".join(row))

if __name__ == "__main__":
    main()
