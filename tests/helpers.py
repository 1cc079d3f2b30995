"""Shared test/bench helpers: oracle loader (the CHECKER, never the product), libzstd loader,
deterministic synthetic corpora.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs import this module."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_SO = os.path.join(EMU_DIR, "libb2c_emu.so")


def build_oracle():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)
    return ORACLE_SO


def build_emu():
    subprocess.run(["make", "-s", "-C", EMU_DIR], check=True)
    return EMU_SO


_emu = None


def emu():
    """ctypes handle of the SIMT-emulated kernels (tests/emu; built on demand) with argument types set."""
    global _emu
    if _emu is not None:
        return _emu
    build_emu()
    E = ctypes.CDLL(EMU_SO)
    c = ctypes
    E.emu_zstd_encode.argtypes = [c.c_void_p, c.c_uint64, c.c_void_p, c.c_uint32, c.c_void_p, c.c_uint64, c.c_void_p,
                                  c.c_uint32, c.c_void_p, c.c_void_p, c.c_void_p, c.c_uint32]
    E.emu_zstd_encode_lv.argtypes = E.emu_zstd_encode.argtypes + [c.c_int, c.c_int]
    E.emu_zstd_decode.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p, c.c_uint32, c.c_void_p, c.c_void_p, c.c_void_p,
                                  c.c_void_p]
    E.emu_zstd_decode_mode.argtypes = E.emu_zstd_decode.argtypes + [c.c_int, c.c_void_p]
    E.emu_s2_encode.argtypes = [c.c_void_p, c.c_uint64, c.c_void_p, c.c_uint32, c.c_void_p, c.c_uint64, c.c_void_p, c.c_int]
    E.emu_s2_encode_lv.argtypes = E.emu_s2_encode.argtypes + [c.c_int, c.c_int]
    E.emu_s2_decode.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p, c.c_uint32, c.c_void_p, c.c_void_p, c.c_void_p,
                                c.c_void_p]
    E.emu_huf_compress.argtypes = [c.c_void_p, c.c_uint64, c.c_void_p, c.c_uint32, c.c_void_p, c.c_uint64, c.c_void_p, c.c_int]
    E.emu_huf_decompress.argtypes = [c.c_void_p, c.c_uint64, c.c_void_p, c.c_uint32, c.c_void_p, c.c_uint64, c.c_void_p,
                                     c.c_void_p, c.c_int]
    E.emu_huf_read_table.argtypes = [c.c_void_p, c.c_uint64, c.c_void_p, c.c_uint32, c.c_void_p, c.c_void_p]
    _emu = E
    return E


_oracle = None


def oracle():
    """ctypes handle of oracle/liboracle.so (built on demand)."""
    global _oracle
    if _oracle is not None:
        return _oracle
    if not os.path.exists(ORACLE_SO):
        build_oracle()
    L = ctypes.CDLL(ORACLE_SO)
    c = ctypes
    L.orc_xxh64.restype = c.c_uint64
    L.orc_xxh64.argtypes = [c.c_char_p, c.c_size_t, c.c_uint64]
    L.orc_zstd_decode_all.restype = c.c_int64
    L.orc_zstd_decode_all.argtypes = [c.c_char_p, c.c_size_t, c.c_void_p, c.c_size_t]
    L.orc_zstd_encode_all.restype = c.c_int64
    L.orc_zstd_encode_all.argtypes = [c.c_char_p, c.c_size_t, c.c_int, c.c_int, c.c_void_p, c.c_size_t]
    L.orc_zstd_max_encoded_size.restype = c.c_size_t
    L.orc_zstd_max_encoded_size.argtypes = [c.c_size_t, c.c_int, c.c_int]
    L.orc_zstd_encode_block.restype = c.c_int64
    L.orc_zstd_encode_block.argtypes = [c.c_char_p, c.c_size_t, c.c_char_p, c.c_size_t, c.c_void_p, c.c_size_t,
                                        c.c_int, c.c_void_p, c.c_size_t]
    L.orc_huf_compress_oneshot.restype = c.c_int64
    L.orc_huf_compress_oneshot.argtypes = [c.c_char_p, c.c_size_t, c.c_int, c.c_uint, c.c_void_p, c.c_size_t,
                                           c.c_void_p]
    L.orc_zstd_predef_table.restype = c.c_int
    L.orc_zstd_predef_table.argtypes = [c.c_int, c.c_void_p, c.c_int]
    _oracle = L
    return L


def oracle_decode(data, cap):
    L = oracle()
    out = ctypes.create_string_buffer(max(cap, 1))
    r = L.orc_zstd_decode_all(bytes(data), len(data), out, cap)
    return r, out.raw[: max(r, 0)]


def oracle_encode(data, level=1, crc=True):
    L = oracle()
    cap = L.orc_zstd_max_encoded_size(len(data), level, 1 if crc else 0) + 64
    out = ctypes.create_string_buffer(cap)
    r = L.orc_zstd_encode_all(bytes(data), len(data), level, 1 if crc else 0, out, cap)
    return r, out.raw[: max(r, 0)]


def oracle_encode_block(org, lits, triples, last=1):
    """blockEnc.encode on a fresh blockEnc for the given literals + (litLen, matchLen-3, offset)."""
    L = oracle()
    tri = np.ascontiguousarray(triples, dtype=np.uint32)
    cap = len(org) + 1024
    out = ctypes.create_string_buffer(cap)
    r = L.orc_zstd_encode_block(bytes(org), len(org), bytes(lits), len(lits), tri.ctypes.data, len(tri), last, out, cap)
    return r, out.raw[: max(r, 0)]


_libzstd = None


def libzstd():
    """System libzstd 1.5.5 (independent spec decoder; no header needed)."""
    global _libzstd
    if _libzstd is None:
        Z = ctypes.CDLL("libzstd.so.1")
        Z.ZSTD_decompress.restype = ctypes.c_size_t
        Z.ZSTD_decompress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
        Z.ZSTD_isError.restype = ctypes.c_uint
        Z.ZSTD_isError.argtypes = [ctypes.c_size_t]
        Z.ZSTD_compress.restype = ctypes.c_size_t
        Z.ZSTD_compress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
        Z.ZSTD_compressBound.restype = ctypes.c_size_t
        Z.ZSTD_compressBound.argtypes = [ctypes.c_size_t]
        _libzstd = Z
    return _libzstd


def libzstd_decode(data, n):
    Z = libzstd()
    out = ctypes.create_string_buffer(max(n, 1))
    r = Z.ZSTD_decompress(out, n, bytes(data), len(data))
    if Z.ZSTD_isError(r):
        return None
    return out.raw[:r]


def golden(name):
    with open(os.path.join(GOLDEN, name), "rb") as f:
        return f.read()


# ------------------------------------------------------------------ synthetic corpora
_LETTERS = np.frombuffer(b"etaoinshrdlcumwfgypbvkjxqz", dtype=np.uint8)
_LETTER_P = np.array([12.7, 9.1, 8.2, 7.5, 7.0, 6.7, 6.3, 6.1, 6.0, 4.3, 4.0, 2.8, 2.8, 2.4, 2.4, 2.2, 2.0, 2.0,
                      1.9, 1.5, 1.0, 0.8, 0.15, 0.15, 0.1, 0.07])
_LETTER_P = _LETTER_P / _LETTER_P.sum()


SYNTH_VOCAB = 32768      # words
SYNTH_ZIPF = 0.9         # exponent; calibrated so libzstd -1 on 64 KiB chunks gives ~0.43 (Twain: 0.45)
SYNTH_MAXW = 12
_SEP_P = [0.86, 0.05, 0.07, 0.02]  # " " / ", " / ". " / "\n"


def _synth_tables(seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    wl = rng.integers(2, SYNTH_MAXW + 1, size=SYNTH_VOCAB)
    W = rng.choice(_LETTERS, size=(SYNTH_VOCAB, SYNTH_MAXW), p=_LETTER_P)
    p = 1.0 / np.arange(1, SYNTH_VOCAB + 1) ** SYNTH_ZIPF
    cdf = np.cumsum(p / p.sum())
    return rng, wl, W, cdf


def synth_text(nbytes, seed=0x9E3779B9):
    """Deterministic enwik-like text: 32768-word vocabulary (English letter frequencies, lengths 2..12),
    Zipf(0.9) word choice, sentence punctuation and capitalisation.  numpy-vectorised."""
    rng, wl, W, cdf = _synth_tables(seed)
    ntok = int(nbytes / 6.0) + 1024
    out, total = [], 0
    sep_cdf = np.cumsum(_SEP_P)
    while total < nbytes:
        ids = np.minimum(np.searchsorted(cdf, rng.random(ntok)), SYNTH_VOCAB - 1)
        sepk = np.minimum(np.searchsorted(sep_cdf, rng.random(ntok)), 3)
        tok = np.zeros((ntok, SYNTH_MAXW + 2), dtype=np.uint8)
        tok[:, :SYNTH_MAXW] = W[ids]
        ln = wl[ids]
        cap = np.zeros(ntok, dtype=bool)
        cap[1:] = sepk[:-1] == 2          # capitalise the word following ". "
        tok[cap, 0] -= 32
        mask = np.arange(SYNTH_MAXW + 2)[None, :] < ln[:, None]
        rows = np.arange(ntok)
        tok[rows, ln] = np.array([32, 44, 46, 10], dtype=np.uint8)[sepk]
        mask[rows, ln] = True
        has2 = (sepk == 1) | (sepk == 2)
        tok[rows[has2], ln[has2] + 1] = 32
        mask[rows[has2], ln[has2] + 1] = True
        flat = tok[mask]
        out.append(flat)
        total += flat.size
    return np.concatenate(out)[:nbytes].tobytes()


def synth_text_torch(nbytes, device, seed=0x9E3779B9, piece=64 << 20):
    """Same construction on a torch device (bench.py builds 1 GiB this way in well under a second on
    a GPU).  Returns a uint8 tensor of nbytes.  Not bit-identical to synth_text (different RNG)."""
    import torch
    _, wl, W, cdf = _synth_tables(seed)
    g = torch.Generator(device=device)
    g.manual_seed(seed & 0x7FFFFFFF)
    wl_t = torch.from_numpy(wl.astype(np.int64)).to(device)
    W_t = torch.from_numpy(W.copy()).to(device)
    cdf_t = torch.from_numpy(cdf).to(device)
    sep_cdf = torch.tensor(np.cumsum(_SEP_P), device=device, dtype=torch.float64)
    sepch = torch.tensor([32, 44, 46, 10], dtype=torch.uint8, device=device)
    out = torch.empty(nbytes, dtype=torch.uint8, device=device)
    pos = 0
    while pos < nbytes:
        want = min(piece, nbytes - pos)
        ntok = int(want / 6.0) + 4096
        ids = torch.searchsorted(cdf_t, torch.rand(ntok, generator=g, device=device, dtype=torch.float64)).clamp_(max=SYNTH_VOCAB - 1)
        sepk = torch.searchsorted(sep_cdf, torch.rand(ntok, generator=g, device=device, dtype=torch.float64)).clamp_(max=3)
        tok = torch.zeros((ntok, SYNTH_MAXW + 2), dtype=torch.uint8, device=device)
        tok[:, :SYNTH_MAXW] = W_t[ids]
        ln = wl_t[ids]
        cap = torch.zeros(ntok, dtype=torch.bool, device=device)
        cap[1:] = sepk[:-1] == 2
        tok[cap, 0] -= 32
        mask = torch.arange(SYNTH_MAXW + 2, device=device)[None, :] < ln[:, None]
        rows = torch.arange(ntok, device=device)
        tok[rows, ln] = sepch[sepk]
        mask[rows, ln] = True
        has2 = (sepk == 1) | (sepk == 2)
        tok[rows[has2], ln[has2] + 1] = 32
        mask[rows[has2], ln[has2] + 1] = True
        flat = tok[mask]
        take = min(flat.numel(), nbytes - pos)
        out[pos:pos + take] = flat[:take]
        pos += take
    return out


def synth_chunks(kind, n, size=65536, seed=1):
    """n chunks of `size` bytes of a named synthetic kind (text / random / zeros / mixed)."""
    if kind == "text":
        base = synth_text(n * size, seed)
        return [base[i * size:(i + 1) * size] for i in range(n)]
    rng = np.random.Generator(np.random.PCG64(seed))
    if kind == "random":
        return [rng.integers(0, 256, size=size, dtype=np.uint8).tobytes() for _ in range(n)]
    if kind == "zeros":
        return [bytes(size) for _ in range(n)]
    raise ValueError(kind)
