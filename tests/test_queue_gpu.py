"""The host boundary as its cited callers use it: concurrent one-block calls coalesced by the
library's queue (zstd/encoder.go:717-729 "can be called concurrently"; s2/writer.go:1052-1064 one customEnc call per
goroutine), and the packed call on ordinary pageable memory (a Go slice is not pinned)."""
import threading

import numpy as np
import pytest
import torch

import helpers as H
from test_oracle_s2 import s2_decode as orc_s2_decode

pytestmark = pytest.mark.gpu


def test_queue_16_concurrent_callers():
    from compress_b200 import zstd
    q = zstd.Queue(max_batch=512, linger_us=300)
    enc = zstd.Encoder(max_chunks=64)
    enc2 = zstd.Encoder(level=2, max_chunks=64)
    tw = H.golden("twain.txt")
    rng = np.random.default_rng(5)
    nthreads, per = 16, 24
    work = []
    for t in range(nthreads):
        items = []
        for k in range(per):
            n = int(rng.integers(0, 65537))
            o = int(rng.integers(0, len(tw) - 65536))
            items.append(tw[o:o + n])
        work.append(items)
    results = [[None] * per for _ in range(nthreads)]
    errors = []

    def run(t):
        try:
            for k, blk in enumerate(work[t]):
                kind = (t + k) % 4
                if kind == 0:
                    f = q.EncodeAll(blk)
                    results[t][k] = ("z1", f, q.DecodeAll(f, max_size=len(blk) + 64))
                elif kind == 1:
                    results[t][k] = ("z2", q.EncodeAll(blk, level=2), None)
                elif kind == 2:
                    e = q.S2Encode(blk)
                    results[t][k] = ("s2", e, q.S2Decode(e, max_size=len(blk) + 64))
                else:
                    results[t][k] = ("sn", q.S2Encode(blk, snappy=True), None)
        except Exception as ex:       # surfaced below: a thread's exception would otherwise vanish
            errors.append((t, repr(ex)))

    threads = [threading.Thread(target=run, args=(t,)) for t in range(nthreads)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    calls, batches = q.stats()
    assert calls == sum(2 if (t + k) % 2 == 0 else 1 for t in range(nthreads) for k in range(per))
    assert batches < calls, (calls, batches)      # the calls really were coalesced
    for t in range(nthreads):
        ref1 = enc.encode_chunks(work[t])
        ref2 = enc2.encode_chunks(work[t])
        for k, blk in enumerate(work[t]):
            kind, out, back = results[t][k]
            if kind == "z1":
                assert out == ref1[k] and back == blk
            elif kind == "z2":
                assert out == ref2[k]
            else:
                n, dec = orc_s2_decode(out, len(blk))
                assert n == len(blk) and dec == blk
                if back is not None:
                    assert back == blk
    # an input larger than the level's block size becomes one multi-block frame (frame mode), as EncodeAll's does
    big = tw[:200000]
    f = q.EncodeAll(big)
    assert f == enc.encode_frames([big])[0] and H.libzstd_decode(f, len(big)) == big
    assert q.DecodeAll(f, max_size=len(big) + 64) == big
    q.close(); enc.close(); enc2.close()


def test_packed_call_on_pageable_memory():
    """b2c_zstd_encode_packed on ordinary (pageable) numpy buffers gives the bytes of the pinned path."""
    from compress_b200 import zstd
    enc = zstd.Encoder(max_chunks=128)
    n, last = 700, 4321
    src = H.synth_text_torch(n * 65536, "cuda", seed=3)[: (n - 1) * 65536 + last].cpu()
    pinned = src.pin_memory()
    buf_p, total_p, sizes_p, offs_p = enc.encode_packed(pinned)
    pageable_in = src.numpy().copy()
    out = np.empty(pageable_in.size + n * 32 + 64, dtype=np.uint8)
    buf, total, sizes, offs = enc.encode_packed(torch.from_numpy(pageable_in), dst=torch.from_numpy(out))
    assert total == total_p and (sizes == sizes_p).all() and (offs == offs_p).all()
    assert bytes(out[:total]) == bytes(buf_p[:total_p].numpy())
    r, dec = H.oracle_decode(bytes(out[:total]), pageable_in.size + 64)
    assert r == pageable_in.size and dec == pageable_in.tobytes()
    # error path: a destination that is too small is reported, and the context stays usable
    small = torch.empty(1000, dtype=torch.uint8)
    with pytest.raises(zstd.B2CError):
        enc.encode_packed(pinned, dst=small)
    buf2, total2, _, _ = enc.encode_packed(pinned)
    assert total2 == total_p and bytes(buf2[:total2].numpy()) == bytes(buf_p[:total_p].numpy())
    enc.close()
