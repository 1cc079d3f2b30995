#!/usr/bin/env python3
"""bench.py -- zstd SpeedFastest encode throughput (input GB/s) on N B200s, BASELINE.json's metric.

A step = one pass of the hot path over one batch: 16384 synthetic-text chunks of 64 KiB (1 GiB) per GPU
(BASELINE config 2).  `value` is timed with CUDA events around K launches with inputs resident in HBM
(inputs are 8x larger than L2, so no cache flush is needed); `e2e` goes through the host-buffer C-ABI call
with pinned host input/output (H2D + encode + D2H inside the timed region).  Multi-GPU: one process per
GPU, each with its own 1 GiB (weak scaling), no collective on the data path; time = max over ranks.
`--impl reference` times the CPU oracle (the restatement of the reference's Go path; Go is not installed
here) on all host cores over a bounded sample of the same workload.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CHUNK = 65536
NCHUNKS = 16384  # 1 GiB per GPU
METRIC = "zstd SpeedFastest encode GB/s (input)"


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop = False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        import statistics
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [names[k] for k in range(4) if any(len(r) > 2 + k and r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def host_cores():
    """Host cores this process may really use: CPU affinity capped by the cgroup CPU quota."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = max(1, min(n, int(float(q) / float(per) + 0.999)))
    except Exception:
        pass
    return n


def cpu_reference_rate(sample_chunks, threads, seconds=0.0):
    """Oracle (C restatement of zstd.Encoder.EncodeAll, SpeedFastest) on `threads` host threads, one pooled encoder
    per thread (zstd/encoder.go:90-99); each thread loops over its share of the sample inside one C call until
    `seconds` have passed.  Returns (GB/s, wall seconds, ratio)."""
    import helpers as H
    H.build_oracle()
    L = H.oracle()
    from concurrent.futures import ThreadPoolExecutor
    c = ctypes
    cap = L.orc_zstd_max_encoded_size(CHUNK, 1, 1) + 64
    L.orc_zstd_cctx_new.restype = c.c_void_p
    L.orc_zstd_bench_chunks.restype = c.c_int64
    L.orc_zstd_bench_chunks.argtypes = [c.c_void_p, c.c_void_p, c.c_size_t, c.c_size_t, c.c_int, c.c_int, c.c_void_p,
                                        c.c_size_t, c.c_double, c.c_void_p]
    assert all(len(x) == CHUNK for x in sample_chunks)
    threads = max(1, min(threads, len(sample_chunks)))
    blob = np.frombuffer(b"".join(sample_chunks), dtype=np.uint8)
    bufs = [c.create_string_buffer(cap) for _ in range(threads)]
    ctxs = [L.orc_zstd_cctx_new() for _ in range(threads)]
    per = len(sample_chunks) // threads
    done = (c.c_uint64 * threads)()

    def work(t):
        lo = t * per
        cnt = per if t < threads - 1 else len(sample_chunks) - lo
        r = L.orc_zstd_bench_chunks(ctxs[t], blob.ctypes.data + lo * CHUNK, CHUNK, cnt, 1, 1, bufs[t], cap,
                                    float(seconds), c.byref(done, 8 * t))  # ctypes releases the GIL
        assert r > 0
        return r
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        outs = list(ex.map(work, range(threads)))
    dt = time.perf_counter() - t0
    nbytes = sum(int(d) for d in done)
    return nbytes / dt / 1e9, dt, sum(outs) / (len(sample_chunks) * CHUNK)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nchunks", type=int, default=NCHUNKS)
    ap.add_argument("--e2e-chunks", type=int, default=8288, help="chunks per e2e step (7 batches of 8 x 148 chunks, 518 MiB)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warm = max(args.warmup, 3)
    config = {"workload": "zstd SpeedFastest (level 1), %d x 64 KiB independent chunks of synthetic enwik-like text per GPU, "
                          "one frame per chunk, CRC on" % args.nchunks,
              "chunk_bytes": CHUNK, "chunks_per_gpu": args.nchunks, "l2": "inputs (1 GiB/GPU) larger than L2, no flush needed",
              "parallelism": "chunks sharded over %d GPU(s), no collective" % world}

    import helpers as H
    nthreads = host_cores()

    if args.impl == "reference":
        if rank != 0:
            return
        # bounded sample: each step = every host thread encoding its share of 2048 chunks (128 MiB) of the same
        # synthetic text over and over for ~4 s
        sample = H.synth_chunks("text", 2048, seed=77)
        steps = max(1, args.steps)
        for _ in range(min(args.warmup, 1)):
            cpu_reference_rate(sample, nthreads, 0.5)
        rates = [cpu_reference_rate(sample, nthreads, 4.0) for _ in range(steps)]
        gbs = sum(r[0] * r[1] for r in rates) / sum(r[1] for r in rates)
        line = {"metric": METRIC, "value": gbs, "unit": "GB/s", "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * sum(r[1] for r in rates) / steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic", "impl": "reference", "config": config,
                "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": nthreads, "kind": "port",
                                 "sample": "2048 distinct 64 KiB chunks, re-encoded for ~4 s per step; oracle C restatement "
                                           "of the Go encoder (Go toolchain absent), one EncodeAll per chunk on a pooled "
                                           "encoder, %d threads" % nthreads,
                                 "ratio": rates[-1][2]},
                "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from compress_b200 import zstd
    enc = zstd.Encoder(device=local_rank, max_chunks=1184)  # host-path batch: 8 chunks per SM
    n = args.nchunks
    src = H.synth_text_torch(n * CHUNK, dev, seed=1000 + rank)
    dst = torch.empty((n, zstd.SLOT), dtype=torch.uint8, device=dev)
    outs = torch.empty((n,), dtype=torch.int64, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warm):
        enc.encode_device(src, dst=dst, out_sizes=outs)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = enc.launches
    enc.profile(True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for k in range(args.steps):
        enc.encode_device(src, dst=dst, out_sizes=outs)
        ev[k + 1].record()
    barrier()
    step_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    total_ms = ev[0].elapsed_time(ev[args.steps])
    kernel_launches = enc.launches - launches0
    kms, kcalls = enc.profile_read()
    enc.profile(False)
    outs_h = outs.cpu().numpy()
    assert (outs_h > 0).all(), "encode error"
    out_bytes = int(outs_h.sum())

    # ---- secondary: decode of the frames just produced (SURVEY 8d "decode GB/s (output bytes)"), device-resident
    dec = zstd.Decoder(device=local_rank)
    dsz = outs.to(torch.int32)
    dout = torch.empty((n, CHUNK), dtype=torch.uint8, device=dev)
    dres = torch.empty((n,), dtype=torch.int64, device=dev)
    for _ in range(2):
        dec.decode_device(dst, dsz, src_stride=zstd.SLOT, dst=dout, dst_cap=CHUNK, out_sizes=dres)
    torch.cuda.synchronize()
    d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d0.record()
    for _ in range(3):
        dec.decode_device(dst, dsz, src_stride=zstd.SLOT, dst=dout, dst_cap=CHUNK, out_sizes=dres)
    d1.record()
    torch.cuda.synchronize()
    dec_ms = d0.elapsed_time(d1) / 3
    assert bool((dres == CHUNK).all()) and torch.equal(dout.view(-1), src), "decode mismatch"
    del dout

    # ---- secondary (BASELINE config 3): S2 / Snappy block encode + decode of the same chunks, device-resident
    from compress_b200 import s2 as s2mod
    s2c = s2mod.Codec(device=local_rank)
    s2res = {}
    s2dst = torch.empty((n, s2mod.SLOT), dtype=torch.uint8, device=dev)
    s2sz = torch.empty((n,), dtype=torch.int64, device=dev)
    dout = torch.empty((n, CHUNK), dtype=torch.uint8, device=dev)
    for name, snappy in (("s2", False), ("snappy", True)):
        for _ in range(2):
            s2c.encode_device(src, snappy=snappy, dst=s2dst, out_sizes=s2sz)
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(3):
            s2c.encode_device(src, snappy=snappy, dst=s2dst, out_sizes=s2sz)
        a1.record()
        torch.cuda.synchronize()
        ems = a0.elapsed_time(a1) / 3
        s2out = int(s2sz.sum())
        s2i = s2sz.to(torch.int32)
        for _ in range(2):
            s2c.decode_device(s2dst, s2i, src_stride=s2mod.SLOT, dst=dout, dst_cap=CHUNK, out_sizes=dres)
        torch.cuda.synchronize()
        a0.record()
        for _ in range(3):
            s2c.decode_device(s2dst, s2i, src_stride=s2mod.SLOT, dst=dout, dst_cap=CHUNK, out_sizes=dres)
        a1.record()
        torch.cuda.synchronize()
        dms = a0.elapsed_time(a1) / 3
        assert bool((dres == CHUNK).all()) and torch.equal(dout.view(-1), src), "s2 decode mismatch"
        s2res[name] = {"encode_gbs": n * CHUNK / (ems / 1e3) / 1e9, "encode_ms": ems, "ratio": s2out / (n * CHUNK),
                       "decode_gbs": n * CHUNK / (dms / 1e3) / 1e9, "decode_ms": dms,
                       "encode_roofline_frac": (n * CHUNK + s2out) / (ems / 1e3) / 1e9 / hbm_peak()[0]}
    del dout, s2dst

    # ---- secondary (BASELINE config 4): standalone huff0 Compress4X / Decompress4X, 262143-byte blocks of the same text
    from compress_b200 import huff0 as hufmod
    hc = hufmod.Codec(device=local_rank)
    hb, hstride = 262143, 262144
    hn = (n * CHUNK) // hstride
    hsz = torch.full((hn,), hb, dtype=torch.int32, device=dev)
    hdst = torch.empty((hn, hstride), dtype=torch.uint8, device=dev)
    hout = torch.empty((hn,), dtype=torch.int64, device=dev)
    for _ in range(2):
        hc.compress_device(src, hstride, hsz, True, dst=hdst, out_sizes=hout)
    torch.cuda.synchronize()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for _ in range(3):
        hc.compress_device(src, hstride, hsz, True, dst=hdst, out_sizes=hout)
    a1.record()
    torch.cuda.synchronize()
    hc_ms = a0.elapsed_time(a1) / 3
    assert int(hout.min()) > 0
    hcsz = hout.to(torch.int32)
    hdec = torch.empty((hn, hstride), dtype=torch.uint8, device=dev)
    hres = torch.empty((hn,), dtype=torch.int64, device=dev)
    for _ in range(2):
        hc.decompress_device(hdst.view(-1), hstride, hcsz, hsz, hstride, True, dst=hdec, out_sizes=hres)
    torch.cuda.synchronize()
    a0.record()
    for _ in range(3):
        hc.decompress_device(hdst.view(-1), hstride, hcsz, hsz, hstride, True, dst=hdec, out_sizes=hres)
    a1.record()
    torch.cuda.synchronize()
    hd_ms = a0.elapsed_time(a1) / 3
    assert bool((hres == hb).all()) and torch.equal(hdec[:, :hb], src.view(hn, hstride)[:, :hb]), "huff0 mismatch"
    huf_res = {"blocks": hn, "block_bytes": hb, "compress4x_gbs": hn * hb / (hc_ms / 1e3) / 1e9, "compress_ms": hc_ms,
               "ratio": float(hout.sum()) / (hn * hb), "decompress4x_gbs": hn * hb / (hd_ms / 1e3) / 1e9, "decompress_ms": hd_ms}
    del hdec, hdst

    # ---- end to end through the host-buffer C-ABI call (pinned host in/out)
    ne = min(args.e2e_chunks, n)
    host_in = src[: ne * CHUNK].cpu().pin_memory()
    host_out = torch.empty(ne * CHUNK + ne * 32 + 64, dtype=torch.uint8, pin_memory=True)
    for _ in range(2):
        enc.encode_packed(host_in, dst=host_out)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(e2e_steps):
        _, e_total, _, _ = enc.encode_packed(host_in, dst=host_out)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    sampler.stop = True
    sampler.join(timeout=2)

    # max over ranks
    from compress_b200 import shard
    total_ms_max, e2e_s_max = shard.max_over_ranks([total_ms, e2e_s], device=dev)
    in_bytes = n * CHUNK
    value = world * in_bytes * args.steps / (total_ms_max / 1e3) / 1e9
    e2e_val = world * ne * CHUNK / e2e_s_max / 1e9
    peak, peak_kind = hbm_peak()
    # dominant kernel of the pipeline, from the CUDA events recorded around every kernel of the timed steps
    dom = max(kms, key=kms.get)
    dom_s = kms[dom] / max(kcalls, 1) / 1e3
    achieved = (in_bytes + out_bytes) / dom_s / 1e9
    step_s = (sum(step_ms) / len(step_ms)) / 1e3
    traffic = None
    tj = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tj):
        try:
            traffic = json.load(open(tj)).get(dom)
        except Exception:
            traffic = None
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
                "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8", "data": "synthetic", "config": config, "ratio": out_bytes / in_bytes,
                "hbm_input_fraction": (value / world) / peak,
                "gpu_launches": int(kernel_launches),
                "clocks": sampler.summary(),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "peak_kind": peak_kind, "kernel": dom,
                             "algorithmic_bytes_per_launch": in_bytes + out_bytes,
                             "kernel_ms_per_step": {k: v / max(kcalls, 1) for k, v in kms.items()},
                             "pipeline_frac": (in_bytes + out_bytes) / step_s / 1e9 / peak},
                "e2e": {"value": e2e_val, "unit": "GB/s", "h2d_bytes_per_step": ne * CHUNK, "d2h_bytes_per_step": int(e_total),
                        "api": "b2c_zstd_encode_packed (pinned host in/out; H2D, kernels and D2H on three streams, two slots)", "chunks_per_step": ne}}
        line["decode"] = {"value": in_bytes / (dec_ms / 1e3) / 1e9, "unit": "GB/s (output bytes, this rank)", "ms": dec_ms,
                          "roofline_frac": (in_bytes + out_bytes) / (dec_ms / 1e3) / 1e9 / peak,
                          "note": "b2c_zstd_decode_kernel on the frames produced above; verified equal to the input"}
        line["s2"] = s2res
        line["huff0"] = huf_res
        if not args.no_cpu_baseline and world == 1:
            sample = H.synth_chunks("text", 2048, seed=77)
            gbs, dt, ratio = cpu_reference_rate(sample, nthreads, 10.0)
            g1, d1, _ = cpu_reference_rate(sample[:64], 1, 3.0)
            line["cpu_baseline"] = {"value": gbs, "unit": "GB/s", "cores": nthreads, "kind": "port",
                                    "sample": "2048 distinct 64 KiB synthetic-text chunks re-encoded for %.1f s on %d threads "
                                              "(oracle C restatement of the Go encoder, one EncodeAll per chunk, pooled "
                                              "encoders); 1 thread: %.3f GB/s" % (dt, nthreads, g1),
                                    "ratio": ratio}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
