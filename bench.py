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


def cpu_reference_rate(sample_chunks, threads):
    """Oracle (C restatement of zstd.Encoder.EncodeAll, SpeedFastest) on `threads` host threads, one encoder
    state per call like the reference's encoder pool.  Returns (GB/s, seconds)."""
    import helpers as H
    H.build_oracle()
    L = H.oracle()
    from concurrent.futures import ThreadPoolExecutor
    cap = L.orc_zstd_max_encoded_size(CHUNK, 1, 1) + 64
    bufs = [ctypes.create_string_buffer(cap) for _ in range(threads)]
    L.orc_zstd_cctx_new.restype = ctypes.c_void_p
    L.orc_zstd_encode_all_ctx.restype = ctypes.c_int64
    L.orc_zstd_encode_all_ctx.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_void_p, ctypes.c_size_t]
    ctxs = [L.orc_zstd_cctx_new() for _ in range(threads)]  # one pooled encoder per thread (zstd/encoder.go:90-99)

    def work(t):
        tot = 0
        for i in range(t, len(sample_chunks), threads):
            c = sample_chunks[i]
            r = L.orc_zstd_encode_all_ctx(ctxs[t], c, len(c), 1, 1, bufs[t], cap)  # ctypes releases the GIL
            assert r > 0
            tot += r
        return tot
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        outs = list(ex.map(work, range(threads)))
    dt = time.perf_counter() - t0
    nbytes = sum(len(c) for c in sample_chunks)
    return nbytes / dt / 1e9, dt, sum(outs) / nbytes


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nchunks", type=int, default=NCHUNKS)
    ap.add_argument("--e2e-chunks", type=int, default=8192, help="chunks per e2e step (512 MiB default)")
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
    nthreads = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return
        # bounded sample: each step = 2048 chunks (128 MiB) of the same synthetic text, all host threads
        sample = H.synth_chunks("text", 256, seed=77) * 8
        steps = max(1, args.steps)
        for _ in range(min(args.warmup, 1)):
            cpu_reference_rate(sample[:256], nthreads)
        t0 = time.perf_counter()
        rates = [cpu_reference_rate(sample, nthreads) for _ in range(steps)]
        dt = time.perf_counter() - t0
        gbs = sum(len(c) for c in sample) * steps / sum(r[1] for r in rates) / 1e9
        line = {"metric": METRIC, "value": gbs, "unit": "GB/s", "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * sum(r[1] for r in rates) / steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic", "impl": "reference", "config": config,
                "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": nthreads, "kind": "port",
                                 "sample": "2048 x 64 KiB chunks per step; oracle C restatement of the Go encoder "
                                           "(Go toolchain absent), one EncodeAll per chunk, %d threads" % nthreads,
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
    enc = zstd.Encoder(device=local_rank, max_chunks=2048)
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
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for k in range(args.steps):
        enc.encode_device(src, dst=dst, out_sizes=outs)
        ev[k + 1].record()
    barrier()
    step_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    total_ms = ev[0].elapsed_time(ev[args.steps])
    kernel_launches = enc.launches - launches0
    outs_h = outs.cpu().numpy()
    assert (outs_h > 0).all(), "encode error"
    out_bytes = int(outs_h.sum())

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
    t = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max, e2e_s_max = float(t[0]), float(t[1])
    in_bytes = n * CHUNK
    value = world * in_bytes * args.steps / (total_ms_max / 1e3) / 1e9
    e2e_val = world * ne * CHUNK / e2e_s_max / 1e9
    peak, peak_kind = hbm_peak()
    avg_launch_s = (sum(step_ms) / len(step_ms)) / 1e3
    achieved = (in_bytes + out_bytes) / avg_launch_s / 1e9
    traffic = None
    tj = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tj):
        try:
            traffic = json.load(open(tj)).get("b2c_zstd_encode_kernel")
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
                             "traffic": traffic, "peak_kind": peak_kind, "kernel": "b2c_zstd_encode_kernel",
                             "algorithmic_bytes_per_launch": in_bytes + out_bytes},
                "e2e": {"value": e2e_val, "unit": "GB/s", "h2d_bytes_per_step": ne * CHUNK, "d2h_bytes_per_step": int(e_total),
                        "api": "b2c_zstd_encode_packed (pinned host in/out, double-buffered)", "chunks_per_step": ne}}
        if not args.no_cpu_baseline and world == 1:
            sample = H.synth_chunks("text", 128, seed=77) * 8
            gbs, dt, ratio = cpu_reference_rate(sample, nthreads)
            g1, d1, _ = cpu_reference_rate(sample[:128], 1)
            line["cpu_baseline"] = {"value": gbs, "unit": "GB/s", "cores": nthreads, "kind": "port",
                                    "sample": "1024 x 64 KiB synthetic-text chunks, oracle C restatement of the Go encoder, "
                                              "one EncodeAll per chunk on %d threads (%.1f s); 1 thread: %.3f GB/s" % (nthreads, dt, g1),
                                    "ratio": ratio}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
