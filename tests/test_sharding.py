"""N>1 host logic (chunk ranges, stream offsets, max-over-ranks timing) with the gloo backend, world_size 2."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nchunks, q):
    sys.path.insert(0, ROOT)
    # shard.py has no dependency on the CUDA library; import it without the package __init__ (which loads the .so)
    import importlib.util
    spec = importlib.util.spec_from_file_location("b2c_shard", os.path.join(ROOT, "compress_b200", "shard.py"))
    shard = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(shard)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard.chunk_range(rank, world, nchunks)
    # stand-in for the encoded frames of my chunks: chunk i encodes to 100 + i bytes
    sizes = [100 + i for i in range(lo, hi)]
    off, total, totals = shard.stream_offsets(sum(sizes))
    tmax = shard.max_over_ranks([1.0 + rank, 5.0 - rank])
    q.put((rank, lo, hi, off, total, totals, tmax))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_sharding():
    world, nchunks = 2, 11
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nchunks, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, off0, tot0, totals0, tm0), (r1, lo1, hi1, off1, tot1, totals1, tm1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 6, 6, 11)
    want0 = sum(100 + i for i in range(0, 6))
    want1 = sum(100 + i for i in range(6, 11))
    assert off0 == 0 and off1 == want0 and tot0 == tot1 == want0 + want1 and totals0 == [want0, want1]
    assert tm0 == tm1 == [2.0, 5.0]


def test_chunk_range_covers_everything():
    import importlib.util
    spec = importlib.util.spec_from_file_location("b2c_shard", os.path.join(ROOT, "compress_b200", "shard.py"))
    shard = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(shard)
    for world in (1, 2, 3, 4, 8):
        for n in (0, 1, 7, 8, 16384, 16385):
            spans = [shard.chunk_range(r, world, n) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_numa_binding_helpers(tmp_path):
    """bind_to_gpu_numa reads the GPU's NUMA node and that node's CPU list from sysfs (a fake tree here)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("b2c_shard", os.path.join(ROOT, "compress_b200", "shard.py"))
    shard = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(shard)
    assert shard._parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    dev = tmp_path / "bus/pci/devices/0000:1b:00.0"
    dev.mkdir(parents=True)
    (dev / "numa_node").write_text("1\n")
    node = tmp_path / "devices/system/node/node1"
    node.mkdir(parents=True)
    mine = sorted(os.sched_getaffinity(0))
    (node / "cpulist").write_text(",".join(str(c) for c in mine[:1]) + "\n")
    assert shard.gpu_numa_node("00000000:1B:00.0", str(tmp_path)) == 1
    before = os.sched_getaffinity(0)
    try:
        assert shard.bind_to_gpu_numa("0000:1b:00.0", str(tmp_path)) == 1
        assert os.sched_getaffinity(0) == {mine[0]}
    finally:
        os.sched_setaffinity(0, before)
    (dev / "numa_node").write_text("-1\n")
    assert shard.bind_to_gpu_numa("0000:1b:00.0", str(tmp_path)) is None
