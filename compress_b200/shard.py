"""Multi-GPU sharding of independent chunks: contiguous chunk ranges per rank, no data-path
collective.  The only exchange is the per-rank byte total needed to place each rank's frames in one output stream
(the analogue of the reference's in-order jobFlusher, zstd/enc_jobs.go:180-224)."""
import torch
import torch.distributed as dist


def chunk_range(rank, world, nchunks):
    """Rank r gets chunks [lo, hi): contiguous, sizes differ by at most one, order preserved."""
    base, extra = divmod(nchunks, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def stream_offsets(local_total_bytes, device="cpu"):
    """All ranks learn where their frames start in the concatenated stream.
    Returns (my_offset, total_bytes, per_rank_totals list)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0, int(local_total_bytes), [int(local_total_bytes)]
    world, rank = dist.get_world_size(), dist.get_rank()
    mine = torch.tensor([int(local_total_bytes)], dtype=torch.int64, device=device)
    allv = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine)
    totals = [int(t.item()) for t in allv]
    return sum(totals[:rank]), sum(totals), totals


def max_over_ranks(values, device="cpu"):
    """Timing rule of the bench contract: the job's time is the slowest rank's."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


# ---- host placement: one process per GPU, bound to the GPU's NUMA node -----------------------------------------
def _parse_cpulist(text):
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus += list(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def gpu_numa_node(pci_bus_id, sysfs="/sys"):
    """NUMA node of a GPU from its PCI address ("0000:1b:00.0"); None when the platform does not say (-1 / missing)."""
    import os
    bdf = pci_bus_id.lower()
    if len(bdf.split(":")[0]) == 8:          # nvidia-smi style 00000000:1B:00.0
        bdf = bdf[4:]
    try:
        node = int(open(os.path.join(sysfs, "bus/pci/devices", bdf, "numa_node")).read())
    except (OSError, ValueError):
        return None
    return node if node >= 0 else None


def bind_to_gpu_numa(pci_bus_id, sysfs="/sys"):
    """Restrict this process (and the pinned buffers it allocates from now on: first touch) to the CPUs of the
    GPU's NUMA node, intersected with the CPUs it is allowed to use.  The end-to-end path is bound by host-memory
    traffic of the H2D/D2H copies; with 8 ranks on a two-socket box an unbound rank copies across the socket link.
    Returns the node, or None when nothing was changed."""
    import os
    node = gpu_numa_node(pci_bus_id, sysfs)
    if node is None or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        cpus = set(_parse_cpulist(open(os.path.join(sysfs, "devices/system/node/node%d/cpulist" % node)).read()))
    except OSError:
        return None
    allowed = cpus & set(os.sched_getaffinity(0))
    if not allowed:
        return None
    os.sched_setaffinity(0, allowed)
    return node
