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
