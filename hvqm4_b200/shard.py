"""Multi-GPU sharding helpers: streams are independent, so ranks share nothing on the data path.

stream s of a job with `per_gpu` streams per rank lives on rank s // per_gpu; its bitstream seed is
base_seed + s.  The only cross-rank operations are the timing barrier and the max-over-ranks of the
timed interval (bench.py); both go through torch.distributed (nccl on GPUs, gloo in CPU tests).
"""
from __future__ import annotations


def rank_streams(rank: int, world: int, per_gpu: int):
    """Global stream ids owned by `rank` (contiguous slab)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank * per_gpu, (rank + 1) * per_gpu))


def stream_seed(base_seed: int, stream_id: int, distinct: int | None = None) -> int:
    """Seed of a stream's synthetic bitstream; with `distinct` set, seeds repeat cyclically inside a rank's slab."""
    return base_seed + (stream_id if distinct is None else stream_id % distinct)


def max_over_ranks(value: float, dist=None, device=None) -> float:
    """Whole-job time = slowest rank."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def job_throughput(units_per_rank: int, world: int, seconds: float) -> float:
    """Aggregate units/s of the whole job (all ranks process units_per_rank in `seconds`)."""
    return units_per_rank * world / seconds
