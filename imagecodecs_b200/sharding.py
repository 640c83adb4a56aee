"""Image-index sharding used by bench.py and mirrored by jpeg_gpu_encode_batch.

Every image is an independent JPEG stream (own DC predictors, own bit cursor, own headers:
jpeg_enc.h:1085-1091), so multi-GPU is plain partitioning: rank r of G takes the contiguous
index range [r*n/G, (r+1)*n/G).  There is no data-path collective; the only communication is
the timing reduction (max over ranks) of the benchmark.
"""


def shard_range(n, world, rank):
    """Contiguous [lo, hi) of rank's images; same arithmetic as jpeg_gpu_api.cpp (encode_batch)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    return (rank * n) // world, ((rank + 1) * n) // world


def max_over_ranks(value, device=None):
    """max of a python float over the default process group (identity when not initialised)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
