"""Data-parallel plumbing: one process per GPU, NCCL over NVLink via torch.distributed (gloo on CPU in tests).

Training shards by batch: every rank runs the three steps on its own shard with LOCAL
BatchNorm / latent-standardisation / gradient-norm statistics (what wrapping the reference in DDP would do) and the
parameter gradients are averaged before each optimiser step.  Tile synthesis needs no collective.
"""
import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise the default process group from RANK / WORLD_SIZE / MASTER_* (torchrun); returns (rank, world)."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world


def allreduce_mean_(tensors, bucket_bytes=256 << 20):
    """In-place mean over ranks of a list of same-dtype tensors, bucketed into flat buffers."""
    if not tensors:
        return
    world = dist.get_world_size()
    bucket, size = [], 0
    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        flat = torch.cat([t.reshape(-1) for t in bucket]) if len(bucket) > 1 else bucket[0].reshape(-1)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(world)
        if len(bucket) > 1:
            off = 0
            for t in bucket:
                n = t.numel()
                t.copy_(flat[off:off + n].view_as(t))
                off += n
        bucket, size = [], 0
    for t in tensors:
        nb = t.numel() * t.element_size()
        if size + nb > bucket_bytes and bucket:
            flush()
        bucket.append(t)
        size += nb
    flush()


def shard_range(n, rank, world):
    """Contiguous split of n units over `world` ranks (tile synthesis: independent units, no collective)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
