"""Multi-GPU partitioning of the denoising path.

DualDiff's step shards by scene with NO data-path collective: cross-view attention regroups '(b n)' per scene
(networks/blocks.py:196-197), so a scene never reads another scene's activations.  This mirrors how the reference
runs multi-GPU inference (`accelerator.prepare(val_dataloader)`, perception/data_prepare/val_set_gen.py:121;
manual `shard_volumn`, tools/downstream_v3_batched.py:120).  One process per GPU; the process group is used only
for the barrier and the max-over-ranks timing reduction.
"""
import torch


def shard_scenes(total_scenes: int, rank: int, world_size: int) -> range:
    """contiguous, balanced partition of scene indices [0, total_scenes)"""
    base, rem = divmod(total_scenes, world_size)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def reduce_max_ms(ms: float, device=None) -> float:
    """max over ranks of a device-measured duration (every multi-GPU number is timed as the slowest rank)"""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])
