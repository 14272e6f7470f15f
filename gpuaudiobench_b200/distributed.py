"""Track sharding and the stereo mix-bus reduce — the only multi-GPU logic the path has.

Tracks are independent units (own input, IR, history / delay line), so GPU g of G owns the
contiguous range [g*Tg/G, (g+1)*Tg/G) and no data-path collective is needed except the sum of the
per-GPU stereo mix bus float[2][B].  One process per GPU; torch.distributed
supplies the plumbing (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_tracks(total_tracks, world_size, rank):
    """Contiguous track range [t0, t1) of `rank`; ranges tile [0, total_tracks) exactly."""
    t0 = total_tracks * rank // world_size
    t1 = total_tracks * (rank + 1) // world_size
    return t0, t1


def default_mix_gains(total_tracks, t0, t1):
    """Constant-power pan from the GLOBAL track index, bus scale 1/sqrt(Tg) (same formula as the
    engine's default, csrc/engine.cu set_default_gains)."""
    t = torch.arange(t0, t1, dtype=torch.float64)
    theta = (t + 0.5) / total_tracks * (torch.pi / 2)
    scale = 1.0 / (total_tracks ** 0.5)
    return torch.stack([torch.cos(theta) * scale, torch.sin(theta) * scale], dim=1).to(torch.float32)


def reduce_mix_bus(mix, group=None, all_ranks=True):
    """Sum the per-GPU bus partials [2][B] in place.  all_ranks: all-reduce (every rank gets the
    bus), else reduce to rank 0.  4 KiB at B = 512: latency-bound, enqueued on the current stream."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return mix
    if all_ranks:
        dist.all_reduce(mix, op=dist.ReduceOp.SUM, group=group)
    else:
        dist.reduce(mix, dst=0, op=dist.ReduceOp.SUM, group=group)
    return mix


def stitch_sample_major(column_tiles, total_tracks):
    """Host stitch of per-rank [B][Tg] column tiles (each rank wrote only its own columns)."""
    out = torch.zeros_like(column_tiles[0])
    for tile in column_tiles:
        out += tile
    assert out.shape[1] == total_tracks
    return out
