"""N>1 host logic on CPU (gloo, world_size 2): track sharding, global-index IR/gain generation,
sample-major column tiles and the one collective of the path (sum of the stereo mix bus).  The
per-rank "engine output" here is the oracle (allowed in tests): what is under test is the sharding
and reduction plumbing that bench.py runs over NCCL on the GPU box."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gpuaudiobench_b200.distributed import default_mix_gains, reduce_mix_bus, shard_tracks, stitch_sample_major


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, Tg, B, L, outdir):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    sys.path.insert(0, os.path.dirname(here))
    from oracle_lib import Oracle
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    o = Oracle()
    t0, t1 = shard_tracks(Tg, world, rank)
    x_all = o.generate_input(Tg * B).reshape(Tg, B)          # every rank can derive the global job
    h = o.generate_ir(Tg, L, "accel", t0, t1)                # ... but only builds its own tracks' IRs
    y_local = o.r2(x_all[t0:t1], h, L, B, t1 - t0)           # [B][T_local], the engine's block 0
    tile = torch.zeros(B, Tg)
    tile[:, t0:t1] = torch.from_numpy(y_local)               # column tile of the global [B][Tg] matrix
    gains = default_mix_gains(Tg, t0, t1).double()           # [T_local][2]
    bus = (gains.T @ torch.from_numpy(y_local).double().T).float()  # per-GPU partial [2][B]
    bus = reduce_mix_bus(bus)                                # the only collective
    tiles = [torch.zeros(B, Tg) for _ in range(world)]
    dist.all_gather(tiles, tile)
    if rank == 0:
        np.save(os.path.join(outdir, "bus.npy"), bus.numpy())
        np.save(os.path.join(outdir, "stitched.npy"), stitch_sample_major(tiles, Tg).numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_tile_exactly():
    for Tg in (1, 7, 128, 4096):
        for world in (1, 2, 3, 8):
            ranges = [shard_tracks(Tg, world, r) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == Tg
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    assert [shard_tracks(4096, 8, r) for r in (0, 7)] == [(0, 512), (3584, 4096)]


def test_default_gains_are_constant_power_and_match_engine_formula():
    g = default_mix_gains(4096, 0, 4096).double()
    assert torch.allclose((g ** 2).sum(dim=1), torch.full((4096,), 1.0 / 4096, dtype=torch.float64), atol=1e-9)
    part = default_mix_gains(4096, 512, 1024)
    assert torch.equal(part, default_mix_gains(4096, 0, 4096)[512:1024])


@pytest.mark.timeout(300)
def test_two_rank_sharded_job_matches_single_rank(tmp_path, oracle):
    Tg, B, L, world = 10, 64, 100, 2
    mp.spawn(_worker, args=(world, _free_port(), Tg, B, L, str(tmp_path)), nprocs=world, join=True)
    x = oracle.generate_input(Tg * B)
    full = oracle.r2(x, oracle.generate_ir(Tg, L, "accel"), L, B, Tg)       # [B][Tg], one process
    stitched = np.load(tmp_path / "stitched.npy")
    assert np.array_equal(stitched, full), "sharded column tiles must reproduce the global matrix bit-for-bit"
    gains = default_mix_gains(Tg, 0, Tg).double().numpy()
    want_bus = gains.T @ full.astype(np.float64).T
    bus = np.load(tmp_path / "bus.npy")
    assert np.abs(bus - want_bus).max() <= 1e-6 * np.abs(want_bus).max()
