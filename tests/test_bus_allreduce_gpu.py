"""The engine's own bus all-reduce kernel (csrc/bus_allreduce.cu) on 2 real GPUs (skipped on a
1-GPU box): bit-identical result on every rank, equal to the rank-ordered fp32 sum, over many
epochs (both slot parities), and the NCCL fallback agrees."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, outdir):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.dirname(here))
    import torch.distributed as dist
    from gpuaudiobench_b200.distributed import BusAllReduce
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    B = 512
    bus = torch.zeros(2, B, device="cuda")
    red = BusAllReduce(bus)
    results = []
    for epoch in range(1, 8):
        g = torch.Generator(device="cpu").manual_seed(100 * epoch + rank)
        bus.copy_(torch.rand(2, B, generator=g) - 0.5)
        red()
        torch.cuda.synchronize()
        results.append(bus.cpu().numpy().copy())
    red.check()
    np.save(os.path.join(outdir, f"r{rank}.npy"), np.stack(results))
    with open(os.path.join(outdir, f"kind{rank}.txt"), "w") as f:
        f.write(red.kind)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_own_bus_allreduce_two_gpus(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    assert np.array_equal(r0, r1), "every rank must hold the bit-identical bus"
    for e, epoch in enumerate(range(1, 8)):
        parts = [(torch.rand(2, 512, generator=torch.Generator().manual_seed(100 * epoch + r)) - 0.5).numpy() for r in range(world)]
        want = parts[0].astype(np.float32) + parts[1].astype(np.float32)  # rank order, fp32
        assert np.array_equal(r0[e], want)
    kind = open(tmp_path / "kind0.txt").read()
    print("collective used:", kind)
    assert kind.startswith("own"), kind
