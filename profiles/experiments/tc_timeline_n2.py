# Two (or more) ranks, C2 shape per rank: where the multi-GPU step's extra microseconds are.  Per block: the phase
# stamps of the tensor-core kernel (b200conv_tc_trace), the bus exchange stamps (b200conv_bus_trace: push, summed bus)
# and the step's event time, all relative to the kernel's first CTA start.  Run under torchrun.
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

os.environ["B200CONV_TC_TRACE"] = "1"
os.environ["B200CONV_BUS_TRACE"] = "1"
sys.path.insert(0, os.getcwd())
import gpuaudiobench_b200 as g
from gpuaudiobench_b200 import synth
from gpuaudiobench_b200.distributed import EngineBusGroup

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
T, B, L = 128, 512, 16384
st = torch.cuda.current_stream(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
x = torch.from_numpy(synth.make_input(8 * T * B, seed=rank).reshape(8, T, B)).to(dev)
y = torch.zeros(T, B, device=dev)
mix = torch.zeros(2, B, device=dev)
e = g.ConvEngine(T, B, L, g.ALGO_DIRECT_TC, device=lr, track_offset=rank * T, total_tracks=world * T)
e.load_ir(synth.make_ir(world * T, L, rank * T, rank * T + T))
bus = EngineBusGroup(e, mix)
gate = torch.zeros(1, device=dev)
rows = []
for cold in (False, True):
    for k in range(14):
        if cold:
            flush.fill_(k)
        dist.all_reduce(gate)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        e.process(x[k % 8].data_ptr(), y.data_ptr(), mix.data_ptr(), stream=st.cuda_stream)
        b.record(st)
        torch.cuda.synchronize()
        if k < 4:
            continue
        tr = e.tc_trace().astype(np.int64)
        t0 = tr[:, 0].min()
        bt = e.bus_trace(1).astype(np.int64)[0]
        rows.append({"cold": cold, "event_us": a.elapsed_time(b) * 1e3, "own_med": float(np.median(tr[:, 2] - t0)) / 1e3,
                     "mma_done_med": float(np.median(tr[:, 4] - t0)) / 1e3, "epi_done_med": float(np.median(tr[:, 5] - t0)) / 1e3,
                     "epi_done_max": float((tr[:, 5] - t0).max()) / 1e3, "push_cta0": float(bt[0] - t0) / 1e3,
                     "summed_cta0": float(bt[1] - t0) / 1e3, "t0_abs_ns": int(t0)})
out = {}
for cold in (False, True):
    sel = [r for r in rows if r["cold"] == cold]
    out["cold" if cold else "warm"] = {k: round(float(np.median([r[k] for r in sel])), 2) for k in sel[0] if k not in ("cold", "t0_abs_ns")}
# start skew between ranks: first-CTA start times (globaltimer is node-wide only approximately; reported as is)
t0s = torch.tensor([r["t0_abs_ns"] for r in rows], dtype=torch.int64, device=dev)
allt = [torch.empty_like(t0s) for _ in range(world)]
dist.all_gather(allt, t0s)
sk = (torch.stack(allt).max(0).values - torch.stack(allt).min(0).values).double().cpu().numpy() / 1e3
out["start_skew_us_median_max"] = [round(float(np.median(sk)), 2), round(float(sk.max()), 2)]
if rank == 0:
    print(json.dumps(out))
torch.cuda.synchronize()
dist.barrier()
bus.close()
e.close()
dist.destroy_process_group()
