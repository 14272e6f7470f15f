# Where does the N=2 penalty of the short C2 step come from?  Run under torchrun with 2 ranks.
# Variants: (a) flush + step with the in-kernel exchange, (b) flush + step WITHOUT a bus (no exchange at all),
# (c) flush + step with a bus but the engine detached (local bus only), (d) as (a) with a barrier-aligned start
# of every step (host sync + dist.barrier before each step: removes accumulated skew), (e) back-to-back, no flush.
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.getcwd())
import gpuaudiobench_b200 as g
from gpuaudiobench_b200 import synth
from gpuaudiobench_b200.distributed import EngineBusGroup

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.environ["B200CONV_BUS_TRACE"] = "1"
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev = torch.device("cuda", lr)
st = torch.cuda.current_stream(dev)
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
algo, T, B, L = {"c2": (g.ALGO_DIRECT, 128, 512, 16384), "c4": (g.ALGO_UPOLS, 512, 512, 96000)}[wl]
e = g.ConvEngine(T, B, L, algo, device=lr, track_offset=T * rank, total_tracks=T * world)
e.load_ir(synth.make_ir(T * world, L, T * rank, T * rank + T))
x = torch.from_numpy(synth.make_input(8 * T * B, seed=rank).reshape(8, T, B)).to(dev)
y = torch.zeros(T, B, device=dev)
mix = torch.zeros(2, B, device=dev)
w = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
r = torch.zeros(64 << 20, dtype=torch.float32, device=dev)
K = 200


def run(mixp, flush=True, align=False):
    for k in range(40):
        e.process(x[k % 8].data_ptr(), y.data_ptr(), mixp, stream=st.cuda_stream)
    torch.cuda.synchronize()
    dist.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for k in range(K):
        if flush:
            w.fill_(k & 255)
            r.sum()
        if align:
            torch.cuda.synchronize()
            dist.barrier()
        ev[k][0].record(st)
        e.process(x[k % 8].data_ptr(), y.data_ptr(), mixp, stream=st.cuda_stream)
        ev[k][1].record(st)
    torch.cuda.synchronize()
    lat = np.array([a.elapsed_time(b) for a, b in ev]) * 1e3
    t = torch.tensor([lat.mean(), np.median(lat), np.percentile(lat, 99)], device=dev, dtype=torch.float64)
    allt = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    return [[round(float(v), 1) for v in a] for a in allt]


res = {"workload": wl, "columns": "per rank [mean, median, p99] us"}
res["c_detached_local_bus"] = run(mix.data_ptr())
res["b_no_bus"] = run(0)
grp = EngineBusGroup(e, mix)
res["kind"] = grp.kind[:40]
res["a_flush_exchange"] = run(mix.data_ptr())
if wl == "c2":
    tr = e.bus_trace(K).astype(np.int64)  # the K steps just timed
    mine = torch.from_numpy(tr).to(dev)
    both = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(both, mine)
    a, b = both[0].cpu().numpy(), both[1].cpu().numpy()
    res["trace_wait_us_rank0_[mean,median,p90]"] = [round(float(v) / 1e3, 2) for v in ((a[:, 1] - a[:, 0]).mean(), np.median(a[:, 1] - a[:, 0]), np.percentile(a[:, 1] - a[:, 0], 90))]
    res["trace_wait_us_rank1_[mean,median,p90]"] = [round(float(v) / 1e3, 2) for v in ((b[:, 1] - b[:, 0]).mean(), np.median(b[:, 1] - b[:, 0]), np.percentile(b[:, 1] - b[:, 0], 90))]
    d = (a[:, 0] - b[:, 0]) / 1e3  # ready-time difference between the ranks (if the two %globaltimers agree)
    res["ready_time_rank0_minus_rank1_us_[mean,std,mean_abs]"] = [round(float(d.mean()), 2), round(float(d.std()), 2), round(float(np.abs(d - d.mean()).mean()), 2)]
res["e_back_to_back_exchange"] = run(mix.data_ptr(), flush=False)
res["d_aligned_every_step"] = run(mix.data_ptr(), align=True)

grp.check()
grp.close()
if rank == 0:
    print(json.dumps(res))
dist.barrier()
dist.destroy_process_group()
