# Where do the +4 us of a multi-GPU C2 step come from?  Per rank, L2 flushed between steps, ranks released together:
#   a) engine alone (no peer mapping attached), with bus        b) bus group attached, launch WITHOUT a bus (no remote traffic)
#   c) bus group attached, with bus (the bench's step)
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

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
gate = torch.zeros(1, device=dev)


def timed(e, mixp, n=100, gated=True):
    for k in range(10):
        e.process(x[k % 8].data_ptr(), y.data_ptr(), mixp, stream=st.cuda_stream)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    torch.cuda.synchronize()
    dist.barrier()
    for k, (a, b) in enumerate(ev):
        flush.fill_(k & 255)
        if gated:
            dist.all_reduce(gate)
        a.record(st)
        e.process(x[k % 8].data_ptr(), y.data_ptr(), mixp, stream=st.cuda_stream)
        b.record(st)
    torch.cuda.synchronize()
    t = torch.tensor([float(np.median([a.elapsed_time(b) for a, b in ev])) * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return round(float(t.item()), 2)


res = {}
e = g.ConvEngine(T, B, L, g.ALGO_DIRECT_TC, device=lr, track_offset=rank * T, total_tracks=world * T)
e.load_ir(synth.make_ir(world * T, L, rank * T, rank * T + T))
res["a_alone_bus_us"] = timed(e, mix.data_ptr())
res["a_alone_nobus_us"] = timed(e, 0)
bus = EngineBusGroup(e, mix)
res["b_attached_nobus_us"] = timed(e, 0)
res["c_attached_bus_us"] = timed(e, mix.data_ptr())
res["c_attached_bus_ungated_us"] = timed(e, mix.data_ptr(), gated=False)
for dbg, name in ((3, "d_no_remote_push_no_remote_poll_us"), (7, "f_local_only_collected_at_once_us"), (4, "g_collected_at_once_us")):
    os.environ["B200CONV_BUS_DEBUG"] = str(dbg)   # (read at every launch; results are wrong on purpose)
    res[name] = timed(e, mix.data_ptr())
# h) as f, but this rank's OWN buffer is ordinary device memory instead of the symmetric-memory allocation
plain = torch.zeros(bus.buf.numel(), dtype=torch.float32, device=dev)
ptrs = [int(p) for p in bus.hdl.buffer_ptrs]
real = list(ptrs)
ptrs[rank] = plain.data_ptr()
os.environ["B200CONV_BUS_DEBUG"] = "7"
e.attach_bus(ptrs, rank, world)
res["h_local_only_plain_own_buffer_us"] = timed(e, mix.data_ptr())
torch.cuda.synchronize()
dist.barrier()
e.attach_bus(real, rank, world)
os.environ["B200CONV_BUS_DEBUG"] = "0"
torch.cuda.synchronize()
dist.barrier()
bus.buf.zero_()
torch.cuda.synchronize()
dist.barrier()
bus.check()
if rank == 0:
    print(json.dumps(res))
torch.cuda.synchronize()
dist.barrier()
bus.close()
e.close()
dist.destroy_process_group()
