# Device time of the tensor-core direct engine at C2 (128 x 512 x 16384) with phases switched off
# (B200CONV_TC_DEBUG: 1 no MMAs, 2 no pending-ring traffic, 4 no tap-image load), beside the FFMA engine.
# The host is kept AHEAD of the GPU (a ~3 ms sleep kernel is queued first) so that the event bracket holds
# device time only: with 10 us kernels, back-to-back event pairs otherwise measure the Python call (~35 us).
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.getcwd())
import gpuaudiobench_b200 as g
from gpuaudiobench_b200 import synth

T, B, L = 128, 512, 16384
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream(dev)
flushbuf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res = {}


def timed(e, x, y, mixp, n=40, cold=False):
    for k in range(10):
        e.process(x[k % 8].data_ptr(), y.data_ptr(), mixp, stream=st.cuda_stream)
    torch.cuda.synchronize()
    if cold:
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        for k in range(n):
            flushbuf.fill_(k & 255)
            ev[k][0].record(st)
            e.process(x[k % 8].data_ptr(), y.data_ptr(), mixp, stream=st.cuda_stream)
            ev[k][1].record(st)
        torch.cuda.synchronize()
        return float(np.median([a.elapsed_time(b) for a, b in ev])) * 1e3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(6_000_000)
    e0.record(st)
    for k in range(n):
        e.process(x[k % 8].data_ptr(), y.data_ptr(), mixp, stream=st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


x = torch.from_numpy(synth.make_input(8 * T * B).reshape(8, T, B)).to(dev)
y = torch.zeros(T, B, device=dev)
mix = torch.zeros(2, B, device=dev)
ir = synth.make_ir(T, L, 0, T)
for dbg in (0, 1, 2, 4, 7):
    os.environ["B200CONV_TC_DEBUG"] = str(dbg)
    e = g.ConvEngine(T, B, L, g.ALGO_DIRECT_TC)
    e.load_ir(ir)
    res[f"tc_debug{dbg}_warm_nobus_us"] = timed(e, x, y, 0)
    res[f"tc_debug{dbg}_warm_bus_us"] = timed(e, x, y, mix.data_ptr())
    if dbg == 0:
        res["tc_cold_nobus_us"] = timed(e, x, y, 0, cold=True)
        res["tc_cold_bus_us"] = timed(e, x, y, mix.data_ptr(), cold=True)
    e.close()
os.environ["B200CONV_TC_DEBUG"] = "0"
e = g.ConvEngine(T, B, L, g.ALGO_DIRECT, flags=g.engine.FLAG_FFMA_ONLY)
e.load_ir(ir)
res["ffma_warm_nobus_us"] = timed(e, x, y, 0)
res["ffma_warm_bus_us"] = timed(e, x, y, mix.data_ptr())
res["ffma_cold_bus_us"] = timed(e, x, y, mix.data_ptr(), cold=True)
e.close()
print(json.dumps(res))
