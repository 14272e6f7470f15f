# Tensor-core FIR against the FFMA kernel at small blocks (128 tracks x 16384 taps), L2 flushed, with the bus.
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.getcwd())
import gpuaudiobench_b200 as g
from gpuaudiobench_b200 import synth

T, L = 128, 16384
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res = {}
for B in (128, 256):
    x = torch.from_numpy(synth.make_input(8 * T * B).reshape(8, T, B)).to(dev)
    y = torch.zeros(T, B, device=dev)
    mix = torch.zeros(2, B, device=dev)
    ir = synth.make_ir(T, L, 0, T)
    for name, algo, flags in (("tc", g.ALGO_DIRECT_TC, 0), ("ffma", g.ALGO_DIRECT, g.engine.FLAG_FFMA_ONLY)):
        e = g.ConvEngine(T, B, L, algo, flags=flags)
        e.load_ir(ir)
        for k in range(10):
            e.process(x[k % 8].data_ptr(), y.data_ptr(), mix.data_ptr(), stream=st.cuda_stream)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(60)]
        for k, (a, b) in enumerate(ev):
            flush.fill_(k & 255)
            a.record(st)
            e.process(x[k % 8].data_ptr(), y.data_ptr(), mix.data_ptr(), stream=st.cuda_stream)
            b.record(st)
        torch.cuda.synchronize()
        res[f"B{B}_{name}_us"] = round(float(np.median([a.elapsed_time(b) for a, b in ev])) * 1e3, 2)
        e.close()
        print(json.dumps(res), flush=True)
