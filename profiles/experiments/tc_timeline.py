# Where the time of one tensor-core FIR launch goes (C2 shape): %globaltimer stamps of every CTA's first item
# (B200CONV_TC_TRACE=1, b200conv_tc_trace), relative to the earliest CTA start; L2 flushed before the launch.
import json
import os
import sys

import numpy as np
import torch

os.environ["B200CONV_TC_TRACE"] = "1"
sys.path.insert(0, os.getcwd())
import gpuaudiobench_b200 as g
from gpuaudiobench_b200 import synth

T, B, L = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (128, 512, 16384)))
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
x = torch.from_numpy(synth.make_input(8 * T * B).reshape(8, T, B)).to(dev)
y = torch.zeros(T, B, device=dev)
mix = torch.zeros(2, B, device=dev)
os.environ["B200CONV_TC_DEBUG"] = os.environ.get("TC_DEBUG", "0")
e = g.ConvEngine(T, B, L, g.ALGO_DIRECT_TC)
e.load_ir(synth.make_ir(T, L, 0, T))
names = ["start", "band", "own", "tree", "mma_done", "epi_done", "mma_first", "mma_last", "staged", "epi_b0", "epi_b4"]
out = {}
cases = (("warm_nobus", 0, False), ("warm_bus", mix.data_ptr(), False), ("cold_bus", mix.data_ptr(), True))
if os.environ.get("TC_DEBUG", "0") != "0":
    cases = cases[:1]
for label, mixp, cold in cases:
    rows = []
    for k in range(12):
        if cold:
            flush.fill_(k)
        e.process(x[k % 8].data_ptr(), y.data_ptr(), mixp, stream=st.cuda_stream)
        torch.cuda.synchronize()
        if k >= 4:
            tr = e.tc_trace().astype(np.int64)
            t0 = tr[:, 0].min()
            rel = (tr - t0) / 1e3
            grp0 = np.arange(len(tr)) < T   # item order is (column group, track): the first T CTAs carry group 0
            row = {"launch_spread_us": float(rel[:, 0].max())}
            for gname, sel in (("g0", grp0), ("g1", ~grp0)):
                for s, nm in enumerate(names):
                    if nm in ("own", "tree") and (gname == "g1" or (nm == "tree" and not mixp)):
                        continue
                    v = rel[sel, s]
                    if v.size == 0:
                        continue
                    row[f"{gname}_{nm}_med"] = float(np.median(v))
                    row[f"{gname}_{nm}_max"] = float(v.max())
            rows.append(row)
    out[label] = {k: round(float(np.median([r[k] for r in rows])), 2) for k in rows[0]}
print(json.dumps(out, indent=1))
