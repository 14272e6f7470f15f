"""True incremental cost of the stereo-bus kernel in the un-profiled pipeline: step time with and without d_mix."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import gpuaudiobench_b200 as g
from gpuaudiobench_b200 import synth
for name, algo, T, B, L, layout in (("c2", g.ALGO_DIRECT, 128, 512, 16384, g.OUT_TRACK_MAJOR), ("c3", g.ALGO_UPOLS, 1024, 256, 65536, g.OUT_SAMPLE_MAJOR),
                                    ("c4", g.ALGO_UPOLS, 512, 512, 96000, g.OUT_TRACK_MAJOR)):
    e = g.ConvEngine(T, B, L, algo, layout)
    e.load_ir(synth.make_ir(T, L))
    x = torch.from_numpy(synth.make_input(8 * T * B).reshape(8, T, B)).cuda()
    y = torch.zeros(B * T, device="cuda"); mix = torch.zeros(2, B, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    res = {}
    for label, mp in (("with_bus", mix.data_ptr()), ("no_bus", 0)):
        for k in range(30): e.process(x[k % 8].data_ptr(), y.data_ptr(), mp)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(100)]
        torch.cuda.synchronize()
        for k, (a, b) in enumerate(ev):
            flush.fill_(k & 255); a.record(); e.process(x[k % 8].data_ptr(), y.data_ptr(), mp); b.record()
        torch.cuda.synchronize()
        res[label] = float(np.median([a.elapsed_time(b) for a, b in ev])) * 1e3
    print(name, {k: round(v, 1) for k, v in res.items()}, "bus cost us", round(res["with_bus"] - res["no_bus"], 1))
    e.close()
