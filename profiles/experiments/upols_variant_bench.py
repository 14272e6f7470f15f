"""Time the UPOLS block (C3 or C4 shard) for a given libb200conv build / env knobs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import gpuaudiobench_b200 as g
from gpuaudiobench_b200 import engine, synth
if len(sys.argv) > 1:
    engine.LIB_PATH = os.path.abspath(sys.argv[1])
T, B, L = (int(v) for v in os.environ.get("SHAPE", "1024,256,65536").split(","))
e = g.ConvEngine(T, B, L, g.ALGO_UPOLS)
e.load_ir(synth.make_ir(T, L))
x = torch.from_numpy(synth.make_input(8 * T * B).reshape(8, T, B)).cuda()
y = torch.zeros(T, B, device="cuda"); mix = torch.zeros(2, B, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
flush_r = torch.zeros(64 << 20, dtype=torch.float32, device="cuda")  # read pass: no dirty lines left for the step
for k in range(40): e.process(x[k % 8].data_ptr(), y.data_ptr(), mix.data_ptr())
torch.cuda.synchronize(); e.set_profiling(True)
for k in range(100):
    flush.fill_(k & 255); _ = flush_r.sum(); e.process(x[k % 8].data_ptr(), y.data_ptr(), mix.data_ptr())
torch.cuda.synchronize(); q = e.query()
n = q["stage_count"]; ms = [m / q["stage_calls"] for m in q["stage_ms"][:n]]
d = q["dominant_stage"]
print(os.path.basename(engine.LIB_PATH), (T, B, L), {k: os.environ[k] for k in os.environ if k.startswith("B200CONV")},
      dict(zip(q["stage_name"], [round(m * 1e3, 1) for m in ms])), "us; GB/s", round(q["alg_bytes_per_block"] / ms[d] / 1e6, 0), "S", e.query()["partitions"])
