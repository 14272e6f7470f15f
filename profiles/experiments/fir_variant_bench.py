"""Time the direct FIR kernel (C2) for a given libb200conv build / env knobs: prints per-stage ms."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import gpuaudiobench_b200 as g
from gpuaudiobench_b200 import engine, synth
if len(sys.argv) > 1:
    engine.LIB_PATH = os.path.abspath(sys.argv[1])
T, B, L = (int(v) for v in os.environ.get("SHAPE", "128,512,16384").split(","))
e = g.ConvEngine(T, B, L, g.ALGO_DIRECT)
e.load_ir(synth.make_ir(T, L))
x = torch.from_numpy(synth.make_input(8 * T * B).reshape(8, T, B)).cuda()
y = torch.zeros(T, B, device="cuda"); mix = torch.zeros(2, B, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for k in range(40): e.process(x[k % 8].data_ptr(), y.data_ptr(), mix.data_ptr())
torch.cuda.synchronize(); e.set_profiling(True)
for k in range(100):
    flush.fill_(k & 255); e.process(x[k % 8].data_ptr(), y.data_ptr(), mix.data_ptr())
torch.cuda.synchronize(); q = e.query()
n = q["stage_count"]; d = q["dominant_stage"]
ms = [m / q["stage_calls"] for m in q["stage_ms"][:n]]
print(os.path.basename(engine.LIB_PATH), (T, B, L), {k: os.environ[k] for k in os.environ if k.startswith("B200CONV")},
      dict(zip(q["stage_name"], [round(m * 1e3, 2) for m in ms])), "us; FIR TFLOP/s", round(q["flops_per_block"] / ms[d] / 1e9, 2))
