# Multi-GPU legs of the round-2 campaign: N=$1 ranks on one box (gpurun --gpus N).  Default bench line (C2 + C3 + C4
# under "also", parity leg at every N) and the direct-form half of the config-5 sweep.
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
$TR bench.py --gpus $N --steps 200 --warmup 5 2> gpurun_out/r2D_bench_n$N.err | grep "^{" > gpurun_out/r2D_bench_n$N.json; echo "bench rc=$?"
if [ "$2" = "sweep" ]; then
  $TR bench.py --gpus $N --sweep --sweep-filter direct --steps 1000 --warmup 5 2> gpurun_out/r2D_sweep_direct_n$N.err | grep "^{" > gpurun_out/r2D_sweep_direct_n$N.json; echo "sweep rc=$?"
fi
