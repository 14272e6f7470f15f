# Round-2 single-GPU measurement campaign (one B200 via gpurun), final kernels.  Outputs land in gpurun_out/r2D_*;
# the summaries committed under profiles/ are made from them with profiles/summarize_ncu.py.
python -m pytest tests -m gpu -q > gpurun_out/r2D_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2D_pytest.log; tail -3 gpurun_out/r2D_pytest.log
python bench.py --steps 200 --warmup 5 2> gpurun_out/r2D_bench_n1.err | grep "^{" > gpurun_out/r2D_bench_n1.json; echo "bench rc=$?"
python bench.py --workload c2ffma --no-also --steps 200 --warmup 5 2>/dev/null | grep "^{" > gpurun_out/r2D_bench_c2ffma_n1.json; echo "ffma rc=$?"
python bench.py --sweep --steps 1000 --warmup 5 2> gpurun_out/r2D_sweep_n1.err | grep "^{" > gpurun_out/r2D_sweep_n1.json; echo "sweep rc=$?"
python bench.py --workload c4 --latency 1000 2>/dev/null | grep "^{" > gpurun_out/r2D_latency_c4_n1.json
python bench.py --workload c2 --latency 1000 2>/dev/null | grep "^{" > gpurun_out/r2D_latency_c2_n1.json
python bench.py --strip --steps 100 2>/dev/null | grep "^{" > gpurun_out/r2D_strip.json; echo "strip rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | grep "^{" > gpurun_out/r2D_reference_arm.json; echo "ref rc=$?"
python profiles/experiments/tc_phase_timing.py > gpurun_out/r2D_tc_phases.json 2>/dev/null; echo "phases rc=$?"
python profiles/experiments/tc_vs_ffma_small.py 2>/dev/null | tail -1 > gpurun_out/r2D_tc_vs_ffma_small.json
for shape in "128 512 16384" "128 128 16384" "128 1024 16384"; do python profiles/experiments/tc_timeline.py $shape 2>/dev/null > "gpurun_out/r2D_tc_timeline_$(echo $shape | tr ' ' '_').json"; done; echo "timeline rc=$?"
for w in c2 c3 c4; do ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2D_${w}_launches.csv python bench.py --workload $w --no-also --steps 6 --warmup 3 > /dev/null 2>&1; echo "ncu list $w rc=$?"; done
ncu --set full --import-source on --clock-control none -k regex:tc_toeplitz -c 2 -o gpurun_out/r2D_tc python bench.py --workload c2 --no-also --steps 3 --warmup 3 > /dev/null 2>&1; echo "ncu tc rc=$?"
