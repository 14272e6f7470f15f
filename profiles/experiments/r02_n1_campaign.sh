python -m pytest tests -m gpu -q > gpurun_out/r2C_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2C_pytest.log; tail -3 gpurun_out/r2C_pytest.log
python bench.py --steps 200 --warmup 5 2> gpurun_out/r2C_bench_n1.err | grep "^{" > gpurun_out/r2C_bench_n1.json; echo "bench rc=$?"
python bench.py --workload c2ffma --no-also --steps 200 --warmup 5 2>/dev/null | grep "^{" > gpurun_out/r2C_bench_c2ffma_n1.json; echo "ffma rc=$?"
python bench.py --sweep --steps 1000 --warmup 5 2> gpurun_out/r2C_sweep_n1.err | grep "^{" > gpurun_out/r2C_sweep_n1.json; echo "sweep rc=$?"
python bench.py --workload c4 --latency 1000 2>/dev/null | grep "^{" > gpurun_out/r2C_latency_c4_n1.json
python bench.py --workload c2 --latency 1000 2>/dev/null | grep "^{" > gpurun_out/r2C_latency_c2_n1.json
python bench.py --strip --steps 100 2>/dev/null | grep "^{" > gpurun_out/r2C_strip.json; echo "strip rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | grep "^{" > gpurun_out/r2C_reference_arm.json; echo "ref rc=$?"
for w in c2 c3 c4; do ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2C_${w}_launches.csv python bench.py --workload $w --no-also --steps 6 --warmup 3 > /dev/null 2>&1; echo "ncu list $w rc=$?"; done
ncu --set full --import-source on --clock-control none -k regex:tc_toeplitz -c 2 -o gpurun_out/r2C_tc python bench.py --workload c2 --no-also --steps 3 --warmup 3 > /dev/null 2>&1; echo "ncu tc rc=$?"
ncu --set full --import-source on --clock-control none -k regex:upols_fused -s 200 -c 2 -o gpurun_out/r2C_upols_c4 python bench.py --workload c4 --no-also --steps 3 --warmup 3 > /dev/null 2>&1; echo "ncu c4 rc=$?"
ncu --set full --import-source on --clock-control none -k regex:upols_fused -s 270 -c 2 -o gpurun_out/r2C_upols_c3 python bench.py --workload c3 --no-also --steps 3 --warmup 3 > /dev/null 2>&1; echo "ncu c3 rc=$?"
ncu --set full --import-source on --clock-control none -k regex:fir_direct -s 40 -c 2 -o gpurun_out/r2C_fir python bench.py --workload c2ffma --no-also --steps 3 --warmup 3 > /dev/null 2>&1; echo "ncu fir rc=$?"
