#!/bin/bash
# r02c round-end evidence (under gpurun): the full GPU parity suite + smoke + default bench line, then the ncu launch list of
# the bench command (after it exited 0 without ncu).
mkdir -p gpurun_out
bash tools/gpu_check.sh --no-bench; echo "gpu_check rc=$?"
timeout 1200 python bench.py > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02c_bench_reference_arm.json 2> gpurun_out/r02c_bench_reference_arm.err; echo "reference arm rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/plain_bench.json 2> gpurun_out/plain_bench.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02c_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
