#!/bin/bash
# One GPU visit: parity tests, bench line, ncu launch list of the same command, one --set full capture of the draw kernel,
# throughput of the other BASELINE configs.
TAG=${1:-r1x}
export PYTHONDONTWRITEBYTECODE=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" ; tail -3 gpurun_out/pytest_$TAG.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_$TAG.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref_$TAG.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_$TAG.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nuts_chain_kernel --launch-skip 3 -c 1 -f -o gpurun_out/prof_$TAG python scratch/prof_run.py > gpurun_out/prof_$TAG.log 2>&1; echo "ncu full rc=$?"
timeout 600 python scratch/configs_run.py > gpurun_out/configs_$TAG.log 2>&1; echo "configs rc=$?"; cat gpurun_out/configs_$TAG.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_leapfrog --csv --log-file gpurun_out/plane_leapfrog_$TAG.csv python scratch/plane_leapfrog_run.py > gpurun_out/plane_leapfrog_$TAG.log 2>&1; echo "plane ncu rc=$?"; tail -1 gpurun_out/plane_leapfrog_$TAG.log
