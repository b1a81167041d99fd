#!/bin/bash
# round-2 single-GPU check: GPU test suite, Fredholm tensor-core modes (+ ncu), bench N=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/r02_pytest_gpu.log
bash profiles/fredholm_tc_run.sh > gpurun_out/r02_fredholm_modes.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench_rc=$?" >> gpurun_out/r02_bench_n1.err
tail -n 15 gpurun_out/r02_pytest_gpu.log; cat gpurun_out/r02_fredholm_modes.log | tail -n 40; tail -n 5 gpurun_out/r02_bench_n1.err; cut -c1-300 gpurun_out/r02_bench_n1.json
