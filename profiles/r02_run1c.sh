#!/bin/bash
# round-2 single-GPU follow-up 2: full GPU suite after the lean enqueue path, CGLS capture cost per capture mode,
# host enqueue cost, bench N=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_pytest_gpu.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/r02c_pytest_gpu.log
for mode in thread_local global relaxed; do
  B2_CGLS_CAPTURE_MODE=$mode timeout 200 python profiles/cgls_capture.py > gpurun_out/r02c_cgls_capture_$mode.log 2>&1
done
timeout 300 python profiles/host_enqueue.py > gpurun_out/r02c_host_enqueue.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err; echo "bench1_rc=$?" >> gpurun_out/r02c_bench_n1.err
tail -n 6 gpurun_out/r02c_pytest_gpu.log
for mode in thread_local global relaxed; do tail -n 3 gpurun_out/r02c_cgls_capture_$mode.log | cut -c1-1500; done
head -c 2500 gpurun_out/r02c_host_enqueue.log
tail -n 3 gpurun_out/r02c_bench_n1.err; cut -c1-300 gpurun_out/r02c_bench_n1.json
