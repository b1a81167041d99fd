#!/bin/bash
# round-2 single-GPU follow-up 3: row-splitting GEMV kernel (tests, old vs new on the per-GPU panels of the
# "32768-vec" case), Fredholm operator enqueue, bench N=1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "gemv or matrixmult or MatrixMult or cgls or fredholm or Fredholm" > gpurun_out/r02d_pytest.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/r02d_pytest.log
for sp in 1 0; do B2_GEMV_SPLIT=$sp timeout 200 python profiles/gemv_rows.py > gpurun_out/r02d_gemv_rows_$sp.log 2>&1; done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r02d_bench_n1.json 2> gpurun_out/r02d_bench_n1.err; echo "bench1_rc=$?" >> gpurun_out/r02d_bench_n1.err
tail -n 5 gpurun_out/r02d_pytest.log; tail -n 1 gpurun_out/r02d_gemv_rows_1.log; tail -n 1 gpurun_out/r02d_gemv_rows_0.log
tail -n 3 gpurun_out/r02d_bench_n1.err; cut -c1-300 gpurun_out/r02d_bench_n1.json
