#!/bin/bash
# round-2 single-GPU follow-up: single-pass pack kernel of the Fredholm path (old vs new), Fredholm tests,
# host enqueue cost of one MPIFirstDerivative.matvec, bench N=1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "fredholm or Fredholm or mdc or MDC or cgls or cg_" > gpurun_out/r02b_pytest_fredholm.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/r02b_pytest_fredholm.log
for ps in 1 0; do
  B2_FREDHOLM_PACK_SMALL=$ps timeout 200 python profiles/fredholm_tc_check.py --time > gpurun_out/fr_pack${ps}.log 2>&1; echo "rc=$?" >> gpurun_out/fr_pack${ps}.log
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/fr_launches_pack1.csv python profiles/fredholm_tc_check.py --time --time-only > gpurun_out/fr_ncu_pack1.log 2>&1
timeout 300 python profiles/host_enqueue.py > gpurun_out/r02b_host_enqueue.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err; echo "bench1_rc=$?" >> gpurun_out/r02b_bench_n1.err
tail -n 6 gpurun_out/r02b_pytest_fredholm.log
for f in gpurun_out/fr_pack1.log gpurun_out/fr_pack0.log; do echo == $f; grep -c OK $f; grep "FAIL\|us \|rc=\|Error\|error" $f | head -12; done
head -c 3000 gpurun_out/r02b_host_enqueue.log
tail -n 3 gpurun_out/r02b_bench_n1.err; cut -c1-300 gpurun_out/r02b_bench_n1.json
