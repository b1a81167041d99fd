#!/bin/bash
# round-2 two-GPU check: GPU test suite (incl. 2-rank worker), Fredholm modes + ncu, multi worker, bench N=2 and N=1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/r02_pytest_gpu.log
bash profiles/fredholm_tc_run.sh > gpurun_out/r02_fredholm_modes.log 2>&1
B2_PARITY_FULL=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/multi_worker.py > gpurun_out/r02_multi2.log 2>&1; echo "multi2_rc=$?" >> gpurun_out/r02_multi2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench2_rc=$?" >> gpurun_out/r02_bench_n2.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench1_rc=$?" >> gpurun_out/r02_bench_n1.err
tail -n 12 gpurun_out/r02_pytest_gpu.log; tail -n 25 gpurun_out/r02_fredholm_modes.log; tail -n 12 gpurun_out/r02_multi2.log; tail -n 4 gpurun_out/r02_bench_n2.err; cut -c1-600 gpurun_out/r02_bench_n2.json
