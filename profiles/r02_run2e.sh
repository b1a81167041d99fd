#!/bin/bash
# round-2 two-GPU follow-up: M = 1 MatrixMult diagnosis, parity worker, bench N = 2 (refresh after the graph-pool fix)
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 profiles/m1_diag.py > gpurun_out/r02_m1_diag_n2.log 2>&1
B2_PARITY_FULL=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/multi_worker.py > gpurun_out/r02_multi2.log 2>&1; echo "multi2_rc=$?" >> gpurun_out/r02_multi2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench2_rc=$?" >> gpurun_out/r02_bench_n2.err
grep -v "^\*\|OMP" gpurun_out/r02_m1_diag_n2.log | head -c 3000; tail -n 4 gpurun_out/r02_multi2.log; tail -n 3 gpurun_out/r02_bench_n2.err; cut -c1-400 gpurun_out/r02_bench_n2.json
