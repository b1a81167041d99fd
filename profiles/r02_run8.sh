#!/bin/bash
# round-2 eight-GPU confirmation: bench --gpus 8 (parity preamble + timings), then the parity worker at P = 8
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo "bench8_rc=$?" >> gpurun_out/r02_bench_n8.err
B2_PARITY_FULL=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 tests/multi_worker.py > gpurun_out/r02_multi8.log 2>&1; echo "multi8_rc=$?" >> gpurun_out/r02_multi8.log
tail -n 4 gpurun_out/r02_bench_n8.err; cut -c1-600 gpurun_out/r02_bench_n8.json; grep "MULTI_WORKER\|multi8_rc\|Error" gpurun_out/r02_multi8.log | head -12
