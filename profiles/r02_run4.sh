#!/bin/bash
# round-2 four-GPU check: bench --gpus 4 (parity preamble on the 2 x 2 / 1 x 4 / 4 x 1 grids + timings), parity worker at P = 4
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err; echo "bench4_rc=$?" >> gpurun_out/r02_bench_n4.err
B2_PARITY_FULL=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29631 tests/multi_worker.py > gpurun_out/r02_multi4.log 2>&1; echo "multi4_rc=$?" >> gpurun_out/r02_multi4.log
tail -n 4 gpurun_out/r02_bench_n4.err; cut -c1-600 gpurun_out/r02_bench_n4.json; grep "MULTI_WORKER\|multi4_rc\|Error" gpurun_out/r02_multi4.log | head -12
