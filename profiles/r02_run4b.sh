#!/bin/bash
# round-2 four-GPU re-check after the graph-pool / lean-enqueue changes: bench --gpus 4, parity worker at P = 4,
# host enqueue cost of the peer-halo stencil path
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err; echo "bench4_rc=$?" >> gpurun_out/r02_bench_n4.err
B2_PARITY_FULL=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29631 tests/multi_worker.py > gpurun_out/r02_multi4.log 2>&1; echo "multi4_rc=$?" >> gpurun_out/r02_multi4.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29633 profiles/host_enqueue.py > gpurun_out/r02_host_enqueue_n4.log 2>&1
tail -n 4 gpurun_out/r02_bench_n4.err; cut -c1-600 gpurun_out/r02_bench_n4.json; grep "MULTI_WORKER\|multi4_rc\|Error" gpurun_out/r02_multi4.log | head -12; grep enqueue_us gpurun_out/r02_host_enqueue_n4.log
