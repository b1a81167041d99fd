#!/bin/bash
# Run ON THE GPU BOX (via gpurun): ncu launch list of the bench command + full captures of
# the dominant kernels.  Outputs go to gpurun_out/ (merged back), summaries are then
# extracted here into profiles/ by profiles/summarise.py.
set -x
mkdir -p gpurun_out
NCU=$(command -v ncu || echo /usr/local/cuda/bin/ncu)
# 1. every launch of one short bench run with its device time (shares, not absolutes)
$NCU --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
# 2. full capture of the stencil kernel (headline) -- 3 launches after warm-up
$NCU --set full --clock-control none --import-source on -k regex:stencil_vec_kernel -s 3 -c 3 \
    -o gpurun_out/prof_stencil python bench.py --steps 3 --warmup 3 --no-extras --no-cpu > /dev/null 2>&1
# 3. full capture of the tcgen05 GEMM and the reductions / gemv
$NCU --set full --clock-control none --import-source on -k regex:"gemm_bf16_tc_kernel|reduce_kernel|gemv_n_kernel|gemv_t_kernel|gemm_simt" -c 12 \
    -o gpurun_out/prof_extras python profiles/run_kernels_once.py > gpurun_out/run_kernels_once.log 2>&1
ls -la gpurun_out
