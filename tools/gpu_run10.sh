#!/bin/bash
cd $GRAFT_REPO_ROOT
BENCH_FLUSH_MODE=none PFGPU_POST_TRACE=1 timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-second > gpurun_out/r10_bench_noflush.json 2> gpurun_out/r10_bench_noflush.err
PFGPU_POST_TRACE=1 timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-second > gpurun_out/r10_bench.json 2> gpurun_out/r10_bench.err
cat gpurun_out/r10_bench_noflush.err gpurun_out/r10_bench.err; grep -h -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.]*\|"avg_launch_ms": [0-9.]*' gpurun_out/r10_bench_noflush.json gpurun_out/r10_bench.json
