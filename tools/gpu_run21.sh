#!/bin/bash
cd $GRAFT_REPO_ROOT
for v in 1 0; do
BENCH_VERBOSE=1 BENCH_KERNEL_TIMER=$v timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-second > gpurun_out/r21_t$v.json 2> gpurun_out/r21_t$v.err
echo "timer=$v"; cat gpurun_out/r21_t$v.err | tail -3; grep -h -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.]*\|"avg_launch_ms": [0-9.]*\|"no_flush[a-z_]*": [0-9.e+]*' gpurun_out/r21_t$v.json
done
BENCH_VERBOSE=1 BENCH_KERNEL_TIMER=0 BENCH_FLUSH_MODE=none timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-second > gpurun_out/r21_nf.json 2> gpurun_out/r21_nf.err
echo "noflush"; tail -3 gpurun_out/r21_nf.err; grep -h -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.]*' gpurun_out/r21_nf.json
