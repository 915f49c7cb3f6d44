#!/bin/bash
cd $GRAFT_REPO_ROOT
for v in 0 1; do
PFGPU_EKF_VARIANT=$v PFGPU_POST_TRACE=1 timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r3_bench_v$v.json 2> gpurun_out/r3_bench_v$v.err
done
cat gpurun_out/r3_bench_v*.err
