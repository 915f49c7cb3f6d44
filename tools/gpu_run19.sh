#!/bin/bash
cd $GRAFT_REPO_ROOT
PFGPU_POST_TRACE=1 timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r19_bench.json 2> gpurun_out/r19_bench.err
cat gpurun_out/r19_bench.err; grep -h -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.]*\|"avg_launch_ms": [0-9.]*' gpurun_out/r19_bench.json
