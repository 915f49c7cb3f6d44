#!/bin/bash
cd $GRAFT_REPO_ROOT
PFGPU_POST_TRACE=1 BENCH_FLUSH_MODE=none timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-second > gpurun_out/r35_nf.json 2> gpurun_out/r35_nf.err
cat gpurun_out/r35_nf.err; grep -h -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.]*' gpurun_out/r35_nf.json | head -3
PFGPU_POST_TRACE=1 timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-second > gpurun_out/r35_fl.json 2> gpurun_out/r35_fl.err
grep "timeline" gpurun_out/r35_fl.err; grep -h -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.]*' gpurun_out/r35_fl.json | head -3
