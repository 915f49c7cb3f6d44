#!/bin/bash
cd $GRAFT_REPO_ROOT
BENCH_EMPTY_OBS=1 PFGPU_POST_TRACE=1 timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-second > gpurun_out/r15_empty.json 2> gpurun_out/r15_empty.err
BENCH_EMPTY_OBS=1 BENCH_FLUSH_MODE=none PFGPU_POST_TRACE=1 timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-second > gpurun_out/r15_empty_nf.json 2> gpurun_out/r15_empty_nf.err
cat gpurun_out/r15_empty.err gpurun_out/r15_empty_nf.err; grep -h -o '"ms_per_step": [0-9.]*' gpurun_out/r15_empty.json gpurun_out/r15_empty_nf.json
