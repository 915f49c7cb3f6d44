#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fastslam" > gpurun_out/r25_pytest_fs.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r25_pytest_fs.log; tail -3 gpurun_out/r25_pytest_fs.log
for e in 1 0; do
PFGPU_EARLY_LAUNCH=$e BENCH_VERBOSE=1 timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r25_e$e.json 2> gpurun_out/r25_e$e.err
echo "early=$e"; tail -2 gpurun_out/r25_e$e.err; grep -h -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.]*\|"avg_launch_ms": [0-9.]*\|"value_steady_state_no_flush": [0-9.e+]*' gpurun_out/r25_e$e.json
done
