#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fastslam" > gpurun_out/r7_pytest_fs.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r7_pytest_fs.log
for hh in 3 2 1; do
PFGPU_EKF_HELPERS=$hh PFGPU_POST_TRACE=1 timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r7_bench_h$hh.json 2> gpurun_out/r7_bench_h$hh.err
done
tail -5 gpurun_out/r7_pytest_fs.log; cat gpurun_out/r7_bench_h3.err; grep -h -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.]*\|"avg_launch_ms": [0-9.]*' gpurun_out/r7_bench_h*.json
