#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fastslam and not sharded" > gpurun_out/r2_pytest_fs.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_fs.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sharded_in_process" > gpurun_out/r2_pytest_shard.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_shard.log
for v in 0 1 3; do
PFGPU_EKF_VARIANT=$v PFGPU_POST_TRACE=1 timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_bench_v$v.json 2> gpurun_out/r2_bench_v$v.err
done
PFGPU_POST_NT=512 PFGPU_POST_TRACE=1 timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_bench_nt512.json 2> gpurun_out/r2_bench_nt512.err
tail -5 gpurun_out/r2_pytest_fs.log; tail -5 gpurun_out/r2_pytest_shard.log; grep -h -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.]*\|"avg_launch_ms": [0-9.]*' gpurun_out/r2_bench_*.json
