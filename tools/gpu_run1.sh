#!/bin/bash
# first GPU run of the fs3 engine
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r1_smi.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fastslam" > gpurun_out/r1_pytest_fs.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r1_pytest_fs.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r1_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r1_smoke.log
PFGPU_POST_TRACE=1 timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err
echo "bench rc=$?" >> gpurun_out/r1_bench.err
tail -5 gpurun_out/r1_pytest_fs.log; tail -3 gpurun_out/r1_smoke.log; cat gpurun_out/r1_bench.json
