#!/bin/bash
cd $GRAFT_REPO_ROOT
# post kernel: two launches (one that resamples among them, hopefully); EKF kernel: one launch
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fs3_post -s 12 -c 3 -o gpurun_out/r3_post python bench.py --steps 4 --warmup 12 --no-cpu-baseline > gpurun_out/r3_ncu_post.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fs3_ekf -s 12 -c 1 -o gpurun_out/r3_ekf python bench.py --steps 4 --warmup 12 --no-cpu-baseline > gpurun_out/r3_ncu_ekf.log 2>&1
ls -la gpurun_out/*.ncu-rep
