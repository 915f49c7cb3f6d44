#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/r45_launches.csv python bench.py --steps 8 --warmup 12 --no-cpu-baseline --no-second > gpurun_out/r45_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fs3_ekf -s 12 -c 1 -o gpurun_out/r45_ekf python bench.py --steps 4 --warmup 12 --no-cpu-baseline --no-second > gpurun_out/r45_ncu_ekf.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fs3_post -s 12 -c 3 -o gpurun_out/r45_post python bench.py --steps 4 --warmup 12 --no-cpu-baseline --no-second > gpurun_out/r45_ncu_post.log 2>&1
ls -la gpurun_out/r45*
