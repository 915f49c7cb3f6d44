#!/bin/bash
# source-level profile of the post kernel on the 2^20-particle configuration (where its per-value phases dominate)
cd $GRAFT_REPO_ROOT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fs3_post -s 8 -c 2 -o gpurun_out/r20_post_c4 python bench.py --config c4 --steps 4 --warmup 6 --no-cpu-baseline --no-second > gpurun_out/r20_ncu_post_c4.log 2>&1
ls -la gpurun_out/r20*
