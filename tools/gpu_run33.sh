#!/bin/bash
cd $GRAFT_REPO_ROOT
python bench.py > gpurun_out/r33_bench_default.json 2> gpurun_out/r33_bench_default.err; tail -2 gpurun_out/r33_bench_default.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r33_bench_default.json')); print('default', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['roofline']['traffic_source'][:30])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pf3_post -s 20 -c 2 -o gpurun_out/r33_pf3 python bench.py --workload pf --particles 65536 --threshold 0.5 --steps 8 --warmup 12 --no-cpu-baseline > gpurun_out/r33_ncu_pf3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fs2_propose -s 12 -c 1 -o gpurun_out/r33_fs2 python bench.py --variant 2 --steps 4 --warmup 12 --no-cpu-baseline --no-second > gpurun_out/r33_ncu_fs2.log 2>&1
ls -la gpurun_out/r33*
