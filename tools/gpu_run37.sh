#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fastslam or pf or mcl" > gpurun_out/r37_pytest_fs.log 2>&1; tail -2 gpurun_out/r37_pytest_fs.log
for i in 1 2; do
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-second > gpurun_out/r37_b$i.json 2> gpurun_out/r37_b$i.err
python - <<PY
import json
d=json.load(open('gpurun_out/r37_b$i.json')); print('run $i', d['value'], d['ms_per_step'], 'noflush', d['value_steady_state_no_flush'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])
PY
done
