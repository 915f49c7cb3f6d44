#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fastslam" > gpurun_out/r44_pytest_fs.log 2>&1; tail -2 gpurun_out/r44_pytest_fs.log
for e in 1 0 1 0; do
PFGPU_EARLY_LAUNCH=$e timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-second > gpurun_out/r44_e$e.json 2> gpurun_out/r44_e$e.err
python - <<PY
import json
d=json.load(open('gpurun_out/r44_e$e.json')); print('early=$e', '%.4e'%d['value'], '%.2f us'%(d['ms_per_step']*1e3), 'noflush %.4e'%d['value_steady_state_no_flush'], 'e2e %.4e'%d['e2e']['value'])
PY
done
