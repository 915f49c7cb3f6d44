#!/bin/bash
cd $GRAFT_REPO_ROOT
for cfg in "256 128" "512 64" "512 128" "256 64" "512 32"; do
  set -- $cfg
  PFGPU_POST_NT=$1 PFGPU_POST_TILES=$2 timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-second > gpurun_out/r43_$1_$2.json 2> gpurun_out/r43_$1_$2.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r43_$1_$2.json')); print('NT=$1 tiles=$2', '%.4e'%d['value'], '%.2f us'%(d['ms_per_step']*1e3), 'noflush %.4e'%d['value_steady_state_no_flush'])
except Exception as e: print('NT=$1 tiles=$2 ERR', e)
PY
done
