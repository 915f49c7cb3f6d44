#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pf or mcl" > gpurun_out/r47_pytest.log 2>&1; tail -2 gpurun_out/r47_pytest.log
for p in 10 14 16 18; do
  n=$((1<<p))
  timeout 600 python bench.py --workload pf --particles $n --threshold 0.5 --steps 60 --warmup 10 --no-cpu-baseline > gpurun_out/r47_pf_$p.json 2> gpurun_out/r47_pf_$p.err
done
timeout 600 python bench.py --workload pf --particles 1000 --threshold 0.5 --steps 200 --warmup 20 > gpurun_out/r47_pf_c1.json 2> gpurun_out/r47_pf_c1.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r47_*.json')):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], d['config'].get('particles'), '%.3e p-steps/s' % d['value'], '%.4f ms' % d['ms_per_step'], 'e2e %.3e' % d['e2e']['value'], 'cpu', d.get('cpu_baseline',{}).get('value'))
    except Exception as e: print(f, 'ERR', e)
PY
