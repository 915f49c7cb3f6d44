#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pf or mcl or mirror or xsum" > gpurun_out/r29_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r29_pytest.log; tail -12 gpurun_out/r29_pytest.log
for p in 10 14 16 18 20; do
  n=$((1<<p))
  for f in 1 0; do
    PFGPU_PF_FUSED=$f timeout 600 python bench.py --workload pf --particles $n --threshold 0.5 --steps 60 --warmup 10 --no-cpu-baseline > gpurun_out/r29_pf_${p}_f$f.json 2> gpurun_out/r29_pf_${p}_f$f.err
  done
done
timeout 600 python bench.py --workload pf --particles 1000 --threshold 0.5 --steps 200 --warmup 20 > gpurun_out/r29_pf_c1.json 2> gpurun_out/r29_pf_c1.err
timeout 900 python bench.py --workload mcl --particles $((1<<20)) --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r29_mcl_c2.json 2> gpurun_out/r29_mcl_c2.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r29_*.json')):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], d['config'].get('particles'), '%.3e p-steps/s' % d['value'], '%.4f ms' % d['ms_per_step'], 'e2e %.3e' % d['e2e']['value'], 'launches/step %.1f' % (d['gpu_launches']/d['steps']), 'cpu', d.get('cpu_baseline',{}).get('value'))
    except Exception as e: print(f, 'ERR', e)
PY
