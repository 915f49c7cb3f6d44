#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pf or mcl or fastslam2 or mirror" > gpurun_out/r28_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r28_pytest.log; tail -4 gpurun_out/r28_pytest.log
for p in 10 12 14 16 18 20; do
  n=$((1<<p))
  for gph in 1 0; do
    PFGPU_PF_GRAPH=$gph timeout 600 python bench.py --workload pf --particles $n --threshold 0.5 --steps 60 --warmup 10 --no-cpu-baseline > gpurun_out/r28_pf_${p}_g$gph.json 2> gpurun_out/r28_pf_${p}_g$gph.err
  done
done
timeout 600 python bench.py --variant 2 --steps 100 --warmup 10 --no-second --no-cpu-baseline > gpurun_out/r28_fs2_c3.json 2> gpurun_out/r28_fs2_c3.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r28_*.json')):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], d['config'].get('particles'), '%.3e p-steps/s' % d['value'], '%.4f ms' % d['ms_per_step'], 'e2e %.3e' % d['e2e']['value'], 'launches/step %.1f' % (d['gpu_launches']/d['steps']))
    except Exception as e: print(f, 'ERR', e)
PY
