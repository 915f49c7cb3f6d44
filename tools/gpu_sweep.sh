#!/bin/bash
# BASELINE config 5: particle-count sweep of the PF step (C1 model, 5 landmarks), thresholds 1.0 (resample every step) and 0.5;
# + config 2 (MCL, 2^20 particles x 360 beams).  One GPU.  JSON lines -> gpurun_out/sweep3_*.json
cd $GRAFT_REPO_ROOT
for thr in 1.0 0.5; do
  for p in 10 12 14 16 18 20 22 24; do
    n=$((1<<p))
    timeout 600 python bench.py --workload pf --particles $n --threshold $thr --steps 60 --warmup 10 --no-cpu-baseline > gpurun_out/sweep3_pf_${thr}_$p.json 2> gpurun_out/sweep3_pf_${thr}_$p.err
  done
done
timeout 900 python bench.py --workload mcl --particles $((1<<20)) --steps 30 --warmup 5 > gpurun_out/sweep3_mcl_c2.json 2> gpurun_out/sweep3_mcl_c2.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/sweep3_*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f.split('/')[-1], d['config']['particles'], '%.3e p-steps/s' % d['value'], '%.4f ms' % d['ms_per_step'], 'launches/step %.1f' % (d['gpu_launches']/d['steps']), 'main kernel frac %.3f' % r['frac'])
    except Exception as e: print(f, 'ERR', e)
PY
# FastSLAM 2.0 on the C3 shape (informational: same engine, proposal kernel in front of the EKF launch)
timeout 600 python bench.py --variant 2 --steps 100 --warmup 10 --no-second > gpurun_out/sweep3_fs2_c3.json 2> gpurun_out/sweep3_fs2_c3.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/sweep3_fs2_c3.json')); print('fs2 c3', '%.3e'%d['value'], d['ms_per_step'], d['e2e']['value'], d.get('cpu_baseline',{}).get('value'))
PY
