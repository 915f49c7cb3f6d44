#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/final3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final3_pytest.log; tail -3 gpurun_out/final3_pytest.log
python bench.py > gpurun_out/final3_bench.json 2> gpurun_out/final3_bench.err
PFGPU_POST_TRACE=1 BENCH_FLUSH_MODE=none timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-second > gpurun_out/final3_trace.json 2> gpurun_out/final3_trace.err; cat gpurun_out/final3_trace.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/final3_bench.json')); print('ours', d['value'], d['ms_per_step'], 'noflush', d['value_steady_state_no_flush'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'c4', d['c4_strong']['value'], d['c4_strong']['ekf_roofline_frac'], 'cpu', d['cpu_baseline']['value'])
PY
bash tools/gpu_ncu2.sh > /dev/null 2>&1
ls gpurun_out/r45* | head
