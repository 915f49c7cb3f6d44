#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pf or mcl" > gpurun_out/r32_pytest.log 2>&1; tail -2 gpurun_out/r32_pytest.log
eval timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -k 'pf_step_paths_agree' -x -q -m gpu > gpurun_out/san2_racecheck_ty_py__k__pf_step_paths_agree_.log 2>&1
grep -E 'RACECHECK SUMMARY|passed|failed' gpurun_out/san2_racecheck_ty_py__k__pf_step_paths_agree_.log
