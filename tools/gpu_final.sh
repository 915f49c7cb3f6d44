#!/bin/bash
# final validation of the round: the whole GPU suite, compute-sanitizer over the kernels added late (fused PF tail, FastSLAM 2.0 proposal)
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/final_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/final_pytest.log; tail -4 gpurun_out/final_pytest.log
rm -f gpurun_out/san2_summary.txt
for tool in memcheck racecheck synccheck; do
  for what in "tests/test_gpu_parity.py -k 'pf_step_paths_agree'" "tests/test_gpu_parity.py -k 'fastslam2_trajectory_bit_exact and 64'" "tests/test_gpu_parity.py -k 'fastslam2_edge'"; do
    tag=$(echo "$what" | sed 's/[^a-zA-Z0-9]/_/g' | cut -c20-70)
    eval timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest $what -x -q -m gpu > gpurun_out/san2_${tool}_${tag}.log 2>&1
    echo "== $tool $what: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/san2_${tool}_${tag}.log | tr '\n' ' ')" | tee -a gpurun_out/san2_summary.txt
  done
done
( time python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" ) 2>&1 | tail -4
python bench.py --steps 100 --warmup 10 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -2 gpurun_out/final_bench.err; cut -c1-300 gpurun_out/final_bench.json
