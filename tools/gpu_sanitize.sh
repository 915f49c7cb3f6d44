#!/bin/bash
# compute-sanitizer (memcheck, racecheck, synccheck, initcheck) over smoke-sized runs of every kernel family; summaries -> gpurun_out/
cd $GRAFT_REPO_ROOT
for tool in memcheck racecheck synccheck; do
  for what in "tests/test_gpu_parity.py -k 'fastslam_trajectory_bit_exact and 1000'" "tests/test_gpu_parity.py -k 'sharded_in_process_edge'" "tests/test_gpu_parity.py -k 'pf_trajectory_bit_exact and 1000'" "tests/test_gpu_parity.py -k 'mcl_kld_adaptive'"; do
    tag=$(echo "$what" | sed 's/[^a-zA-Z0-9]/_/g' | cut -c20-70)
    eval timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python -m pytest $what -x -q -m gpu > gpurun_out/san_${tool}_${tag}.log 2>&1
    echo "== $tool $what: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/san_${tool}_${tag}.log | tr '\n' ' ')" | tee -a gpurun_out/san_summary.txt
  done
done
