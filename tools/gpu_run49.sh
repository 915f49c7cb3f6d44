#!/bin/bash
cd $GRAFT_REPO_ROOT
rm -f gpurun_out/san3_summary.txt
for tool in memcheck racecheck; do
  for what in "tests/test_gpu_parity.py -k 'fastslam_trajectory_bit_exact and 1000'" "tests/test_gpu_parity.py -k 'sharded_in_process_edge'" "tests/test_gpu_parity.py -k 'pf_trajectory_bit_exact and 1000'"; do
    tag=$(echo "$what" | sed 's/[^a-zA-Z0-9]/_/g' | cut -c20-70)
    eval timeout 600 compute-sanitizer --tool $tool --print-limit 20 python -m pytest $what -x -q -m gpu > gpurun_out/san3_${tool}_${tag}.log 2>&1
    echo "== $tool $what: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/san3_${tool}_${tag}.log | tr '\n' ' ')" | tee -a gpurun_out/san3_summary.txt
  done
done
grep -h "Race reported" -A1 gpurun_out/san3_racecheck_*.log | grep -o "fs3.cuh:[0-9]*\|pf3.cuh:[0-9]*\|xsum.cuh:[0-9]*\|pf_kernels.cuh:[0-9]*" | sort | uniq -c
