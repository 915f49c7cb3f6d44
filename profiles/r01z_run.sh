#!/bin/bash
# round-1 final validation on one B200: tests, bench (our arm + reference arm), smoke, ncu launch list and two full captures
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r01z
(time timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3) 2>&1 | grep -v "^$" | tail -5
PFGPU_POST_TRACE=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r01z/bench_traced.json 2> gpurun_out/r01z/bench_traced.err
grep "phase times\|chain detail" gpurun_out/r01z/bench_traced.err
timeout 600 python bench.py > gpurun_out/r01z/bench.json 2> gpurun_out/r01z/bench.err; cut -c1-260 gpurun_out/r01z/bench.json
timeout 600 python bench.py --impl reference > gpurun_out/r01z/bench_reference.json 2> gpurun_out/r01z/bench_reference.err; cut -c1-200 gpurun_out/r01z/bench_reference.json
timeout 300 python bench.py --workload mcl --no-cpu-baseline > gpurun_out/r01z/bench_mcl.json 2>/dev/null; cut -c1-200 gpurun_out/r01z/bench_mcl.json
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01z/launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fs_ekf_kernel -s 4 -c 1 -o gpurun_out/r01z/ekf python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fs_post_kernel -s 4 -c 1 -o gpurun_out/r01z/post python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/r01z | tail -12
