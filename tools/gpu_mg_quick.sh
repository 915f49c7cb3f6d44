#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/q2_bench.out 2> gpurun_out/q2_bench.err
echo rc=$?; grep -h "^{\"metric\"" gpurun_out/q2_bench.out gpurun_out/q2_bench.err | tail -1 | cut -c1-900
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29632 bench.py --impl reference --gpus 2 --steps 20 --warmup 3 > gpurun_out/q2_ref.out 2> gpurun_out/q2_ref.err
echo rc=$?; grep -h "\"impl\"" gpurun_out/q2_ref.out gpurun_out/q2_ref.err | tail -1 | cut -c1-300
