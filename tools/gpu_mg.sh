#!/bin/bash
# usage: bash tools/gpu_mg.sh N   — multi-GPU parity tests (up to N GPUs) + bench at N ranks; logs under gpurun_out/
N=${1:-2}
cd $GRAFT_REPO_ROOT
nvidia-smi -L > gpurun_out/mg${N}_gpus.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/mg${N}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/mg${N}_pytest.log
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29617 bench.py --gpus $n --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/mg${N}_bench_n$n.out 2> gpurun_out/mg${N}_bench_n$n.err; grep -h "^{\"metric\"" gpurun_out/mg${N}_bench_n$n.out gpurun_out/mg${N}_bench_n$n.err | tail -1 > gpurun_out/mg${N}_bench_n$n.json
    echo "bench n=$n rc=$?" >> gpurun_out/mg${N}_pytest.log
  fi
done
timeout 600 python bench.py --gpus 1 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/mg${N}_bench_n1.json 2> gpurun_out/mg${N}_bench_n1.err
# stage trace of the post kernel at N ranks (its own short run: the stamps cost a little)
PFGPU_POST_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29618 bench.py --gpus $N --steps 60 --warmup 10 --no-cpu-baseline > gpurun_out/mg${N}_trace.out 2> gpurun_out/mg${N}_trace.err
grep -h "fs3_post_kernel\|per RESAMPLE\|leader chain\|classify split" gpurun_out/mg${N}_trace.err gpurun_out/mg${N}_trace.out > gpurun_out/mg${N}_stage_trace.txt
tail -6 gpurun_out/mg${N}_pytest.log
for f in gpurun_out/mg${N}_bench_n*.json; do python - <<PY
import json
try:
    d=json.load(open("$f")); q=d.get("c4_strong",{})
    print("$f", "c3 %.3e %.4f ms" % (d["value"], d["ms_per_step"]), "c4 %.3e %.4f ms" % (q.get("value",0), q.get("ms_per_step",0)))
except Exception as e: print("$f", "ERR", e)
PY
done
