#!/bin/bash
# bench only (no tests) at 1, 2 and 4 ranks on a 4-GPU box: refresh of the scaling lines with the final build
cd $GRAFT_REPO_ROOT
timeout 600 python bench.py --gpus 1 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/mg4b_n1.json 2> gpurun_out/mg4b_n1.err
for n in 2 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2964$n bench.py --gpus $n --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/mg4b_n$n.out 2> gpurun_out/mg4b_n$n.err
  grep -h "^{\"metric\"" gpurun_out/mg4b_n$n.out gpurun_out/mg4b_n$n.err | tail -1 > gpurun_out/mg4b_n$n.json
done
for f in gpurun_out/mg4b_n*.json; do python - <<PY
import json
try:
    d=json.load(open("$f")); q=d.get("c4_strong",{})
    print("$f", "c3 %.3e %.4f ms" % (d["value"], d["ms_per_step"]), "c4 %.3e %.4f ms" % (q.get("value",0), q.get("ms_per_step",0)), "e2e %.3e" % d["e2e"]["value"])
except Exception as e: print("$f", "ERR", e)
PY
done
