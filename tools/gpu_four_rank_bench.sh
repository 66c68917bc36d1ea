#!/bin/bash
# gpurun --gpus 4: the N = 4 bench line (weak scaling, NCCL all-reduce + AdamW inside the captured step) and its exit code.
mkdir -p gpurun_out
t0=$(date +%s)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 20 --warmup 5 \
  > gpurun_out/r02l_bench_4gpu.json 2> gpurun_out/r02l_bench_4gpu.err
echo "torchrun N=4 rc=$? wall=$(( $(date +%s) - t0 )) s"
python - <<'PY'
import json
l = [x for x in open("gpurun_out/r02l_bench_4gpu.json") if x.startswith("{")]
if l:
    d = json.loads(l[-1]); print("N=4", d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"].get("allreduce_in_graph"))
PY
