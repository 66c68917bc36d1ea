#!/bin/bash
# gpurun --gpus 8: the N = 8 bench line (weak scaling, NCCL all-reduce + AdamW inside the captured step) and its exit code.
mkdir -p gpurun_out
t0=$(date +%s)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 5 \
  > gpurun_out/r02k_bench_8gpu.json 2> gpurun_out/r02k_bench_8gpu.err
echo "torchrun N=8 rc=$? wall=$(( $(date +%s) - t0 )) s"
python - <<'PY'
import json
l = [x for x in open("gpurun_out/r02k_bench_8gpu.json") if x.startswith("{")]
if l:
    d = json.loads(l[-1]); print("N=8", d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"].get("allreduce_in_graph"))
PY
