#!/bin/bash
# gpurun --gpus 2: the 2-rank NCCL test and the N = 2 bench (NCCL all-reduce + AdamW inside the captured step, then the
# round-1 arrangement for comparison).
mkdir -p gpurun_out
python -m pytest tests/test_multi_gpu.py -m gpu -q -s --tb=short > gpurun_out/r02g_multigpu_test.log 2>&1
tail -5 gpurun_out/r02g_multigpu_test.log | cut -c1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 \
  > gpurun_out/r02g_bench_2gpu.json 2> gpurun_out/r02g_bench_2gpu.err
grep -E "capture|failed|Error" gpurun_out/r02g_bench_2gpu.err | head -5
python - <<'PY'
import json
l = [x for x in open("gpurun_out/r02g_bench_2gpu.json") if x.startswith("{")]
if l:
    d = json.loads(l[-1]); print("N=2", d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"])
PY
P2R_GRAPH_ALLREDUCE=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 \
  > gpurun_out/r02g_bench_2gpu_eager_allreduce.json 2>/dev/null
python - <<'PY'
import json
l = [x for x in open("gpurun_out/r02g_bench_2gpu_eager_allreduce.json") if x.startswith("{")]
if l:
    d = json.loads(l[-1]); print("N=2, all-reduce after the replay", d["value"], d["ms_per_step"])
PY
