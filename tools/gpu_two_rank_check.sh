#!/bin/bash
# gpurun --gpus 2: the 2-rank NCCL test and the N = 2 bench (NCCL all-reduce + AdamW inside the captured step), exit codes
# included (the teardown of the process group must not turn a finished measurement into a failed torchrun).
mkdir -p gpurun_out
python -m pytest tests/test_multi_gpu.py tests/test_native_ops_gpu.py tests/test_ref_ext_gpu.py -m gpu -q -s --tb=short > gpurun_out/r02j_multigpu_test.log 2>&1
tail -3 gpurun_out/r02j_multigpu_test.log | cut -c1-400
t0=$(date +%s)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 \
  > gpurun_out/r02j_bench_2gpu.json 2> gpurun_out/r02j_bench_2gpu.err
echo "torchrun N=2 rc=$? wall=$(( $(date +%s) - t0 )) s"
grep -E "capture|failed|Error|teardown|no progress" gpurun_out/r02j_bench_2gpu.err | head -5
python - <<'PY'
import json
l = [x for x in open("gpurun_out/r02j_bench_2gpu.json") if x.startswith("{")]
if l:
    d = json.loads(l[-1]); print("N=2", d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"].get("allreduce_in_graph"))
PY
t0=$(date +%s)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 \
  > gpurun_out/r02j_bench_2gpu_ref.json 2>/dev/null
echo "reference arm N=2 rc=$? wall=$(( $(date +%s) - t0 )) s: $(head -c 300 gpurun_out/r02j_bench_2gpu_ref.json)"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
