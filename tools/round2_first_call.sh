#!/bin/bash
# First GPU call of the next round (everything written in the CPU-only fourth session of round 1 gets its first real run):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'
# 1. the non-gating tests, verbosely and WITHOUT the xfail cushion; 2. the 1000-scene eval test with its timing line;
# 3. the bench (headline + data-path variants + experiments + census); 4. ncu: launch list of a step and a full capture of
# the kernels that have none yet.  Outputs under gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests/test_model_gpu.py tests/test_dataloader_gpu.py tests/test_demo_sequence.py -m gpu -q --runxfail -k "fused or pipelined or demo" \
    > gpurun_out/r02_fused_tests.log 2>&1
python -m pytest tests/test_zz_eval_1k_gpu.py -m gpu -q -s --runxfail > gpurun_out/r02_eval1k.log 2>&1
python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"make_batch|detection_loss|gmm_mix|vote_tail" \
    -o gpurun_out/r02_new_kernels python tools/ncu_new_kernels.py > gpurun_out/r02_ncu_new.log 2>&1
tail -5 gpurun_out/r02_fused_tests.log gpurun_out/r02_eval1k.log
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02_bench.json") if l.startswith("{")][-1])
print(json.dumps({k: d.get(k) for k in ("value", "ms_per_step", "e2e", "experiments", "first_step_loss")}, indent=1)[:3000])
print("census:", json.dumps(d.get("census"))[:1500])
print("data_path:", json.dumps(d.get("data_path"))[:1200])
PY
