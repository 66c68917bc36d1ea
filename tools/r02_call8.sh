#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rA --tb=short -s > gpurun_out/r02i_gputests.log 2>&1
grep -E "passed|failed" gpurun_out/r02i_gputests.log | tail -2
grep -E "^FAILED|^ERROR" gpurun_out/r02i_gputests.log | cut -c1-300 | head
python bench.py --steps 20 --warmup 5 > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err
tail -c 300 gpurun_out/r02i_bench.err
python - <<'PY'
import json
l = [x for x in open("gpurun_out/r02i_bench.json") if x.startswith("{")]
if l:
    d = json.loads(l[-1])
    print("BENCH", d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("census") or {}).get("kernels"),
          (d.get("census") or {}).get("torch_glue_kernels"), d.get("first_step_loss"))
    print("SA ops", [(r["shape"], r["impl"], round(r["fps_us"], 1)) for r in d.get("sa_operators") or []], d.get("sa_module_forward"))
    print("fwd", d.get("forward_only"))
PY
