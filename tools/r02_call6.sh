#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rA --tb=short -s > gpurun_out/r02f_gputests.log 2>&1
grep -E "passed|failed" gpurun_out/r02f_gputests.log | tail -2
grep -E "relative L2|per-proposal|loss curve|sa_fused gradients|^FAILED|^ERROR|gradient tensors bit" gpurun_out/r02f_gputests.log | cut -c1-1100 | head -30
python bench.py --steps 20 --warmup 5 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err
tail -c 300 gpurun_out/r02f_bench.err
python - <<'PY'
import json
l = [x for x in open("gpurun_out/r02f_bench.json") if x.startswith("{")]
if l:
    d = json.loads(l[-1])
    print("BENCH", d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("census") or {}).get("kernels"),
          (d.get("census") or {}).get("torch_glue_kernels"), d.get("first_step_loss"))
PY
for i in 1 2 3; do
  P2R_FUSED_COLSUM=1 P2R_BENCH_SUPERVISE=0 P2R_BENCH_STALL_S=50 P2R_BENCH_TRACE_AFTER_S=45 P2R_BENCH_GDB=1 P2R_BENCH_DEBUG=1 \
  P2R_BENCH_DATA_PATH=0 P2R_BENCH_CENSUS=0 P2R_E2E_PIPELINED=0 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline \
    > gpurun_out/r02f_colsum_$i.json 2> gpurun_out/r02f_colsum_$i.err
  echo "colsum attempt $i rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02f_colsum_$i.json | head -1) $(grep -o '"first_step_loss": [0-9.]*' gpurun_out/r02f_colsum_$i.json | head -1)"
  grep -E "no progress|cuda-gdb|Kernel|illegal" gpurun_out/r02f_colsum_$i.err | head -8
done
P2R_FUSED_RESADD=0 P2R_BENCH_SUPERVISE=0 P2R_BENCH_DATA_PATH=0 P2R_BENCH_CENSUS=0 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02f_noresadd.json 2>/dev/null
echo "no-resadd: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02f_noresadd.json | head -1)"
du -sh gpurun_out
