#!/bin/bash
# GPU call 5 of round 2: full GPU suite, bench with legs, the fused-colsum stall probe, ncu of the small kernels (CSV made
# on the box: a .ncu-rep above 64 MiB would block the copy-back of gpurun_out/).
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rA --tb=short -s > gpurun_out/r02e_gputests.log 2>&1
grep -E "passed|failed" gpurun_out/r02e_gputests.log | tail -2
grep -E "relative L2|fp32 boxes|loss curve|^FAILED|^ERROR|gradient tensors bit" gpurun_out/r02e_gputests.log | cut -c1-900 | head -30
python bench.py --steps 20 --warmup 5 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err
tail -c 300 gpurun_out/r02e_bench.err
python - <<'PY'
import json
l = [x for x in open("gpurun_out/r02e_bench.json") if x.startswith("{")]
if l:
    d = json.loads(l[-1])
    print("BENCH", d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("census") or {}).get("kernels"),
          (d.get("census") or {}).get("torch_glue_kernels"))
PY
for i in 1 2 3; do
  P2R_FUSED_COLSUM=1 P2R_BENCH_SUPERVISE=0 P2R_BENCH_STALL_S=50 P2R_BENCH_TRACE_AFTER_S=45 P2R_BENCH_GDB=1 P2R_BENCH_DEBUG=1 \
  P2R_BENCH_DATA_PATH=0 P2R_BENCH_CENSUS=0 P2R_E2E_PIPELINED=0 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline \
    > gpurun_out/r02e_colsum_$i.json 2> gpurun_out/r02e_colsum_$i.err
  echo "colsum attempt $i rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02e_colsum_$i.json | head -1)"
  grep -E "no progress|cuda-gdb|Kernel|kernel" gpurun_out/r02e_colsum_$i.err | head -12
done
timeout 420 ncu --set full --clock-control none -k regex:"$(python tools/ncu_small_kernels.py --regex)" -c 170 \
  -o gpurun_out/r02_small_kernels python tools/ncu_small_kernels.py > gpurun_out/r02_ncu_small.log 2>&1
tail -2 gpurun_out/r02_ncu_small.log
ncu -i gpurun_out/r02_small_kernels.ncu-rep --page raw --csv > gpurun_out/r02_small_kernels_raw.csv 2>/dev/null
ls -la gpurun_out/r02_small_kernels.ncu-rep
if [ $(stat -c %s gpurun_out/r02_small_kernels.ncu-rep 2>/dev/null || echo 0) -gt 30000000 ]; then rm -f gpurun_out/r02_small_kernels.ncu-rep; fi
du -sh gpurun_out
