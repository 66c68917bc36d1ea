#!/bin/bash
# One GPU call that regenerates the evidence kept under profiles/ (gpurun --timeout 2400 -- "bash tools/gpu_round_evidence.sh"):
# the full GPU suite, the bench line, the multi-stream stress runs, an ncu --set full pass over the GEMMs / streaming BatchNorm
# kernels (CSV made on the box: a large .ncu-rep cannot travel) and the ncu launch list of an eager window of the bench.
# (tools/ncu_small_kernels.py + tools/summarize_ncu.py do the same for every other kernel.)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rA --tb=short -s > gpurun_out/r02h_gputests.log 2>&1
grep -E "passed|failed" gpurun_out/r02h_gputests.log | tail -2
grep -E "^FAILED|^ERROR" gpurun_out/r02h_gputests.log | cut -c1-300 | head
python bench.py --steps 20 --warmup 5 > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err
tail -c 300 gpurun_out/r02h_bench.err
python - <<'PY'
import json
l = [x for x in open("gpurun_out/r02h_bench.json") if x.startswith("{")]
if l:
    d = json.loads(l[-1])
    print("BENCH", d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("census") or {}).get("kernels"),
          (d.get("census") or {}).get("torch_glue_kernels"), d.get("first_step_loss"))
    dp = d.get("data_path") or {}
    print("make_batch", dp.get("kernel_ms"), dp.get("kernel_frac_of_hbm_peak"), (dp.get("variants") or {}))
PY
timeout 200 python tools/stress_multistream.py 2>&1 | grep -E "STRESS|Error|error" | cut -c1-300
P2R_FUSED_COLSUM=1 timeout 200 python tools/stress_multistream.py 2>&1 | grep -E "STRESS|Error|error" | cut -c1-300
P2R_GCN_PAIR_DW=0 timeout 200 python tools/stress_multistream.py 2>&1 | grep -E "STRESS|Error|error" | cut -c1-300
echo "stress rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm2_bf16|gemm2_dw|gemm_bf16_kernel|stream_bn" -c 48 \
  -o gpurun_out/r02_hot_kernels python tools/ncu_round.py > gpurun_out/r02_ncu_hot.log 2>&1
tail -2 gpurun_out/r02_ncu_hot.log
ncu -i gpurun_out/r02_hot_kernels.ncu-rep --page raw --csv > gpurun_out/r02_hot_kernels_raw.csv 2>/dev/null
ls -la gpurun_out/r02_hot_kernels.ncu-rep
if [ $(stat -c %s gpurun_out/r02_hot_kernels.ncu-rep 2>/dev/null || echo 0) -gt 40000000 ]; then rm -f gpurun_out/r02_hot_kernels.ncu-rep; fi
timeout 500 ncu --set full --clock-control none -k regex:"$(python tools/ncu_small_kernels.py --regex)" -c 260 \
  -o gpurun_out/r02_small_kernels python tools/ncu_small_kernels.py > gpurun_out/r02_ncu_small.log 2>&1
tail -2 gpurun_out/r02_ncu_small.log
ncu -i gpurun_out/r02_small_kernels.ncu-rep --page raw --csv > gpurun_out/r02_small_kernels_raw.csv 2>/dev/null
rm -f gpurun_out/r02_small_kernels.ncu-rep
timeout 120 python tools/tconv_bench.py > gpurun_out/r02_tconv_bench.txt 2>&1
P2R_TCONV_HALO=0 timeout 120 python tools/tconv_bench.py >> gpurun_out/r02_tconv_bench.txt 2>&1
P2R_TCONV_TRACE=1 timeout 120 python tools/tconv_bench.py > gpurun_out/r02_tconv_timeline.txt 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02_launches.csv \
  env P2R_BENCH_SUPERVISE=0 P2R_BENCH_DATA_PATH=0 P2R_BENCH_CENSUS=0 P2R_CUDA_GRAPH=0 P2R_E2E_PIPELINED=0 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_launches_run.log 2>&1
ls -la gpurun_out/r02_launches.csv
du -sh gpurun_out
