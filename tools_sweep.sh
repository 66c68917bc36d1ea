mkdir -p gpurun_out
for cfg in "74" "48" "32" "24" "16"; do
  P2R_DW_PAIRS=$cfg timeout 200 python bench.py --no-cpu-baseline --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('dw_pairs=$cfg', 'ms/step', round(d['ms_per_step'],3), 'gemm_ms', round(d['roofline']['avg_launch_ms'],4))"
done > gpurun_out/sweep.txt 2>&1
cat gpurun_out/sweep.txt
