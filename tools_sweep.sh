mkdir -p gpurun_out
for cfg in "0 0" "1 1" "1 0"; do
  set -- $cfg
  P2R_GCN_PAIR_DW=$1 P2R_GCN_DW_INLINE=$2 timeout 200 python bench.py --no-cpu-baseline --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('pair_dw=$1 inline=$2', 'ms/step', round(d['ms_per_step'],3), 'gemm_ms', round(d['roofline']['avg_launch_ms'],4))"
done > gpurun_out/sweep.txt 2>&1
cat gpurun_out/sweep.txt
