mkdir -p gpurun_out
for cfg in "1 1" "1 0" "0 1"; do
  set -- $cfg
  P2R_GCN_PAIR=$1 P2R_FUSED_STATS=$2 timeout 200 python bench.py --no-cpu-baseline --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('pair=$1 stats=$2', 'ms/step', round(d['ms_per_step'],3), 'gemm_ms', round(d['roofline']['avg_launch_ms'],4), 'frac', round(d['roofline']['frac'],3))"
done > gpurun_out/sweep.txt 2>&1
cat gpurun_out/sweep.txt
