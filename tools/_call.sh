mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dense_ops_gpu.py tests/test_model_gpu.py tests/test_gemm_sm100_gpu.py -q -m gpu -x 2>&1 | tail -3
export P2R_BENCH_VARIANTS=0 P2R_BENCH_EXPERIMENTS=0 P2R_BENCH_LEGS=0
for v in 1 0 1; do
P2R_FUSED_COLSUM1=$v timeout 400 python bench.py --no-cpu-baseline 2>gpurun_out/err_$v.log | tail -1 > gpurun_out/run_$v.json
python - <<PY
import json
d=json.loads(open("gpurun_out/run_$v.json").read())
print("COLSUM1=$v", d["value"], d["ms_per_step"], d.get("first_step_loss"), d["config"].get("retry"), d["census"]["kernels"], d["census"]["kernel_time_us"])
PY
done
