mkdir -p gpurun_out
timeout 250 python tools/stress_multistream.py 2>&1 | grep -E "STRESS|Error|error" | cut -c1-300
export P2R_BENCH_VARIANTS=0 P2R_BENCH_EXPERIMENTS=0 P2R_BENCH_LEGS=0
for i in 1 2 3; do
timeout 400 python bench.py --no-cpu-baseline 2>gpurun_out/err_$i.log | tail -1 > gpurun_out/run_$i.json
python - <<PY
import json
d=json.loads(open("gpurun_out/run_$i.json").read())
print("run $i", d["value"], d["ms_per_step"], d.get("first_step_loss"), d["config"].get("retry"), d["roofline"]["other_kernels"][0]["avg_launch_ms"], d["roofline"]["other_kernels"][0]["executed_over_dense"])
PY
done
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_gemm_sm100_gpu.py tests/test_bf16_parity_gpu.py -q -m gpu -x 2>&1 | tail -3
