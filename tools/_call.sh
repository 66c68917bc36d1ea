mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dense_ops_gpu.py tests/test_model_gpu.py tests/test_gemm_sm100_gpu.py tests/test_sa_fused_gpu.py -q -m gpu -x 2>&1 | tail -4
timeout 300 python tools/glue_trace.py --top 100 > gpurun_out/glue_trace.txt 2>gpurun_out/glue_trace.err; tail -3 gpurun_out/glue_trace.err
export P2R_BENCH_VARIANTS=0 P2R_BENCH_EXPERIMENTS=0 P2R_BENCH_LEGS=0
timeout 500 python bench.py --no-cpu-baseline 2>gpurun_out/sel_err.log | tail -1 > gpurun_out/sel.json
python - <<PY
import json
d=json.loads(open("gpurun_out/sel.json").read())
c=d["census"]
print(d["value"], d["ms_per_step"], d.get("first_step_loss"), {k:c[k] for k in ['kernels','kernel_time_us','span_us','idle_us','overlapped_us','torch_glue_kernels','torch_glue_time_us']})
PY
