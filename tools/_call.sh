mkdir -p gpurun_out
timeout 300 python tools/dw_splits_bench.py > gpurun_out/dw_splits.txt 2>gpurun_out/dw_splits.err; tail -3 gpurun_out/dw_splits.err
