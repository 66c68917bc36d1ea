"""GEMM microbench (not the bench.py contract): our tcgen05 kernel vs cuBLAS (torch.matmul) on the hot-path shapes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root (run as `python tools/<name>.py`)

import json
import sys

import torch

from pose2room_b200 import gemm_sm100


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / iters


def main():
    dev = torch.device("cuda:0")
    shapes = [("gcn fwd", 32768, 1600, 1600), ("tcn fwd", 819200, 64, 192), ("mlp 64", 819200, 64, 64),
              ("conv_joint", 16384, 256, 1600), ("sa mlp", 65536, 256, 256)]
    for name, M, N, K in shapes:
        x = torch.randn(M, K, device=dev).bfloat16()
        w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
        dy = torch.randn(M, N, device=dev).bfloat16()
        flops = 2.0 * M * N * K
        rec = {"shape": name, "M": M, "N": N, "K": K}
        for bn in ([0, 128, 160, 256] if N >= 256 else [0]):
            if bn == 160 and N % 160:
                continue
            t = timeit(lambda: gemm_sm100.gemm(x, w, block_n=bn))
            rec["fwd_bn%d_ms" % bn] = t
            rec["fwd_bn%d_tflops" % bn] = flops / t / 1e9
        t = timeit(lambda: x @ w.t())
        rec["cublas_fwd_ms"], rec["cublas_fwd_tflops"] = t, flops / t / 1e9
        t = timeit(lambda: gemm_sm100.gemm(dy, w, False, True))
        rec["dx_ms"], rec["dx_tflops"] = t, flops / t / 1e9
        t = timeit(lambda: dy @ w)
        rec["cublas_dx_ms"] = t
        for sp in [1, 2, 4, 8, 16, 64]:
            if (M + sp - 1) // sp < 256:
                continue
            t = timeit(lambda: gemm_sm100.gemm(dy, x, True, True, out_dtype=torch.float32, splits=sp))
            rec["dw_s%d_ms" % sp] = t
        t = timeit(lambda: dy.t() @ x)
        rec["cublas_dw_ms"] = t
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
