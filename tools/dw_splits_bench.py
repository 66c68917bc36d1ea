"""Weight-gradient GEMM dW[N,K] = dz^T . x (both operands MN-major, split-K with fp32 atomics) against the number of
splits, for the tall shapes of the step.  Diagnostic: picks the split rule in gemm_sm100._Backend.linear_dw."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pose2room_b200 import gemm_sm100

dev = torch.device("cuda:0")
shapes = [(819200, 64, 64), (655360, 64, 64), (16384, 256, 256), (16384, 264, 256), (16384, 256, 1600), (4096, 128, 128),
          (4096, 128, 256), (65536, 256, 256)]
for m, n, k in shapes:
    dz = torch.randn(m, n, device=dev).bfloat16()
    x = torch.randn(m, k, device=dev).bfloat16()
    ref = (dz[:4096].float().t() @ x[:4096].float()) if m <= 4096 else None
    row = []
    for splits in (1, 2, 4, 8, 16, 32, 37, 74, 100, 148, 200, 296, 444, 592):
        if splits > max(1, m // 64):
            continue
        f = lambda: gemm_sm100.gemm(dz, x, True, True, out_dtype=torch.float32, splits=splits)
        for _ in range(3):
            out = f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            f()
        e1.record()
        torch.cuda.synchronize()
        row.append((splits, e0.elapsed_time(e1) * 100))
        if ref is not None:
            assert torch.allclose(out, ref, rtol=2e-2, atol=2.0), (splits, (out - ref).abs().max())
    best = min(row, key=lambda r: r[1])
    print("M=%d N=%d K=%d  best splits=%d %.1f us |" % (m, n, k, best[0], best[1]), " ".join("%d:%.1f" % r for r in row), flush=True)
