"""Launches the hot kernels at their BASELINE shapes a few times each (for `ncu --set full -k regex:...` captures)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root (run as `python tools/<name>.py`)

import torch
import torch.nn as nn

from pose2room_b200 import gemm_sm100, ops

dev = torch.device("cuda:0")
gemm_sm100.install()
M, V, C = 32768, 25, 64
x = torch.randn(M, V * C, device=dev).bfloat16()
w = (torch.randn(V * C, V * C, device=dev) / 40).bfloat16()
dy = torch.randn(M, V * C, device=dev).bfloat16()
for _ in range(4):
    y = gemm_sm100.gemm(x, w)                                              # graph-conv forward  (BLOCK_N 160, TN)
    dx = gemm_sm100.gemm(dy, w, False, True)                               # input gradient      (BLOCK_N 256, NT)
    dw = gemm_sm100.gemm(dy, x, True, True, out_dtype=torch.float32)       # weight gradient     (MN-major both)
rows = torch.randn(M * V, C, device=dev).bfloat16().requires_grad_(True)
bn = nn.BatchNorm2d(C).to(dev)
conv = nn.Conv2d(C, C, (3, 1), (1, 1), (1, 0)).to(dev)
for _ in range(4):
    h = ops.batchnorm_act(rows, bn, relu=True)
    t = ops.temporal_conv(h.reshape(32, 1024, V, C), conv.weight, conv.bias)
    out = ops.batchnorm_act(t, bn, relu=True, residual=rows)
    out.backward(torch.ones_like(out))
torch.cuda.synchronize()
print("done")
