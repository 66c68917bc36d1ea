"""Launches each hot kernel of the train step once or twice at its BASELINE shape (for one `ncu --set full` capture)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root (run as `python tools/<name>.py`)

import numpy as np
import torch
import torch.nn as nn

from pose2room_b200 import gemm_sm100, ops
from pose2room_b200.p2rnet.graph import layout_for_joints, spatial_adjacency

dev = torch.device("cuda:0")
gemm_sm100.install()
M, V, C = 32768, 25, 64
A = np.array(spatial_adjacency(layout_for_joints(V), max_hop=5))
sp = gemm_sm100.BlockSparsity((np.abs(A).sum(0) > 0).T)
x = torch.randn(M, V * C, device=dev).bfloat16()
w = (torch.randn(V * C, V * C, device=dev) / 40).bfloat16()
dy = torch.randn(M, V * C, device=dev).bfloat16()
bias = torch.randn(V * C, device=dev)
for _ in range(2):
    st = torch.zeros(16, 2, 64, dtype=torch.float64, device=dev)
    y = gemm_sm100.gemm_pair(x, w, bias=bias, block_n=256, kb_list=sp.kb_list(256, False, dev), stats=st)   # graph conv fwd
    dx = gemm_sm100.gemm_pair(dy, w, block_n=256, kb_list=sp.kb_list(256, True, dev))                        # input gradient
    dx = gemm_sm100.gemm_pair(dy, w, block_n=256, kb_list=sp.kb_list(256, True, dev), accumulate_into=dx)    # ... added onto the residual gradient
    dw = gemm_sm100.gemm(dy, x, True, True, out_dtype=torch.float32, block_n=128, tile_mask=sp.tile_mask(128, 128, dev))
    dw = gemm_sm100.gemm_pair_dw(dy, x, sp.tile_list(256, 256, dev))                                          # the default since round 2
rows = torch.randn(M * V, C, device=dev).bfloat16().requires_grad_(True)
res = torch.randn(M * V, C, device=dev).bfloat16()
bn = nn.BatchNorm2d(C).to(dev)
conv = nn.Conv2d(C, C, (3, 1), (1, 1), (1, 0)).to(dev)
for _ in range(2):
    h = ops.batchnorm_act(rows, bn, relu=True)                       # stats + affine (streaming kernels)
    yy, s2 = ops.temporal_conv(h.reshape(32, 1024, V, C), conv.weight, conv.bias, want_stats=True)
    out = ops.batchnorm_act(yy, bn, relu=True, residual=res, sums=s2)
    out.backward(torch.randn_like(out))
    rows.grad = None
torch.cuda.synchronize()
