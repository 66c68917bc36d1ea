"""Launches the graph-conv GEMM variants a few times each for an `ncu --set full` capture (diagnostic)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root (run as `python tools/<name>.py`)

import numpy as np
import torch

from pose2room_b200 import gemm_sm100
from pose2room_b200.p2rnet.graph import layout_for_joints, spatial_adjacency

dev = torch.device("cuda:0")
M, N = 32768, 1600
A = np.array(spatial_adjacency(layout_for_joints(25), max_hop=5))
sp = gemm_sm100.BlockSparsity((np.abs(A).sum(0) > 0).T)
x = torch.randn(M, N, device=dev).bfloat16()
w = (torch.randn(N, N, device=dev) / 40).bfloat16()
bias = torch.randn(N, device=dev)
for _ in range(3):
    gemm_sm100.gemm_pair(x, w, bias=bias, block_n=256)
    gemm_sm100.gemm_pair(x, w, bias=bias, block_n=256, kb_list=sp.kb_list(256, False, dev))
    gemm_sm100.gemm(x, w, bias=bias, block_n=128, kb_list=sp.kb_list(128, False, dev))
    y = x @ w.t()
torch.cuda.synchronize()
