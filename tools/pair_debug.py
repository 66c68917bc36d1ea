"""Isolate what bounds the CTA-pair GEMM: full kernel vs no stores / no MMAs / no bias (diagnostic)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root (run as `python tools/<name>.py`)

import json
import torch
from pose2room_b200 import gemm_sm100
from gemm_bench import timeit
dev = torch.device("cuda:0")
M, N = 32768, 1600
x = torch.randn(M, N, device=dev).bfloat16()
w = (torch.randn(N, N, device=dev) / 40).bfloat16()
bias = torch.randn(N, device=dev)
rec = {}
for bn in (256, 128):
    for name, fl in (("full", 0), ("nostore", 256), ("nomma", 512), ("nobias", 1024), ("nostore_nobias", 1280), ("nomma_nostore_nobias", 1792)):
        rec["bn%d_%s" % (bn, name)] = round(timeit(lambda: gemm_sm100.gemm_pair(x, w, bias=bias, block_n=bn, _debug_flags=fl)) * 1e3, 1)
rec["cublas"] = round(timeit(lambda: x @ w.t()) * 1e3, 1)
print(json.dumps(rec))
