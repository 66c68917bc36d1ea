"""BN kernel microbench on the block-activation shape [819200, 64] bf16 (diagnostic)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root (run as `python tools/<name>.py`)

import os, sys, json
import torch
import torch.nn as nn
from pose2room_b200 import ops, _lib

dev = torch.device("cuda:0")
M, C = 819200, 64
x = torch.randn(M, C, device=dev).bfloat16()
res = torch.randn(M, C, device=dev).bfloat16()
dy = torch.randn(M, C, device=dev).bfloat16()
bn = nn.BatchNorm1d(C).to(dev)


def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / n * 1e3

DT = 1
s = torch.zeros(2, C, dtype=torch.float64, device=dev)
stats = torch.rand(4, C, device=dev) + 0.5
y = torch.empty_like(x); dx = torch.empty_like(x); dres = torch.empty_like(x)
st = torch.cuda.current_stream().cuda_stream
out = {"ctas_per_sm": os.environ.get("P2R_COLREDUCE_CTAS_PER_SM", "8")}
out["col_stats_us"] = t(lambda: _lib.call("p2r_col_stats", x.data_ptr(), DT, M, C, s[0].data_ptr(), s[1].data_ptr(), st))
out["col_bwd_stats_relu2_us"] = t(lambda: _lib.call("p2r_col_bwd_stats", dy.data_ptr(), x.data_ptr(), None, DT, M, C, stats[0].data_ptr(), stats[1].data_ptr(), 2, s[0].data_ptr(), s[1].data_ptr(), stats[2].data_ptr(), stats[3].data_ptr(), st))
out["col_bwd_stats_relu1_us"] = t(lambda: _lib.call("p2r_col_bwd_stats", dy.data_ptr(), x.data_ptr(), res.data_ptr(), DT, M, C, stats[0].data_ptr(), stats[1].data_ptr(), 1, s[0].data_ptr(), s[1].data_ptr(), stats[2].data_ptr(), stats[3].data_ptr(), st))
out["affine_us"] = t(lambda: _lib.call("p2r_affine_act", x.data_ptr(), DT, M, C, stats[2].data_ptr(), stats[3].data_ptr(), None, 1, y.data_ptr(), None, st))
out["affine_res_us"] = t(lambda: _lib.call("p2r_affine_act", x.data_ptr(), DT, M, C, stats[2].data_ptr(), stats[3].data_ptr(), res.data_ptr(), 1, y.data_ptr(), None, st))
out["bwd_apply_relu2_us"] = t(lambda: _lib.call("p2r_bn_bwd_apply", dy.data_ptr(), x.data_ptr(), None, DT, M, C, stats[0].data_ptr(), stats[1].data_ptr(), stats[2].data_ptr(), s[0].data_ptr(), s[1].data_ptr(), 2, dx.data_ptr(), None, stats[3].data_ptr(), None, 0, st))
out["bwd_apply_relu1_res_us"] = t(lambda: _lib.call("p2r_bn_bwd_apply", dy.data_ptr(), x.data_ptr(), res.data_ptr(), DT, M, C, stats[0].data_ptr(), stats[1].data_ptr(), stats[2].data_ptr(), s[0].data_ptr(), s[1].data_ptr(), 1, dx.data_ptr(), dres.data_ptr(), stats[3].data_ptr(), None, 0, st))
out["ideal_us_per_105MB_pass"] = 104.9e6 / 6548e9 * 1e6
print(json.dumps(out))
