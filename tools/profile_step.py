"""Per-kernel time table of one train step (torch.profiler / CUPTI). Diagnostic only -- not a bench number."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root (run as `python tools/<name>.py`)

import sys

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

from pose2room_b200 import gemm_sm100, synthetic
from pose2room_b200.config import P2RConfig
from pose2room_b200.p2rnet import P2RNet

precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dev = torch.device("cuda:0")
if precision == "bf16":
    gemm_sm100.install()
torch.manual_seed(42)
np.random.seed(42)
net = P2RNet(P2RConfig(mode="train", joint_num=25, num_frames=1024, precision=precision))
net.load_state_dict(synthetic.deterministic_state_dict(net.state_dict(), seed=7))
net = net.to(dev).train()
opt = torch.optim.AdamW(net.parameters(), lr=1e-3, fused=True)
data = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in synthetic.make_batch(B, 1024, 25, seed=1234).items()}


def step():
    opt.zero_grad(set_to_none=True)
    ep = net(data)
    loss = net.loss(ep, data)["total"]
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(3):
    step()
torch.cuda.synchronize()
print("wall ms/step (no profiler): %.2f" % ((time.perf_counter() - t0) / 3 * 1e3))
t0 = time.perf_counter()
for _ in range(3):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
print("python-side ms/step (launch only): %.2f" % ((t1 - t0) / 3 * 1e3))
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    step()
    torch.cuda.synchronize()
from torch.autograd import DeviceType
rows = {}
for e in prof.events():
    if e.device_type == DeviceType.CUDA:
        k = e.name
        t, c = rows.get(k, (0.0, 0))
        rows[k] = (t + e.device_time, c + 1)
rows = [(k, t, c) for k, (t, c) in rows.items()]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print("total device time %.2f ms over %d kernel names" % (tot / 1e3, len(rows)))
for k, t, c in rows[:int(sys.argv[3]) if len(sys.argv) > 3 else 45]:
    print("%8.3f ms %5.1f%% x%-4d %s" % (t / 1e3, 100 * t / tot, c, k[:110]))
mine = sum(t for k, t, c in rows if any(s in k for s in ["gemm_bf16", "colreduce", "bn_", "affine_act", "sgemm_kernel", "smallk",
           "embed_sum", "fps_kernel", "ball_query", "group_", "maxpool", "nn_distance", "uniform_seed", "gather_points",
           "relu_bwd", "colsum_wide", "temporal_unfold", "decode_boxes", "nms3d"]))
print("hand-written kernels: %.2f ms (%.1f%%), torch/library glue: %.2f ms" % (mine / 1e3, 100 * mine / tot, (tot - mine) / 1e3))

print("---- largest torch glue ops by device time (with shapes)")
glue = [(e.key, str(e.input_shapes)[:90], e.device_time_total, e.count) for e in prof.key_averages(group_by_input_shape=True)
        if e.device_time_total > 0 and e.key.startswith("aten::")]
glue.sort(key=lambda r: -r[2])
for k, sh, t, c in glue[:28]:
    print("%8.3f ms x%-3d %-28s %s" % (t / 1e3, c, k, sh))
