"""The kernels added in the second half of round 2 at their BASELINE shapes, twice each, for one `ncu --set full` pass
(imported by tools/ncu_small_kernels.py; runnable on its own):

    ncu --set full --clock-control none -k regex:"embed_l1|coord_moments|select_rows|colsum_any|stream_colsum_period|gcn_" \
        -o gpurun_out/new_kernels python tools/ncu_new_kernels.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.nn as nn

import bench
from pose2room_b200 import _lib, gemm_sm100, ops
from pose2room_b200.p2rnet.graph import layout_for_joints, spatial_adjacency

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
_lib.load()
B, T, J = bench.B_PER_GPU, bench.T_FRAMES, bench.JOINTS

# ---- round-2 additions: first embed layer on fp32 coordinates, seed-frame pick, odd-width / periodic column sums,
# ---- the graph-conv weight build and its gradient fold ----------------------------------------------------------------
xyz = torch.randn(B * T * J, 3, device=dev)
w3 = (torch.randn(64, 3, device=dev) / 3 ** 0.5).requires_grad_(True)
bn_e = nn.BatchNorm1d(64).to(dev).train()
frames = torch.randn(B, T, J * 64, device=dev).bfloat16().requires_grad_(True)
picks = torch.sort(torch.randint(0, T, (B, 512), device=dev), dim=1)[0]
odd = torch.randn(16384, 259, device=dev).bfloat16()
wide = torch.randn(B * T, J * 64, device=dev).bfloat16()
A_e = torch.tensor(spatial_adjacency(layout_for_joints(J), max_hop=5), dtype=torch.float32, device=dev)
cw = torch.randn(A_e.shape[0] * 64, 64, device=dev) / 8
cb = torch.randn(A_e.shape[0] * 64, device=dev)
for _ in range(2):
    y_e = ops.embed_l1(xyz, w3, bn_e)
    y_e.backward(torch.randn_like(y_e))
    sel = ops.select_rows(frames, picks)
    sel.backward(torch.randn_like(sel))
    ops._col_sum(odd)
    ops._col_sum(wide)
    built = ops._gcn_build(cw, cb, A_e)
    dwe = torch.randn(J * 64, J * 64, device=dev)
    dbe = torch.randn(J * 64, device=dev)
    d_w, d_b, d_a = torch.empty_like(cw), torch.empty_like(cb), torch.empty_like(A_e)
    _lib.call("p2r_gcn_reduce_weight_grad", dwe.data_ptr(), dbe.data_ptr(), cw.data_ptr(), cb.data_ptr(), A_e.data_ptr(),
              A_e.shape[0], J, 64, 64, d_w.data_ptr(), d_b.data_ptr(), d_a.data_ptr(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()

# ---- AdamW over the 131 parameter tensors of P2RNet (shapes from the model itself) ----------------------------------
from pose2room_b200.config import P2RConfig
from pose2room_b200.optim import AdamW
from pose2room_b200.p2rnet import P2RNet

net = P2RNet(P2RConfig(mode="train", joint_num=J, num_frames=T, precision="bf16")).to(dev)
params = [p for p in net.parameters() if p.requires_grad]
for p in params:
    p.grad = torch.randn_like(p)
opt = AdamW(params, lr=1e-3)
for _ in range(2):
    opt.step()
torch.cuda.synchronize()
print("new kernels done")
