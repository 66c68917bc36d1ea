"""Launches the kernels that have no ncu capture yet at their BASELINE shapes, twice each (for one `ncu --set full` pass):
make_batch (both data-movement variants), detection_loss (+grad), gmm_mix (+grad), vote_tail (+grad).  Diagnostic only.

    ncu --set full --clock-control none --import-source on -k regex:"make_batch|detection_loss|gmm_mix|vote_tail" \
        -o gpurun_out/new_kernels python tools/ncu_new_kernels.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root (run as `python tools/<name>.py`)

import torch

import bench
from pose2room_b200 import _lib, dataloader as DL
from pose2room_b200.config import P2RConfig
from pose2room_b200.p2rnet.loss import BoxNetDetectionLoss
from pose2room_b200.p2rnet.mdn import MixtureDensityHead, Struct, _FusedGMMPredict
from tests.test_loss_math import make_case

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
_lib.load()
B, T, J = bench.B_PER_GPU, bench.T_FRAMES, bench.JOINTS

# ---- sample -> batch, both variants ----------------------------------------------------------------------------
store, ds = bench._data_path_store(dev, B)
jd, vd, fsd = store.device_arrays(dev)
ids = torch.arange(B, dtype=torch.int32, device=dev)
params = torch.from_numpy(ds.host_side(list(range(B)), [DL.draw_augmentation() for _ in range(B)])[0]).to(dev)
outs = (torch.empty(B, T, J, 3, device=dev), torch.empty(B, T, J, 9, device=dev),
        torch.empty(B, T, J, dtype=torch.int64, device=dev))
for variant in (1, 2, 1, 2):
    _lib.call("p2r_make_batch_variant", variant, jd.data_ptr(), vd.data_ptr(), fsd.data_ptr(), ids.data_ptr(),
              params.data_ptr(), B, T, J, 3, outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(),
              torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()

# ---- detection loss at B=32, S=512, P=128 ------------------------------------------------------------------------
os.environ["P2R_FUSED_LOSS"] = "1"
crit = BoxNetDetectionLoss(1, 0, P2RConfig(mode="train", joint_num=J))
est, gt, sem_obj = make_case(1, B=B, T=T, J=J, S=512, P=128)
for _ in range(2):
    so = sem_obj.to(dev).requires_grad_(True)
    e = {k: (v.to(dev).requires_grad_(v.is_floating_point()) if isinstance(v, torch.Tensor) else v) for k, v in est.items()}
    e["objectness_scores"], e["sem_cls_scores"] = so[..., 0:2], so[..., 2:]
    g = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in gt.items()}
    crit(e, g, None)["total"].backward()
torch.cuda.synchronize()

# ---- mixture heads: 4096 rows, 100 components --------------------------------------------------------------------
for d, dt in ((3, torch.float32), (2, torch.float64)):
    head = MixtureDensityHead(Struct(input_dim=8, num_gaussian=100, out_dim=d, n_samples=1, central_tendency="mean",
                                     mu_bias_init=torch.randn(100, d).to(dt))).to(dev)
    for _ in range(2):
        logits = torch.randn(B * 128, 100, device=dev).bfloat16().requires_grad_(True)
        eps = head.mu.data.new(B * 128, 100, 1, d).normal_()
        _FusedGMMPredict.apply(logits, head.mu, head.log_sigma, eps).sum().backward()
torch.cuda.synchronize()

# ---- vote tail: 16 384 seed rows, 256 channels ---------------------------------------------------------------------
from pose2room_b200.p2rnet.vote_center import _VoteTail
for dt in (torch.bfloat16, torch.float32):
    for _ in range(2):
        net = torch.randn(B * 512, 259, device=dev).to(dt).requires_grad_(True)
        skel = torch.randn(B, 512, J, 3, device=dev)
        sf = torch.randn(B, 512, 256, device=dev, requires_grad=True)
        xyz, feat = _VoteTail.apply(net, skel[:, :, 0], sf)
        (xyz.sum() + feat.sum()).backward()
torch.cuda.synchronize()
print("done")
