"""Launches every kernel of the path that is NOT a GEMM or a streaming BatchNorm pass at its BASELINE shape, twice each, for
one `ncu --set full` pass: the nine pointnet2 operators (live shape of ProposalNet: 32 x 512 votes -> 128 proposals, r 0.3,
16 samples, 256 channels), their channel-last twins, nn_distance at the three call sites of the loss, the seed sampling,
knn / graph offset, the eval kernels (decode_boxes, nms3d, box3d_iou on 125 scenes), make_batch (both variants), the fused
detection loss / mixture heads / vote tail, embed_sum.  Diagnostic only.

    ncu --set full --clock-control none --import-source on -k regex:"$(python tools/ncu_small_kernels.py --regex)" \
        -o gpurun_out/small_kernels python tools/ncu_small_kernels.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root (run as `python tools/<name>.py`)

REGEX = ("fps_kernel|ball_query|group_points|group_rows|gather_points|three_nn|three_interpolate|nn_distance|uniform_seed|"
         "knn_kernel|graph_offset|decode_boxes|nms3d|box3d_iou|make_batch|detection_loss|gmm_mix|vote_tail|maxpool_rows|"
         "embed_sum|smallk|sa_fused|embed_l1|coord_moments|select_rows|colsum_any|stream_colsum_period|gcn_dA|gcn_dw_kernel|"
         "gcn_build_weight")
if "--regex" in sys.argv:
    print(REGEX)
    sys.exit(0)

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

# ---- the nine pointnet2 operators + channel-last twins at the live shape of ProposalNet ----------------------------
from pose2room_b200 import ap_helper, ext, geometry, ops, synthetic
xyz = torch.from_numpy(synthetic.make_cloud(B, 512, seed=1)).to(dev)
feats = torch.randn(B, 256, 512, device=dev)
rows = torch.randn(B, 512, 256, device=dev).bfloat16().requires_grad_(True)
for _ in range(2):
    idx = ext.furthest_point_sampling(xyz, 128)
    new_xyz = ext.gather_points(xyz.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
    ext.gather_points_grad(torch.randn(B, 3, 128, device=dev), idx, 512)
    bq = ext.ball_query(new_xyz, xyz, 0.3, 16)
    g = ext.group_points(feats, bq)
    ext.group_points_grad(torch.randn_like(g), bq, 512)
    d2, i3 = ext.three_nn(xyz, new_xyz)
    w = torch.softmax(-d2, dim=-1).contiguous()
    up = ext.three_interpolate(torch.randn(B, 256, 128, device=dev), i3, w)
    ext.three_interpolate_grad(torch.randn_like(up), i3, w, 128)
    grouped = ops.group_rows(rows, bq)
    pooled = ops.maxpool_rows(grouped.reshape(B * 128, 16, 256))
    pooled.float().sum().backward()
torch.cuda.synchronize()

# ---- the fused set-abstraction kernel (gather -> 256-256-256 MLP -> max over 16) at the live shape -------------------
from pose2room_b200 import gemm_sm100
gemm_sm100.install()
conv1, conv2 = torch.nn.Conv2d(256, 256, 1).to(dev), torch.nn.Conv2d(256, 256, 1).to(dev)
with torch.no_grad():
    for _ in range(2):
        ops.sa_fused(rows.detach(), bq, conv1, conv2)
for _ in range(2):
    ops.sa_fused(rows.detach().requires_grad_(True), bq, conv1, conv2).float().sum().backward()
gemm_sm100.uninstall()
torch.cuda.synchronize()

# ---- nn_distance at the three call sites of the loss (loss.py:64,105,128), seed sampling, knn, graph offset ----------
joints = synthetic.make_batch(B, T, J, seed=1234)["input_joints"].to(dev)
for _ in range(2):
    geometry.nn_distance(torch.randn(B, 128, 3, device=dev), torch.randn(B, 10, 3, device=dev))
    geometry.nn_distance(torch.randn(B * 512, 3, 3, device=dev), torch.randn(B * 512, J, 3, device=dev))
    hip = joints[:, :, 0]
    out = torch.empty(B, 512, dtype=torch.int64, device=dev)
    _lib.call("p2r_uniform_seed_inds", hip.data_ptr(), int(hip.stride(1)), B, T, 512, out.data_ptr(),
              torch.cuda.current_stream().cuda_stream)
    hip_cf = hip.transpose(1, 2).contiguous()
    nbr = geometry.knn(hip_cf, 20)
    geometry.get_graph_offset(hip_cf, k=20, idx=nbr)
    sk = torch.randn(B * T, J, 64, device=dev).bfloat16()
    ops.embed_sum(sk, torch.randn(B * T, 20, 64, device=dev).bfloat16())
torch.cuda.synchronize()

# ---- round-2 additions (first embed layer, seed-frame pick, odd-width / periodic column sums, graph-conv weight build and
# ---- gradient fold): tools/ncu_new_kernels.py, also runnable on its own
import ncu_new_kernels  # noqa: E402,F401

# ---- eval kernels on 125 scenes of the 1000-scene set (BASELINE config #5) --------------------------------------------
cfg = P2RConfig(mode="test", joint_num=J).eval_config
est_e, gt_e = synthetic.make_eval_batch(20240, 0, 125)
est_e = {k: v.to(dev) for k, v in est_e.items()}
for _ in range(2):
    eval_dict, parsed = ap_helper.parse_predictions(est_e, {"input_joints": gt_e["input_joints"].to(dev)}, cfg)
    eval_dict = ap_helper.assembly_pred_map_cls(eval_dict, parsed, cfg)
    gts = ap_helper.assembly_gt_map_cls(ap_helper.parse_groundtruths(gt_e, cfg))
    calc = ap_helper.APCalculator(0.25)
    calc.step(eval_dict["batch_pred_map_cls"], gts)
    calc.compute_metrics()
torch.cuda.synchronize()
print("done")
