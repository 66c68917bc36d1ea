"""GPU parity: knn / nn_distance / box decode / NMS / OBB IoU / parse_predictions / AP through the C ABI."""
import numpy as np
import pytest
import torch

from oracle import geometry_ref as G
from oracle.pointnet2_ref import knn_ref

pytestmark = pytest.mark.gpu


def test_knn_vs_oracle_and_reference_golden(cuda, golden_pointnet2):
    from pose2room_b200 import geometry
    g = golden_pointnet2
    x = torch.from_numpy(g["knn_x"])
    idx = geometry.knn(x.to(cuda), 8).cpu()
    assert idx.dtype == torch.int64
    assert np.array_equal(np.sort(idx.numpy(), -1), np.sort(g["knn_idx"], -1))     # reference (torch.topk) sets
    assert np.array_equal(np.sort(idx.numpy(), -1), np.sort(knn_ref(x, 8).numpy(), -1))
    off = geometry.get_graph_offset(x.to(cuda), k=8, idx=torch.from_numpy(g["knn_idx"]).to(cuda))
    assert np.array_equal(off.cpu().numpy(), g["knn_offset"])
    # live configuration of the backbone: (B,3,T) hip trajectory, k = 20, T = 1024
    rng = np.random.default_rng(0)
    xt = torch.from_numpy(np.cumsum(rng.normal(0, 0.05, size=(4, 3, 1024)), axis=2).astype(np.float32))
    got = geometry.knn(xt.to(cuda), 20).cpu().numpy()
    want = knn_ref(xt, 20).numpy()
    assert (np.sort(got, -1) == np.sort(want, -1)).mean() > 0.999   # fp32 near-ties at the k-th neighbour
    assert (got[..., 0] == np.arange(1024)[None]).all()


def test_graph_offset_grad(cuda):
    from pose2room_b200 import geometry
    x = torch.randn(2, 6, 50, device=cuda, requires_grad=True)
    idx = torch.randint(0, 50, (2, 50, 5), device=cuda)
    out = geometry.get_graph_offset(x, idx=idx)
    w = torch.randn_like(out)
    (out * w).sum().backward()
    x2 = x.detach().clone().requires_grad_(True)
    xt = x2.transpose(1, 2)                                                     # (B,N,6)
    ref = (torch.gather(xt[:, None].expand(-1, 50, -1, -1), 2, idx[..., None].expand(-1, -1, -1, 6)) - xt[:, :, None])
    (ref.reshape(2, 50, 5, 2, 3) * w).sum().backward()
    assert torch.allclose(out, ref.reshape(2, 50, 5, 2, 3))
    assert torch.allclose(x.grad, x2.grad, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("case", ["demo", "a", "b"])
@pytest.mark.parametrize("mode,kw", [("l2", {}), ("l1s", dict(l1smooth=True)), ("l1", dict(l1=True))])
def test_nn_distance_bit_exact_vs_reference_goldens(cuda, golden_geometry, case, mode, kw):
    from pose2room_b200 import geometry
    g = golden_geometry
    r = geometry.nn_distance(torch.from_numpy(g["nnd_%s_pc1" % case]).to(cuda),
                             torch.from_numpy(g["nnd_%s_pc2" % case]).to(cuda), **kw)
    for key, got in zip(["d1", "i1", "d2", "i2"], r):
        assert np.array_equal(got.cpu().numpy(), g["nnd_%s_%s_%s" % (case, mode, key)]), key


def test_nn_distance_live_shapes_and_grad(cuda):
    from pose2room_b200 import geometry
    gen = torch.Generator().manual_seed(0)
    # the three live call sites of models/loss.py:64,105,128 at BASELINE shapes
    # (+ a per-GPU batch of 160: B * num_seeds = 81 920 clouds, beyond the 65 535 limit of grid.y the batch used to ride on)
    for (B, N, M) in [(32, 128, 10), (32 * 512, 3, 25), (1, 128, 7), (160 * 512, 3, 25)]:
        a = torch.randn(B, N, 3, generator=gen)
        b = torch.randn(B, M, 3, generator=gen)
        want = G.nn_distance(a.numpy(), b.numpy())
        got = geometry.nn_distance(a.to(cuda), b.to(cuda))
        for x, y in zip(got, want):
            assert np.array_equal(x.cpu().numpy(), y)
    for kw in [{}, dict(l1smooth=True, delta=0.7), dict(l1=True)]:
        a = torch.randn(3, 40, 3, generator=gen).to(cuda).requires_grad_(True)
        b = torch.randn(3, 9, 3, generator=gen).to(cuda).requires_grad_(True)
        d1, _, d2, _ = geometry.nn_distance(a, b, **kw)
        w1, w2 = torch.randn_like(d1), torch.randn_like(d2)
        ((d1 * w1).sum() + (d2 * w2).sum()).backward()
        a2, b2 = a.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
        diff = a2[:, :, None] - b2[:, None]
        if kw.get("l1smooth"):
            ab = diff.abs()
            q = torch.clamp(ab, max=0.7)
            e = 0.5 * q ** 2 + 0.7 * (ab - q)
        elif kw.get("l1"):
            e = diff.abs()
        else:
            e = diff ** 2
        dist = e.sum(-1)
        ((dist.min(2).values * w1).sum() + (dist.min(1).values * w2).sum()).backward()
        assert torch.allclose(a.grad, a2.grad, rtol=1e-5, atol=1e-6)
        assert torch.allclose(b.grad, b2.grad, rtol=1e-5, atol=1e-6)


def test_decode_boxes_and_iou_vs_reference_goldens(cuda, golden_geometry):
    from pose2room_b200 import geometry
    g = golden_geometry
    th = torch.from_numpy(g["box_theta"])
    heading = torch.stack([torch.sin(th), torch.cos(th)], -1)[None]
    center = torch.from_numpy(g["box_center"])[None]
    lsize = torch.log(torch.from_numpy(g["box_size"]))[None]
    corners, aabb, _ = geometry.decode_boxes(center.to(cuda), lsize.to(cuda), heading.to(cuda),
                                             torch.zeros(1, 4, 3, device=cuda))
    # exp(log(s)) in fp32 is within an ulp or two of s; corners follow
    assert np.allclose(corners[0].cpu().numpy(), g["box_corners"], rtol=0, atol=2e-6)
    assert np.allclose(aabb[0, :, :3].cpu().numpy(), g["box_corners"].min(1), atol=2e-6)
    assert np.allclose(aabb[0, :, 3:].cpu().numpy(), g["box_corners"].max(1), atol=2e-6)
    i3, i2 = geometry.box3d_iou_matrix(g["box_corners"], g["box_corners"])
    ok = ~np.isnan(g["box_iou3d"]) & ~np.eye(24, dtype=bool)
    assert np.allclose(i3.cpu().numpy()[ok], g["box_iou3d"][ok], atol=1e-9)
    assert np.allclose(i2.cpu().numpy()[ok], g["box_iou2d"][ok], atol=1e-9)
    # Appendix D known answers (generated by the reference)
    size = np.array([1.0, 2.0, 0.5])
    A = G.get_3d_box(size, 0.3, np.array([0.0, 1.0, 0.0]))
    Bx = G.get_3d_box(size, 0.3, np.array([0.2, 1.0, 0.1]))
    C = G.get_3d_box(size, -1.2, np.array([0.1, 1.5, 0.0]))
    assert np.allclose(geometry.box3d_iou(A, Bx), (0.40762005092979847, 0.4076200509297984), atol=1e-12)
    assert np.allclose(geometry.box3d_iou(A, C), (0.23148291421667108, 0.33445040365524775), atol=1e-12)


@pytest.mark.parametrize("t", range(4))
def test_nms_selection_exact_vs_reference_goldens(cuda, golden_geometry, t):
    from pose2room_b200 import geometry
    g = golden_geometry
    boxes = g["nms%d_boxes" % t]
    assert geometry.nms_3d_faster(boxes[:, :7], 0.10) == g["nms%d_pick" % t].tolist()
    assert geometry.nms_3d_faster(boxes[:, :7], 0.25, old_type=True) == g["nms%d_pick_old" % t].tolist()
    assert geometry.nms_3d_faster_samecls(boxes, 0.10) == g["nms%d_pick_cls" % t].tolist()
    assert geometry.nms_2d_faster(boxes[:, [0, 1, 3, 4, 6]], 0.10) == g["nms%d_pick_2d" % t].tolist()


def test_nms_random_vs_oracle_many(cuda):
    from pose2room_b200 import geometry
    rng = np.random.default_rng(5)
    for k in [1, 2, 31, 128, 200, 1000]:
        lo = rng.normal(0, 1.0, size=(k, 3))
        boxes = np.concatenate([lo, lo + rng.uniform(0.05, 1.5, size=(k, 3)), rng.uniform(size=(k, 1))], 1)
        assert geometry.nms_3d_faster(boxes, 0.1) == G.nms_3d_faster(boxes, 0.1)


class _DC:
    origin_joint_id = 0
    contact_dist_thresh = 1.0
    num_class = 22


CFG = dict(dataset_config=_DC, remove_far_box=True, use_3d_nms=True, cls_nms=False, nms_iou=0.10,
           use_old_type_nms=False, per_class_proposal=True, conf_thresh=0.05, sample_cls=False)


def test_parse_predictions_and_map_vs_reference_goldens(cuda, golden_geometry):
    """ap_helper API end to end on the GPU against what the reference produced for the same inputs:
    pred_mask exact, corners to fp32-exp rounding, per-class AP and mAP@0.25/0.5."""
    from pose2room_b200 import ap_helper
    g = golden_geometry
    est = dict(center=torch.from_numpy(g["pp_center"]).to(cuda), size=torch.from_numpy(g["pp_size"]).to(cuda),
               heading=torch.from_numpy(g["pp_heading"]).to(cuda),
               objectness_scores=torch.from_numpy(g["pp_objectness"]).to(cuda),
               sem_cls_scores=torch.from_numpy(g["pp_sem"]).to(cuda))
    data = dict(input_joints=torch.from_numpy(g["pp_input_joints"]).to(cuda))
    eval_dict, parsed = ap_helper.parse_predictions(est, data, CFG)
    assert np.array_equal(eval_dict["pred_mask"], g["pp_pred_mask"])
    assert np.allclose(parsed["pred_corners_3d"], g["pp_corners"], atol=2e-6)
    assert np.allclose(parsed["obj_prob"], g["pp_obj_prob"], atol=1e-6)
    assert np.array_equal(parsed["pred_sem_cls"], g["pp_pred_sem_cls"])
    eval_dict = ap_helper.assembly_pred_map_cls(eval_dict, parsed, CFG)
    gt = {k: torch.from_numpy(g["pp_gt_" + k]) for k in ["box_label_mask", "sem_cls_label", "center_label", "size", "heading"]}
    gt_map = ap_helper.assembly_gt_map_cls(ap_helper.parse_groundtruths(gt, CFG))
    for thr in [0.25, 0.5]:
        calc = ap_helper.APCalculator(thr)
        calc.step(eval_dict["batch_pred_map_cls"], gt_map)
        m = calc.compute_metrics()
        want = g["ap_%d" % int(thr * 100)]
        for c in range(22):
            key = "%d Average Precision" % c
            if np.isnan(want[c]):
                assert key not in m or np.isnan(m[key])
            else:
                assert abs(m[key] - want[c]) < 1e-6, (thr, c, m[key], want[c])
        assert abs(m["mAP"] - np.nanmean(want)) < 1e-6


def test_graph_features_and_fp_module_vs_reference_goldens(cuda, golden_pointnet2):
    """get_graph_feature / get_graph_feature_cross (VN-DGCNN edge features) and PointnetFPModule against outputs of the
    reference's own code (dead in the live P2RNet path, named by north_star)."""
    import torch.nn as nn
    from pose2room_b200 import geometry
    from pose2room_b200.pointnet2_modules import PointnetFPModule
    g = golden_pointnet2
    x = torch.from_numpy(g["gf_x"]).to(cuda)
    idx = torch.from_numpy(g["gf_idx"]).to(cuda)
    assert np.array_equal(geometry.get_graph_feature(x, k=8, idx=idx).cpu().numpy(), g["gf_feature"])
    assert np.allclose(geometry.get_graph_feature_cross(x, k=8, idx=idx).cpu().numpy(), g["gf_cross"], atol=1e-6)
    C = g["feats"].shape[1]
    fp = PointnetFPModule(mlp=[C + C, 16, 8], bn=False).to(cuda)
    with torch.no_grad():
        fp.mlp[0].weight.copy_(torch.from_numpy(g["fp_w0"]))
        fp.mlp[0].bias.copy_(torch.from_numpy(g["fp_b0"]))
        fp.mlp[2].weight.copy_(torch.from_numpy(g["fp_w1"]))
        fp.mlp[2].bias.copy_(torch.from_numpy(g["fp_b1"]))
    out = fp(torch.from_numpy(g["xyz"]).to(cuda), torch.from_numpy(g["new_xyz"]).to(cuda),
             torch.from_numpy(g["feats"]).to(cuda), torch.from_numpy(g["ti_kfeat"]).to(cuda))
    assert np.allclose(out.detach().cpu().numpy(), g["fp_out"], rtol=1e-4, atol=1e-5)
