"""Generate tests/golden/*.npz by EXECUTING the unmodified reference (from /root/reference) on CPU.

Run in the build container only (the reference tree does not exist on the GPU box):

    python tests/golden/make_golden.py [geometry] [pointnet2] [model]

The native ops have no CPU path in the reference, so the reference's Python wrappers / modules
(pointnet2_utils.py, pointnet2_modules.py, ProposalNet) run on top of oracle.pointnet2_ref.RefExt;
everything else (nn_distance, NMS, box IoU, parse_predictions with scipy Delaunay, AP, the whole
P2RNet forward / loss / generate) is the reference's own code and arithmetic.
"""
import os
import os.path as osp
import sys

import numpy as np
import torch

ROOT = osp.dirname(osp.dirname(osp.dirname(osp.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = osp.dirname(osp.abspath(__file__))

from oracle import pointnet2_ref, ref_import  # noqa: E402
from pose2room_b200 import synthetic  # noqa: E402


def random_boxes(rng, n, spread=1.5):
    center = rng.normal(0, spread, size=(n, 3))
    size = rng.uniform(0.2, 1.7, size=(n, 3))
    theta = rng.uniform(-np.pi, np.pi, size=n)
    return center, size, theta


def geometry(ns):
    rng = np.random.default_rng(7)
    out = {}
    # --- nn_distance: the reference's own known-answer case + random cases in 3 modes
    np.random.seed(0)
    pc1 = np.random.random((1, 5, 3)).astype(np.float32)
    pc2 = np.random.random((1, 6, 3)).astype(np.float32)
    cases = [("demo", pc1, pc2)]
    cases.append(("a", rng.normal(size=(3, 128, 3)).astype(np.float32), rng.normal(size=(3, 10, 3)).astype(np.float32)))
    cases.append(("b", rng.normal(size=(64, 3, 3)).astype(np.float32), rng.normal(size=(64, 25, 3)).astype(np.float32)))
    for name, a, b in cases:
        out["nnd_%s_pc1" % name], out["nnd_%s_pc2" % name] = a, b
        for mode, kw in [("l2", {}), ("l1s", dict(l1smooth=True)), ("l1", dict(l1=True))]:
            r = ns.nn_distance.nn_distance(torch.from_numpy(a), torch.from_numpy(b), **kw)
            for key, t in zip(["d1", "i1", "d2", "i2"], r):
                out["nnd_%s_%s_%s" % (name, mode, key)] = t.numpy()
    # --- boxes: corners + pairwise IoU
    c, s, th = random_boxes(rng, 24, spread=0.6)
    corners = np.stack([ns.pc_utils.get_3d_box(s[i].astype(np.float32), float(th[i]), c[i].astype(np.float32))
                        for i in range(24)])
    out["box_center"], out["box_size"], out["box_theta"], out["box_corners"] = \
        c.astype(np.float32), s.astype(np.float32), th, corners
    iou3 = np.zeros((24, 24))
    iou2 = np.zeros((24, 24))
    with np.errstate(all="ignore"):
        for i in range(24):
            for j in range(24):
                try:  # the reference divides by zero on (near-)parallel coincident edges, e.g. i == j
                    iou3[i, j], iou2[i, j] = ns.box_util.box3d_iou(corners[i], corners[j])
                except Exception:
                    iou3[i, j] = iou2[i, j] = np.nan
    out["box_iou3d"], out["box_iou2d"] = iou3, iou2
    # --- NMS
    for t in range(4):
        k = [128, 128, 37, 5][t]
        lo = rng.normal(0, 1.0, size=(k, 3))
        hi = lo + rng.uniform(0.2, 1.5, size=(k, 3))
        score = rng.permutation(k) / k + rng.uniform(0, 1e-3, size=k)
        cls = rng.integers(0, 4, size=k).astype(np.float64)
        boxes = np.concatenate([lo, hi, score[:, None], cls[:, None]], axis=1)
        out["nms%d_boxes" % t] = boxes
        out["nms%d_pick" % t] = np.array(ns.nms.nms_3d_faster(boxes[:, :7], 0.10), np.int64)
        out["nms%d_pick_old" % t] = np.array(ns.nms.nms_3d_faster(boxes[:, :7], 0.25, old_type=True), np.int64)
        out["nms%d_pick_cls" % t] = np.array(ns.nms.nms_3d_faster_samecls(boxes, 0.10), np.int64)
        out["nms%d_pick_2d" % t] = np.array(ns.nms.nms_2d_faster(boxes[:, [0, 1, 3, 4, 6]], 0.10), np.int64)
    # --- parse_predictions (reference: scipy Delaunay far-box test + numpy NMS) on synthetic network outputs
    B, K, T, J = 4, 128, 256, 25
    data = synthetic.make_batch(B, T, J, seed=99)
    hip = data["input_joints"][:, :, 0].numpy()
    center = np.zeros((B, K, 3), np.float32)
    lsize = np.zeros((B, K, 3), np.float32)
    heading = np.zeros((B, K, 2), np.float64)
    for b in range(B):
        fr = rng.integers(0, T, size=K)
        center[b] = hip[b, fr] + rng.normal(0, 0.8, size=(K, 3))
        lsize[b] = np.log(rng.uniform(0.1, 2.0, size=(K, 3)))
        thb = rng.uniform(-np.pi, np.pi, size=K)
        heading[b, :, 0], heading[b, :, 1] = np.sin(thb) * 0.9, np.cos(thb) * 0.9
    center[:, :8] += 30.0  # far away: must be dropped by remove_far_box
    lsize[:, 8:12, 0] = np.log(0.005)  # degenerate size: dropped
    lsize[:, 12:14, 1] = np.log(12.0)
    objectness = rng.normal(size=(B, K, 2)).astype(np.float32)
    sem = rng.normal(size=(B, K, 22)).astype(np.float32)
    end_points = dict(center=torch.from_numpy(center), size=torch.from_numpy(lsize),
                      heading=torch.from_numpy(heading), objectness_scores=torch.from_numpy(objectness),
                      sem_cls_scores=torch.from_numpy(sem))
    _, cfg = ref_import.build_reference_model(mode="test", joint_num=25, num_frames=T)
    eval_dict, parsed = ns.ap_helper.parse_predictions(end_points, {"input_joints": data["input_joints"]},
                                                       cfg.eval_config)
    out.update(pp_input_joints=data["input_joints"].numpy(), pp_center=center, pp_size=lsize, pp_heading=heading,
               pp_objectness=objectness, pp_sem=sem, pp_pred_mask=eval_dict["pred_mask"],
               pp_corners=parsed["pred_corners_3d"], pp_obj_prob=parsed["obj_prob"],
               pp_sem_probs=parsed["sem_cls_probs"], pp_pred_sem_cls=parsed["pred_sem_cls"])
    # --- AP: reference APCalculator on the parsed predictions against the synthetic GT
    eval_dict = ns.ap_helper.assembly_pred_map_cls(eval_dict, parsed, cfg.eval_config)
    gts = ns.ap_helper.parse_groundtruths(data, cfg.eval_config)
    gt_map = ns.ap_helper.assembly_gt_map_cls(gts)
    for thr in [0.25, 0.5]:
        pred_all = {i: eval_dict["batch_pred_map_cls"][i] for i in range(B)}
        gt_all = {i: gt_map[i] for i in range(B)}
        # eval_det_multiprocessing_wo_mesh uses Pool(10); call the per-class routine directly, same maths
        pred, gt = {}, {}
        for img in pred_all:
            for cls, box, score in pred_all[img]:
                pred.setdefault(cls, {}).setdefault(img, []).append((box, score))
                gt.setdefault(cls, {}).setdefault(img, [])
        for img in gt_all:
            for cls, box in gt_all[img]:
                gt.setdefault(cls, {}).setdefault(img, []).append(box)
        aps = np.full(22, np.nan)
        with np.errstate(all="ignore"):
            for cls in gt:
                if cls in pred:
                    aps[cls] = ns.eval_det.eval_det_cls_wo_mesh(pred[cls], gt[cls], thr,
                                                                get_iou_func=ns.eval_det.get_iou_obb)[2]
                else:
                    aps[cls] = 0
        out["ap_%d" % int(thr * 100)] = aps
    for k in ["box_label_mask", "sem_cls_label", "center_label", "size", "heading"]:
        out["pp_gt_" + k] = data[k].numpy()
    np.savez_compressed(osp.join(OUT, "geometry.npz"), **out)
    print("geometry.npz:", len(out), "arrays")


def pointnet2(ns):
    """Reference Python wrappers/modules (pointnet2_utils.py, pointnet2_modules.py) over RefExt."""
    rng = np.random.default_rng(11)
    out = {}
    B, N, C, P, S = 2, 96, 8, 16, 8
    xyz = torch.from_numpy(synthetic.make_cloud(B, N, seed=3))
    feats = torch.from_numpy(rng.normal(size=(B, C, N)).astype(np.float32)).requires_grad_(True)
    pu = ns.pointnet2_utils
    inds = pu.furthest_point_sample(xyz, P)
    new_xyz = pu.gather_operation(xyz.transpose(1, 2).contiguous(), inds).transpose(1, 2).contiguous()
    grouper = pu.QueryAndGroup(0.4, S, use_xyz=True, ret_grouped_xyz=True, normalize_xyz=True)
    new_feats, gxyz = grouper(xyz, new_xyz, feats)
    w = torch.from_numpy(rng.normal(size=tuple(new_feats.shape)).astype(np.float32))
    (new_feats * w).sum().backward()
    out.update(xyz=xyz.numpy(), feats=feats.detach().numpy(), fps=inds.numpy(), new_xyz=new_xyz.numpy(),
               qg_features=new_feats.detach().numpy(), qg_xyz=gxyz.detach().numpy(), qg_w=w.numpy(),
               qg_feats_grad=feats.grad.numpy())
    # three_nn / three_interpolate / FP module
    unknown = xyz
    known = new_xyz
    kfeat = torch.from_numpy(rng.normal(size=(B, C, P)).astype(np.float32)).requires_grad_(True)
    dist, idx = pu.three_nn(unknown, known)
    dr = 1.0 / (dist + 1e-8)
    weight = dr / dr.sum(dim=2, keepdim=True)
    interp = pu.three_interpolate(kfeat, idx, weight)
    w2 = torch.from_numpy(rng.normal(size=tuple(interp.shape)).astype(np.float32))
    (interp * w2).sum().backward()
    out.update(tnn_dist=dist.numpy(), tnn_idx=idx.numpy(), ti_kfeat=kfeat.detach().numpy(), ti_weight=weight.numpy(),
               ti_out=interp.detach().numpy(), ti_w=w2.numpy(), ti_grad=kfeat.grad.numpy())
    # SA-Votes module (the live configuration in miniature: use_xyz False, normalize_xyz True, bn False)
    torch.manual_seed(5)
    sa = ns.pointnet2_modules.PointnetSAModuleVotes(npoint=P, radius=0.4, nsample=S, mlp=[C, 16, 12], use_xyz=False,
                                                    normalize_xyz=True, bn=False)
    f2 = feats.detach().clone().requires_grad_(True)
    sxyz, sfeat, sinds = sa(xyz, f2)
    w3 = torch.from_numpy(rng.normal(size=tuple(sfeat.shape)).astype(np.float32))
    (sfeat * w3).sum().backward()
    out.update(sa_w0=sa.mlp_module[0].weight.detach().numpy(), sa_b0=sa.mlp_module[0].bias.detach().numpy(),
               sa_w1=sa.mlp_module[2].weight.detach().numpy(), sa_b1=sa.mlp_module[2].bias.detach().numpy(),
               sa_xyz=sxyz.numpy(), sa_feat=sfeat.detach().numpy(), sa_inds=sinds.numpy(), sa_w=w3.numpy(),
               sa_feats_grad=f2.grad.numpy(), sa_w0_grad=sa.mlp_module[0].weight.grad.numpy())
    # knn + graph offset (dead in the live path, named by north_star): reference torch implementation
    xk = torch.from_numpy(rng.normal(size=(2, 3, 64)).astype(np.float32))
    kidx = ns.vn_dgcnn_util.knn(xk, 8)
    goff = ns.vn_dgcnn_util.get_graph_offset(xk, k=8, idx=kidx)
    out.update(knn_x=xk.numpy(), knn_idx=kidx.numpy(), knn_offset=goff.numpy())
    xg = torch.from_numpy(rng.normal(size=(2, 2, 3, 64)).astype(np.float32))        # (B, d, 3, N)
    gidx = ns.vn_dgcnn_util.knn(xg.view(2, -1, 64), 8)
    out.update(gf_x=xg.numpy(), gf_idx=gidx.numpy(),
               gf_feature=ns.vn_dgcnn_util.get_graph_feature(xg, k=8, idx=gidx).numpy(),
               gf_cross=ns.vn_dgcnn_util.get_graph_feature_cross(xg, k=8, idx=gidx).numpy())
    # feature-propagation module (three_nn + three_interpolate + shared MLP without BN)
    torch.manual_seed(6)
    fp = ns.pointnet2_modules.PointnetFPModule(mlp=[C + C, 16, 8], bn=False)
    fp_out = fp(unknown, known, feats.detach(), kfeat.detach())
    out.update(fp_w0=fp.mlp[0].weight.detach().numpy(), fp_b0=fp.mlp[0].bias.detach().numpy(),
               fp_w1=fp.mlp[2].weight.detach().numpy(), fp_b1=fp.mlp[2].bias.detach().numpy(),
               fp_out=fp_out.detach().numpy())
    np.savez_compressed(osp.join(OUT, "pointnet2.npz"), **out)
    print("pointnet2.npz:", len(out), "arrays")


if __name__ == "__main__":
    which = sys.argv[1:] or ["geometry", "pointnet2", "model"]
    ns = ref_import.import_reference(pointnet2_ref.RefExt)
    if "geometry" in which:
        geometry(ns)
    if "pointnet2" in which:
        pointnet2(ns)
    if "model" in which:
        from tests.golden import make_golden_model
        make_golden_model.main(ns)
