"""Eval post-processing of P2RNet on the B200 kernels, behind the reference's `ap_helper` API.

Mirrors /root/reference/net_utils/ap_helper.py: parse_predictions (:133-255), parse_groundtruths
(:257-292), assembly_pred_map_cls (:294-350), assembly_gt_map_cls (:402-432) and APCalculator
(:24-131, with eval_det.py:259-343 / :93-123 for the per-class AP).  Same argument and return
structure (dicts of numpy arrays / lists of (class, corners(8,3), score) tuples), so
models/p2rnet/testing.py and test_epoch.py consume the results unchanged.

What moved to the GPU: box decode + 8 corners + the far-box test (analytic point-in-box instead
of 128 scipy Delaunay triangulations per scene -- 75 % of the reference's eval time), the AABB
hull, the greedy 3-D NMS for the whole batch in one launch, and the oriented-box IoU matrix the AP
matching needs (one launch per class instead of a Python call + Qhull per pair).
Only the `.cpu()` of the finished results remains on the host.
"""
import numpy as np
import torch

from . import geometry


def _cfg_get(cfg, key, default=None):
    return cfg[key] if key in cfg else default


def parse_predictions(est_data, gt_data, config_dict):
    """Decode network outputs to oriented boxes and suppress overlaps (ap_helper.py:133-255).

    est_data: center (B,K,3), size (B,K,3) log, heading (B,K,2) (sin,cos), objectness_scores (B,K,2),
    sem_cls_scores (B,K,C) CUDA tensors; gt_data['input_joints'] (B,T,J,3).
    Returns (eval_dict {'pred_mask': u8 (B,K)}, parsed {'pred_corners_3d' f64 (B,K,8,3),
    'sem_cls_probs', 'obj_prob', 'pred_sem_cls'}) as numpy, like the reference."""
    dataset_config = config_dict["dataset_config"]
    center = est_data["center"].detach()
    log_size = est_data["size"].detach()
    heading = est_data["heading"].detach()
    sem = est_data["sem_cls_scores"].detach()
    if _cfg_get(config_dict, "sample_cls", False):
        raise NotImplementedError("sample_cls=True (stochastic class sampling) is not part of the live config")
    pred_sem_cls = torch.argmax(sem, -1)
    sem_probs = torch.softmax(sem.float(), dim=-1)
    obj_prob = torch.softmax(est_data["objectness_scores"].detach().float(), dim=-1)[:, :, 1]

    joints = gt_data["input_joints"]
    if not joints.is_cuda:
        joints = joints.to(center.device, non_blocking=True)
    hip = joints[:, :, dataset_config.origin_joint_id, 0:3]
    corners, aabb, nonempty = geometry.decode_boxes(center, log_size, heading, hip,
                                                    contact=dataset_config.contact_dist_thresh)
    if not _cfg_get(config_dict, "remove_far_box", True):
        nonempty = torch.ones_like(nonempty)
    if not _cfg_get(config_dict, "use_3d_nms", True):
        raise NotImplementedError("2-D NMS branch (use_3d_nms=False) is config-dead in the reference YAML")
    cls = pred_sem_cls if _cfg_get(config_dict, "cls_nms", _cfg_get(config_dict, "use_cls_nms", False)) else None
    keep, _ = geometry.nms3d_batched(aabb, obj_prob.double(), nonempty, cls, config_dict["nms_iou"],
                                     _cfg_get(config_dict, "use_old_type_nms", False))
    pred_mask = keep.cpu().numpy()
    # the reference asserts len(pick) > 0 per scene (ap_helper.py:230)
    assert (pred_mask.sum(axis=1) > 0).all(), "a scene lost every proposal in remove_far_box / NMS"
    eval_dict = {"pred_mask": pred_mask}
    parsed = {"pred_corners_3d": corners.cpu().numpy(), "sem_cls_probs": sem_probs.cpu().numpy(),
              "obj_prob": obj_prob.cpu().numpy(), "pred_sem_cls": pred_sem_cls.cpu().numpy()}
    return eval_dict, parsed


def parse_groundtruths(gt_data, config_dict):
    """GT labels -> oriented corners (ap_helper.py:257-292)."""
    center = gt_data["center_label"][:, :, 0:3].detach()
    dev = center.device if center.is_cuda else torch.device("cuda", torch.cuda.current_device())
    center = center.to(dev).float()
    log_size = gt_data["size"].detach().to(dev).float()
    heading = gt_data["heading"].detach().to(dev)
    mask = gt_data["box_label_mask"].detach().cpu().numpy()
    b, k2, _ = center.shape
    # decode_boxes wants a hip trajectory for the far-box test; GT boxes skip that test
    dummy_hip = center[:, :1, :].contiguous()
    corners, _, _ = geometry.decode_boxes(center, log_size, heading, dummy_hip, contact=0.0)
    corners = corners.cpu().numpy()
    corners[mask == 0] = 0.0
    return {"sem_cls_label": gt_data["sem_cls_label"], "gt_corners_3d": corners, "box_label_mask": mask}


def assembly_pred_map_cls(eval_dict, parsed_predictions, config_dict, mesh_outputs=None, voxel_size=0.047):
    """ap_helper.py:294-350 (mesh branch out of scope)."""
    assert mesh_outputs is None, "mesh evaluation is outside the hot path"
    corners = parsed_predictions["pred_corners_3d"]
    sem_probs = parsed_predictions["sem_cls_probs"]
    obj_prob = parsed_predictions["obj_prob"]
    pred_mask = eval_dict["pred_mask"]
    pred_sem_cls = parsed_predictions["pred_sem_cls"]
    bsize, n_prop = pred_sem_cls.shape
    out = []
    for i in range(bsize):
        sel = [j for j in range(n_prop) if pred_mask[i, j] == 1 and obj_prob[i, j] > config_dict["conf_thresh"]]
        if config_dict["per_class_proposal"]:
            cur = []
            for c in range(config_dict["dataset_config"].num_class):
                cur += [(c, corners[i, j], sem_probs[i, j, c] * obj_prob[i, j]) for j in sel]
        else:
            cur = [(pred_sem_cls[i, j].item(), corners[i, j], obj_prob[i, j]) for j in sel]
        out.append(cur)
    eval_dict["batch_pred_map_cls"] = out
    return eval_dict


def assembly_gt_map_cls(parsed_gts, mesh_outputs=None, voxel_size=0.047):
    """ap_helper.py:402-432 (mesh branch out of scope)."""
    assert mesh_outputs is None
    sem = parsed_gts["sem_cls_label"]
    corners = parsed_gts["gt_corners_3d"]
    mask = parsed_gts["box_label_mask"]
    return [[(sem[i, j].item(), corners[i, j]) for j in range(corners.shape[1]) if mask[i, j] == 1]
            for i in range(sem.shape[0])]


def _voc_ap(rec, prec):
    mrec = np.concatenate(([0.0], rec, [1.0]))
    mpre = np.concatenate(([0.0], prec, [0.0]))
    for i in range(mpre.size - 1, 0, -1):
        mpre[i - 1] = np.maximum(mpre[i - 1], mpre[i])
    i = np.where(mrec[1:] != mrec[:-1])[0]
    return np.sum((mrec[i + 1] - mrec[i]) * mpre[i + 1])


# Pairwise IoU is only ever needed between a detection and the ground-truth boxes of ITS OWN scene.  Scenes are packed
# into chunks whose (detections x ground truths) rectangle stays below this many entries, one IoU launch per chunk:
# the waste is the off-diagonal part of each chunk instead of the whole (all detections x all ground truths) matrix,
# which at 10 k scenes would be ~2 GB of float64 per class.
_IOU_CHUNK_ENTRIES = 1 << 22


def _same_scene_iou_rows(pred, gt):
    """pred: {img: [(box, score), ...]}, gt: {img: [box, ...]} of one class -> {img: (n_det_img, n_gt_img) float64
    ndarray} for the scenes that have both detections and ground truth."""
    imgs = [img for img in pred if len(pred[img]) and len(gt.get(img, ()))]
    out = {}
    start = 0
    while start < len(imgs):
        end, np_, ng_ = start, 0, 0
        while end < len(imgs):
            a, b = len(pred[imgs[end]]), len(gt[imgs[end]])
            if end > start and (np_ + a) * (ng_ + b) > _IOU_CHUNK_ENTRIES:
                break
            np_, ng_, end = np_ + a, ng_ + b, end + 1
        chunk = imgs[start:end]
        det = np.stack([box for img in chunk for box, _ in pred[img]])
        gtb = np.stack([box for img in chunk for box in gt[img]])
        iou = geometry.box3d_iou_matrix(det, gtb)[0].cpu().numpy()
        r = c = 0
        for img in chunk:
            a, b = len(pred[img]), len(gt[img])
            out[img] = iou[r:r + a, c:c + b]
            r, c = r + a, c + b
        start = end
    return out


def _eval_class(pred, gt, ovthresh):
    """eval_det.py:259-343 for one class; the OBB IoUs come from the GPU, a launch per chunk of scenes."""
    npos = sum(len(v) for v in gt.values())
    claimed = {img: [False] * len(v) for img, v in gt.items()}
    ids, conf, local = [], [], []
    for img, lst in pred.items():
        for j, (box, score) in enumerate(lst):
            ids.append(img)
            conf.append(score)
            local.append(j)                      # row of this detection in its scene's IoU block
    nd = len(ids)
    tp = np.zeros(nd)
    fp = np.zeros(nd)
    if nd:
        blocks = _same_scene_iou_rows(pred, gt)
        order = np.argsort(-np.asarray(conf))
        for d, k in enumerate(order):
            img = ids[k]
            ovmax, jmax = -np.inf, -1
            if img in blocks:
                row = blocks[img][local[k]]
                for j in range(row.shape[0]):  # first strict maximum, like the reference's `if iou > ovmax`
                    if row[j] > ovmax:
                        ovmax, jmax = row[j], j
            if ovmax > ovthresh and not claimed[img][jmax]:
                tp[d] = 1.0
                claimed[img][jmax] = True
            else:
                fp[d] = 1.0
    fp = np.cumsum(fp)
    tp = np.cumsum(tp)
    with np.errstate(divide="ignore", invalid="ignore"):
        rec = tp / float(npos)
    prec = tp / np.maximum(tp + fp, np.finfo(np.float64).eps)
    return rec, prec, _voc_ap(rec, prec)


class APCalculator(object):
    """VOC-style AP over accumulated scenes (ap_helper.py:24-92)."""

    def __init__(self, ap_iou_thresh=0.25, class2type_map=None, evaluate_mesh=False):
        assert not evaluate_mesh, "mesh evaluation is outside the hot path"
        self.ap_iou_thresh = ap_iou_thresh
        self.class2type_map = class2type_map
        self.reset()

    def step(self, batch_pred_map_cls, batch_gt_map_cls):
        assert len(batch_pred_map_cls) == len(batch_gt_map_cls)
        for p, g in zip(batch_pred_map_cls, batch_gt_map_cls):
            self.gt_map_cls[self.scan_cnt] = g
            self.pred_map_cls[self.scan_cnt] = p
            self.scan_cnt += 1

    def merge(self, other_state):
        """Fold another rank's (pred_map_cls, gt_map_cls) in (after all_gather_object); the reference
        has no cross-rank gather -- this is what makes sharded eval exact."""
        preds, gts = other_state
        for k in sorted(preds):
            self.gt_map_cls[self.scan_cnt] = gts[k]
            self.pred_map_cls[self.scan_cnt] = preds[k]
            self.scan_cnt += 1

    def compute_metrics(self):
        pred, gt = {}, {}
        for img, lst in self.pred_map_cls.items():
            for cls, box, score in lst:
                pred.setdefault(cls, {}).setdefault(img, []).append((box, score))
                gt.setdefault(cls, {}).setdefault(img, [])
        for img, lst in self.gt_map_cls.items():
            for cls, box in lst:
                gt.setdefault(cls, {}).setdefault(img, []).append(box)
        rec, ap = {}, {}
        for cls in gt:
            if cls in pred:
                rec[cls], _, ap[cls] = _eval_class(pred[cls], gt[cls], self.ap_iou_thresh)
            else:
                rec[cls], ap[cls] = 0, 0
        ret = {}
        name = (lambda k: self.class2type_map[k]) if self.class2type_map else str
        for key in sorted(ap):
            ret["%s Average Precision" % name(key)] = ap[key]
        vals = [v for v in ap.values() if not np.isnan(v)]
        ret["mAP"] = np.mean(vals) if vals else float("nan")
        recs = []
        for key in sorted(ap):
            try:
                r = rec[key][-1]
            except Exception:
                r = 0
            ret["%s Recall" % name(key)] = r
            recs.append(r)
        recs = [r for r in recs if not np.isnan(r)]
        ret["AR"] = np.mean(recs) if recs else float("nan")
        return ret

    def reset(self):
        self.gt_map_cls = {}
        self.pred_map_cls = {}
        self.scan_cnt = 0
