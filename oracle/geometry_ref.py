"""numpy restatement of the reference's loss-side and eval-side geometry (CPU oracle).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg, never by pose2room_b200/.

Restates (citations into /root/reference):
  nn_distance / huber_loss        net_utils/nn_distance.py:15-61
  nms_2d_faster / nms_3d_faster / nms_3d_faster_samecls   net_utils/nms.py:7-119
  head2rot / get_3d_box           utils/pc_utils.py:22-27,50-67 ; get_box_corners utils/tools.py:33-51
  softmax                         net_utils/libs.py:75-80
  box3d_iou (+polygon_clip, poly_area, box3d_vol)         net_utils/box_util.py:17-118
  parse_predictions (decode, remove_far_box, 3-D NMS)     net_utils/ap_helper.py:133-255
  voc_ap / eval_det_cls           net_utils/eval_det.py:93-123,259-343

Pinned against: demo_nn_distance (nn_distance.py:63-94, the reference's only known answer),
the survey-generated answers of SURVEY.md Appendix D, and tests/golden/*.npz produced by running
the reference itself (tests/golden/make_golden.py).  Two deliberate differences from the
reference, both documented where they occur: the point-in-box test is analytic instead of
scipy Delaunay, and the clipped-polygon area is a shoelace sum instead of scipy ConvexHull.volume
(identical up to rounding for the convex polygons Sutherland-Hodgman emits).
"""
import numpy as np


# ----------------------------------------------------------------------------- losses
def huber(err, delta=1.0):
    a = np.abs(err)
    q = np.minimum(a, np.float32(delta)).astype(err.dtype)
    lin = a - q
    return (np.float32(0.5) * q * q + np.float32(delta) * lin).astype(err.dtype)


def nn_distance(pc1, pc2, l1smooth=False, delta=1.0, l1=False):
    """pc1 (B,N,C), pc2 (B,M,C) float32 -> dist1 (B,N), idx1 (B,N) i64, dist2 (B,M), idx2 (B,M) i64.
    Per-pair cost summed over C in index order (fp32, no fused multiply-add), first minimal index
    wins ties (torch.min semantics, nn_distance.py:59-60)."""
    pc1 = np.asarray(pc1, np.float32)
    pc2 = np.asarray(pc2, np.float32)
    diff = pc1[:, :, None, :] - pc2[:, None, :, :]
    if l1smooth:
        e = huber(diff, delta)
    elif l1:
        e = np.abs(diff)
    else:
        e = diff * diff
    dist = e[..., 0].copy()
    for c in range(1, e.shape[-1]):
        dist = dist + e[..., c]
    idx1 = np.argmin(dist, axis=2).astype(np.int64)
    idx2 = np.argmin(dist, axis=1).astype(np.int64)
    dist1 = np.min(dist, axis=2)
    dist2 = np.min(dist, axis=1)
    return dist1, idx1, dist2, idx2


# ----------------------------------------------------------------------------- boxes
def head2rot(theta):
    c, s = np.cos(theta), np.sin(theta)
    return np.array([[c, 0.0, -s], [0.0, 1.0, 0.0], [s, 0.0, c]], dtype=np.float64)


_SIGNS = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1],
                   [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], dtype=np.float64)


def box_corners(center, vectors):
    """tools.py:33-51: corner i = center + s0*v0 + s1*v1 + s2*v2 with the sign table above,
    evaluated left to right like the reference ((c -/+ v0) -/+ v1) -/+ v2."""
    center = np.asarray(center, np.float64)
    out = np.empty((8, 3), np.float64)
    for i, (s0, s1, s2) in enumerate(_SIGNS):
        out[i] = ((center + s0 * vectors[0]) + s1 * vectors[1]) + s2 * vectors[2]
    return out


def get_3d_box(size, theta, center):
    vectors = np.diag(np.asarray(size, np.float64) / 2.0).dot(head2rot(float(theta)))
    return box_corners(center, vectors)


def softmax(x):
    p = np.exp(x - np.max(x, axis=-1, keepdims=True))
    return p / np.sum(p, axis=-1, keepdims=True)


def points_in_obb(points, center, size_half, theta):
    """Analytic stand-in for scipy Delaunay(box).find_simplex(p) >= 0 (libs.py:97-101):
    |R (p - c)| <= half extents, component-wise."""
    R = head2rot(float(theta))
    local = (np.asarray(points, np.float64) - np.asarray(center, np.float64)) @ R.T
    return np.all(np.abs(local) <= np.asarray(size_half, np.float64)[None, :], axis=1)


def _clip(subject, clipper):
    """Sutherland-Hodgman, same inside test / intersection formula as box_util.py:22-61."""
    out = list(subject)
    cp1 = clipper[-1]
    for cp2 in clipper:
        inp, out = out, []
        if not inp:
            return None
        s = inp[-1]

        def inside(p):
            return (cp2[0] - cp1[0]) * (p[1] - cp1[1]) > (cp2[1] - cp1[1]) * (p[0] - cp1[0])

        def cross(s, e):
            dc = (cp1[0] - cp2[0], cp1[1] - cp2[1])
            dp = (s[0] - e[0], s[1] - e[1])
            n1 = cp1[0] * cp2[1] - cp1[1] * cp2[0]
            n2 = s[0] * e[1] - s[1] * e[0]
            n3 = 1.0 / (dc[0] * dp[1] - dc[1] * dp[0])
            return ((n1 * dp[0] - n2 * dc[0]) * n3, (n1 * dp[1] - n2 * dc[1]) * n3)

        for e in inp:
            if inside(e):
                if not inside(s):
                    out.append(cross(s, e))
                out.append(e)
            elif inside(s):
                out.append(cross(s, e))
            s = e
        cp1 = cp2
        if len(out) == 0:
            return None
    return out


def _shoelace(poly):
    x = np.array([p[0] for p in poly])
    y = np.array([p[1] for p in poly])
    return 0.5 * np.abs(np.dot(x, np.roll(y, 1)) - np.dot(y, np.roll(x, 1)))


def box3d_iou(c1, c2):
    """box_util.py:90-118.  Returns (iou3d, iou2d)."""
    perm = [7, 6, 2, 3, 4, 5, 1, 0]
    a = np.asarray(c1, np.float64)[perm]
    b = np.asarray(c2, np.float64)[perm]
    r1 = [(a[i, 0], a[i, 2]) for i in range(3, -1, -1)]
    r2 = [(b[i, 0], b[i, 2]) for i in range(3, -1, -1)]
    area1, area2 = _shoelace(r1), _shoelace(r2)
    inter = _clip(r1, r2)
    inter_area = _shoelace(inter) if inter is not None and len(inter) >= 3 else 0.0
    iou2d = inter_area / (area1 + area2 - inter_area)
    ymax = min(a[0, 1], b[0, 1])
    ymin = max(a[4, 1], b[4, 1])
    inter_vol = inter_area * max(0.0, ymax - ymin)

    def vol(c):
        return (np.sqrt(np.sum((c[0] - c[1]) ** 2)) * np.sqrt(np.sum((c[1] - c[2]) ** 2)) *
                np.sqrt(np.sum((c[0] - c[4]) ** 2)))
    return inter_vol / (vol(a) + vol(b) - inter_vol), iou2d


# ----------------------------------------------------------------------------- NMS
def _nms(lo, hi, score, thr, old_type, cls=None):
    """Shared body of nms.py:7-119: sort ascending by score, repeatedly take the last, drop
    everything whose overlap with it is > thr (strict)."""
    vol = np.prod(hi - lo, axis=1)
    order = np.argsort(score)
    pick = []
    while order.size:
        i = order[-1]
        pick.append(int(i))
        rest = order[:-1]
        ext = np.maximum(0, np.minimum(hi[i], hi[rest]) - np.maximum(lo[i], lo[rest]))
        inter = np.prod(ext, axis=1) if ext.shape[1] == 2 else ext[:, 0] * ext[:, 1] * ext[:, 2]
        o = inter / vol[rest] if old_type else inter / (vol[i] + vol[rest] - inter)
        if cls is not None:
            o = o * (cls[i] == cls[rest])
        order = rest[~(o > thr)]
    return pick


def nms_2d_faster(boxes, thr, old_type=False):
    boxes = np.asarray(boxes)
    return _nms(boxes[:, 0:2], boxes[:, 2:4], boxes[:, 4], thr, old_type)


def nms_3d_faster(boxes, thr, old_type=False):
    boxes = np.asarray(boxes)
    return _nms(boxes[:, 0:3], boxes[:, 3:6], boxes[:, 6], thr, old_type)


def nms_3d_faster_samecls(boxes, thr, old_type=False):
    boxes = np.asarray(boxes)
    return _nms(boxes[:, 0:3], boxes[:, 3:6], boxes[:, 6], thr, old_type, cls=boxes[:, 7])


# ----------------------------------------------------------------------------- eval post-processing
def parse_predictions(center, log_size, heading_sincos, objectness, sem_cls, hip_traj,
                      nms_iou=0.10, contact=1.0, old_type=False):
    """ap_helper.py:133-255 with the test YAML's live switches (use_3d_nms, not cls_nms,
    remove_far_box, sample_cls False; p2rnet_test.yaml:33-45).
    Inputs are the network outputs as numpy (float32, heading possibly float64) and the hip
    trajectory (B,T,3).  Returns dict(pred_mask u8 (B,K), corners f64 (B,K,8,3), obj_prob,
    sem_cls_probs, pred_sem_cls, nonempty)."""
    center = np.asarray(center)
    B, K = center.shape[:2]
    import torch  # the reference decodes with torch.exp / torch.atan2 (ap_helper.py:153-155); mirror it bit for bit
    size = torch.exp(torch.as_tensor(np.asarray(log_size))).numpy()
    hs = torch.as_tensor(np.asarray(heading_sincos))
    theta = torch.atan2(hs[..., 0], hs[..., 1]).numpy()
    sem_probs = softmax(np.asarray(sem_cls))
    pred_cls = np.argmax(np.asarray(sem_cls), -1)
    obj_prob = softmax(np.asarray(objectness))[:, :, 1]
    corners = np.zeros((B, K, 8, 3))
    nonempty = np.ones((B, K))
    for i in range(B):
        for j in range(K):
            corners[i, j] = get_3d_box(size[i, j], theta[i, j], center[i, j])
            if np.any(size[i, j] < 0.01) or np.any(size[i, j] > 10):
                nonempty[i, j] = 0
                continue
            half = (size[i, j] / 2. + contact).astype(np.float64)  # float32 arithmetic like ap_helper.py:192
            if not points_in_obb(hip_traj[i], center[i, j], half, theta[i, j]).any():
                nonempty[i, j] = 0
    pred_mask = np.zeros((B, K), np.uint8)
    for i in range(B):
        keep = np.where(nonempty[i] == 1)[0]
        lo = corners[i].min(axis=1)
        hi = corners[i].max(axis=1)
        boxes = np.concatenate([lo, hi, obj_prob[i][:, None]], axis=1)
        if keep.size:
            pick = nms_3d_faster(boxes[keep], nms_iou, old_type)
            pred_mask[i, keep[pick]] = 1
    return dict(pred_mask=pred_mask, corners=corners, obj_prob=obj_prob, sem_cls_probs=sem_probs,
                pred_sem_cls=pred_cls, nonempty=nonempty)


# ----------------------------------------------------------------------------- AP
def voc_ap(rec, prec):
    """eval_det.py:93-123 with use_07_metric=False: area under the monotone precision envelope."""
    mrec = np.concatenate(([0.0], rec, [1.0]))
    mpre = np.concatenate(([0.0], prec, [0.0]))
    for i in range(mpre.size - 1, 0, -1):
        mpre[i - 1] = max(mpre[i - 1], mpre[i])
    i = np.where(mrec[1:] != mrec[:-1])[0]
    return float(np.sum((mrec[i + 1] - mrec[i]) * mpre[i + 1]))


def eval_det_cls(pred, gt, ovthresh, iou_fn=None):
    """eval_det.py:259-343 restated.  pred: {img: [(corners, score)]}, gt: {img: [corners]}.
    Greedy matching in descending score order; each GT box may be claimed once."""
    iou_fn = iou_fn or (lambda a, b: box3d_iou(a, b)[0])
    npos = 0
    claimed = {}
    for img, boxes in gt.items():
        npos += len(boxes)
        claimed[img] = [False] * len(boxes)
    for img in pred:
        if img not in gt:
            gt[img] = []
            claimed[img] = []
    ids, conf, bb = [], [], []
    for img, lst in pred.items():
        for box, score in lst:
            ids.append(img)
            conf.append(score)
            bb.append(box)
    conf = np.array(conf)
    order = np.argsort(-conf)
    tp = np.zeros(len(ids))
    fp = np.zeros(len(ids))
    for d, k in enumerate(order):
        img = ids[k]
        best, jbest = -np.inf, -1
        for j, g in enumerate(gt[img]):
            iou = iou_fn(bb[k], g)
            if iou > best:
                best, jbest = iou, j
        if best > ovthresh and not claimed[img][jbest]:
            tp[d] = 1.0
            claimed[img][jbest] = True
        else:
            fp[d] = 1.0
    fp = np.cumsum(fp)
    tp = np.cumsum(tp)
    with np.errstate(divide="ignore", invalid="ignore"):
        rec = tp / float(npos)  # npos == 0 -> nan, exactly like the reference (class then drops out of mAP)
    prec = tp / np.maximum(tp + fp, np.finfo(np.float64).eps)
    return rec, prec, voc_ap(rec, prec)


def eval_map(batch_pred_map_cls, batch_gt_map_cls, ovthresh):
    """APCalculator.step + compute_metrics (ap_helper.py:39-92, eval_det.py:424-473) restated:
    per-class AP over all scenes, mAP = mean over classes that appear in the GT."""
    pred, gt = {}, {}
    for img, (plist, glist) in enumerate(zip(batch_pred_map_cls, batch_gt_map_cls)):
        for cls, box, score in plist:
            pred.setdefault(cls, {}).setdefault(img, []).append((box, score))
            gt.setdefault(cls, {})
        for cls, box in glist:
            gt.setdefault(cls, {}).setdefault(img, []).append(box)
            pred.setdefault(cls, {})
    ap = {}
    for cls in gt:
        if cls in pred and len(pred[cls]):
            _, _, ap[cls] = eval_det_cls(pred[cls], gt[cls], ovthresh)
        else:
            ap[cls] = 0.0
    vals = [v for v in ap.values() if not np.isnan(v)]  # ap_helper.py:79 filters NaN classes
    return ap, (float(np.mean(vals)) if vals else float("nan"))
