"""CPU: oracle/geometry_ref.py against the reference's own known answers and committed goldens."""
import numpy as np
import pytest

from oracle import geometry_ref as G


def test_demo_nn_distance_known_answer():
    # net_utils/nn_distance.py:63-94 with np.random.seed(0)
    np.random.seed(0)
    pc1 = np.random.random((1, 5, 3)).astype(np.float32)
    pc2 = np.random.random((1, 6, 3)).astype(np.float32)
    d1, i1, _, _ = G.nn_distance(pc1, pc2)
    assert np.allclose(d1[0], [0.1058, 0.0842, 0.1167, 0.0197, 0.2633], atol=5e-5)
    assert i1[0].tolist() == [5, 2, 1, 5, 5]
    d1, i1, _, _ = G.nn_distance(pc1, pc2, l1smooth=True)
    assert np.allclose(d1[0], [0.0529, 0.0421, 0.0583, 0.0099, 0.1317], atol=5e-5)
    assert i1[0].tolist() == [5, 2, 1, 5, 5]


@pytest.mark.parametrize("case", ["demo", "a", "b"])
@pytest.mark.parametrize("mode,kw", [("l2", {}), ("l1s", dict(l1smooth=True)), ("l1", dict(l1=True))])
def test_nn_distance_bit_exact_vs_reference_goldens(golden_geometry, case, mode, kw):
    g = golden_geometry
    r = G.nn_distance(g["nnd_%s_pc1" % case], g["nnd_%s_pc2" % case], **kw)
    for key, got in zip(["d1", "i1", "d2", "i2"], r):
        assert np.array_equal(got, g["nnd_%s_%s_%s" % (case, mode, key)]), key


def test_appendix_d_known_answers():
    size = np.array([1.0, 2.0, 0.5])
    A = G.get_3d_box(size, 0.3, np.array([0.0, 1.0, 0.0]))
    B = G.get_3d_box(size, 0.3, np.array([0.2, 1.0, 0.1]))
    C = G.get_3d_box(size, -1.2, np.array([0.1, 1.5, 0.0]))
    D = G.get_3d_box(size, 0.3, np.array([5.0, 1.0, 0.0]))
    assert np.allclose(G.box3d_iou(A, B), (0.40762005092979847, 0.4076200509297984), atol=1e-12)
    assert np.allclose(G.box3d_iou(A, C), (0.23148291421667108, 0.33445040365524775), atol=1e-12)
    assert G.box3d_iou(A, D) == (0.0, 0.0)
    boxes = np.array([[0, 0, 0, 1, 1, 1, .9], [.1, .1, .1, 1.1, 1.1, 1.1, .8], [2, 2, 2, 3, 3, 3, .7], [0, 0, 0, 1, 1, 1, .9]])
    assert G.nms_3d_faster(boxes, 0.1) == [3, 2]


def test_boxes_and_iou_vs_reference_goldens(golden_geometry):
    g = golden_geometry
    for i in range(24):
        c = G.get_3d_box(g["box_size"][i], g["box_theta"][i], g["box_center"][i])
        assert np.array_equal(c, g["box_corners"][i])
    # i == j (coincident parallel edges) is excluded: there the reference's clipper divides by zero and
    # returns garbage (IoU 2.13, 11.66 ... in this very fixture) or raises inside Qhull.
    ok = ~np.isnan(g["box_iou3d"]) & ~np.eye(24, dtype=bool)
    got3 = np.zeros((24, 24))
    got2 = np.zeros((24, 24))
    for i in range(24):
        for j in range(24):
            if ok[i, j]:
                got3[i, j], got2[i, j] = G.box3d_iou(g["box_corners"][i], g["box_corners"][j])
    assert np.allclose(got3[ok], g["box_iou3d"][ok], atol=1e-9)
    assert np.allclose(got2[ok], g["box_iou2d"][ok], atol=1e-9)
    assert (g["box_iou3d"][ok] > 0.05).sum() > 20  # the fixture exercises real overlaps


@pytest.mark.parametrize("t", range(4))
def test_nms_selection_exact_vs_reference_goldens(golden_geometry, t):
    g = golden_geometry
    boxes = g["nms%d_boxes" % t]
    assert G.nms_3d_faster(boxes[:, :7], 0.10) == g["nms%d_pick" % t].tolist()
    assert G.nms_3d_faster(boxes[:, :7], 0.25, old_type=True) == g["nms%d_pick_old" % t].tolist()
    assert G.nms_3d_faster_samecls(boxes, 0.10) == g["nms%d_pick_cls" % t].tolist()
    assert G.nms_2d_faster(boxes[:, [0, 1, 3, 4, 6]], 0.10) == g["nms%d_pick_2d" % t].tolist()


def test_parse_predictions_vs_reference_goldens(golden_geometry):
    g = golden_geometry
    hip = g["pp_input_joints"][:, :, 0]
    r = G.parse_predictions(g["pp_center"], g["pp_size"], g["pp_heading"], g["pp_objectness"], g["pp_sem"], hip)
    assert np.allclose(r["corners"], g["pp_corners"], atol=1e-12)
    assert np.allclose(r["obj_prob"], g["pp_obj_prob"], atol=1e-7)
    assert np.array_equal(r["pred_sem_cls"], g["pp_pred_sem_cls"])
    # analytic point-in-box vs scipy Delaunay: identical NMS selection on this fixture
    assert np.array_equal(r["pred_mask"], g["pp_pred_mask"])
    assert r["pred_mask"][:, :14].sum() == 0 and r["pred_mask"].sum() > 8


def test_ap_vs_reference_goldens(golden_geometry):
    g = golden_geometry
    hip = g["pp_input_joints"][:, :, 0]
    r = G.parse_predictions(g["pp_center"], g["pp_size"], g["pp_heading"], g["pp_objectness"], g["pp_sem"], hip)
    B, K = r["pred_mask"].shape
    pred_map, gt_map = [], []
    for i in range(B):
        cur = []
        for c in range(22):
            cur += [(c, r["corners"][i, j], r["sem_cls_probs"][i, j, c] * r["obj_prob"][i, j])
                    for j in range(K) if r["pred_mask"][i, j] == 1 and r["obj_prob"][i, j] > 0.05]
        pred_map.append(cur)
        gts = []
        for j in range(10):
            if g["pp_gt_box_label_mask"][i, j] == 1:
                hs = g["pp_gt_heading"][i, j]
                box = G.get_3d_box(np.exp(g["pp_gt_size"][i, j]), np.arctan2(hs[0], hs[1]), g["pp_gt_center_label"][i, j])
                gts.append((int(g["pp_gt_sem_cls_label"][i, j]), box))
        gt_map.append(gts)
    for thr in [0.25, 0.5]:
        ap, m = G.eval_map(pred_map, gt_map, thr)
        want = g["ap_%d" % int(thr * 100)]
        for c in range(22):
            if np.isnan(want[c]):
                assert c not in ap or np.isnan(ap[c])
            else:
                assert abs(ap[c] - want[c]) < 1e-9, (thr, c)
        assert abs(m - np.nanmean(want)) < 1e-9


@pytest.mark.needs_reference
def test_oracle_vs_live_reference_random():
    """In the build container: run the reference itself next to the oracle on fresh random inputs."""
    import torch
    from oracle import pointnet2_ref, ref_import
    ns = ref_import.import_reference(pointnet2_ref.RefExt)
    rng = np.random.default_rng(123)
    for _ in range(3):
        a = rng.normal(size=(2, 40, 3)).astype(np.float32)
        b = rng.normal(size=(2, 9, 3)).astype(np.float32)
        for kw in [{}, dict(l1smooth=True, delta=0.5), dict(l1=True)]:
            ref = ns.nn_distance.nn_distance(torch.from_numpy(a), torch.from_numpy(b), **kw)
            got = G.nn_distance(a, b, **kw)
            for x, y in zip(ref, got):
                assert np.array_equal(x.numpy(), y)
        k = 60
        lo = rng.normal(size=(k, 3))
        boxes = np.concatenate([lo, lo + rng.uniform(0.1, 2, size=(k, 3)), rng.uniform(size=(k, 1))], 1)
        assert ns.nms.nms_3d_faster(boxes, 0.1) == G.nms_3d_faster(boxes, 0.1)
