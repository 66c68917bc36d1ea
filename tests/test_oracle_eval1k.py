"""CPU: the numpy oracle of the eval path (oracle/geometry_ref.py) against what the UNMODIFIED reference produced on the
1000-scene evaluation set of BASELINE.json config #5 (tests/golden/eval1k.npz): NMS / far-box selection exact, corner
checksums, per-class AP and mAP at IoU 0.25 / 0.5.  The product's GPU path is held to the same fixture by
tests/test_zz_eval_1k_gpu.py."""
import numpy as np

from oracle import geometry_ref as G
from pose2room_b200 import synthetic
from tests import eval1k_helpers as H


def test_fixture_is_the_full_non_trivial_set():
    g, mask = H.load()
    assert int(g["n_scenes"]) == 1000 and mask.shape == (1000, H.K)
    assert 0.2 < float(g["map_25"]) < 0.8 and 0.1 < float(g["map_50"]) < float(g["map_25"])     # far from 0 and 1
    kept = mask.sum(1)
    assert kept.min() >= 1 and 3 < kept.mean() < 40                # NMS removed the duplicates, something survived
    assert mask[:, :6].sum() == 0                                  # far / degenerate proposals never survive


def _oracle_run(g, n, chunk=50):
    masks, checks, pred_map, gt_map = [], [], [], []
    for start in range(0, n, chunk):
        est, gt = synthetic.make_eval_batch(int(g["seed"]), start, min(chunk, n - start))
        r = G.parse_predictions(est["center"].numpy(), est["size"].numpy(), est["heading"].numpy(),
                                est["objectness_scores"].numpy(), est["sem_cls_scores"].numpy(),
                                gt["input_joints"][:, :, 0].numpy())
        masks.append(r["pred_mask"])
        checks.append(np.abs(r["corners"]).sum(axis=(1, 2, 3)))
        for i in range(r["pred_mask"].shape[0]):
            cur = []
            for c in range(22):
                cur += [(c, r["corners"][i, j], r["sem_cls_probs"][i, j, c] * r["obj_prob"][i, j])
                        for j in range(H.K) if r["pred_mask"][i, j] == 1 and r["obj_prob"][i, j] > 0.05]
            pred_map.append(cur)
            gts = []
            for j in range(10):
                if gt["box_label_mask"][i, j] == 1:
                    hs = gt["heading"][i, j].numpy()
                    gts.append((int(gt["sem_cls_label"][i, j]),
                                G.get_3d_box(np.exp(gt["size"][i, j].numpy()), np.arctan2(hs[0], hs[1]),
                                             gt["center_label"][i, j].numpy())))
            gt_map.append(gts)
    return np.concatenate(masks), np.concatenate(checks), pred_map, gt_map


def test_oracle_selection_and_ap_match_the_reference_on_all_1000_scenes():
    """~40 s: analytic point-in-box + numpy NMS + polygon-clipping IoU (the oracle) against scipy Delaunay + Qhull (the
    reference) -- selection bit-exact on 128 000 proposals, per-class AP / mAP to 1e-9 at both thresholds, and the same for
    the first-`subset`-scenes numbers the fixture also carries."""
    g, mask = H.load()
    n, sub = int(g["n_scenes"]), int(g["subset"])
    got_mask, checks, pred_map, gt_map = _oracle_run(g, n)
    assert np.array_equal(got_mask, mask)
    assert np.allclose(checks, g["corner_abs_sum"], rtol=1e-12)
    assert [len(p) for p in pred_map] == g["n_pred"].tolist()
    for thr in (0.25, 0.5):
        tag = "%d" % int(thr * 100)
        ap, m = G.eval_map(pred_map, gt_map, thr)
        H.check_ap(ap, m, g["ap_" + tag], float(g["map_" + tag]), 1e-9)
        ap, m = G.eval_map(pred_map[:sub], gt_map[:sub], thr)
        H.check_ap(ap, m, g["ap_sub_" + tag], float(g["map_sub_" + tag]), 1e-9)
